"""Developer probe: throughput of the 3D kernel on C4-like scenes."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eradiate_b200 import scenes
from eradiate_b200.kernel import mi_load_dict, render

t0 = time.perf_counter()
kd = scenes.config_c4(spp=1 << 12, lai=3.0, radius=0.1, size=(25.0, 25.0, 2.0), padding=2, n_vza=32, film=(64, 64))
sc = mi_load_dict(kd)
print("load", time.perf_counter() - t0, "s; leaves", sc.flat.leaf_groups[0].disks.shape[0], "instances", len(sc.flat.instances))
for sensor, spp in ((0, 1 << 16), (0, 1 << 18), (1, 1 << 9), (1, 1 << 11)):
    t0 = time.perf_counter()
    bmp = render(sc, sensor=sensor, seed=3, spp=spp)
    dt = time.perf_counter() - t0
    st = bmp.stats
    print(f"sensor {sensor} spp {spp}: {st['n_paths']/st['device_ms']/1e3:.1f} Mpaths/s device, wall {dt:.3f} s, "
          f"surf/path {st['n_surface']/st['n_paths']:.2f} scat/path {st['n_scatter']/st['n_paths']:.2f} "
          f"trips {(st['trips_main']+st['trips_nee'])/st['n_paths']:.1f} mean L {np.array(bmp)[..., 0].mean():.5f}")
