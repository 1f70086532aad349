"""Developer probe: device time of small single-pixel renders (launch/drain overheads)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eradiate_b200 import scenes
from eradiate_b200.kernel import mi_load_dict
from eradiate_b200.kernel._render import _device_scene

for npix in (1, 32):
    vza = np.linspace(-60.0, 60.0, npix) if npix > 1 else np.array([0.0])
    sc = mi_load_dict(scenes.atmosphere_scene(sensor={"type": "mdistant", "vza": vza, "vaa": 0.0}, spp=16))
    dev = _device_scene(sc)
    dev.render(0, 1, 1 << 16)
    for lg in (14, 16, 18, 20, 22):
        spp = (1 << lg) // npix if npix > 1 else 1 << lg
        best, wall = 1e9, 1e9
        for r in range(5):
            t0 = time.perf_counter()
            st = dev.render(0, 10 + r, spp)[3]
            wall = min(wall, time.perf_counter() - t0)
            best = min(best, st.device_ms)
        n = npix * spp
        print(f"npix={npix:3d} paths=2^{lg}: device {best*1e3:8.1f} us  wall {wall*1e6:8.1f} us  "
              f"{n/best/1e3:8.1f} Mpaths/s (device)")
