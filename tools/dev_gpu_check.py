"""Developer scratch check (GPU box): smoke + quick throughput probe of C1/C2."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g

g.build()
g.smoke()
from eradiate_b200 import scenes
from eradiate_b200.kernel import mi_load_dict, render

for name, sd, spp in (("C1", scenes.config_c1(), 1 << 22), ("C2", scenes.config_c2(), 1 << 16), ("C2", scenes.config_c2(), 1 << 20)):
    sc = mi_load_dict(sd)
    render(sc, 0, 1, 1024)
    t = time.time()
    bmp = render(sc, 0, 2, spp)
    dt = time.time() - t
    st = bmp.stats
    npaths = st["n_paths"]
    print(f"{name} spp={spp}: wall {dt*1e3:.1f} ms, device {st['device_ms']:.2f} ms, "
          f"{npaths/st['device_ms']/1e3:.1f} Mpaths/s, K={ (st['trips_main']+st['trips_nee'])/npaths:.2f} {st}")
    print("  L:", np.array2string(bmp.raw['sum_l'].ravel()[:8]/spp, precision=5))
