"""GPU probe: plane-parallel C2-like scene, volpath (global majorant) vs piecewise_volpath."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
g.build_cuda()
from eradiate_b200 import scenes
from eradiate_b200.kernel import mi_load_dict
from eradiate_b200.kernel._render import _device_scene

spp = 1 << 20
sens = {"type": "mdistant", "vza": np.linspace(-75.0, 75.0, 32), "vaa": 0.0}
for kw in (dict(), dict(aerosol=True), dict(stokes=True, phase={"type": "rayleigh_polarized"})):
    for integ in ("volpath", "piecewise_volpath"):
        sc = mi_load_dict(scenes.atmosphere_scene(geometry="plane_parallel", integrator=integ, sensor=sens, **kw))
        dev = _device_scene(sc)
        dev.render(0, 1, 1 << 14)
        best = 1e9
        for rep in range(3):
            out = dev.render(0, 2 + rep, spp)
            st = out[-1]
            best = min(best, st.device_ms)
        l = out[1]
        k = (st.trips_main + st.trips_nee) / st.n_paths
        print(f"{str(sorted(kw))[:30]:30s} {integ:18s}: {best:8.3f} ms {32*spp/best/1e3:9.1f} Mpaths/s "
              f"L0={l[0]/spp:.5f} L16={l[16]/spp:.5f} K={k:.2f}", flush=True)
