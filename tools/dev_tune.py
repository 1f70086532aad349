"""Tuning probe (GPU box): C2 throughput vs scheduler threshold."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
g.build_cuda()
from eradiate_b200 import scenes
from eradiate_b200.kernel import mi_load_dict
from eradiate_b200.kernel._render import _device_scene

spp = 1 << 20
sc = mi_load_dict(scenes.config_c2(spp=spp))
dev = _device_scene(sc)
dev.render(0, 1, 1 << 14)
for tw in [int(x) for x in (sys.argv[1:] or [1, 4, 8, 12, 16, 20, 24])]:
    os.environ["ERTB_TW"] = str(tw)
    best = 1e9
    for rep in range(3):
        wl, l, l2, st = dev.render(0, 2 + rep, spp)
        best = min(best, st.device_ms)
    print(f"TW={tw:3d}: {best:8.3f} ms  {32*spp/best/1e3:8.1f} Mpaths/s  L0={l[0]/spp:.5f}")
