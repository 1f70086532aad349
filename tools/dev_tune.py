"""Tuning probe (GPU box): C2 throughput for kernel variants / scheduler thresholds.
usage: dev_tune.py legacy:TW ... pool:TW:TWI ..."""
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
for spec in (sys.argv[1:] or ["legacy:4", "pool:32:16"]):
    parts = spec.split(":")
    os.environ["ERTB_KERNEL"] = parts[0]
    if parts[0] == "legacy":
        os.environ["ERTB_TW"] = parts[1]
    else:
        os.environ["ERTB_POOL_TW"] = parts[1]; os.environ["ERTB_POOL_TWI"] = parts[2]
    best = 1e9
    for rep in range(3):
        wl, l, l2, st = dev.render(0, 2 + rep, spp)
        best = min(best, st.device_ms)
    k = (st.trips_main + st.trips_nee) / st.n_paths
    print(f"{spec:14s}: {best:8.3f} ms  {32*spp/best/1e3:8.1f} Mpaths/s  L0={l[0]/spp:.5f} K={k:.2f} n={st.n_paths}")
