#!/usr/bin/env python
"""Developer probe: where the time of one end-to-end step (mi_render of one context) goes.
ERTB_TIMING=1 makes the library print the host time of each part of ertb_render."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from eradiate_b200 import scenes
from eradiate_b200.kernel import KernelContext, SeedState, mi_load_dict, mi_render, mi_traverse, render
scene = mi_load_dict(scenes.config_c2(spp=1 << 20, n_vza=32))
ms = mi_traverse(scene, scenes.spectral_update_map(1200, spherical=True))
ctx = KernelContext(w=550.0)
mode = sys.argv[1] if len(sys.argv) > 1 else "mi_render"
def step(i):
    if mode == "mi_render":
        res = mi_render(ms, [ctx], spp=1 << 20, seed_state=SeedState(i))
        return np.array(res[ctx.si.as_hashable]["measure"])
    if mode == "render_only":
        return np.array(render(scene, sensor=0, seed=i, spp=1 << 20, stats=False))
    if mode == "render_sleep":
        time.sleep(0.002)
        return np.array(render(scene, sensor=0, seed=i, spp=1 << 20, stats=False))
    if mode == "update_irradiance_only":
        ms.parameters.update({"illumination.irradiance.value": 1.8})
        return np.array(render(scene, sensor=0, seed=i, spp=1 << 20, stats=False))
for i in range(3): step(i)
t = time.perf_counter()
for i in range(30): step(i)
print(mode, "ms/step", (time.perf_counter() - t) / 30 * 1e3, flush=True)
