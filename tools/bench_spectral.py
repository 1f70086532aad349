#!/usr/bin/env python
"""
Spectral-loop benchmark: many contexts x one sensor through ``mi_render``,
once with the reference's strictly sequential update -> render -> read-back loop
(``_render.py:433-468``) and once through the pipelined ``ertb_batch_*`` entry points.

    python tools/bench_spectral.py [--contexts 256] [--spp 1048576] [--pixels 1]

Prints one JSON line with contexts/s and Mpaths/s of both loops. Needs a GPU.
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from eradiate_b200 import scenes  # noqa: E402
from eradiate_b200.kernel import KernelContext, mi_load_dict, mi_render, mi_traverse  # noqa: E402
from eradiate_b200.kernel._render import SeedState  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--contexts", type=int, default=256)
    ap.add_argument("--spp", type=int, default=1 << 20)
    ap.add_argument("--pixels", type=int, default=1)
    ap.add_argument("--repeats", type=int, default=3)
    args = ap.parse_args()

    vza = np.linspace(-60.0, 60.0, args.pixels) if args.pixels > 1 else np.array([0.0])
    kdict = scenes.atmosphere_scene(
        geometry="spherical_shell", atmosphere="afgl",
        sensor={"type": "mdistant", "vza": vza, "vaa": 0.0}, spp=args.spp)
    mi_scene = mi_traverse(mi_load_dict(kdict), scenes.spectral_update_map(1200, spherical=True))
    ctxs = [KernelContext(w=w) for w in np.linspace(400.0, 1000.0, args.contexts)]
    paths = args.contexts * args.pixels * args.spp

    out = {}
    for name, flag in (("sequential", False), ("pipelined", True)):
        mi_render(mi_scene, ctxs[:8], spp=args.spp, seed_state=SeedState(0), pipelined=flag)  # warm-up
        best = float("inf")
        for _ in range(args.repeats):
            t0 = time.perf_counter()
            res = mi_render(mi_scene, ctxs, spp=args.spp, seed_state=SeedState(0), pipelined=flag)
            best = min(best, time.perf_counter() - t0)
        assert len(res) == args.contexts
        out[name] = {"s": best, "contexts_per_s": args.contexts / best, "Mpaths_per_s": paths / best / 1e6}
    # host-only cost of the loop (parameter update + table flattening, no render)
    t0 = time.perf_counter()
    for ctx in ctxs:
        mi_scene.parameters.update(mi_scene.umap_template.render(ctx))
        mi_scene.obj._device_scene.sync()
    out["host_update_only_s"] = time.perf_counter() - t0
    out["config"] = {"contexts": args.contexts, "spp": args.spp, "pixels": args.pixels,
                     "scene": "AFGL-shaped molecular + RPV, spherical shell, mdistant"}
    out["speedup"] = out["sequential"]["s"] / out["pipelined"]["s"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
