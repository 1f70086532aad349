"""
Generates tests/golden/oracle_renders.json: per-pixel mean / variance-of-the-mean /
loop-trip counters of the CPU oracle for the scene battery at a fixed seed.

These fixtures are outputs of the oracle restatement -- pinned against the reference's golden
vectors by tests/test_oracle_golden.py / test_oracle_system.py and against renders of the reference
itself (oracle/_ref, tests/golden/reference_renders.json) by tests/test_oracle_vs_reference.py.
Run:  python tools/make_golden.py [--only-missing]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from eradiate_b200.kernel import mi_load_dict  # noqa: E402
from oracle import oracle  # noqa: E402
from tests.scene_battery import battery  # noqa: E402

SPP = 1 << 17
SEED = 20261017


def main():
    path = os.path.join(ROOT, "tests", "golden", "oracle_renders.json")
    out = {"spp": SPP, "seed": SEED, "scenes": {}}
    if "--only-missing" in sys.argv and os.path.exists(path):  # keep the fixtures already committed
        out = json.load(open(path))
    for name, d in battery().items():
        if name in out["scenes"]:
            continue
        sc = mi_load_dict(d)
        desc = sc.flat.build_desc()
        npix = desc.sensors[0].width * desc.sensors[0].height
        spp = SPP if npix <= 8 else SPP // 2
        if npix > 256:  # full-size films (C3 32x32 at ~475 loop trips per path): 2^23 paths in total
            spp = max(1 << 10, (1 << 23) // npix)
        stokes = None
        if desc.polarized:
            wl, l, l2, stokes, st = oracle.render_stokes(desc, 0, SEED, spp)
        else:
            wl, l, l2, st = oracle.render(desc, 0, SEED, spp)
        mean = l / spp
        var = np.maximum(l2 / spp - mean**2, 0) / spp
        out["scenes"][name] = {
            "spp": spp,
            "mean_wl": (wl / spp).tolist(),
            "mean": mean.tolist(),
            "var_of_mean": var.tolist(),
            "trips_main_per_path": st["trips_main"] / st["n_paths"],
            "trips_nee_per_path": st["trips_nee"] / st["n_paths"],
            "scatter_per_path": st["n_scatter"] / st["n_paths"],
            "surface_per_path": st["n_surface"] / st["n_paths"],
            # free flights actually sampled (no stencil-crossing iterations, no zero-weight shadow rays):
            # what the CUDA kernels count as loop trips
            "flights_main_per_path": st["flights_main"] / st["n_paths"],
            "flights_nee_per_path": st["flights_nee"] / st["n_paths"],
        }
        if stokes is not None:
            out["scenes"][name]["stokes"] = (stokes / spp).tolist()
            out["scenes"][name]["m2"] = (l2 / spp).tolist()
        print(f"{name:40s} mean[0]={mean[0]:.6f} K={(st['trips_main']+st['trips_nee'])/st['n_paths']:.2f}")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
