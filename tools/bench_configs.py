#!/usr/bin/env python
"""
Throughput of the BASELINE.json configurations other than the headline (C2 is bench.py's job):
C1, C3 (full size: hdistant 32x32, spp 2^22 -> 4.3e9 paths, reduced with --quick), C4 (canopy) and one
band of C5 (polarized ocean + aerosol). Prints one JSON line per configuration. Needs a GPU.
"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eradiate_b200 import scenes
from eradiate_b200.kernel import mi_load_dict, render


def run(name, kd, spp, sensor=0, repeats=2):
    sc = mi_load_dict(kd)
    render(sc, sensor=sensor, seed=1, spp=max(16, spp >> 6))  # warm-up (tables, BVH upload)
    best = None
    for r in range(repeats):
        t0 = time.perf_counter()
        bmp = render(sc, sensor=sensor, seed=2 + r, spp=spp)
        wall = time.perf_counter() - t0
        st = bmp.stats
        if best is None or st["device_ms"] < best[0]["device_ms"]:
            best = (st, wall)
    st, wall = best
    img = np.array(bmp)[..., 0]
    print(json.dumps({
        "config": name, "paths": st["n_paths"], "device_ms": round(st["device_ms"], 3),
        "Mpaths_per_s": round(st["n_paths"] / st["device_ms"] / 1e3, 1), "wall_s": round(wall, 4),
        "loop_trips_per_path": round((st["trips_main"] + st["trips_nee"]) / st["n_paths"], 2),
        "n_bands": st["n_bands"], "mean_radiance": float(img.mean()), "finite": bool(np.isfinite(img).all()),
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    q = 6 if args.quick else 0
    run("C1 homogeneous+Lambertian pp, mdistant 1, spp 4096 (x4096 for a measurable launch)", scenes.config_c1(), 4096 * 4096 >> q)
    run("C3 AFGL+aerosol tabphase, spherical shell, hdistant 32x32, spp 2^22", scenes.config_c3(), (1 << 22) >> q)
    os.environ["ERTB_MAJORANT"] = "global"
    run("C3, reference's single global majorant (ERTB_MAJORANT=global), spp 2^18", scenes.config_c3(), (1 << 18) >> q)
    del os.environ["ERTB_MAJORANT"]
    run("C5 band @550 nm: polarized ocean + AFGL + polarized aerosol, spherical shell, mdistant 1, spp 2^20 (x64)",
        scenes.config_c5(), (1 << 26) >> q)
    c4 = scenes.config_c4(spp=16)
    run("C4 canopy 25x25x2 m LAI 3 x 5x5 + AFGL + RPV, mdistant 32, spp 2^18", c4, (1 << 18) >> q, sensor=0)
    run("C4 canopy, perspective 64x64 inside the atmosphere, spp 2^11", c4, (1 << 11) >> q, sensor=1)


if __name__ == "__main__":
    main()
