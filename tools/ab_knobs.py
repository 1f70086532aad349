#!/usr/bin/env python
"""
A/B runs of one configuration under different developer knobs (environment variables read by the CUDA
library, optionally another build through ERTB_LIB): one subprocess per setting, prints one JSON line each.

    python tools/ab_knobs.py --config c2 --spp-log2 20 -- ERTB_MAJORANT=global -- ERTB_BAND_PENALTY=0.1 -- ""

Needs a GPU.  Each line: kernel ms (best of --repeats), Mpaths/s, loop trips per path, bands, film mean.
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(config, spp, repeats, sensor):
    sys.path.insert(0, ROOT)
    import numpy as np
    from eradiate_b200 import scenes
    from eradiate_b200.kernel import mi_load_dict, render

    cfgs = {"c1": scenes.config_c1, "c2": scenes.config_c2, "c3": scenes.config_c3,
            "c4": lambda **k: scenes.config_c4(spp=16), "c5": scenes.config_c5}
    if config in cfgs:
        kd = cfgs[config]()
    else:  # a scene of the test battery
        from tests.scene_battery import battery
        kd = battery()[config]
    sc = mi_load_dict(kd)
    render(sc, sensor=sensor, seed=1, spp=max(16, spp >> 6))
    best = None
    for r in range(repeats):
        bmp = render(sc, sensor=sensor, seed=2 + r, spp=spp)
        st = bmp.stats
        if best is None or st["device_ms"] < best["device_ms"]:
            best = st
    img = np.array(bmp)[..., 0]
    print(json.dumps({
        "knobs": os.environ.get("ERTB_AB_LABEL", ""), "config": config, "paths": best["n_paths"],
        "device_ms": round(best["device_ms"], 3), "Mpaths_per_s": round(best["n_paths"] / best["device_ms"] / 1e3, 1),
        "trips_per_path": round((best["trips_main"] + best["trips_nee"]) / best["n_paths"], 3),
        "n_bands": best["n_bands"], "mean": float(img.mean()),
        # last render, per pixel: mean and variance of the mean (for z-scores between settings)
        "px_mean": (bmp.raw["sum_l"].ravel() / spp).tolist(),
        "px_var": (np.maximum(bmp.raw["sum_l2"].ravel() / spp - (bmp.raw["sum_l"].ravel() / spp) ** 2, 0) / spp).tolist(),
    }), flush=True)


def main():
    if os.environ.get("ERTB_AB_CHILD"):
        a = json.loads(os.environ["ERTB_AB_CHILD"])
        return child(a["config"], a["spp"], a["repeats"], a["sensor"])
    argv = sys.argv[1:]
    groups, cur = [], []
    if "--" in argv:
        i = argv.index("--")
        argv, rest = argv[:i], argv[i + 1:]
        for tok in rest:
            if tok == "--":
                groups.append(cur)
                cur = []
            else:
                cur.append(tok)
        groups.append(cur)
    else:
        groups = [[]]
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--spp-log2", type=int, default=20)
    ap.add_argument("--repeats", type=int, default=3)
    ap.add_argument("--sensor", type=int, default=0)
    args = ap.parse_args(argv)
    results = []
    for g in groups:
        env = dict(os.environ)
        label = []
        for kv in g:
            if kv and "=" in kv:
                k, v = kv.split("=", 1)
                env[k] = v
                label.append(kv)
        env["ERTB_AB_LABEL"] = " ".join(label) or "(default)"
        env["ERTB_AB_CHILD"] = json.dumps({"config": args.config, "spp": 1 << args.spp_log2, "repeats": args.repeats,
                                           "sensor": args.sensor})
        r = subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, check=False, capture_output=True, text=True)
        sys.stderr.write(r.stderr[-2000:])
        for line in r.stdout.splitlines():
            try:
                d = json.loads(line)
            except ValueError:
                print(line)
                continue
            results.append(d)
            short = {k: v for k, v in d.items() if not k.startswith("px_")}
            if len(results) > 1 and len(d["px_mean"]) <= 4096:
                import numpy as np
                a, b = results[0], d
                z = (np.array(b["px_mean"]) - np.array(a["px_mean"])) / np.sqrt(np.array(b["px_var"]) + np.array(a["px_var"]) + 1e-300)
                short["z_vs_first_max"] = round(float(np.abs(z).max()), 2)
                short["z_vs_first_mean"] = round(float(z.mean()), 2)
            print(json.dumps(short), flush=True)


if __name__ == "__main__":
    main()
