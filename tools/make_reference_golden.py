#!/usr/bin/env python
"""
Renders scenes of the battery with THE REFERENCE ITSELF -- the Mitsuba fork + Eradiate plugins built
from /root/reference by oracle/build_ref.sh into oracle/_ref -- from the very dictionaries
``eradiate_b200.scenes`` / ``tests/scene_battery.py`` hand to ``mi_load_dict``, and writes

    tests/golden/reference_renders.json

(per-pixel mean, variance of the mean from the `moment` integrator's m2 channel, Stokes components for
polarized scenes, and the parameter keys ``mitsuba.traverse`` publishes for each scene).  These
fixtures pin (i) the C oracle port (tests/test_oracle_vs_reference.py), (ii) the CUDA path directly
(tests/test_gpu_parity.py::test_render_matches_reference_fixture) and (iii) the key set of
``mi_traverse``.  Variant: scalar_mono_double (polarized scenes: scalar_mono_polarized_double).

The reference cannot travel to the GPU box as a pin (only as the CPU arm of bench.py), hence committed
fixtures.  Run here:   python tools/make_reference_golden.py [--only NAME ...] [--spp-log2 N]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from oracle import ref  # noqa: E402  (adds oracle/_ref/python to sys.path, imports mitsuba)
from tests.scene_battery import battery  # noqa: E402

SEED = 20261018
# Every scene of the battery is rendered: total paths = pixels x spp ~ 2^BUDGET, reduced for scenes whose
# walk takes hundreds of loop trips per path (K from the oracle fixtures) and for one-pixel films (the scalar
# reference renders a pixel on ONE thread, integrator.cpp:203-213).  The scalar double-precision reference
# does ~0.6 Mpaths/s on 8 cores at K = 13.
BUDGET = 24
OVERRIDE = {  # log2(total paths) for the BASELINE configurations
    "c1_homogeneous_lambertian_pp": 23,
    "c2_afgl_rpv_spherical": 26,
    "c3_afgl_aerosol_tab_hdistant": 24,
    "c2_full_size_32vza": 26,
    "c3_full_film_32x32": 24,
    "c5_polarized_ocean_aerosol_reduced": 24,
}


def spp_for(name: str, npix: int, k_oracle: float) -> int:
    log2 = OVERRIDE.get(name, BUDGET)
    if name not in OVERRIDE:
        if k_oracle > 100:
            log2 -= 3
        if npix == 1:
            log2 -= 2
    return max(1 << 10, (1 << log2) // max(npix, 1))


def render_scene(name: str, d: dict, spp: int) -> dict:
    polarized = ref.is_polarized(d)
    mi = ref.mitsuba("scalar_mono_polarized_double" if polarized else "scalar_mono_double")
    # optimize=False as eradiate.kernel.mi_load_dict does (_render.py:186): no merging of identical objects,
    # so the published parameter keys do not depend on two textures happening to hold the same value
    scene = mi.load_dict(ref.to_mitsuba(mi, d), optimize=False)
    keys = sorted(mi.traverse(scene).keys())
    t0 = time.perf_counter()
    mi.render(scene, sensor=0, seed=SEED, spp=spp)
    dt = time.perf_counter() - t0
    chans = ref.film_channels(mi, scene.sensors()[0])
    out = {"spp": spp, "seconds": round(dt, 2), "variant": mi.variant(), "traverse_keys": keys}
    m1 = chans["nested.Y"] if "nested.Y" in chans else chans["Y"]
    out["mean"] = m1.ravel().tolist()
    out["mean_wl"] = chans["Y"].ravel().tolist()
    if "m2_nested.Y" in chans:
        m2 = chans["m2_nested.Y"]
        out["var_of_mean"] = (np.maximum(m2 - m1 * m1, 0.0) / spp).ravel().tolist()
    if "S0.R" in chans:
        out["stokes"] = [chans[f"S{k}.R"].ravel().tolist() for k in range(4)]
    npix = m1.size
    print(f"{name:44s} {npix:5d} px  spp 2^{int(np.log2(spp)):2d}  {dt:7.1f} s  "
          f"{npix * spp / dt / 1e6:6.3f} Mpaths/s  mean[0] = {m1.ravel()[0]:.6g}", flush=True)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", nargs="*", default=None)
    ap.add_argument("--spp-log2", type=int, default=None, help="override the per-scene sample count")
    ap.add_argument("--list", action="store_true")
    ap.add_argument("--keys-only", action="store_true", help="refresh only the traverse key lists (no rendering)")
    args = ap.parse_args()
    path = os.path.join(ROOT, "tests", "golden", "reference_renders.json")
    out = {"seed": SEED, "generator": "tools/make_reference_golden.py", "reference": ref.describe(), "scenes": {}}
    if os.path.exists(path):
        old = json.load(open(path))
        out["scenes"], out["refused"] = old.get("scenes", {}), old.get("refused", {})
    bat = battery()
    if args.list:
        print("\n".join(bat.keys()))
        return
    if args.keys_only:
        for name, rec in out["scenes"].items():
            polarized = ref.is_polarized(bat[name])
            mi = ref.mitsuba("scalar_mono_polarized_double" if polarized else "scalar_mono_double")
            rec["traverse_keys"] = sorted(mi.traverse(mi.load_dict(ref.to_mitsuba(mi, bat[name]), optimize=False)).keys())
        with open(path, "w") as f:
            json.dump(out, f, indent=1)
        print("refreshed the key lists of", len(out["scenes"]), "scenes")
        return
    oracle_gold = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_renders.json")))["scenes"]
    out.setdefault("refused", {})
    for name in bat:
        if args.only is not None and name not in args.only:
            continue
        if args.only is None and (name in out["scenes"] or name in out["refused"]):
            continue
        g = oracle_gold.get(name, {})
        k = g.get("trips_main_per_path", 10.0) + g.get("trips_nee_per_path", 4.0)
        spp = (1 << args.spp_log2) if args.spp_log2 else spp_for(name, len(g.get("mean", [0])), k)
        try:
            out["scenes"][name] = render_scene(name, bat[name], spp)
        except Exception as e:  # a scene the reference refuses is a finding, not a crash
            print(f"!! {name}: reference failed: {type(e).__name__}: {e}", file=sys.stderr)
            out["scenes"].pop(name, None)
            out["refused"][name] = f"{type(e).__name__}: {e}"[:500]
        with open(path, "w") as f:
            json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
