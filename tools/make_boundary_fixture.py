"""
Runs THE REFERENCE'S OWN boundary functions -- `eradiate.kernel.mi_load_dict`, `mi_traverse` (with an update-map
template whose parameters are located by `SearchSceneParameter(node_type=mi.Medium, ...)` look-ups, as Eradiate's
scene elements do) and the `mi_render` spectral loop (src/eradiate/kernel/_render.py:186-468) -- on top of the
reference Mitsuba compiled into oracle/_ref, on the kernel dictionaries and update maps of eradiate_b200/scenes.py,
and records what they do: the parameter ids every look-up resolved to, the parameter table left after
`drop_parameters`, the structure of the result (`{spectral index: {sensor id: Bitmap}}`, channel names and pixel
formats of `Bitmap.split()`), and the films of a three-context loop -> tests/golden/reference_boundary.json.

Run here (needs /root/reference and oracle/_ref):  python tools/make_boundary_fixture.py
"""

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from eradiate_b200 import scenes  # noqa: E402
from eradiate_b200.kernel import KernelContext, SeedState  # noqa: E402
from oracle import ref, ref_eradiate  # noqa: E402

WAVELENGTHS = [440.0, 550.0, 670.0]
SEED = 20261019


def cases():
    """name -> (kernel dict, update map, spp)."""
    n = 120
    return {
        # the C2 scene as BASELINE.json describes it (1200 layers), four view angles
        "c2_reduced_spectral_loop": (scenes.config_c2(spp=4, n_vza=4), scenes.spectral_update_map(1200, spherical=True), 1 << 19),
        "plane_parallel_two_sensors": (
            scenes.atmosphere_scene(geometry="plane_parallel", n_layers=n, sza=35.0, saa=20.0,
                                    sensor={"type": "mdistant", "vza": [-40.0, 0.0, 40.0], "vaa": 20.0},
                                    extra_sensors=[{"type": "hdistant", "film_resolution": (2, 2)}]),
            scenes.spectral_update_map(n, spherical=False), 1 << 18),
        # polarized variant: the `stokes` wrapper around `moment`-less volpath (AOV layers S0 .. S3), C5-like
        "c5_reduced_polarized_loop": (scenes.config_c5(spp=4, n_vza=3, n_layers=n, geometry="plane_parallel"),
                                      scenes.spectral_update_map_c5(n, spherical=False), 1 << 17),
    }


def film_stats(mi, bmp, spp):
    out = {"pixel_format": str(bmp.pixel_format()).split(".")[-1], "channels": {}}
    for name, sub in bmp.split():
        a = np.array(sub, dtype=np.float64)
        if a.ndim == 2:
            a = a[..., None]
        out["channels"][name] = {"pixel_format": str(sub.pixel_format()).split(".")[-1], "shape": list(a.shape),
                                 "first": a[..., 0].ravel().tolist()}
    return out


def main():
    out = {"generator": "tools/make_boundary_fixture.py", "reference": ref.describe(), "wavelengths": WAVELENGTHS,
           "seed": SEED, "cases": {}}
    for name, (kdict, umap, spp) in cases().items():
        variant = "scalar_mono_polarized_double" if ref.is_polarized(kdict) else "scalar_mono_double"
        rd, kd, mi = ref_eradiate.kernel(variant)
        mi_obj = rd.mi_load_dict(ref.to_mitsuba(mi, kdict))
        wrapper = rd.mi_traverse(mi_obj, ref_eradiate.translate_umap(umap, variant))
        resolved = {k: p.parameter_id for k, p in wrapper.umap_template.items()}
        kept = sorted(wrapper.parameters.keys())
        ctxs = [KernelContext(w=w) for w in WAVELENGTHS]
        results = rd.mi_render(wrapper, ctxs, spp=spp, seed_state=rd.SeedState(SEED))
        films = {}
        for siah, per_sensor in results.items():
            films[repr(float(siah))] = {sid: film_stats(mi, bmp, spp) for sid, bmp in per_sensor.items()}
        out["cases"][name] = {"spp": spp, "variant": variant, "resolved_parameter_ids": resolved, "parameters": kept,
                              "result_keys": [float(k) for k in results.keys()],
                              "sensor_ids": [list(v.keys()) for v in results.values()][0], "films": films}
        print(name, resolved, len(kept), list(films.keys()))
    # the seed sequence of the loop (rng.py): one SeedState.next() per (context, sensor)
    rd, _, _ = ref_eradiate.kernel("scalar_mono_double")
    ss = rd.SeedState(SEED)
    out["seed_sequence_first6"] = [int(ss.next().squeeze()) for _ in range(6)]
    ours = SeedState(SEED)
    assert out["seed_sequence_first6"] == [int(ours.next().squeeze()) for _ in range(6)]
    path = os.path.join(ROOT, "tests", "golden", "reference_boundary.json")
    with open(path, "w") as fh:
        json.dump(out, fh)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
