"""
The boundary against THE REFERENCE'S OWN boundary code: tests/golden/reference_boundary.json records what
`eradiate.kernel.mi_load_dict / mi_traverse / mi_render` (src/eradiate/kernel/_render.py, executed from the reference
tree on top of the reference Mitsuba built into oracle/_ref, tools/make_boundary_fixture.py + oracle/ref_eradiate.py)
did with the kernel dictionaries and update maps of eradiate_b200/scenes.py -- `SearchSceneParameter(node_type=
mi.Medium, ...)` look-ups, the spectral loop, the seed sequence, the structure of the result.  The functions of
eradiate_b200.kernel must do the same with the same inputs.
"""

import json
import os

import numpy as np
import pytest

from eradiate_b200.kernel import KernelContext, SeedState, mi_load_dict, mi_render, mi_traverse
from tests.test_oracle_vs_reference import _is_passive
from tests.util import sidak_ok, z_scores
from tools.make_boundary_fixture import cases

REF = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_boundary.json")))


@pytest.mark.parametrize("name", list(REF["cases"].keys()))
def test_traverse_resolves_the_update_map_as_the_reference_does(name):
    kdict, umap, _ = cases()[name]
    ref = REF["cases"][name]
    w = mi_traverse(mi_load_dict(kdict), umap)
    # _render.py:333-371: every SearchSceneParameter look-up ends in the same parameter id
    assert {k: p.parameter_id for k, p in w.umap_template.items()} == ref["resolved_parameter_ids"]
    # and the parameter table holds the reference's keys (minus the passive ones this kernel has no use for)
    used = set(ref["resolved_parameter_ids"].values())  # (a blendphase `weight.data` is no mask bitmap)
    want = {k for k in ref["parameters"] if k in used or not _is_passive(k)}
    assert set(w.parameters.keys()) == want
    # rendering the template for a context gives updates for exactly the resolved ids (_kernel_dict.py:276-314)
    upd = w.umap_template.render(KernelContext(w=550.0))
    assert set(upd.keys()) == set(ref["resolved_parameter_ids"].values())


def test_seed_sequence_of_the_loop_matches_the_reference():
    ss = SeedState(REF["seed"])  # rng.py: one next() per (context, sensor)
    assert [int(ss.next().squeeze()) for _ in range(6)] == REF["seed_sequence_first6"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(REF["cases"].keys()))
def test_mi_render_loop_matches_the_reference_loop(name):
    """Same dict, same update map, same contexts, same spp, same SeedState: same result structure
    ({spectral index: {sensor id: Bitmap}}, channel names, pixel formats, shapes) and statistically the same films
    (each side's variance from its own m2_nested channel, `moment` integrator)."""
    kdict, umap, spp = cases()[name]
    ref = REF["cases"][name]
    assert spp == ref["spp"]
    w = mi_traverse(mi_load_dict(kdict), umap)
    results = mi_render(w, [KernelContext(w=wl) for wl in REF["wavelengths"]], spp=spp, seed_state=SeedState(REF["seed"]))
    assert [float(k) for k in results.keys()] == ref["result_keys"]
    for siah, per_sensor in results.items():
        rf = ref["films"][repr(float(siah))]
        assert list(per_sensor.keys()) == ref["sensor_ids"] == list(rf.keys())
        for sid, bmp in per_sensor.items():
            assert str(bmp.pixel_format()).split(".")[-1] == rf[sid]["pixel_format"]
            split = dict(bmp.split())
            assert list(split.keys()) == list(rf[sid]["channels"].keys())
            vals = {}
            for cname, sub in split.items():
                rc = rf[sid]["channels"][cname]
                a = np.array(sub, dtype=np.float64)
                if a.ndim == 2:
                    a = a[..., None]
                assert str(sub.pixel_format()).split(".")[-1] == rc["pixel_format"] and list(a.shape) == rc["shape"]
                vals[cname] = a[..., 0].ravel()
            assert np.array_equal(vals["<root>"], vals["nested"])  # box filter: the two are the same estimator
            stokes = [k for k in vals if k.startswith("S")]
            assert stokes == ([] if ref.get("variant") != "scalar_mono_polarized_double" else ["S0", "S1", "S2", "S3"])
            m_g, m_r = vals["nested"], np.array(rf[sid]["channels"]["nested"]["first"])
            v_g = np.maximum(vals["m2_nested"] - m_g**2, 0.0) / spp
            v_r = np.maximum(np.array(rf[sid]["channels"]["m2_nested"]["first"]) - m_r**2, 0.0) / spp
            ok, zc = sidak_ok(z_scores(m_g, v_g, m_r, v_r), alpha=0.001 / 6)  # six (context, sensor) films per case
            assert ok, (name, siah, sid, m_g, m_r)
            for k in stokes:  # Stokes components: |S_k| <= I, compared on the scale of the noise of I
                s_r = np.array(rf[sid]["channels"][k]["first"])
                assert np.all(np.abs(vals[k] - s_r) <= 5.0 * np.sqrt(v_g + v_r) + 1e-12), (name, siah, k, vals[k], s_r)
