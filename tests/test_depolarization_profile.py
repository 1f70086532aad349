"""
Layer-dependent Rayleigh depolarization (MI/src/phase/rayleigh.cpp:48,79 and
ERP/phase/rayleigh_polarized.cpp: `depolarization` is a *volume* evaluated at the interaction; Eradiate
emits it as a gridvolume on the atmosphere's grid, scenes/phase/_rayleigh.py:98-131).

The host flattener carries a per-layer profile as the blend of two constant-depolarization leaves
(``_Flat.phase_leaves``).  These tests state why that is the same phase function (value and sampling density),
that the flattener produces it, and that the renders are sensitive to the profile -- the pin itself is the
reference render of the three ``*depolarization_profile*`` scenes (test_oracle_vs_reference.py,
test_gpu_parity.py::test_render_matches_reference_fixture).
"""

import json
import os

import numpy as np
import pytest

from eradiate_b200 import scenes
from eradiate_b200.kernel import mi_load_dict, mi_traverse
from tests.scene_battery import battery
from tests.util import stats_from_sums


def rayleigh_value(ct, rho):
    """rayleigh.cpp:61-71."""
    r1, r2 = (1.0 - rho) / (1.0 + rho / 2.0), (1.0 + rho) / (1.0 - rho)
    return 3.0 / (16.0 * np.pi) * r1 * (r2 + ct * ct)


def rayleigh_mueller(ct, rho):
    """ERP/phase/rayleigh_polarized.cpp:78-110, scattering-plane frame: the six non-zero entries."""
    r1, r2, r3 = (1.0 - rho) / (1.0 + rho / 2.0), (1.0 + rho) / (1.0 - rho), (1.0 - 2.0 * rho) / (1.0 - rho)
    k = 3.0 / (16.0 * np.pi) * r1
    return np.array([k * (r2 + ct * ct), k * (ct * ct - 1.0), k * (ct * ct + 1.0), k * 2.0 * ct, k * 2.0 * ct * r3])


def test_two_leaf_blend_is_the_layer_phase_function():
    rng = np.random.default_rng(5)
    for _ in range(200):
        lo, hi = np.sort(rng.uniform(0.0, 0.95, 2))
        rho = rng.uniform(lo, hi)
        b = lambda r: r / (2.0 + r)  # noqa: E731
        w = (b(rho) - b(lo)) / (b(hi) - b(lo))
        assert 0.0 <= w <= 1.0
        ct = rng.uniform(-1.0, 1.0, 7)
        assert np.allclose((1 - w) * rayleigh_value(ct, lo) + w * rayleigh_value(ct, hi), rayleigh_value(ct, rho),
                           rtol=1e-12)
        assert np.allclose((1 - w) * rayleigh_mueller(ct, lo) + w * rayleigh_mueller(ct, hi),
                           rayleigh_mueller(ct, rho), rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("geometry", ["plane_parallel", "spherical_shell"])
def test_flattener_splits_a_gridded_profile(geometry):
    n = 12
    rho = np.linspace(0.02, 0.3, n)
    sc = mi_load_dict(scenes.atmosphere_scene(geometry=geometry, n_layers=n, aerosol=True, aerosol_phase="hg",
                                              phase={"type": "rayleigh", "depolarization": rho}))
    leaves = sc.flat.phase_leaves(n)
    assert [ph.type for ph, _ in leaves] == ["rayleigh", "rayleigh", "hg"]
    assert np.isclose(leaves[0][0].rho, 0.02) and np.isclose(leaves[1][0].rho, 0.3)
    w = np.stack([p for _, p in leaves]).astype(np.float64)
    assert np.allclose(w.sum(axis=0), 1.0, atol=1e-6)
    mol = w[0] + w[1]
    ct = np.linspace(-1, 1, 9)[:, None]
    blend = (w[0] * rayleigh_value(ct, 0.02) + w[1] * rayleigh_value(ct, 0.3)) / mol
    assert np.allclose(blend, rayleigh_value(ct, rho[None, :].astype(np.float32)), rtol=2e-6)
    # the volume is published as Mitsuba publishes it, and an update is re-flattened (a uniform profile included)
    params = mi_traverse(sc).parameters
    key = [k for k in params.keys() if "depolarization" in k and k.endswith("data")]
    assert len(key) == 1
    params.update({key[0]: np.full_like(params[key[0]], 0.1)})
    leaves = sc.flat.phase_leaves(n)
    assert len(leaves) == 3 and np.all(leaves[1][1] == 0.0) and np.isclose(leaves[0][0].rho, 0.1)
    with pytest.raises(RuntimeError, match="Depolarization factor"):
        params.update({key[0]: np.full_like(params[key[0]], 1.0)})
        sc.flat.build_desc()


def test_scalar_depolarization_keeps_one_leaf():
    sc = mi_load_dict(scenes.atmosphere_scene(n_layers=8, phase={"type": "rayleigh", "depolarization": 0.03}))
    leaves = sc.flat.phase_leaves(8)
    assert len(leaves) == 1 and leaves[0][0].type == "rayleigh"


@pytest.mark.parametrize("name", ["rayleigh_depolarization_profile_pp",
                                  "polarized_rayleigh_depolarization_profile_spherical"])
def test_renders_are_sensitive_to_the_profile(oracle, name):
    """The committed films (which agree with the reference's) differ significantly from a render with the
    column-mean depolarization: the fixtures do resolve the per-layer lookup."""
    d = battery()[name]
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "oracle_renders.json")))["scenes"][name]
    sc = mi_load_dict(d)
    params = mi_traverse(sc).parameters
    key = [k for k in params.keys() if "depolarization" in k and k.endswith("data")][0]
    params.update({key: np.full_like(params[key], float(np.mean(params[key])))})
    spp = 1 << 16
    out = oracle.render(sc.flat.build_desc(), 0, 11, spp)
    mean, var = stats_from_sums(out[1], out[2], spp)
    gm, gv = np.array(gold["mean"]), np.array(gold["var_of_mean"])
    z = np.abs(mean - gm) / np.sqrt(var + gv)
    assert (z ** 2).sum() > 50.0, z  # chi^2 over the 5 pixels (P < 1e-8 under "no difference")


def test_profile_below_a_multiphase_node_with_mis():
    """multiphase.cpp:176-200 (use_mis): the mixture weight sum_j w_j value_j / sum_j w_j pdf_j over the split leaves
    is the reference's own weight for the layer's depolarization, because the two halves share the sampling density."""
    n = 10
    rho = np.linspace(0.05, 0.4, n)
    vol = lambda v: scenes._volume(np.asarray(v, float), False, scenes.EARTH_RADIUS, scenes.TOA, 1.0e9)  # noqa: E731
    phase = {"type": "multiphase", "use_mis": True,
             "phase0": {"type": "rayleigh", "depolarization": vol(rho)}, "weight0": vol(np.full(n, 2.0)),
             "phase1": {"type": "hg", "g": 0.7}, "weight1": vol(np.linspace(0.5, 1.5, n))}
    sc = mi_load_dict(scenes.atmosphere_scene(geometry="plane_parallel", n_layers=n, phase=phase))
    leaves = sc.flat.phase_leaves(n)
    assert [ph.type for ph, _ in leaves] == ["rayleigh", "rayleigh", "hg"] and sc.flat._phase_mis
    w = np.stack([p for _, p in leaves]).astype(np.float64)
    assert np.allclose(w.sum(axis=0), 1.0, atol=1e-6)
    ct = np.linspace(-1.0, 1.0, 11)[:, None]
    pdf_ray = 3.0 / (16.0 * np.pi) * (1.0 + ct * ct)
    g = 0.7
    hg = (1.0 - g * g) / (4.0 * np.pi * (1.0 + g * g + 2.0 * g * ct) ** 1.5)  # value == pdf; either cosine convention
    num = w[0] * rayleigh_value(ct, leaves[0][0].rho) + w[1] * rayleigh_value(ct, leaves[1][0].rho) + w[2] * hg
    den = (w[0] + w[1]) * pdf_ray + w[2] * hg
    w_mol = 2.0 / (2.0 + np.linspace(0.5, 1.5, n))
    ref = (w_mol * rayleigh_value(ct, rho.astype(np.float32)) + (1.0 - w_mol) * hg) / (w_mol * pdf_ray + (1.0 - w_mol) * hg)
    assert np.allclose(num / den, ref, rtol=5e-6)
    assert sc.flat.build_desc().phase_mis == 1
