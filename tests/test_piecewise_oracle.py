"""
Pins the oracle's restatement of the piecewise medium (ERP/media/piecewise.cpp:183-429)
on the golden vectors of the reference's own test, ERP/tests/media/test_piecewise.py, and
checks the piecewise integrator (ERP/integrators/piecewise_volpath.cpp) against the
null-collision one on the same scene.  CPU only.
"""

import numpy as np
import pytest

from eradiate_b200 import scenes
from eradiate_b200.kernel import mi_load_dict
from tests.util import sidak_ok, stats_from_sums, z_scores

H, N, INTEGRAL, LBD = 100000.0, 10, 3.0, 8300.0
HALF_WIDTH = 50000.0  # test_piecewise.py:25-27: the medium cube spans [-5e4, 5e4]^2 x [0, 1e5]


@pytest.fixture(scope="module")
def medium_desc():
    """test_piecewise.py:6-45 create_medium_dict: 10 exponential layers, albedo 0.8."""
    z = np.linspace(0.0, H, N, endpoint=False)
    ext = (INTEGRAL / LBD) * np.exp(-z / LBD)
    d = scenes.atmosphere_scene(geometry="plane_parallel", atmosphere="afgl", n_layers=N, toa=H,
                                integrator="piecewise_volpath")
    sc = mi_load_dict(d)
    assert sc.flat.medium.type == "piecewise"
    sc.flat.medium.children["sigma_t"].values["data"][:] = ext.reshape(-1, 1, 1, 1).astype(np.float32)
    sc.flat.medium.children["albedo"].values["data"][:] = 0.8
    return sc, sc.flat.build_desc()


def _sample(oracle, desc, o, d, samples):
    n = len(samples)
    return oracle.piecewise_sample(desc, np.tile(o, (n, 1)), np.tile(d, (n, 1)), samples, half_width=HALF_WIDTH)


@pytest.mark.parametrize("origin,direction,samples,gt_dists,gt_trs,gt_pdfs", [
    # test01_sample_distances_down (:48-80)
    ([0, 0, H], [0, 0, -1], [0.0, 0.25, 0.5, 0.75, 0.99],
     [0.0, 74578.93910366, 82117.51365975, 88515.28423884, 98460.51859237],
     [1.0, 0.75, 0.5, 0.25, 0.01],
     [7.06034444e-09, 2.43563215e-05, 5.41709938e-05, 2.70854969e-05, 3.61445783e-06]),
    # test02_sample_distances_up (:83-116)
    ([0, 0, 0], [0, 0, 1], [0.0, 0.25, 0.5, 0.75, 0.99],
     [0.0, 7.95920400e02, 1.91770720e03, 3.83541440e03, 1.91443066e04],
     [1.0, 0.75, 0.5, 0.25, 0.01],
     [3.61445783e-04, 2.71084337e-04, 1.80722892e-04, 9.03614458e-05, 1.08341988e-06]),
    # test03_sample_distances_horizontal (:119-152)
    ([0, 0, 15000.0], [0, 1, 0], [0.0, 0.25, 0.5, 0.75, 0.99],
     [0.0, 2655.31469158, 6397.77055374, 12795.54110748, 42505.86749422],
     [1.0, 0.75, 0.5, 0.25, 0.01],
     [1.08341988e-04, 8.12564910e-05, 5.41709940e-05, 2.70854970e-05, 1.08341988e-06]),
    # test04_sample_distances_diag (:155-182): distances only
    ([0, 0, 15000.0], [0.5, 0.5, 0.5], [0.001, 0.25, 0.5, 0.75, 0.99],
     [9.23464998e00, 2.65531470e03, 6.39777058e03, 2.24562174e04, np.inf], None, None),
])
def test_sample_interaction_real_golden(oracle, medium_desc, origin, direction, samples, gt_dists, gt_trs, gt_pdfs):
    _, desc = medium_desc
    t, tr, pdf = _sample(oracle, desc, origin, direction, samples)
    gt = np.asarray(gt_dists)
    assert np.array_equal(np.isinf(t), np.isinf(gt))
    fin = np.isfinite(gt)
    assert np.allclose(t[fin], gt[fin], rtol=1e-5, atol=1e-8), (t, gt)
    if gt_trs is not None:
        assert np.allclose(tr, gt_trs, rtol=1e-5, atol=1e-8), (tr, gt_trs)
        assert np.allclose(pdf, gt_pdfs, rtol=1e-5, atol=1e-8), (pdf, gt_pdfs)


def test_sample_distances_heights_golden(oracle, medium_desc):
    """test05_sample_distances_heights (:185-224): u = 0.3 from seven start altitudes, looking up."""
    _, desc = medium_desc
    heights = [0.0, 9800.0, 10100.0, 27500.0, 40020.0, 75000, 98000.0]
    gt = np.array([986.80067823, 2824.88988516, 3292.12110592, np.inf, np.inf, np.inf, np.inf])
    o = np.array([[0.0, 0.0, h] for h in heights])
    t, _, _ = oracle.piecewise_sample(desc, o, np.tile([0.0, 0.0, 1.0], (7, 1)), 0.3, half_width=HALF_WIDTH)
    assert np.array_equal(np.isinf(t), np.isinf(gt))
    assert np.allclose(t[:3], gt[:3], rtol=1e-5)


@pytest.mark.parametrize("direction,heights,gt_trs,rtol", [
    # test06_eval_transmittance_up (:227-258), scalar_mono_double
    ([0, 0, 1], [0.0, 5000.0, 15000.0, 25000.0, 45000.0, 75000, 100000.0],
     [0.00573247, 0.03493101, 0.36588307, 0.7398143, 0.97331388, 0.99930119, 1.0], 1e-5),
    # test07_eval_transmittance_down (:261-292), scalar_mono (float32 reference arithmetic)
    ([0, 0, -1], [0.0, 9000.0, 15000.0, 29800.0, 45000.0, 70020, 100000.0],
     [1.0, 0.03865759, 0.01566748, 0.00663011, 0.00588964, 0.00573872, 0.00573247], 2e-5),
])
def test_eval_transmittance_golden(oracle, medium_desc, direction, heights, gt_trs, rtol):
    _, desc = medium_desc
    o = np.array([[0.0, 0.0, h] for h in heights])
    tr, _, _ = oracle.piecewise_eval(desc, o, np.tile(np.asarray(direction, float), (len(heights), 1)),
                                     half_width=HALF_WIDTH)
    assert np.allclose(tr, gt_trs, rtol=rtol, atol=1e-8), (tr, gt_trs)


def test_sampled_distance_distribution_matches_transmittance(oracle, medium_desc):
    """Property: P(t > x) of the sampler equals the exact transmittance over [0, x] on an oblique ray."""
    _, desc = medium_desc
    d = np.array([0.3, -0.2, 0.6]); d /= np.linalg.norm(d)
    u = (np.arange(20000) + 0.5) / 20000
    t, tr, pdf = oracle.piecewise_sample(desc, np.tile([10.0, 20.0, 3000.0], (u.size, 1)), np.tile(d, (u.size, 1)), u)
    fin = np.isfinite(t)
    # transmittance reported by the sampler at the sampled point is 1 - u
    assert np.allclose(tr[fin], 1.0 - u[fin], rtol=1e-9)
    # ... and agrees with the evaluator over a segment cut by a surface at the same distance
    k = np.flatnonzero(fin)[::997]
    tr2, _, _ = oracle.piecewise_eval(desc, np.tile([10.0, 20.0, 3000.0], (k.size, 1)), np.tile(d, (k.size, 1)), si_t=t[k])
    assert np.allclose(tr2, tr[k], rtol=1e-9)
    # escape probability = exp(-tau to the top)
    tr_top, _, _ = oracle.piecewise_eval(desc, [[10.0, 20.0, 3000.0]], [d])
    assert abs((~fin).mean() - tr_top[0]) < 1e-4


@pytest.mark.parametrize("kw", [
    dict(),
    dict(aerosol=True, surface={"type": "diffuse", "reflectance": 0.5}, sza=55.0, saa=40.0),
    dict(max_depth=3),
])
def test_piecewise_volpath_matches_volpath(oracle, kw):
    """Both integrators estimate the same radiance; piecewise needs no null collisions."""
    spp = 60000
    sens = {"type": "mdistant", "vza": np.linspace(-70.0, 70.0, 6), "vaa": 30.0}
    out = {}
    for integ in ("volpath", "piecewise_volpath"):
        d = scenes.atmosphere_scene(geometry="plane_parallel", n_layers=40, integrator=integ, sensor=sens, **kw)
        sc = mi_load_dict(d)
        assert sc.flat.medium.type == ("piecewise" if integ == "piecewise_volpath" else "heterogeneous")
        wl, l, l2, st = oracle.render(sc.flat.build_desc(), spp=spp, seed=3)
        out[integ] = (stats_from_sums(l, l2, spp), st)
    (m1, v1), s1 = out["volpath"]
    (m2, v2), s2 = out["piecewise_volpath"]
    z = z_scores(m1, v1, m2, v2)
    assert sidak_ok(z), z
    assert s2["trips_main"] < s1["trips_main"] and s2["trips_nee"] < s1["trips_nee"]


def test_piecewise_medium_under_volpath_is_heterogeneous(oracle):
    """piecewise.cpp inherits Medium::sample_interaction: with volpath it is delta tracking, bit for bit."""
    sens = {"type": "mdistant", "vza": [0.0, 40.0], "vaa": 0.0}
    res = []
    for fm in (True, False):
        d = scenes.atmosphere_scene(geometry="plane_parallel", n_layers=20, integrator="volpath", sensor=sens,
                                    force_majorant=fm)
        sc = mi_load_dict(d)
        res.append(oracle.render(sc.flat.build_desc(), spp=2000, seed=1)[0])
    assert np.array_equal(res[0], res[1])


def test_piecewise_errors():
    sens = {"type": "mdistant", "vza": [0.0], "vaa": 0.0}
    d = scenes.atmosphere_scene(geometry="plane_parallel", n_layers=8, integrator="piecewise_volpath", sensor=sens,
                                force_majorant=True)
    with pytest.raises(RuntimeError, match="sample_interaction_real"):
        mi_load_dict(d).flat
    d = scenes.atmosphere_scene(geometry="plane_parallel", atmosphere="homogeneous", integrator="piecewise_volpath",
                                sensor=sens)
    with pytest.raises(RuntimeError, match="sample_interaction_real"):
        mi_load_dict(d).flat
    d = scenes.atmosphere_scene(geometry="spherical_shell", n_layers=8, integrator="piecewise_volpath", sensor=sens)
    with pytest.raises(RuntimeError, match="sample_interaction_real"):
        mi_load_dict(d).flat
