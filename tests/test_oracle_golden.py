"""
Pins the CPU oracle (oracle/ertb_oracle.c) against every golden vector / known-answer
test the reference's own test-suite holds for the hot path.
CPU only.  Citations: "ERP" = /root/reference/ext/mitsuba/src/eradiate_plugins,
"MI" = /root/reference/ext/mitsuba.

The numpy formulas below are written from the published model equations
(RPV: Rahman, Pinty & Verstraete 1993; RTLS: MODIS BRDF/Albedo ATBD v5, Ross-thick /
Li-sparse-reciprocal kernels), i.e. the same sources the reference's numpy
re-implementations in ERP/tests/bsdfs/test_rpv.py:22-48 and test_rtls.py:9-83 follow.
"""

import numpy as np
import pytest

from eradiate_b200 import scenes
from eradiate_b200.kernel import mi_load_dict


def sph_to_dir(theta, phi):
    theta, phi = np.asarray(theta, float), np.asarray(phi, float)
    return np.stack(
        [np.sin(theta) * np.cos(phi), np.sin(theta) * np.sin(phi), np.cos(theta)], axis=-1
    )


def make_desc(surface=None, phase=None, **kw):
    kw.setdefault("geometry", "plane_parallel")
    kw.setdefault("n_layers", 10)
    d = scenes.atmosphere_scene(surface=surface, phase=phase, **kw)
    sc = mi_load_dict(d)
    return sc, sc.flat.build_desc()


# ------------------------------------------------------------------------------ Hapke
POMMEROL = dict(w=0.526, theta=13.3, b=0.187, c=(1.0 + 0.273) / 2.0, h=0.083, B_0=1.0)


@pytest.mark.parametrize(
    "theta_o_deg,golden",
    [
        (30.0, 0.24746648),   # ERP/tests/bsdfs/test_hapke.py:82-94  (hot spot)
        (-89.0, 0.15426355),  # :97-112 (grazing outgoing direction)
        (80.0, 0.19555340),   # :115-127 (backward)
    ],
)
def test_hapke_golden(oracle, theta_o_deg, golden):
    _, desc = make_desc(surface={"type": "hapke", **POMMEROL})
    ti, to = np.deg2rad(30.0), np.deg2rad(theta_o_deg)
    wi, wo = sph_to_dir(ti, 0.0), sph_to_dir(to, 0.0)
    val = oracle.bsdf_eval(desc, wi, wo)[0] / abs(np.cos(to)) * np.pi
    assert np.allclose(val, golden, rtol=1e-5)


def test_hapke_reciprocity(oracle):
    # ERP/tests/bsdfs/test_hapke.py:255-273
    _, desc = make_desc(surface={"type": "hapke", **POMMEROL})
    rng = np.random.default_rng(1)
    ti, to = rng.uniform(0.05, 1.4, 64), rng.uniform(0.05, 1.4, 64)
    pi_, po = rng.uniform(0, 2 * np.pi, 64), rng.uniform(0, 2 * np.pi, 64)
    wi, wo = sph_to_dir(ti, pi_), sph_to_dir(to, po)
    a = oracle.bsdf_eval(desc, wi, wo) / np.cos(to)
    b = oracle.bsdf_eval(desc, wo, wi) / np.cos(ti)
    assert np.allclose(a, b, rtol=1e-3)


# -------------------------------------------------------------------------------- RPV
def rpv_model(rho_0, rho_c, k, g, theta_i, phi_i, theta_o, phi_o):
    """RPV BRDF (sr^-1, no foreshortening) from the published model equations."""
    ci, co = np.cos(theta_i), np.cos(theta_o)
    si, so = np.sin(theta_i), np.sin(theta_o)
    ti, to = np.tan(theta_i), np.tan(theta_o)
    cdphi = np.cos(phi_i - phi_o)
    # In the (wi, wo both pointing away from the surface) convention the phase angle
    # cosine carries a +: cos g = ci co + si so cos(dphi)
    cosg = ci * co + si * so * cdphi
    F = (1 - g**2) / (1 + g**2 + 2 * g * cosg) ** 1.5
    G = np.sqrt(np.maximum(ti**2 + to**2 - 2 * ti * to * cdphi, 0.0))
    H = 1 + (1 - rho_c) / (1 + G)
    M = (ci * co * (ci + co)) ** (k - 1)
    return rho_0 * M * F * H / np.pi


@pytest.mark.parametrize(
    "rho_0,k,g,rho_c",
    [
        (0.004, 0.543, -0.29, 0.004),   # ERP/tests/bsdfs/test_rpv.py:66-87 parameter sets
        (0.1, 0.9, -0.1, 0.1),
        (0.027685, 0.95, -0.1, 0.027685),  # test_cases/atmospheres.py:100
        (0.2, 0.6, 0.3, 0.5),
    ],
)
def test_rpv_vs_model(oracle, rho_0, k, g, rho_c):
    _, desc = make_desc(surface={"type": "rpv", "rho_0": rho_0, "k": k, "g": g, "rho_c": rho_c})
    rng = np.random.default_rng(0)
    n = 256
    ti, to = rng.uniform(0.0, 1.5, n), rng.uniform(0.0, 1.5, n)
    pi_, po = rng.uniform(0, 2 * np.pi, n), rng.uniform(0, 2 * np.pi, n)
    val = oracle.bsdf_eval(desc, sph_to_dir(ti, pi_), sph_to_dir(to, po))
    ref = rpv_model(rho_0, rho_c, k, g, ti, pi_, to, po) * np.cos(to)
    assert np.allclose(val, ref, rtol=1e-3)  # tolerance of test_rpv.py:87
    assert np.allclose(val, ref, rtol=1e-6)  # params are stored as float32


def test_rpv_default_rho_c_is_rho_0(oracle):
    # rpv.cpp:86-89
    _, d1 = make_desc(surface={"type": "rpv", "rho_0": 0.1, "k": 0.8, "g": -0.2})
    _, d2 = make_desc(surface={"type": "rpv", "rho_0": 0.1, "k": 0.8, "g": -0.2, "rho_c": 0.1})
    wi, wo = sph_to_dir([0.3], [0.1]), sph_to_dir([0.9], [2.0])
    assert oracle.bsdf_eval(d1, wi, wo)[0] == oracle.bsdf_eval(d2, wi, wo)[0]


def test_rpv_degenerate_is_lambertian(oracle):
    # ERP/tests/bsdfs/test_rpv.py:90-112: k=1, g=0, rho_c=1  ==  diffuse
    _, rpv = make_desc(surface={"type": "rpv", "rho_0": 0.3, "k": 1.0, "g": 0.0, "rho_c": 1.0})
    _, dif = make_desc(surface={"type": "diffuse", "reflectance": 0.3})
    rng = np.random.default_rng(3)
    wi = sph_to_dir(rng.uniform(0, 1.5, 100), rng.uniform(0, 6.28, 100))
    wo = sph_to_dir(rng.uniform(0, 1.5, 100), rng.uniform(0, 6.28, 100))
    assert np.allclose(oracle.bsdf_eval(rpv, wi, wo), oracle.bsdf_eval(dif, wi, wo), rtol=1e-6)


@pytest.mark.parametrize("bsdf", [
    {"type": "rpv", "rho_0": 0.1, "k": 0.9, "g": -0.1},
    {"type": "rtls"},
    {"type": "hapke", **POMMEROL},
    {"type": "diffuse", "reflectance": 0.4},
])
def test_sample_weight_is_eval_over_pdf(oracle, bsdf):
    # ERP/tests/bsdfs/test_rpv.py:151-162
    _, desc = make_desc(surface=bsdf)
    rng = np.random.default_rng(4)
    n = 200
    wi = sph_to_dir(rng.uniform(0, 1.4, n), rng.uniform(0, 6.28, n))
    u = rng.uniform(0, 1, (n, 3))
    wo, w = oracle.bsdf_sample(desc, wi, u)
    pdf = wo[:, 2] / np.pi
    ev = oracle.bsdf_eval(desc, wi, wo)
    assert np.allclose(w, ev / pdf, rtol=1e-9)
    # cosine-hemisphere warp
    assert np.allclose(np.linalg.norm(wo, axis=1), 1.0)


def test_bsdf_below_horizon_is_zero(oracle):
    _, desc = make_desc(surface={"type": "rpv", "rho_0": 0.1, "k": 0.9, "g": -0.1})
    up, down = sph_to_dir([0.3], [0.0]), sph_to_dir([2.5], [0.0])
    assert oracle.bsdf_eval(desc, up, down)[0] == 0.0
    assert oracle.bsdf_eval(desc, down, up)[0] == 0.0


# ------------------------------------------------------------------------------- RTLS
def rtls_model(f_iso, f_vol, f_geo, theta_i, phi_i, theta_o, phi_o, h=2.0, r=1.0, b=1.0):
    """Ross-thick / Li-sparse-reciprocal BRDF from the MODIS ATBD kernels."""
    dphi = phi_i - phi_o
    ci, co = np.cos(theta_i), np.cos(theta_o)
    cos_xi = ci * co + np.sin(theta_i) * np.sin(theta_o) * np.cos(dphi)
    xi = np.arccos(cos_xi)
    k_vol = ((np.pi / 2 - xi) * cos_xi + np.sin(xi)) / (ci + co) - np.pi / 4
    tip, top = b / r * np.tan(theta_i), b / r * np.tan(theta_o)
    thi, tho = np.arctan(tip), np.arctan(top)
    cos_xip = np.cos(thi) * np.cos(tho) + np.sin(thi) * np.sin(tho) * np.cos(dphi)
    sec = 1 / np.cos(thi) + 1 / np.cos(tho)
    D = np.sqrt(tip**2 + top**2 - 2 * tip * top * np.cos(dphi))
    cos_t = np.clip(h / b * np.sqrt(D**2 + (tip * top * np.sin(dphi)) ** 2) / sec, -1, 1)
    t = np.arccos(cos_t)
    O = (t - np.sin(t) * cos_t) * sec / np.pi
    k_geo = O - sec + 0.5 * (1 + cos_xip) / (np.cos(thi) * np.cos(tho))
    return (f_iso + f_vol * k_vol + f_geo * k_geo) / np.pi, k_vol, k_geo


def test_rtls_fixed_geometry(oracle):
    # ERP/tests/bsdfs/test_rtls.py:134-160: regression geometry, K_vol only
    ti, to, pi_, po = 0.18430089, 1.46592582, 3.92553451, 0.39280281
    _, desc = make_desc(surface={"type": "rtls", "f_iso": 0.0, "f_vol": 1.0, "f_geo": 0.0})
    val = oracle.bsdf_eval(desc, sph_to_dir([ti], [pi_]), sph_to_dir([to], [po]))[0] / np.cos(to)
    ref, _, _ = rtls_model(0.0, 1.0, 0.0, ti, pi_, to, po)
    assert np.allclose(val, ref, rtol=1e-6)


def test_rtls_defaults(oracle):
    # ERP/bsdfs/rtls.cpp:63-78 and test_rtls.py:117-131 (printed defaults)
    sc, desc = make_desc(surface={"type": "rtls"})
    p = list(desc.bsdf_params)[:6]
    assert np.allclose(p, [0.209741, 0.081384, 0.004140, 2.0, 1.0, 1.0], rtol=1e-6)


@pytest.mark.parametrize("params", [
    dict(f_iso=0.209741, f_vol=0.081384, f_geo=0.004140),
    dict(f_iso=0.3, f_vol=0.2, f_geo=0.05),
    dict(f_iso=0.3, f_vol=0.2, f_geo=0.05, h=1.5, r=1.2, b=0.9),
])
def test_rtls_vs_model(oracle, params):
    _, desc = make_desc(surface={"type": "rtls", **params})
    rng = np.random.default_rng(5)
    n = 256
    ti, to = rng.uniform(0.0, 1.4, n), rng.uniform(0.0, 1.4, n)
    pi_, po = rng.uniform(0, 2 * np.pi, n), rng.uniform(0, 2 * np.pi, n)
    val = oracle.bsdf_eval(desc, sph_to_dir(ti, pi_), sph_to_dir(to, po))
    ref, _, _ = rtls_model(
        params["f_iso"], params["f_vol"], params["f_geo"], ti, pi_, to, po,
        params.get("h", 2.0), params.get("r", 1.0), params.get("b", 1.0),
    )
    assert np.allclose(val, ref * np.cos(to), rtol=1e-3, atol=1e-9)


# ---------------------------------------------------------------------- phase functions
def test_tabphase_eval(oracle):
    # MI/src/phase/tests/test_tabphase.py:13-66
    ref_y = np.array([0.5, 1.0, 1.5])
    ref_x = np.linspace(-1, 1, 3)
    integral = np.trapezoid(ref_y, ref_x)
    _, desc = make_desc(atmosphere="afgl", phase={"type": "tabphase", "values": "0.5, 1.0, 1.5"})
    wi = np.array([0.0, 0.0, -1.0])
    thetas, phis = np.linspace(0, np.pi / 2, 16), np.linspace(0, np.pi, 16)
    wos = np.array([sph_to_dir(t, p) for t in thetas for p in phis])
    cos_graphics = wos @ wi
    ref = 0.5 / np.pi * np.interp(-cos_graphics, ref_x, ref_y) / integral
    val = oracle.phase_eval(desc, 0, cos_graphics)
    assert np.allclose(val, ref)


def test_tabphase_sample_convention(oracle):
    # MI/src/phase/tests/test_tabphase.py:69-93: u=1 -> forward scattering, pdf = 0.5/pi
    _, desc = make_desc(atmosphere="afgl", phase={"type": "tabphase", "values": "0.0, 0.5, 1.0"})
    ct, w, pdf = oracle.phase_sample(desc, 0, [[1.0, 0.0]])
    assert np.allclose(ct, 1.0)  # propagation direction preserved == wo = -wi
    assert np.allclose(pdf, 0.5 / np.pi)
    assert np.allclose(w, 1.0)


def test_distr_regular_matches_numpy(oracle):
    # distr_1d.h:300-620: trapezoid CDF + linear-segment inversion
    rng = np.random.default_rng(6)
    pdf = rng.uniform(0.1, 2.0, 33).astype(np.float32)
    x = np.linspace(-1, 1, 33)
    u = rng.uniform(0, 1, 1000)
    xs, pe, integral = oracle.distr_regular(pdf, u=u, xq=rng.uniform(-1, 1, 1000))
    assert np.allclose(integral, np.trapezoid(pdf.astype(float), x), rtol=1e-12)
    # CDF(xs) == u
    cdf_nodes = np.concatenate([[0], np.cumsum(0.5 * (pdf[1:] + pdf[:-1]) * np.diff(x))])
    i = np.clip(np.searchsorted(x, xs, side="right") - 1, 0, 31)
    t = xs - x[i]
    slope = (pdf[i + 1] - pdf[i]) / (x[i + 1] - x[i])
    cdf = cdf_nodes[i] + pdf[i] * t + 0.5 * slope * t * t
    assert np.allclose(cdf / integral, u, atol=1e-9)


def test_distr_irregular_equals_regular_on_regular_nodes(oracle):
    # distr_1d.h:628-1000 must agree with :300-620 when nodes are equidistant
    rng = np.random.default_rng(7)
    pdf = rng.uniform(0.0, 2.0, 17).astype(np.float32)
    nodes = np.linspace(-1, 1, 17).astype(np.float32)
    u, xq = rng.uniform(0, 1, 500), rng.uniform(-1, 1, 500)
    a = oracle.distr_regular(pdf, u=u, xq=xq)
    b = oracle.distr_irregular(nodes, pdf, u=u, xq=xq)
    assert np.allclose(a[0], b[0], atol=1e-6)
    assert np.allclose(a[1], b[1], atol=1e-6)
    assert np.allclose(a[2], b[2], rtol=1e-6)


@pytest.mark.parametrize("phase", [
    {"type": "rayleigh"},
    {"type": "rayleigh", "depolarization": 0.03},
    {"type": "hg", "g": 0.7},
    {"type": "hg", "g": -0.4},
    {"type": "isotropic"},
    {"type": "tabphase", "values": "0.5, 1.0, 1.5, 3.0, 9.0"},
])
def test_phase_normalisation_and_sampling(oracle, phase):
    """chi^2-style consistency of sample() and eval_pdf() (test_rayleigh.py:10,
    test_hg.py:11, test_tabphase.py:95) done with a histogram in cos(theta)."""
    _, desc = make_desc(atmosphere="afgl", phase=phase)
    # normalisation: 2 pi * int pdf dmu = 1   (graphics cosine c; symmetric integral)
    mu = np.linspace(-1, 1, 20001)
    val = oracle.phase_eval(desc, 0, mu)
    total = 2 * np.pi * np.trapezoid(val, mu)
    rho = phase.get("depolarization", 0.0)
    assert np.allclose(total, 1.0, rtol=2e-4) or rho > 0  # depolarised value is not a pdf
    rng = np.random.default_rng(8)
    n = 400000
    ct, w, pdf = oracle.phase_sample(desc, 0, rng.uniform(0, 1, (n, 2)))
    hist, edges = np.histogram(ct, bins=40, range=(-1, 1))
    # expected mass per bin from the pdf (physics cosine ct = -c)
    fine = np.linspace(-1, 1, 40 * 200 + 1)
    p_fine = oracle.phase_eval(desc, 0, -fine) if rho == 0 else None
    if p_fine is not None:
        mass = 2 * np.pi * np.add.reduceat(
            0.5 * (p_fine[1:] + p_fine[:-1]) * np.diff(fine), np.arange(0, 8000, 200)
        )
        expected = mass * n
        chi2 = np.sum((hist - expected) ** 2 / np.maximum(expected, 1.0))
        assert chi2 < 100.0, chi2  # 39 dof, P(chi2 > 100) ~ 1e-7
        # sampled pdf equals eval at the sampled angle
        assert np.allclose(pdf, oracle.phase_eval(desc, 0, -ct), rtol=1e-9)
    else:
        assert np.allclose(w, oracle.phase_eval(desc, 0, -ct) / pdf, rtol=1e-9)


def test_hg_matches_tabulated_hg(oracle):
    mu, p = scenes.hg_table(0.7, 2001)
    _, d_tab = make_desc(atmosphere="afgl", phase={"type": "tabphase", "values": ",".join(map(str, p))})
    _, d_hg = make_desc(atmosphere="afgl", phase={"type": "hg", "g": 0.7})
    c = np.linspace(-1, 1, 101)
    assert np.allclose(oracle.phase_eval(d_tab, 0, c), oracle.phase_eval(d_hg, 0, c), rtol=2e-3)


# ------------------------------------------------------------------------------ warps
def test_warps(oracle):
    # MI/include/mitsuba/core/warp.h:54-90, :374-433
    assert np.allclose(oracle.warp("uniform_disk_concentric", [0.5], [0.5]), [[0, 0]])
    assert np.allclose(oracle.warp("uniform_disk_concentric", [1.0], [0.5]), [[1, 0]])
    assert np.allclose(oracle.warp("uniform_disk_concentric", [0.5], [1.0]), [[0, 1]], atol=1e-15)
    assert np.allclose(oracle.warp("cosine_hemisphere", [0.5], [0.5]), [[0, 0, 1]])
    assert np.allclose(oracle.warp("uniform_hemisphere", [0.5], [0.5]), [[0, 0, 1]])
    rng = np.random.default_rng(9)
    u, v = rng.uniform(0, 1, 20000), rng.uniform(0, 1, 20000)
    ch = oracle.warp("cosine_hemisphere", u, v)
    uh = oracle.warp("uniform_hemisphere", u, v)
    assert np.allclose(np.linalg.norm(ch, axis=1), 1) and np.allclose(np.linalg.norm(uh, axis=1), 1)
    assert abs(ch[:, 2].mean() - 2 / 3) < 5e-3  # E[cos] under a cosine-weighted density
    assert abs(uh[:, 2].mean() - 0.5) < 5e-3    # E[cos] under a uniform density


# ---------------------------------------------------------------------------- sensors
def test_mdistant_directions(oracle):
    # ERP/tests/sensors/test_mdistant.py:94-125: ray directions == normalised `directions`
    dirs = np.array([[0, 0, -1], [1, 0, -1], [0, 1, -2], [-1, -1, -1]], float)
    d = scenes.atmosphere_scene(geometry="plane_parallel", atmosphere=None,
                                sensor={"type": "mdistant", "vza": [0, 1, 2, 3]})
    d["measure"]["directions"] = ",".join(map(str, dirs.ravel()))
    d["measure"]["target"] = [1.0, 2.0, 3.0]
    sc = mi_load_dict(d)
    desc = sc.flat.build_desc()
    fs = np.array([[(i + 0.5) / 4, 0.5] for i in range(4)])
    o, dd, w = oracle.sensor_ray(desc, 0, fs, np.full((4, 2), 0.5))
    assert np.allclose(dd, dirs / np.linalg.norm(dirs, axis=1, keepdims=True))
    assert np.allclose(w, 1.0)
    # origin = target - d * ray_offset, ray_offset = 2 * bsphere radius (mdistant.cpp:180-190)
    off = np.linalg.norm(o - np.array([1.0, 2.0, 3.0]), axis=1)
    assert np.allclose(off, 2 * sc.flat.bsphere_radius, rtol=1e-6)
    assert np.allclose(np.cross(o - np.array([1.0, 2.0, 3.0]), dd), 0, atol=1e-3)


def test_hdistant_and_distantflux_rays(oracle):
    # ERP/sensors/hdistant.cpp:248-250, distantflux.cpp:163-170
    for ty in ("hdistant", "distantflux"):
        d = scenes.atmosphere_scene(geometry="plane_parallel", atmosphere=None,
                                    sensor={"type": ty, "film_resolution": (4, 4)})
        sc = mi_load_dict(d)
        desc = sc.flat.build_desc()
        rng = np.random.default_rng(10)
        fs = rng.uniform(0, 1, (64, 2))
        o, dd, w = oracle.sensor_ray(desc, 0, fs, rng.uniform(0, 1, (64, 2)))
        h = oracle.warp("uniform_hemisphere", fs[:, 0], fs[:, 1])
        assert np.allclose(dd, -h)
        if ty == "hdistant":
            assert np.allclose(w, 1.0)
        else:
            assert np.allclose(w, h[:, 2] * 2 * np.pi / 16)


def test_target_rectangle_weight(oracle):
    # mdistant.cpp:222-227: weight 1/(pdf*area) == 1 for a uniformly sampled rectangle
    from eradiate_b200.kernel import ScalarTransform4f
    d = scenes.atmosphere_scene(geometry="plane_parallel", atmosphere=None,
                                sensor={"type": "mdistant", "vza": [0.0, 40.0]})
    d["measure"]["target"] = {"type": "rectangle",
                              "to_world": ScalarTransform4f().translate([5, 6, 0]).scale([2, 3, 1])}
    sc = mi_load_dict(d)
    desc = sc.flat.build_desc()
    rng = np.random.default_rng(11)
    ap = rng.uniform(0, 1, (200, 2))
    o, dd, w = oracle.sensor_ray(desc, 0, np.full((200, 2), 0.25), ap)
    off = 2 * sc.flat.bsphere_radius
    tgt = o + dd * off
    assert np.allclose(w, 1.0)
    assert np.all(np.abs(tgt[:, 0] - 5) <= 2 + 1e-3) and np.all(np.abs(tgt[:, 1] - 6) <= 3 + 1e-3)
    assert np.allclose(tgt[:, 2], 0, atol=1e-2)


# ---------------------------------------------------------------------------- volumes
def test_spherical_volume_lookup(oracle):
    # ERP/tests/volumes/test_spherical.py:28-56: r -> layer remap, zero fill outside [rmin, rmax]
    sc, desc = make_desc(geometry="spherical_shell", atmosphere="afgl", n_layers=6)
    R, H = scenes.EARTH_RADIUS, scenes.TOA
    st_ref = np.array(list(desc.sigma_t[0:6]), dtype=np.float64)
    alts = (np.arange(6) + 0.5) * H / 6
    pts = np.array([[0, 0, R + a] for a in alts] + [[R + alts[2], 0, 0]] + [[0, R - 10.0, 0], [0, 0, R + H + 1]])
    st, al = oracle.medium_lookup(desc, pts)
    assert np.allclose(st[:6], st_ref)
    assert st[6] == st_ref[2]
    assert st[7] == 0.0 and st[8] == 0.0  # fillmin / fillmax


def test_grid_nearest_layer_convention_piecewise_golden(oracle):
    """
    ERP/tests/media/test_piecewise.py:6-56: exponential 10-layer medium, distances at
    which the transmittance from the top (looking down) reaches 0.75/0.5/0.25/0.01.
    Pins the nearest-filter layer convention (layer i spans [i, i+1) * dz) and the
    optical-depth integration the free-flight sampler must reproduce in expectation.
    """
    H, n, integral, lbd = 100000.0, 10, 3.0, 8300.0
    z = np.linspace(0.0, H, n, endpoint=False)
    ext = (integral / lbd) * np.exp(-z / lbd)
    d = scenes.atmosphere_scene(geometry="plane_parallel", atmosphere="afgl", n_layers=n, toa=H)
    sc = mi_load_dict(d)
    sc.flat.medium.children["sigma_t"].values["data"][:] = ext.reshape(-1, 1, 1, 1).astype(np.float32)
    desc = sc.flat.build_desc()
    # optical depth from the top along -z using the oracle's own lookup
    zs = np.linspace(H, 0.0, 2_000_001)
    mid = 0.5 * (zs[1:] + zs[:-1])
    st, _ = oracle.medium_lookup(desc, np.stack([np.zeros_like(mid), np.zeros_like(mid), mid], axis=1))
    tau = np.concatenate([[0], np.cumsum(st * (H / 2_000_000))])
    gt_dists = [74578.93910366, 82117.51365975, 88515.28423884, 98460.51859237]
    gt_trs = [0.75, 0.5, 0.25, 0.01]
    for dist, tr in zip(gt_dists, gt_trs):
        got = np.interp(-np.log(tr), tau, H - zs)
        assert np.allclose(got, dist, rtol=2e-5), (got, dist)


# ------------------------------------------------------------------------ ocean_legacy
OCEAN = dict(type="ocean_legacy", component=0, wavelength=1500.0, wind_speed=1.0, wind_direction=90.0,
             chlorinity=19.0, pigmentation=0.3, shadowing=False)


@pytest.mark.parametrize("override,golden", [
    # ERP/tests/bsdfs/test_ocean_legacy.py:52-101 (6SV reference values, rtol 1e-3, atol 1e-4)
    ({}, [1.91408132e-03, -5.44804487e-08, -5.45187861e-08, -5.45187861e-08, -5.45187861e-08]),
    ({"wavelength": 550.0, "wind_speed": 30.0}, [0.11300096, 0.10733355, 0.10722339, 0.1055309, 0.11909937]),
])
def test_ocean_legacy_6sv_golden(oracle, override, golden):
    _, desc = make_desc(surface={**OCEAN, **override})
    vza = np.deg2rad([0.0, 22.475, 44.95, 67.425, 89.9])
    wi = sph_to_dir(vza, np.zeros(5))                      # si.wi = view directions
    wo = sph_to_dir(np.full(5, np.deg2rad(22.475)), np.zeros(5))  # wo = sun direction
    val = oracle.bsdf_eval(desc, wi, wo) * np.pi
    assert np.allclose(val, golden, rtol=1e-3, atol=1e-4), val


def test_ocean_legacy_sample_pdf_consistency(oracle):
    """chi^2-style check of ERP/tests/bsdfs/test_ocean_legacy.py:18-33: the sampled directions
    follow pdf(): E[eval/pdf] over sample() equals the hemispherical integral of eval."""
    _, desc = make_desc(surface={**OCEAN, "wavelength": 550.0, "wind_speed": 10.0, "shadowing": True})
    rng = np.random.default_rng(12)
    wi = sph_to_dir([np.deg2rad(35.0)], [0.4])
    n = 200000
    wo, w = oracle.bsdf_sample(desc, np.repeat(wi, n, axis=0), rng.uniform(0, 1, (n, 3)))
    assert np.all(np.isfinite(w)) and np.all(w >= 0)
    est = w.mean()
    # reference integral by uniform-hemisphere quadrature of eval()
    u = rng.uniform(0, 1, (n, 2))
    d = oracle.warp("uniform_hemisphere", u[:, 0], u[:, 1])
    ref = (oracle.bsdf_eval(desc, np.repeat(wi, n, axis=0), d) * 2 * np.pi)
    assert abs(est - ref.mean()) < 5 * np.sqrt(w.var() / n + ref.var() / n), (est, ref.mean())


# --------------------------------------------------------- ocean_mishchenko / ocean_grasp / maignan
MISHCHENKO = dict(type="ocean_mishchenko", wind_speed=2.0, eta=1.33, k=0.0, ext_ior=1.0)
MAIGNAN = dict(type="maignan", C=5.0, ndvi=0.8, refr_re=1.5, refr_im=0.0, ext_ior=1.0)


def _deg_dir(theta, phi):
    return sph_to_dir([np.deg2rad(theta)], [np.deg2rad(phi)])


@pytest.mark.parametrize("override,vza,sza,saa,golden,rtol,atol", [
    # ERP/tests/bsdfs/test_ocean_mishchenko.py:44-109: Mishchenko's reference code, 550 nm / 2 m/s ...
    ({}, 15.0, 15.0, 180.0,
     [[0.125155, -0.0132689, 0.0, 0.0], [-0.0132689, 0.125155, 0.0, 0.0],
      [0.0, 0.0, -0.124450, 0.0], [0.0, 0.0, 0.0, -0.124450]], 1e-4, 1e-5),
    # ... and 900 nm / 10 m/s / eta 1.39 after a parameter update
    ({"wind_speed": 10.0, "eta": 1.39}, 60.0, 40.0, 180.0,
     [[0.733924e-01, -0.713385e-01, 0.0, 0.0], [-0.713385e-01, 0.733924e-01, 0.0, 0.0],
      [0.0, 0.0, -0.172412e-01, 0.0], [0.0, 0.0, 0.0, -0.172412e-01]], 1e-3, 1e-4),
])
def test_ocean_mishchenko_golden_mueller(oracle, override, vza, sza, saa, golden, rtol, atol):
    _, desc = make_desc(surface={**MISHCHENKO, **override}, stokes=True)
    M = oracle.bsdf_mueller(desc, _deg_dir(vza, 0.0), _deg_dir(sza, saa))[0]
    assert np.allclose(M, golden, rtol=rtol, atol=atol), M
    # unpolarized variants return the (0, 0) entry (ocean_mishchenko.cpp:289-293)
    _, d0 = make_desc(surface={**MISHCHENKO, **override})
    assert np.isclose(oracle.bsdf_eval(d0, _deg_dir(vza, 0.0), _deg_dir(sza, saa))[0], M[0, 0], rtol=1e-12)


def test_ocean_mishchenko_update_equals_fresh_scene(oracle):
    # test_ocean_mishchenko.py:91-96: params["wind_speed"], params["eta.value"] then update()
    from eradiate_b200.kernel import mi_traverse
    sc, _ = make_desc(surface=MISHCHENKO, stokes=True)
    w = mi_traverse(sc)
    keys = {k.split("surface_bsdf.")[-1]: k for k in w.parameters.keys() if k.startswith("surface_bsdf.")}
    assert set(keys) == {"wind_speed", "eta.value", "k.value", "ext_ior.value"}  # ocean_mishchenko.cpp:118-123
    w.parameters.update({keys["wind_speed"]: 10.0, keys["eta.value"]: 1.39})
    M = oracle.bsdf_mueller(sc.flat.build_desc(), _deg_dir(60.0, 0.0), _deg_dir(40.0, 180.0))[0]
    _, fresh = make_desc(surface={**MISHCHENKO, "wind_speed": 10.0, "eta": 1.39}, stokes=True)
    assert np.array_equal(M, oracle.bsdf_mueller(fresh, _deg_dir(60.0, 0.0), _deg_dir(40.0, 180.0))[0])


def _fresnel_unpolarized(n, cos_i):
    sin_t2 = (1.0 - cos_i**2) / n**2
    cos_t = np.sqrt(1.0 - sin_t2)
    rs = (cos_i - n * cos_t) / (cos_i + n * cos_t)
    rp = (n * cos_i - cos_t) / (n * cos_i + cos_t)
    return 0.5 * (rs**2 + rp**2)


def test_ocean_mishchenko_m00_vs_microfacet_model(oracle):
    """(0, 0) entry = D G F / (4 cos_i) written from the published formulas: Beckmann facets with
    alpha^2 = Cox-Munk mean square slope 0.003 + 0.00512 U, Walter et al. 2007 rational Smith Lambda,
    unpolarized Fresnel reflectance at the facet."""
    ws, eta = 7.0, 1.34
    _, desc = make_desc(surface={**MISHCHENKO, "wind_speed": ws, "eta": eta})
    rng = np.random.default_rng(5)
    n = 2000
    wi = sph_to_dir(rng.uniform(0.05, 1.2, n), rng.uniform(0, 2 * np.pi, n))
    wo = sph_to_dir(rng.uniform(0.05, 1.2, n), rng.uniform(0, 2 * np.pi, n))
    alpha = np.sqrt(0.003 + 0.00512 * ws)
    m = wi + wo
    m /= np.linalg.norm(m, axis=1, keepdims=True)
    ct = m[:, 2]
    D = np.exp(-(1.0 - ct**2) / ct**2 / alpha**2) / (np.pi * alpha**2 * ct**4)

    def lam(v):
        a = 1.0 / (alpha * np.sqrt(1.0 - v[:, 2] ** 2) / v[:, 2])
        return np.where(a >= 1.6, 0.0, (1.0 - 1.259 * a + 0.396 * a * a) / (3.535 * a + 2.181 * a * a))

    G = 1.0 / (1.0 + lam(wi) + lam(wo))
    F = _fresnel_unpolarized(eta, np.sum(wi * m, axis=1))
    ref = D * G * F / (4.0 * wi[:, 2])
    got = oracle.bsdf_eval(desc, wi, wo)
    ok = D * ct > 1e-20
    assert np.allclose(got[ok], ref[ok], rtol=1e-6, atol=1e-12), np.max(np.abs(got[ok] / ref[ok] - 1))


@pytest.mark.parametrize("surface", [
    {**MISHCHENKO, "wind_speed": 6.0},
    {"type": "ocean_grasp", "wavelength": 550.0, "wind_speed": 12.0, "eta": 1.336, "water_body_reflectance": 0.05},
])
def test_glint_sample_pdf_consistency(oracle, surface):
    """chi^2-style check (test_ocean_mishchenko.py:15-31, test_ocean_grasp.py:15-31): E[weight] over
    sample() equals the hemispherical integral of eval()."""
    _, desc = make_desc(surface=surface)
    rng = np.random.default_rng(21)
    wi = sph_to_dir([np.deg2rad(40.0)], [0.7])
    n = 200000
    wo, w = oracle.bsdf_sample(desc, np.repeat(wi, n, axis=0), rng.uniform(0, 1, (n, 3)))
    assert np.all(np.isfinite(w)) and np.all(w >= 0)
    # the glint lobe is narrow: integrate eval() by importance sampling a cone around the mirror direction
    mirror = np.array([-wi[0, 0], -wi[0, 1], wi[0, 2]])
    cos_max = np.cos(np.deg2rad(60.0))
    u = rng.uniform(0, 1, (n, 2))
    ct = 1.0 - u[:, 0] * (1.0 - cos_max)
    st, ph = np.sqrt(1.0 - ct**2), 2 * np.pi * u[:, 1]
    a = np.cross(mirror, [0.0, 0.0, 1.0]); a /= np.linalg.norm(a)
    b = np.cross(mirror, a)
    d = ct[:, None] * mirror + st[:, None] * (np.cos(ph)[:, None] * a + np.sin(ph)[:, None] * b)
    cone = oracle.bsdf_eval(desc, np.repeat(wi, n, axis=0), d) * 2 * np.pi * (1.0 - cos_max)
    # outside the cone only the diffuse part of ocean_grasp remains
    d2 = oracle.warp("uniform_hemisphere", u[:, 0], u[:, 1])
    out = np.where(d2 @ mirror < cos_max, oracle.bsdf_eval(desc, np.repeat(wi, n, axis=0), d2) * 2 * np.pi, 0.0)
    ref, var = cone.mean() + out.mean(), cone.var() / n + out.var() / n
    assert abs(w.mean() - ref) < 5 * np.sqrt(w.var() / n + var), (w.mean(), ref)


def test_ocean_grasp_components(oracle):
    """ocean_grasp.cpp:354-455 term by term: Frouin whitecaps x Monahan coverage (oceanprops.h:330-363) and
    the Lambertian water body outside the glint; the glint is Mishchenko's with the exact Smith Lambda
    (erf form, :246-255) instead of the rational fit: same value within the fit's accuracy."""
    wl, ws, wbr = 865.0, 15.0, 0.03
    g = {"type": "ocean_grasp", "wavelength": wl, "wind_speed": ws, "eta": 1.33, "k": 0.0, "ext_ior": 1.0,
         "water_body_reflectance": wbr}
    _, desc = make_desc(surface=g)
    cov = 2.95e-6 * ws**3.52
    wc = cov * 0.22 * np.exp(-1.75 * (wl * 1e-3 - 0.6) ** 0.99)
    # far from the specular direction (backscattering side) only the diffuse terms remain
    wi, wo = _deg_dir(50.0, 0.0), _deg_dir(50.0, 10.0)
    val = oracle.bsdf_eval(desc, wi, wo)[0]
    assert np.isclose(val, (wc + (1.0 - cov) * wbr) * wo[0, 2] / np.pi, rtol=1e-9)
    # in the glint: subtract the diffuse part, compare with ocean_mishchenko
    _, dm = make_desc(surface={**MISHCHENKO, "wind_speed": ws})
    rng = np.random.default_rng(3)
    n = 500
    wi = sph_to_dir(rng.uniform(0.1, 1.0, n), np.zeros(n))
    wo = sph_to_dir(np.arccos(wi[:, 2]) + rng.normal(0, 0.08, n), np.full(n, np.pi) + rng.normal(0, 0.08, n))
    glint = oracle.bsdf_eval(desc, wi, wo) - (wc + (1.0 - cov) * wbr) * wo[:, 2] / np.pi
    ref = (1.0 - cov) * oracle.bsdf_eval(dm, wi, wo)
    ok = ref > 1e-3
    assert ok.sum() > 300 and np.allclose(glint[ok], ref[ok], rtol=2e-2)


@pytest.mark.parametrize("override,wi_deg,wo_deg,golden", [
    # ERP/tests/bsdfs/test_maignan.py:28-58 (evergreen needleleaf, 550 nm) and :60-88 (savanna, 900 nm);
    # the reference values are float32 outputs: compared to 2e-7 absolute
    ({"C": 4.98, "ndvi": 0.8}, (40.0, 0.0), (40.0, 5.0),
     [[1.42013086e-02, 1.48094732e-05, -1.60061711e-06, 0.0],
      [1.48624285e-05, 1.41894910e-02, -5.79237472e-04, 0.0],
      [9.95336109e-07, -5.79236308e-04, -1.41894845e-02, 0.0],
      [0.0, 0.0, 0.0, -1.42013021e-02]]),
    ({"C": 6.66, "ndvi": 0.3}, (0.0, 0.0), (40.0, 170.0),
     [[1.9543538e-02, -3.1128337e-03, -1.1320484e-03, 0.0],
      [3.1127632e-03, -1.9510504e-02, -8.9618654e-05, 0.0],
      [-1.1322410e-03, 9.2017806e-05, 1.9293837e-02, 0.0],
      [0.0, 0.0, 0.0, -1.9260805e-02]]),
])
def test_maignan_golden_mueller(oracle, override, wi_deg, wo_deg, golden):
    _, desc = make_desc(surface={**MAIGNAN, **override}, stokes=True)
    M = oracle.bsdf_mueller(desc, _deg_dir(*wi_deg), _deg_dir(*wo_deg))[0]
    # (wi along the normal runs into the clamp mu <= 0.9999999 of oceanprops.h:476-477, which rounds differently in float32)
    assert np.allclose(M, golden, rtol=2e-5, atol=6e-7), M


def test_maignan_sample_conventions(oracle):
    """maignan.cpp:168-224 as written: cosine-hemisphere sampling, pdf = cos / pi, eval() without the
    foreshortening factor, and a sample weight equal to eval() at the sampled direction (not eval / pdf)."""
    _, desc = make_desc(surface={**MAIGNAN, "ext_ior": 1.000277})
    rng = np.random.default_rng(8)
    n = 1000
    wi = sph_to_dir(rng.uniform(0.0, 1.4, n), rng.uniform(0, 2 * np.pi, n))
    u = rng.uniform(0, 1, (n, 3))
    wo, w = oracle.bsdf_sample(desc, wi, u)
    c = oracle.warp("cosine_hemisphere", u[:, 1], u[:, 2])
    assert np.allclose(wo, c, atol=1e-14)
    assert np.allclose(w, oracle.bsdf_eval(desc, wi, wo), rtol=1e-12)
    # Eq. 21 of Maignan et al. 2009 at one geometry, from the formula
    ti, to, dphi = np.deg2rad(30.0), np.deg2rad(50.0), np.deg2rad(140.0)
    cT = np.cos(ti) * np.cos(to) + np.sin(ti) * np.sin(to) * np.cos(dphi)
    tan_a = np.tan(0.5 * np.arccos(cT))  # alpha = half the angle between the two directions
    F = _fresnel_unpolarized(1.5 / 1.000277, np.cos(0.5 * np.arccos(cT)))
    ref = 5.0 * np.exp(-tan_a) * np.exp(-0.8) * F / (4.0 * (np.cos(ti) + np.cos(to)))
    got = oracle.bsdf_eval(desc, sph_to_dir([ti], [0.3]), sph_to_dir([to], [0.3 + dphi]))[0]
    assert np.isclose(got, ref, rtol=1e-6), (got, ref)


# ---------------------------------------------------------------------------------------- mqdiffuse
MQ_SAMPLE_DATA = np.array([  # ERP/tests/bsdfs/test_mqdiffuse.py:66-89: [cos_theta_i][phi_d][cos_theta_o]
    [np.linspace(0, 1, 5), -np.linspace(0, 1, 5), np.linspace(0, 1, 5)],
    [np.linspace(1, 2, 5), -np.linspace(1, 2, 5), np.linspace(1, 2, 5)],
])


def _mq_desc(data=MQ_SAMPLE_DATA):
    from eradiate_b200.kernel import VolumeGrid
    return make_desc(surface={"type": "mqdiffuse", "grid": VolumeGrid(data)})


@pytest.mark.parametrize("theta_o,phi_o,theta_i,phi_i,expected", [
    # ERP/tests/bsdfs/test_mqdiffuse.py:97-114 (hand-picked entries of the table above)
    [np.pi / 3, 0.0, 0.0, 0.0, 1.5],
    [np.pi * 0.4195693767448338, 0.0, 0.0, 0.0, 1.25],
    [np.pi / 3, np.pi, 0.0, 0.0, -1.5],
    [np.pi / 3, 0.5 * np.pi, 0.0, 0.0, 0.0],
    [np.pi / 3, 1.5 * np.pi, 0.0, 0.0, 0.0],
    [np.pi / 2, 0.0, np.pi / 2, 0.0, 0.0],
    [np.pi / 3, 0.0, np.pi / 2, 0.0, 0.5],
    [np.pi / 3, np.pi, np.pi / 2, 0.0, -0.5],
    [np.pi / 2, np.pi, np.pi / 2, 0.0, 0.0],
    [np.pi / 2, np.pi, 0.0, 0.0, -1.0],
])
def test_mqdiffuse_golden_eval(oracle, theta_o, phi_o, theta_i, phi_i, expected):
    _, desc = _mq_desc()
    val = oracle.bsdf_eval(desc, sph_to_dir([theta_i], [phi_i]), sph_to_dir([theta_o], [phi_o]))[0]
    assert np.isclose(val, expected * np.cos(theta_o), atol=1e-6)  # test_mqdiffuse.py:115-125


def test_mqdiffuse_vs_scipy_interpolation_and_sampling(oracle):
    """eval() against an independent trilinear interpolation (scipy) of a random table on the grid the plugin
    defines (nodes at linspace(0, 1) / linspace(0, 2 pi), mqdiffuse.cpp:94-107); sample(): cosine-hemisphere
    directions, weight = value * pi, where a negative azimuth difference is not wrapped (:125-131) and hence
    reads the phi_d = 0 plane."""
    from scipy.interpolate import RegularGridInterpolator

    rng = np.random.default_rng(17)
    nz, ny, nx = 6, 9, 7
    data = rng.uniform(0.05, 0.3, (nz, ny, nx))
    _, desc = _mq_desc(data)
    interp = RegularGridInterpolator((np.linspace(0, 1, nz), np.linspace(0, 2 * np.pi, ny), np.linspace(0, 1, nx)),
                                     data.astype(np.float32).astype(np.float64))
    n = 4000
    wi = sph_to_dir(rng.uniform(0.0, 1.5, n), rng.uniform(-np.pi, np.pi, n))
    wo = sph_to_dir(rng.uniform(0.0, 1.5, n), rng.uniform(-np.pi, np.pi, n))
    phi_d = np.mod(np.arctan2(wo[:, 1], wo[:, 0]) - np.arctan2(wi[:, 1], wi[:, 0]), 2 * np.pi)
    ref = interp(np.stack([wi[:, 2], phi_d, wo[:, 2]], axis=-1)) * wo[:, 2]
    assert np.allclose(oracle.bsdf_eval(desc, wi, wo), ref, rtol=1e-9, atol=1e-12)
    u = rng.uniform(0, 1, (n, 3))
    wo_s, w = oracle.bsdf_sample(desc, wi, u)
    assert np.allclose(wo_s, oracle.warp("cosine_hemisphere", u[:, 1], u[:, 2]), atol=1e-14)
    diff = np.arctan2(wo_s[:, 1], wo_s[:, 0]) - np.arctan2(wi[:, 1], wi[:, 0])
    phi_s = np.where(diff < 0, 0.0, np.fmod(diff, 2 * np.pi))
    ref_w = interp(np.stack([wi[:, 2], phi_s, wo_s[:, 2]], axis=-1)) * np.pi
    assert (diff < 0).mean() > 0.3 and np.allclose(w, ref_w, rtol=1e-9)


# --------------------------------------------------------------------------------------- multiphase
@pytest.mark.parametrize("weight,g", [(0.2, 0.2), (0.8, 0.2)])
def test_multiphase_mixture_eval(oracle, weight, g):
    # ERP/tests/phase/test_multiphase.py:38-68 (eval of the blend at wi = wo = +z) and :71-110 (the pdf of each
    # component as sample() returns it without MIS)
    phase = {"type": "multiphase", "use_mis": False, "phase0": {"type": "isotropic"}, "weight0": weight,
             "phase1": {"type": "hg", "g": g}, "weight1": 1 - weight}
    _, desc = make_desc(phase=phase)
    w = np.ctypeslib.as_array(desc.phase_weight, shape=(2, desc.n_layers))[:, 0]
    vals = [oracle.phase_eval(desc, leaf, [1.0])[0] for leaf in range(2)]
    inv_four_pi = 1.0 / (4.0 * np.pi)
    assert np.isclose(vals[0], inv_four_pi) and np.isclose(vals[1], inv_four_pi * (1 - g) / (1 + g) ** 2)
    expected = weight * inv_four_pi + (1 - weight) * inv_four_pi * (1.0 - g) / (1.0 + g) ** 2
    assert np.isclose(w[0] * vals[0] + w[1] * vals[1], expected, rtol=1e-6)

