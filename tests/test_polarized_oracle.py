"""
Polarized (Stokes / Mueller) path of the CPU oracle, pinned on physics and on the reference's
own checks (ERP/tests/phase/test_rayleigh_polarized.py, test_tabphase_polarized.py: those tests
compare the plugins with mitsuba.mueller itself, so the independent pins are the published
Rayleigh phase matrix -- Hansen & Travis 1974 eq. 2.15 -- and single-scattering polarisation).
CPU only.
"""

import numpy as np
import pytest

from eradiate_b200 import _abi, scenes
from eradiate_b200.kernel import mi_load_dict


def sph(theta, phi):
    return np.array([np.sin(theta) * np.cos(phi), np.sin(theta) * np.sin(phi), np.cos(theta)])


def rayleigh_matrix(ct, rho=0.0):
    """Hansen & Travis (1974) eq. (2.15): P = Delta * P_Rayleigh + (1 - Delta) * isotropic, written with
    the common factor 3/(16 pi) (for rho = 0 this is the matrix of test_rayleigh_polarized.py:23-40;
    for rho != 0 it equals the plugin's r1 * (r2 + cos^2), rayleigh_polarized.cpp:62-77)."""
    delta = (1.0 - rho) / (1.0 + 0.5 * rho)
    delta_p = (1.0 - 2.0 * rho) / (1.0 - rho)
    a, b, c = ct * ct + 1.0, ct * ct - 1.0, 2.0 * ct
    return 3.0 / (16.0 * np.pi) * np.array([
        [a * delta + (4.0 / 3.0) * (1.0 - delta), b * delta, 0, 0],
        [b * delta, a * delta, 0, 0],
        [0, 0, c * delta, 0],
        [0, 0, 0, c * delta * delta_p]])


def pol_desc(phase, **kw):
    d = scenes.atmosphere_scene(geometry="plane_parallel", atmosphere="afgl", n_layers=10, phase=phase,
                                stokes=True, **kw)
    sc = mi_load_dict(d)
    return sc, sc.flat.build_desc()


@pytest.mark.parametrize("rho", [0.0, 0.0279])
def test_rayleigh_polarized_matrix_invariants(oracle, rho):
    """The rotated matrix R_out M R_in^T keeps M00 and the rotation invariants of eq. 2.15:
    M00, M33, and M01^2+M02^2, M10^2+M20^2; the scattering-plane matrix is recovered when the
    implicit bases happen to lie in the scattering plane."""
    sc, desc = pol_desc({"type": "rayleigh_polarized", "depolarization": rho})
    assert desc.polarized == 1 and desc.phase[0].type == _abi.PHASE_RAYLEIGH_POLARIZED
    rng = np.random.default_rng(0)
    n = 200
    wi = np.array([sph(t, p) for t, p in zip(rng.uniform(0.1, 3.0, n), rng.uniform(0, 6.28, n))])
    wo = np.array([sph(t, p) for t, p in zip(rng.uniform(0.1, 3.0, n), rng.uniform(0, 6.28, n))])
    M, pdf = oracle.phase_mueller(desc, 0, wi, wo)
    for k in range(n):
        ct = -np.dot(wo[k], wi[k])
        ref = rayleigh_matrix(ct, np.float32(rho))
        assert np.allclose(M[k, 0, 0], ref[0, 0], rtol=1e-6)
        assert np.allclose(M[k, 3, 3], ref[3, 3], rtol=1e-6, atol=1e-9)
        assert np.allclose(np.hypot(M[k, 0, 1], M[k, 0, 2]), abs(ref[0, 1]), rtol=1e-6, atol=1e-12)
        assert np.allclose(np.hypot(M[k, 1, 0], M[k, 2, 0]), abs(ref[1, 0]), rtol=1e-6, atol=1e-12)
        assert np.allclose(pdf[k], 3 / (16 * np.pi) * (1 + ct * ct), rtol=1e-9)
        # rows/cols 0 and 3 do not mix with the others
        assert np.allclose(M[k, 0, 3], 0) and np.allclose(M[k, 3, 0], 0) and np.allclose(M[k, 1:3, 3], 0)


def test_tabphase_polarized_equals_rayleigh_polarized(oracle):
    """tabphase_polarized fed with the Rayleigh matrix elements on a fine grid reproduces
    rayleigh_polarized (same rotation code path, tabphase_polarized.cpp:318-368)."""
    mu = np.linspace(-1, 1, 2001)
    m11, m12 = 1 + mu**2, mu**2 - 1
    fmt = lambda a: ",".join(f"{x:.9g}" for x in a)  # noqa: E731
    tab = {"type": "tabphase_polarized", "nodes": fmt(mu), "m11": fmt(m11), "m12": fmt(m12), "m22": fmt(m11),
           "m33": fmt(2 * mu), "m34": fmt(0 * mu), "m44": fmt(2 * mu)}
    _, d_tab = pol_desc(tab)
    _, d_ray = pol_desc({"type": "rayleigh_polarized"})
    rng = np.random.default_rng(1)
    n = 300
    wi = np.array([sph(t, p) for t, p in zip(rng.uniform(0.1, 3.0, n), rng.uniform(0, 6.28, n))])
    wo = np.array([sph(t, p) for t, p in zip(rng.uniform(0.1, 3.0, n), rng.uniform(0, 6.28, n))])
    Mt, pt = oracle.phase_mueller(d_tab, 0, wi, wo)
    Mr, pr = oracle.phase_mueller(d_ray, 0, wi, wo)
    assert np.allclose(Mt, Mr, rtol=2e-5, atol=2e-7)
    assert np.allclose(pt, pr, rtol=2e-5)


def test_tabphase_polarized_sample_convention(oracle):
    # ERP/tests/phase/test_tabphase_polarized.py:152-183: u = 1 -> forward scattering, pdf = 0.5/pi
    _, desc = pol_desc({"type": "tabphase_polarized", "nodes": "-1, 0, 1", "m11": "0.0, 0.5, 1.0"})
    ct, w, pdf = oracle.phase_sample(desc, 0, [[1.0, 0.0]])
    assert np.allclose(ct, 1.0) and np.allclose(pdf, 0.5 / np.pi)


def test_single_scattering_degree_of_polarisation(oracle):
    """Thin Rayleigh layer over a black ground, max_depth 2: DoLP = sin^2(T)/(1+cos^2(T)), the
    polarisation is perpendicular to the scattering plane (Q < 0, U = 0 in the meridian frame for
    principal-plane views), V = 0."""
    d = scenes.atmosphere_scene(geometry="plane_parallel", atmosphere="homogeneous",
                                homogeneous_sigma_t=0.05 / scenes.TOA, phase={"type": "rayleigh_polarized"},
                                surface={"type": "diffuse", "reflectance": 0.0}, sza=30.0, saa=0.0, max_depth=2,
                                sensor={"type": "mdistant", "vza": [-60.0, -30.0, 0.0, 30.0, 60.0], "vaa": 0.0},
                                stokes=True, meridian_align=True)
    sc = mi_load_dict(d)
    spp = 100000
    wl, l, l2, st, _ = oracle.render_stokes(sc.flat.build_desc(), 0, 1, spp)
    I, Q, U, V = st / spp
    vza = np.deg2rad([-60, -30, 0, 30, 60]); sza = np.deg2rad(30)
    view = np.stack([np.sin(vza), 0 * vza, np.cos(vza)], axis=-1)
    cosT = view @ (-np.array([np.sin(sza), 0, np.cos(sza)]))
    dolp = (1 - cosT**2) / (1 + cosT**2)
    assert np.allclose(np.hypot(Q, U) / I, dolp, atol=1e-9)
    assert np.all(Q <= 1e-12) and np.allclose(U, 0, atol=1e-12) and np.allclose(V, 0)
    assert np.allclose(I, l / spp)  # root channel == S0
    # out of the principal plane with the sensor-aligned basis U != 0 but DoLP is frame independent
    d2 = scenes.atmosphere_scene(geometry="plane_parallel", atmosphere="homogeneous",
                                 homogeneous_sigma_t=0.05 / scenes.TOA, phase={"type": "rayleigh_polarized"},
                                 surface={"type": "diffuse", "reflectance": 0.0}, sza=30.0, saa=0.0, max_depth=2,
                                 sensor={"type": "mdistant", "vza": [40.0], "vaa": 70.0},
                                 stokes=True, meridian_align=False)
    sc2 = mi_load_dict(d2)
    _, _, _, st2, _ = oracle.render_stokes(sc2.flat.build_desc(), 0, 1, 20000)
    I2, Q2, U2, V2 = (st2 / 20000)[:, 0]
    v = scenes.angles_to_direction(40.0, 70.0)
    cT = v @ (-np.array([np.sin(sza), 0, np.cos(sza)]))
    assert np.isclose(np.hypot(Q2, U2) / I2, (1 - cT**2) / (1 + cT**2), atol=1e-9) and abs(U2) > 1e-6


def test_polarization_changes_multiple_scattering_intensity(oracle):
    """Scalar vs vector transfer differ in I for a Rayleigh atmosphere (a few percent): the
    polarized code path is really exercised, not a relabelled scalar run."""
    common = dict(geometry="plane_parallel", atmosphere="homogeneous", homogeneous_sigma_t=1.0 / scenes.TOA,
                  surface={"type": "diffuse", "reflectance": 0.0}, sza=60.0,
                  sensor={"type": "mdistant", "vza": [0.0], "vaa": 0.0})
    spp = 400000
    a = mi_load_dict(scenes.atmosphere_scene(phase={"type": "rayleigh"}, **common))
    b = mi_load_dict(scenes.atmosphere_scene(phase={"type": "rayleigh_polarized"}, stokes=True, **common))
    _, la, la2, _ = oracle.render(a.flat.build_desc(), 0, 1, spp)
    _, lb, lb2, stb, _ = oracle.render_stokes(b.flat.build_desc(), 0, 2, spp)
    ma, mb = la[0] / spp, lb[0] / spp
    sig = np.sqrt((la2[0] / spp - ma**2) / spp + (lb2[0] / spp - mb**2) / spp)
    assert abs(ma - mb) > 5 * sig and abs(ma - mb) / ma < 0.15


def test_polarized_flags_and_errors():
    d = scenes.atmosphere_scene(geometry="plane_parallel", n_layers=10, phase={"type": "rayleigh_polarized"},
                                integrator="volpathmis", stokes=True)
    with pytest.raises(RuntimeError, match="does not support polarized mode"):
        mi_load_dict(d)
    sc = mi_load_dict(scenes.atmosphere_scene(geometry="plane_parallel", n_layers=10, stokes=True))
    assert sc.integrator().stokes and sc.integrator().moment and sc.flat.polarized
