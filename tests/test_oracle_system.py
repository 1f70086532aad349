"""
End-to-end known-answer tests of the CPU oracle's render loop, following the
reference's analytic system tests (SURVEY.md 8c "System-level analytic answers").
CPU only; sizes chosen so the whole file runs in well under a minute.
"""

import numpy as np
import pytest

from eradiate_b200 import scenes
from eradiate_b200.kernel import mi_load_dict
from tests.util import stats_from_sums, z_scores, sidak_ok

E0 = 1.8


def run(oracle, d, spp, seed=3):
    sc = mi_load_dict(d)
    desc = sc.flat.build_desc()
    wl, l, l2, st = oracle.render(desc, 0, seed, spp)
    mean, var = stats_from_sums(l, l2, spp)
    return wl / spp, mean, var, st


@pytest.mark.parametrize("geometry", ["plane_parallel", "spherical_shell"])
@pytest.mark.parametrize("rho", [0.0, 0.5, 1.0])
def test_lambertian_brf_no_atmosphere(oracle, geometry, rho):
    # tests/02_system/test_onedim_lambertian_brf.py:112-117 (BRF == rho, spp = 1) and
    # tests/02_system/test_basic.py:96-122 (L = rho E cos(sza) / pi, rtol 1e-3)
    sza = 30.0
    d = scenes.atmosphere_scene(geometry=geometry, atmosphere=None, sza=sza, irradiance=E0,
                                surface={"type": "diffuse", "reflectance": rho},
                                sensor={"type": "mdistant", "vza": [-60.0, -20.0, 0.0, 45.0], "vaa": 0.0})
    wl, mean, var, _ = run(oracle, d, 1)
    brf = np.pi * wl / (E0 * np.cos(np.deg2rad(sza)))
    assert np.allclose(brf, rho, rtol=1e-6, atol=1e-15)  # irradiance is stored as float32


@pytest.mark.parametrize("geometry", ["plane_parallel", "spherical_shell"])
def test_absorbing_atmosphere_beer_lambert(oracle, geometry):
    """Purely absorbing atmosphere above a Lambertian ground: L = rho E mu0/pi * exp(-tau (1/mu0 + 1/muv))
    (vertical sun and view so that plane-parallel and spherical coincide)."""
    n = 50
    d = scenes.atmosphere_scene(geometry=geometry, atmosphere="afgl", n_layers=n, sza=0.0,
                                surface={"type": "diffuse", "reflectance": 0.7},
                                sensor={"type": "mdistant", "vza": [0.0], "vaa": 0.0})
    sc = mi_load_dict(d)
    flat = sc.flat
    sig = flat.medium.children["sigma_t"]
    alb = flat.medium.children["albedo"]
    prof = (np.linspace(2.0, 0.2, n) * 1e-5).astype(np.float32)
    (sig.children["volume"] if sig.type == "sphericalcoordsvolume" else sig).values["data"][:] = prof.reshape(
        (1, 1, -1, 1) if sig.type == "sphericalcoordsvolume" else (-1, 1, 1, 1))
    (alb.children["volume"] if alb.type == "sphericalcoordsvolume" else alb).values["data"][:] = 0.0
    desc = flat.build_desc()
    spp = 200000
    wl, l, l2, st = oracle.render(desc, 0, 5, spp)
    mean, var = stats_from_sums(l, l2, spp)
    tau = float(np.sum(prof.astype(np.float64)) * scenes.TOA / n)
    expected = 0.7 * E0 / np.pi * np.exp(-2.0 * tau)
    z = (mean[0] - expected) / np.sqrt(var[0])
    assert abs(z) < 4.0, (mean, expected, z)
    assert st["n_scatter"] > 0  # absorption events are "real collisions" with albedo 0


def test_single_scattering_homogeneous_slab(oracle):
    """Black ground, homogeneous isotropic slab, max_depth = 2 (single scattering):
    L = (w0 E / 4 pi) * mu0/(mu0+muv) * (1 - exp(-tau (1/mu0 + 1/muv)))  [Chandrasekhar]."""
    tau, w0, sza, vza = 0.8, 0.9, 40.0, 25.0
    H = scenes.TOA
    d = scenes.atmosphere_scene(geometry="plane_parallel", atmosphere="homogeneous",
                                homogeneous_sigma_t=tau / H, homogeneous_albedo=w0,
                                phase={"type": "isotropic"}, sza=sza, max_depth=2,
                                surface={"type": "diffuse", "reflectance": 0.0},
                                sensor={"type": "mdistant", "vza": [vza], "vaa": 70.0})
    spp = 400000
    wl, mean, var, st = run(oracle, d, spp)
    mu0, muv = np.cos(np.deg2rad(sza)), np.cos(np.deg2rad(vza))
    expected = w0 * E0 / (4 * np.pi) * mu0 / (mu0 + muv) * (1 - np.exp(-tau * (1 / mu0 + 1 / muv)))
    z = (mean[0] - expected) / np.sqrt(var[0])
    assert abs(z) < 4.0, (mean, expected, z)


def test_single_scattering_heterogeneous_matches_homogeneous(oracle):
    """Same optical depth spread non-uniformly (null collisions active) must give the same
    single-scattering radiance: exercises delta + ratio tracking with a global majorant."""
    tau, w0, sza, vza = 0.8, 0.9, 40.0, 25.0
    n = 20
    d = scenes.atmosphere_scene(geometry="plane_parallel", atmosphere="afgl", n_layers=n,
                                phase={"type": "isotropic"}, sza=sza, max_depth=2,
                                surface={"type": "diffuse", "reflectance": 0.0},
                                sensor={"type": "mdistant", "vza": [vza], "vaa": 70.0})
    sc = mi_load_dict(d)
    prof = np.exp(-np.arange(n) / 4.0)
    prof *= tau / (prof.sum() * scenes.TOA / n)
    sc.flat.medium.children["sigma_t"].values["data"][:] = prof.reshape(-1, 1, 1, 1).astype(np.float32)
    sc.flat.medium.children["albedo"].values["data"][:] = w0
    spp = 400000
    wl, l, l2, st = oracle.render(sc.flat.build_desc(), 0, 9, spp)
    mean, var = stats_from_sums(l, l2, spp)
    mu0, muv = np.cos(np.deg2rad(sza)), np.cos(np.deg2rad(vza))
    expected = w0 * E0 / (4 * np.pi) * mu0 / (mu0 + muv) * (1 - np.exp(-tau * (1 / mu0 + 1 / muv)))
    z = (mean[0] - expected) / np.sqrt(var[0])
    assert abs(z) < 4.0, (mean, expected, z)
    assert st["trips_main"] > st["n_scatter"] * 2  # null collisions dominate


def test_rpv_degenerate_equals_lambertian_through_atmosphere(oracle):
    # tests/02_system/test_atmosphere_rpv.py:622-764 (rtol 1e-2 at 1e5 spp)
    common = dict(geometry="plane_parallel", atmosphere="afgl", n_layers=60, sza=30.0,
                  sensor={"type": "mdistant", "vza": np.linspace(-60, 60, 7), "vaa": 0.0})
    spp = 100000
    _, m1, v1, _ = run(oracle, scenes.atmosphere_scene(
        surface={"type": "rpv", "rho_0": 0.5, "k": 1.0, "g": 0.0, "rho_c": 1.0}, **common), spp, seed=1)
    _, m2, v2, _ = run(oracle, scenes.atmosphere_scene(
        surface={"type": "diffuse", "reflectance": 0.5}, **common), spp, seed=2)
    ok, zc = sidak_ok(z_scores(m1, v1, m2, v2))
    assert ok
    assert np.allclose(m1, m2, rtol=1e-2)


def test_principal_plane_symmetry_at_nadir_sun(oracle):
    # tests/02_system/test_onedim_symmetry.py:135-138
    d = scenes.atmosphere_scene(geometry="spherical_shell", atmosphere="afgl", n_layers=120, sza=0.0,
                                sensor={"type": "mdistant", "vza": [-60, -30, 30, 60], "vaa": 0.0})
    spp = 100000
    _, m, v, _ = run(oracle, d, spp)
    z = z_scores(m[:2], v[:2], m[:1:-1], v[:1:-1])
    assert np.all(np.abs(z) < 4.0)


def test_irradiance_scaling_and_moment(oracle):
    # tests/02_system/test_irradiance_scaling.py:108-111: radiance is linear in E
    base = dict(geometry="plane_parallel", atmosphere="afgl", n_layers=30,
                sensor={"type": "mdistant", "vza": [10.0], "vaa": 0.0})
    wl1, m1, v1, _ = run(oracle, scenes.atmosphere_scene(irradiance=1.0, **base), 20000, seed=4)
    wl2, m2, v2, _ = run(oracle, scenes.atmosphere_scene(irradiance=3.0, **base), 20000, seed=4)
    assert np.allclose(3.0 * m1, m2, rtol=1e-6)      # same seed -> same paths
    assert np.allclose(9.0 * v1, v2, rtol=1e-5)
    assert np.allclose(wl1, m1)                       # mdistant ray weight == 1


def test_distantflux_lambertian_albedo(oracle):
    """distantflux over a Lambertian ground without atmosphere: the film sums to the
    radiosity  rho * E * mu0  (distantflux.cpp:168-170 weights)."""
    rho, sza = 0.6, 35.0
    d = scenes.atmosphere_scene(geometry="plane_parallel", atmosphere=None, sza=sza,
                                surface={"type": "diffuse", "reflectance": rho},
                                sensor={"type": "distantflux", "film_resolution": (4, 4)})
    spp = 20000
    wl, mean, var, _ = run(oracle, d, spp)
    flux = wl.sum()
    expected = rho * E0 * np.cos(np.deg2rad(sza))
    assert np.allclose(flux, expected, rtol=5e-3)


def test_volpathmis_same_expectation(oracle):
    # MI/src/render/tests/test_renders.py:45-49: volpath scenes are also rendered with volpathmis
    base = dict(geometry="plane_parallel", atmosphere="homogeneous", homogeneous_sigma_t=3.0 / scenes.TOA,
                homogeneous_albedo=0.95, surface={"type": "diffuse", "reflectance": 0.3},
                sensor={"type": "mdistant", "vza": [0.0, 50.0], "vaa": 0.0})
    spp = 60000
    _, m1, v1, s1 = run(oracle, scenes.atmosphere_scene(integrator="volpath", **base), spp, seed=1)
    _, m2, v2, s2 = run(oracle, scenes.atmosphere_scene(integrator="volpathmis", **base), spp, seed=2)
    assert np.all(np.abs(z_scores(m1, v1, m2, v2)) < 4.0)


def test_oracle_reproduces_committed_fixtures(oracle):
    """tests/golden/oracle_renders.json was written by tools/make_golden.py with this very
    oracle; a re-render with the same seed must give the same path set (OpenMP only
    reorders the float64 sums)."""
    import json, os
    from tests.scene_battery import battery
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "oracle_renders.json")))
    b = battery()
    for name in ("no_atmosphere_rpv_spherical", "max_depth_3_rr_2", "homogeneous_spherical_hg"):
        g = gold["scenes"][name]
        sc = mi_load_dict(b[name])
        wl, l, l2, st = oracle.render(sc.flat.build_desc(), 0, gold["seed"], g["spp"])
        assert np.allclose(l / g["spp"], g["mean"], rtol=1e-10)
        assert np.isclose(st["trips_main"] / st["n_paths"], g["trips_main_per_path"], rtol=1e-12)


# ------------------------------------------------------------------------ astroobject (finite solar disc)
def test_astroobject_lambertian_no_atmosphere(oracle):
    """ERP/emitters/astroobject.cpp: radiance E / omega inside a cone of half-angle a about the sun direction.
    A Lambertian ground without atmosphere sees E <cos> = E mu0 (1 + cos a) / 2 (uniform cone about an axis at mu0):
    L = rho E mu0 (1 + cos a) / (2 pi).  Wide disc so that the factor shows (20 deg: 0.9924)."""
    sza, diam, rho = 40.0, 20.0, 0.6
    d = scenes.atmosphere_scene(geometry="plane_parallel", atmosphere=None, sza=sza, saa=25.0, irradiance=E0,
                                angular_diameter=diam, surface={"type": "diffuse", "reflectance": rho},
                                sensor={"type": "mdistant", "vza": [-50.0, 0.0, 30.0], "vaa": 0.0})
    spp = 400000
    _, mean, var, _ = run(oracle, d, spp)
    expected = rho * E0 * np.cos(np.deg2rad(sza)) * (1.0 + np.cos(np.deg2rad(diam / 2))) / (2.0 * np.pi)
    z = (mean - expected) / np.sqrt(var)
    assert np.all(np.abs(z) < 4.0) and np.allclose(mean, expected, rtol=2e-3), (mean, expected, z)
    # ... and it is NOT the directional answer (0.76 % lower), which the test can resolve
    assert np.all(mean < rho * E0 * np.cos(np.deg2rad(sza)) / np.pi * 0.997)


@pytest.mark.parametrize("integrator", ["volpath", "piecewise_volpath"])
def test_astroobject_direct_beam_seen_from_the_ground(oracle, integrator):
    """A radiometer on the ground pointed at the disc through a purely absorbing atmosphere reads the radiance of
    the disc, E / omega, attenuated along the slant path (volpath.cpp:328-346, count_direct for camera rays);
    pointed just outside the disc it reads nothing."""
    n, sza, diam = 40, 35.0, 0.5358
    sun = scenes.angles_to_direction(sza, 0.0)
    off = scenes.angles_to_direction(sza + 0.6 * diam, 0.0)  # 0.6 diameters away from the centre: outside
    inside = scenes.angles_to_direction(sza + 0.3 * diam, 0.0)
    d = scenes.atmosphere_scene(geometry="plane_parallel", atmosphere="afgl", n_layers=n, sza=sza, saa=0.0,
                                irradiance=E0, angular_diameter=diam, integrator=integrator,
                                surface={"type": "diffuse", "reflectance": 0.0},
                                sensor={"type": "mradiancemeter", "medium": {"type": "ref", "id": "medium_atmosphere"},
                                        "origins": [[0.0, 0.0, 1.0]] * 3,
                                        "directions": [list(sun), list(inside), list(off)]})
    sc = mi_load_dict(d)
    flat = sc.flat
    prof = (np.linspace(3.0, 0.3, n) * 1e-6).astype(np.float32)
    flat.medium.children["sigma_t"].values["data"][:] = prof.reshape(-1, 1, 1, 1)
    flat.medium.children["albedo"].values["data"][:] = 0.0
    desc = flat.build_desc()
    spp = 20000
    wl, l, l2, st = oracle.render(desc, 0, 11, spp)
    mean, var = stats_from_sums(l, l2, spp)
    tau = float(np.sum(prof.astype(np.float64)) * scenes.TOA / n)
    omega = 2.0 * np.pi * (1.0 - np.cos(np.deg2rad(diam / 2)))
    for k, dirv in enumerate((sun, inside)):
        expected = E0 / omega * np.exp(-tau / dirv[2])
        tol = 5.0 * np.sqrt(var[k]) + 1e-6 * expected  # delta tracking through an absorber: binary estimator
        assert abs(mean[k] - expected) < tol, (k, mean[k], expected)
    assert mean[2] == 0.0
    # `hide_emitters` (integrator.cpp:29, volpath.cpp:329-330): the disc is not seen by camera rays; nothing
    # scatters in this atmosphere, so the radiometer reads zero everywhere
    d["integrator"]["nested"]["hide_emitters"] = True
    sc = mi_load_dict(d)
    sc.flat.medium.children["sigma_t"].values["data"][:] = prof.reshape(-1, 1, 1, 1)
    sc.flat.medium.children["albedo"].values["data"][:] = 0.0
    _, l, _, _ = oracle.render(sc.flat.build_desc(), 0, 11, 2000)
    assert np.all(l == 0.0)


def test_astroobject_small_disc_matches_directional(oracle):
    """The Sun's 0.54 deg disc changes a TOA radiance by far less than the Monte Carlo noise resolves here: the
    astroobject render (cone-sampled NEE with MIS + emitter hits) agrees with the directional one."""
    kw = dict(geometry="spherical_shell", n_layers=60, sza=45.0, saa=10.0,
              surface={"type": "rpv", "rho_0": 0.1, "k": 0.9, "g": -0.1},
              sensor={"type": "mdistant", "vza": [-60.0, 0.0, 40.0], "vaa": 0.0})
    spp = 200000
    _, m0, v0, _ = run(oracle, scenes.atmosphere_scene(**kw), spp, seed=4)
    _, m1, v1, _ = run(oracle, scenes.atmosphere_scene(angular_diameter=0.5358, **kw), spp, seed=9)
    z = (m1 - m0) / np.sqrt(v0 + v1)
    assert np.all(np.abs(z) < 4.0), z


def test_multiphase_mis_same_expectation(oracle):
    """multiphase.cpp:176-200: with use_mis the weight of a sampled direction is the mixture ratio
    sum_j w_j value_j / sum_j w_j pdf_j instead of the drawn component's own weight.  Both are unbiased: a thick
    medium whose components include a strongly depolarized Rayleigh lobe (value != pdf) renders to the same
    radiance either way, and the two estimators really differ (different noise for the same seed)."""
    def scene(use_mis):
        return scenes.atmosphere_scene(
            geometry="plane_parallel", atmosphere="homogeneous", homogeneous_sigma_t=2.0 / scenes.TOA,
            homogeneous_albedo=0.98, sza=40.0, surface={"type": "diffuse", "reflectance": 0.1},
            sensor={"type": "mdistant", "vza": [-60.0, 0.0, 45.0], "vaa": 0.0},
            phase={"type": "multiphase", "use_mis": use_mis,
                   "phase0": {"type": "rayleigh", "depolarization": 0.4}, "weight0": 2.0,
                   "phase1": {"type": "hg", "g": 0.7}, "weight1": 1.0})
    spp = 150000
    _, m0, v0, _ = run(oracle, scene(False), spp, seed=6)
    _, m1, v1, _ = run(oracle, scene(True), spp, seed=6)
    z = (m1 - m0) / np.sqrt(v0 + v1)
    assert np.all(np.abs(z) < 4.0), z
    assert not np.allclose(m0, m1, rtol=1e-9)
