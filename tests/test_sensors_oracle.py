"""
Pins the oracle's `mradiancemeter` / `mpdistant` sensors (ERP/sensors/mradiancemeter.cpp:147-172,
mpdistant.cpp:214-262) before it checks the CUDA kernels: ray conventions, and closed-form
radiative transfer seen from INSIDE the atmosphere (Beer-Lambert down-looking view at altitude,
single-scattered sky radiance seen from the ground).
"""

import numpy as np
import pytest

from eradiate_b200 import scenes
from eradiate_b200.kernel import mi_load_dict

E0 = 1.8
MED = {"type": "ref", "id": "medium_atmosphere"}


def test_mradiancemeter_rays_are_the_given_rays(oracle):
    org = np.array([[0.0, 0.0, 2.0], [10.0, -3.0, 3000.0], [1.0, 5.0, 1.5e4]])
    dirs = np.array([[0.3, 0.2, 0.9327379], [0.0, 3.0, -4.0], [0.0, 0.0, -1.0]])
    sc = mi_load_dict(scenes.atmosphere_scene(geometry="plane_parallel", n_layers=10,
                                              sensor={"type": "mradiancemeter", "medium": MED, "origins": org, "directions": dirs}))
    d = sc.flat.build_desc()
    fs = np.array([[0.0, 0.3], [0.34, 0.9], [0.66, 0.1], [0.67, 0.5], [0.999, 0.5]])
    o, dd, w = oracle.sensor_ray(d, 0, fs, np.zeros((5, 2)))
    idx = [0, 1, 1, 2, 2]  # Int32(position_sample.x * n), mradiancemeter.cpp:161
    assert np.allclose(o, org[idx], atol=0) and np.allclose(w, 1.0)
    assert np.allclose(dd, dirs[idx] / np.linalg.norm(dirs[idx], axis=1, keepdims=True), atol=1e-15)


def test_mpdistant_rays_image_the_target(oracle):
    lx, ly, z = 6.0, 4.0, 1.5
    tgt = {"type": "rectangle", "to_world": scenes.ScalarTransform4f().translate([1.0, 2.0, z]).scale([0.5 * lx, 0.5 * ly, 1.0])}
    sc = mi_load_dict(scenes.atmosphere_scene(geometry="plane_parallel", atmosphere=None, integrator="path",
                                              sensor={"type": "mpdistant", "vza": 30.0, "vaa": 90.0, "film_resolution": (4, 2),
                                                      "target": tgt, "ray_offset": 100.0}))
    d = sc.flat.build_desc()
    fs = np.array([[0.0, 0.0], [1.0, 1.0], [0.5, 0.5], [0.25, 0.75]])
    o, dd, w = oracle.sensor_ray(d, 0, fs, np.random.default_rng(0).random((4, 2)))  # the aperture sample is unused
    view = np.array([0.0, np.sin(np.radians(30.0)), np.cos(np.radians(30.0))])  # towards the sensor
    assert np.allclose(dd, -view, atol=1e-12) and np.allclose(w, 1.0)
    p = o + 100.0 * dd  # the point of the target each ray goes through: the film sample, mapped on the rectangle
    want = np.stack([1.0 + (2 * fs[:, 0] - 1) * 0.5 * lx, 2.0 + (2 * fs[:, 1] - 1) * 0.5 * ly, np.full(4, z)], axis=1)
    assert np.allclose(p, want, atol=1e-9)


def _render(oracle, d, spp, seed=3):
    wl, l, l2, st = oracle.render(d, 0, seed, spp)
    mean = l / spp
    return mean, np.sqrt(np.maximum(l2 / spp - mean**2, 0) / spp)


def test_beer_lambert_seen_from_inside_the_atmosphere(oracle):
    """Absorbing homogeneous slab (albedo 0) over a Lambertian ground: a radiancemeter at altitude h looking
    down along mu_v sees rho E mu_s / pi x exp(-tau / mu_s) x exp(-sigma h / mu_v); looking up it sees nothing."""
    sigma, rho, sza = 1.5e-5, 0.6, 35.0
    mu_s = np.cos(np.radians(sza))
    hs, mus = np.array([5.0e3, 4.0e4, 1.1e5]), np.array([1.0, 0.8, 0.5])
    dirs = np.stack([np.sqrt(1 - mus**2), np.zeros(3), -mus], axis=1)
    org = np.stack([np.zeros(3), np.zeros(3), hs], axis=1)
    org = np.vstack([org, [[0.0, 0.0, 10.0]]])
    dirs = np.vstack([dirs, [[0.0, 0.6, 0.8]]])
    sc = mi_load_dict(scenes.atmosphere_scene(
        geometry="plane_parallel", atmosphere="homogeneous", homogeneous_sigma_t=sigma, homogeneous_albedo=0.0, sza=sza,
        surface={"type": "diffuse", "reflectance": rho},
        sensor={"type": "mradiancemeter", "medium": MED, "origins": org, "directions": dirs}))
    mean, err = _render(oracle, sc.flat.build_desc(), 300000)  # (absorption is sampled: collisions kill the path)
    want = rho * E0 * mu_s / np.pi * np.exp(-sigma * scenes.TOA / mu_s) * np.exp(-sigma * hs / mus)
    assert np.all(np.abs(mean[:3] - want) < 4.0 * err[:3] + 1e-6 * want), (mean, want, err)
    assert mean[3] == 0.0


@pytest.mark.parametrize("geometry", ["plane_parallel"])
def test_single_scattered_sky_radiance_from_the_ground(oracle, geometry):
    """Isotropic homogeneous slab over a black ground, first order only (max_depth = 2): the sky radiance seen
    from the ground along mu_v is  w0 E / (4 pi mu_v) x (e^(-tau/mu_v) - e^(-tau/mu_s)) / (1/mu_s - 1/mu_v)."""
    tau, w0, sza = 0.6, 0.9, 50.0
    mu_s = np.cos(np.radians(sza))
    mus = np.array([1.0, 0.7, 0.35])
    dirs = np.stack([-np.sqrt(1 - mus**2), 0.3 * np.sqrt(1 - mus**2), mus], axis=1)
    mus = mus / np.linalg.norm(dirs, axis=1)  # (the sensor normalises its directions)
    org = np.tile([0.0, 0.0, 0.5], (3, 1))
    sc = mi_load_dict(scenes.atmosphere_scene(
        geometry=geometry, atmosphere="homogeneous", homogeneous_sigma_t=tau / scenes.TOA, homogeneous_albedo=w0,
        phase={"type": "isotropic"}, sza=sza, max_depth=2, surface={"type": "diffuse", "reflectance": 0.0},
        sensor={"type": "mradiancemeter", "medium": MED, "origins": org, "directions": dirs}))
    mean, err = _render(oracle, sc.flat.build_desc(), 200000)
    want = w0 * E0 / (4 * np.pi * mus) * (np.exp(-tau / mus) - np.exp(-tau / mu_s)) / (1 / mu_s - 1 / mus)
    assert np.all(np.abs(mean - want) < 4.0 * err + 1e-4 * want), (mean, want, err)


def test_sensor_medium_must_match_the_origins():
    with pytest.raises(RuntimeError, match="distant sensors inside a medium"):
        mi_load_dict(scenes.atmosphere_scene(geometry="plane_parallel", n_layers=10,
                                             sensor={"type": "mpdistant", "vza": 0.0, "film_resolution": (2, 2), "medium": MED}))
    with pytest.raises(RuntimeError, match="not equal"):
        mi_load_dict(scenes.atmosphere_scene(geometry="plane_parallel", n_layers=10, sensor={
            "type": "mradiancemeter", "origins": [[0, 0, 1.0], [0, 0, 2.0]], "directions": [[0, 0, 1.0]]}))
