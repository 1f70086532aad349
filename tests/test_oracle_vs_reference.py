"""
The C oracle port pinned on RENDERS OF THE REFERENCE ITSELF.

``tests/golden/reference_renders.json`` holds films rendered by the reference kernel (Eradiate's Mitsuba
fork + plugins, compiled from /root/reference by ``oracle/build_ref.sh``; variant scalar_mono_double, polarized
scenes scalar_mono_polarized_double) from the very dictionaries of ``tests/scene_battery.py``
(``tools/make_reference_golden.py``).  ``tests/golden/oracle_renders.json`` holds the oracle's films of the
same scenes.  Two independent Monte Carlo estimates of the same expectation: every pixel must pass the
reference's own paired z-test with Sidak correction (``test_tools/regression.py:852-893``) and stay
within 4.5 combined sigma.

When ``oracle/_ref`` is present (this container; it also travels to the GPU box) two live checks run as
well: a fresh low-spp render of C2 by both, and the key set ``mitsuba.traverse`` publishes against
``mi_traverse``'s.
"""

import json
import os

import numpy as np
import pytest

from tests.scene_battery import battery
from tests.util import REFERENCE_GROUND_LEAK, leak_bounded, sidak_ok, z_scores

HERE = os.path.dirname(__file__)
REF = json.load(open(os.path.join(HERE, "golden", "reference_renders.json")))
ORA = json.load(open(os.path.join(HERE, "golden", "oracle_renders.json")))
NAMES = [n for n in REF["scenes"] if n in ORA["scenes"]]

# scenes the reference and this repo deliberately treat differently, with the reason
KNOWN_DIFFERENT: dict = {}


def test_reference_fixture_covers_the_baseline_configs():
    for name in ("c1_homogeneous_lambertian_pp", "c2_afgl_rpv_spherical", "c3_afgl_aerosol_tab_hdistant",
                 "c5_polarized_ocean_aerosol_reduced"):
        assert name in REF["scenes"], f"{name}: no reference render committed"
    assert len(NAMES) >= 20


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference_render(name):
    if name in KNOWN_DIFFERENT:
        pytest.skip(KNOWN_DIFFERENT[name])
    r, o = REF["scenes"][name], ORA["scenes"][name]
    rm, om = np.array(r["mean"]), np.array(o["mean"])
    assert rm.shape == om.shape
    rv = np.array(r.get("var_of_mean", np.zeros_like(rm)))
    ov = np.array(o["var_of_mean"])
    if name in REFERENCE_GROUND_LEAK:  # the reference loses rays through the ground here (tests/util.py)
        ok, msg = leak_bounded(om, ov, rm, rv, REFERENCE_GROUND_LEAK[name])
        assert ok, f"{name}: {msg}"
        return
    z = z_scores(om, ov, rm, rv, rel_floor=1e-7)
    ok, zc = sidak_ok(z)
    assert ok and np.all(np.abs(z) <= 4.5), (
        f"{name}: |z| max {np.abs(z).max():.2f} (Sidak bound {zc:.2f})\n oracle    {om}\n reference {rm}")
    # ray-weighted channel (distantflux) -- same statistic on the film total
    rw, ow = np.array(r["mean_wl"]), np.array(o["mean_wl"])
    if not np.allclose(rw, rm):
        s = np.sqrt(rv.sum() + ov.sum())
        assert abs(rw.sum() - ow.sum()) <= 5.0 * s + 1e-3 * abs(rw.sum())


@pytest.mark.parametrize("name", [n for n in NAMES if "stokes" in REF["scenes"][n] and "stokes" in ORA["scenes"][n]])
def test_oracle_matches_reference_stokes(name):
    """Q, U, V of the polarized scenes: the m2 channel only gives the variance of I; the noise of the other
    components is bounded by it (|S_k| <= I sample by sample), so the same sigma is a conservative scale."""
    r, o = REF["scenes"][name], ORA["scenes"][name]
    rs, os_ = np.array(r["stokes"]), np.array(o["stokes"])
    sig = np.sqrt(np.array(r["var_of_mean"]) + np.array(o["var_of_mean"]))
    for k in range(4):
        dz = np.abs(rs[k] - os_[k]) / np.maximum(sig, 1e-12)
        assert np.all(dz <= 5.0), f"{name}: S{k} differs by {dz.max():.2f} sigma\n oracle {os_[k]}\n reference {rs[k]}"


def _ref_or_skip():
    from oracle import ref

    if not ref.available():
        pytest.skip("oracle/_ref (the compiled reference) is not present")
    return ref


def test_live_reference_render_matches_oracle(oracle):
    """Fresh renders by the reference and by the oracle of a reduced C2 (seconds of CPU)."""
    ref = _ref_or_skip()
    from eradiate_b200 import scenes
    from eradiate_b200.kernel import mi_load_dict

    d = scenes.config_c2(spp=1 << 14, n_vza=4)
    mi = ref.mitsuba("scalar_mono_double")
    scene = mi.load_dict(ref.to_mitsuba(mi, d))
    spp = 1 << 15
    mi.render(scene, sensor=0, seed=3, spp=spp)
    ch = ref.film_channels(mi, scene.sensors()[0])
    rm = ch["nested.Y"].ravel()
    rv = np.maximum(ch["m2_nested.Y"].ravel() - rm**2, 0.0) / spp
    _, l, l2, _ = oracle.render(mi_load_dict(d).flat.build_desc(), 0, 11, spp)
    om = l / spp
    ov = np.maximum(l2 / spp - om**2, 0.0) / spp
    z = z_scores(om, ov, rm, rv)
    assert np.all(np.abs(z) <= 4.5), f"z = {z}"


def test_reference_loses_rays_through_the_ground():
    """Evidence for tests/util.py::REFERENCE_GROUND_LEAK: in the reference, camera rays of `hdistant` over the
    default-width plane-parallel scene cross the atmosphere cube's top, then MISS the ground rectangle and hit
    the cube's bottom 1.2 km below it (null BSDF -> the path escapes with L = 0).  No such ray exists for the
    principal-plane `mdistant` directions.  The oracle and the CUDA kernels intersect an analytic slab."""
    ref = _ref_or_skip()
    from eradiate_b200 import scenes

    mi = ref.mitsuba("scalar_mono_double")
    leaks = {}
    for sensor in ({"type": "hdistant", "film_resolution": (3, 3)},
                   {"type": "mdistant", "vza": [-60.0, -30.0, 0.0, 30.0, 60.0], "vaa": 0.0}):
        d = scenes.atmosphere_scene(geometry="plane_parallel", n_layers=10, sensor=sensor, spp=4)
        scene = mi.load_dict(ref.to_mitsuba(mi, d))
        s = scene.sensors()[0]
        rng = np.random.default_rng(1)
        n, leak = 6000, 0
        for _ in range(n):
            ray, _ = s.sample_ray(0.0, 0.0, mi.Point2f(*rng.uniform(0, 1, 2)), mi.Point2f(*rng.uniform(0, 1, 2)))
            si = scene.ray_intersect(ray)
            assert si.is_valid()
            si2 = scene.ray_intersect(si.spawn_ray(ray.d))
            leak += (not si2.is_valid()) or abs(float(si2.p.z)) > 1e-3
        leaks[sensor["type"]] = leak / n
    assert leaks["mdistant"] == 0.0
    assert 0.002 < leaks["hdistant"] < 0.012, leaks


# Keys the reference publishes for objects this kernel keeps fixed after loading: geometry of the analytic
# stencils, film layout, sampler / shutter settings, emitter sampling weights.  They are listed, not hidden:
# the test asserts that the difference of the two key sets is exactly this family.
_PASSIVE = ("to_world", "silhouette_sampling_weight", "sampling_weight", "film.size", "film.crop_size",
            "film.crop_offset", "shutter_open", "shutter_open_time", "allow_thread_reordering",
            # the 12-triangle mesh of the `cube` stencil (an analytic slab here)
            "faces", "vertex_normals", "vertex_positions", "vertex_texcoords",
            # the 3x3 mask bitmap of a CentralPatchSurface and its placement
            "weight.data", "weight.to_uv")


def _is_passive(key: str) -> bool:
    return any(key == p or key.endswith("." + p) for p in _PASSIVE)


@pytest.mark.parametrize("name", list(REF["scenes"].keys()))
def test_traverse_key_set_matches_reference(name):
    """``mi_traverse`` publishes the keys ``mitsuba.traverse`` publishes for the same dict (fixture
    recorded from the reference), so an update map written against the reference resolves here."""
    from eradiate_b200.kernel import mi_load_dict, mi_traverse

    if name not in REF["scenes"]:
        pytest.skip("no reference fixture for this scene")
    want = set(REF["scenes"][name]["traverse_keys"])
    got = set(mi_traverse(mi_load_dict(battery()[name])).parameters.keys())
    missing = {k for k in want - got if not _is_passive(k)}
    extra = {k for k in got - want}
    assert not missing, f"{name}: keys of the reference that mi_traverse does not publish: {sorted(missing)}"
    assert not extra, f"{name}: keys mi_traverse publishes that the reference does not: {sorted(extra)}"
