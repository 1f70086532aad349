"""
Pins the oracle's canopy path (oracle/ertb_oracle_canopy.c + the leaf branches of ertb_oracle.c)
before it is trusted as the checker of the CUDA 3D kernel:

* bilambertian known answers of the reference's own test (ERP/tests/bsdfs/test_bilambertian.py:
  eval = r|t * |cos|/pi by hemisphere, pdf = eval / (r + t), sample/pdf consistency, r = t = 0);
* the uniform-grid ray caster against a brute-force numpy intersection of every disk
  (MI/src/shapes/disk.cpp:388-407 semantics);
* analytic radiative-transfer answers for explicit canopies: Poisson gap fraction exp(-G LAI / mu)
  for planophile and uniform leaf-angle distributions (hot spot and decorrelated geometry), a
  scene-wide horizontal leaf == Lambertian surface, and energy conservation for conservative
  leaves over a white ground (distantflux radiosity == irradiance);
* the perspective sensor ray convention (MI/src/sensors/perspective.cpp:200-236).
"""

import numpy as np
import pytest

from eradiate_b200 import scenes
from eradiate_b200.kernel import mi_load_dict

E0 = 1.8


def _desc(**kw):
    sc = mi_load_dict(scenes.atmosphere_scene(geometry="plane_parallel", **kw))
    return sc, sc.flat.build_desc()


# ------------------------------------------------------------------ bilambertian KATs
def _sph(theta, phi):
    return np.stack([np.cos(phi) * np.sin(theta), np.sin(phi) * np.sin(theta), np.cos(theta)], axis=-1)


@pytest.mark.parametrize("r, t", [(0.2, 0.4), (0.4, 0.2), (0.1, 0.9), (0.9, 0.1), (0.4, 0.6), (0.6, 0.4)])
@pytest.mark.parametrize("wi", [[0, 0, 1], [1, 1, 1], [0, 0, -1], [1, 1, -1]])
def test_bilambertian_eval_pdf_reference_values(oracle, r, t, wi):
    """test_bilambertian.py:21-84."""
    _, d = _desc(atmosphere=None, canopy={"n_leaves": 4, "reflectance": r, "transmittance": t})
    wi = np.asarray(wi, float) / np.linalg.norm(wi)
    th, ph = np.meshgrid(np.linspace(0, np.pi, 60), np.linspace(0, 2 * np.pi, 120), indexing="ij")
    wo = _sph(th.ravel(), ph.ravel())
    wis = np.tile(wi, (wo.shape[0], 1))
    is_reflect = np.sign(wo[:, 2]) == np.sign(wi[2])
    expected = np.where(is_reflect, r, t) * np.abs(wo[:, 2]) / np.pi
    ok = np.abs(wo[:, 2]) > 1e-9  # sign(0) is undefined on the equator
    assert np.allclose(oracle.leaf_bsdf(d, 0, "eval", wis, wo)[ok], expected[ok], rtol=1e-12)
    assert np.allclose(oracle.leaf_bsdf(d, 0, "pdf", wis, wo)[ok], expected[ok] / (r + t), rtol=1e-12)


@pytest.mark.parametrize("r, t", [(0.6, 0.2), (0.2, 0.6), (0.9, 0.1), (1.0, 0.0), (0.0, 1.0), (0.0, 0.0)])
def test_bilambertian_sample_is_consistent_with_pdf(oracle, r, t):
    """test_bilambertian.py:87-122 (chi^2): histogram of sampled directions vs integrated pdf."""
    _, d = _desc(atmosphere=None, canopy={"n_leaves": 4, "reflectance": r, "transmittance": t})
    rng = np.random.default_rng(5)
    n = 200000
    for wi in ([0.3, -0.2, 0.93], [0.1, 0.5, -0.86]):
        wi = np.asarray(wi) / np.linalg.norm(wi)
        wo, w = oracle.leaf_bsdf(d, 0, "sample", np.tile(wi, (n, 1)), u=rng.random((n, 3)))
        if r + t == 0.0:
            assert np.all(w == 0.0)
            continue
        assert np.allclose(w[w > 0], r + t, rtol=1e-6)  # value / pdf (r, t travel as float32)
        refl = np.sign(wo[:, 2]) == np.sign(wi[2])
        assert abs(refl.mean() - r / (r + t)) < 4.0 * np.sqrt(0.25 / n) + 1e-12
        # cosine-weighted: |cos| has density 2c on [0,1] on either side
        hist, _ = np.histogram(np.abs(wo[:, 2]), bins=10, range=(0, 1))
        edges = np.linspace(0, 1, 11)
        assert np.allclose(hist / n, edges[1:] ** 2 - edges[:-1] ** 2, atol=5e-3)
        assert np.allclose(oracle.leaf_bsdf(d, 0, "pdf", np.tile(wi, (n, 1)), wo)[:100] * w[:100],
                           oracle.leaf_bsdf(d, 0, "eval", np.tile(wi, (n, 1)), wo)[:100], rtol=1e-6)


# ------------------------------------------------------------------ ray caster
def _brute_force(disks, offsets, o, d, tmax):
    best = np.full(o.shape[0], np.inf)
    for off in offsets:
        c = disks[:, :3] + off
        n = disks[:, 3:6]
        dn = d @ n.T                                         # [rays, disks]
        with np.errstate(divide="ignore", invalid="ignore"):
            t = np.einsum("rdk,dk->rd", c[None, :, :] - o[:, None, :], n) / dn
        p = o[:, None, :] + t[..., None] * d[:, None, :] - c[None]
        ok = (t >= 0) & (t <= tmax[:, None]) & ((p ** 2).sum(-1) <= disks[:, 6] ** 2)
        best = np.minimum(best, np.where(ok, t, np.inf).min(axis=1))
    return best


def test_grid_ray_caster_equals_brute_force(oracle):
    sc, d = _desc(atmosphere=None, canopy={"lai": 2.0, "radius": 0.12, "size": (3.0, 3.0, 1.0), "padding": 1, "seed": 3})
    f = sc.flat
    disks = f.leaf_groups[0].disks.astype(np.float32).astype(np.float64)
    offsets = [off for _, off in f.instances]
    rng = np.random.default_rng(11)
    n = 3000
    o = np.stack([rng.uniform(-6, 6, n), rng.uniform(-6, 6, n), rng.uniform(-0.5, 2.0, n)], axis=1)
    dirs = rng.normal(size=(n, 3))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    tmax = np.where(rng.random(n) < 0.3, rng.uniform(0.1, 3.0, n), np.inf)
    t, nrm, grp = oracle.canopy_intersect(d, o, dirs, tmax)
    want = _brute_force(disks, offsets, o, dirs, tmax)
    assert np.array_equal(np.isfinite(t), np.isfinite(want))
    hit = np.isfinite(t)
    assert hit.sum() > 500 and (~hit).sum() > 100
    assert np.allclose(t[hit], want[hit], rtol=1e-12, atol=1e-12)
    assert np.all(grp[hit] == 0) and np.all(grp[~hit] == -1)
    assert np.allclose(np.linalg.norm(nrm[hit], axis=1), 1.0, atol=1e-6)


# ------------------------------------------------------------------ analytic canopy answers
def _render(oracle, d, spp, sensor=0, seed=7):
    wl, l, l2, st = oracle.render(d, sensor, seed, spp)
    mean = l / spp
    err = np.sqrt(np.maximum(l2 / spp - mean**2, 0) / spp)
    return mean, err, st, wl / spp


@pytest.mark.parametrize("orientation, G", [("planophile", None), ("uniform", 0.5)])
def test_gap_fraction_of_black_leaves_over_a_white_ground(oracle, orientation, G):
    """Black leaves, Lambertian ground rho = 1, no atmosphere, single bounce: L = E cos(sza)/pi x
    P(sun ray and view ray both reach the ground) with P = exp(-G LAI (1/mu_s + 1/mu_v)) when the two
    rays are decorrelated and exp(-G LAI / mu) in the hot spot (same ray). Planophile leaves: G = mu."""
    lai, sza = 1.0, 30.0
    mu_s = np.cos(np.radians(sza))
    vza = np.array([30.0, 0.0, -55.0])  # +30 in the solar azimuth = hot spot
    _, d = _desc(atmosphere=None, integrator="path", max_depth=2, sza=sza, saa=0.0,
                 surface={"type": "diffuse", "reflectance": 1.0},
                 canopy={"lai": lai, "radius": 0.03, "size": (4.0, 4.0, 1.0), "padding": 3, "orientation": orientation,
                         "reflectance": 0.0, "transmittance": 0.0, "seed": 2},
                 sensor={"type": "mdistant", "vza": vza, "vaa": 0.0})
    mean, err, st, _ = _render(oracle, d, 40000)
    mu_v = np.cos(np.radians(np.abs(vza)))
    if G is None:
        # horizontal leaves: at height h the two rays cross the leaf plane delta(h) apart and a leaf
        # blocks the pair when its centre lies in the UNION of two discs of radius r around the
        # crossings: P = exp(-(LAI / H) / (pi r^2) * int_0^H [2 pi r^2 - lens(delta(h))] dh), which
        # contains both limits (delta = 0: hot spot, delta > 2 r: independent rays)
        r, H = 0.03, 1.0
        h = (np.arange(20000) + 0.5) / 20000 * H
        tan_s = np.tan(np.radians(sza))
        tan_v = np.tan(np.radians(vza))  # signed: positive zeniths look from the solar azimuth
        p = []
        for tv in tan_v:
            delta = np.minimum(h * abs(tan_s - tv), 2 * r)
            lens = 2 * r * r * np.arccos(delta / (2 * r)) - 0.5 * delta * np.sqrt(4 * r * r - delta * delta)
            p.append(np.exp(-lai / (np.pi * r * r) * np.mean(2 * np.pi * r * r - lens)))
        p = np.array(p)
        tol = 0.03
    else:
        p = np.exp(-lai * G / mu_s) * np.exp(-lai * G / mu_v)
        p[0] = np.exp(-lai * G / mu_s)  # hot spot: one ray
        tol = 0.06  # leaf-size correlation near the ground is not in this formula
    want = E0 * mu_s / np.pi * p
    # one realisation of a finite Poisson canopy: a few percent
    assert np.allclose(mean, want, rtol=tol), (mean, want)
    assert st["n_scatter"] == 0


def test_scene_wide_horizontal_leaf_is_a_lambertian_surface(oracle):
    rho = 0.37
    big = {"leaf": {"type": "disk", "bsdf": {"type": "bilambertian", "reflectance": rho, "transmittance": 0.0},
                    "to_world": scenes.ScalarTransform4f().translate([0, 0, 0.5]).scale(1e4)}}
    kd = scenes.atmosphere_scene(geometry="plane_parallel", atmosphere=None, integrator="path", sza=40.0,
                                 surface={"type": "diffuse", "reflectance": 0.0},
                                 sensor={"type": "mdistant", "vza": [-50.0, 0.0, 20.0], "vaa": 30.0,
                                         "target": [0.0, 0.0, 0.5]})
    kd.update(big)
    d = mi_load_dict(kd).flat.build_desc()
    mean, err, _, _ = _render(oracle, d, 64)
    assert np.allclose(mean, rho * E0 * np.cos(np.radians(40.0)) / np.pi, rtol=1e-9)
    assert np.all(err < 1e-8)  # deterministic up to rounding in the m2 difference


def test_conservative_canopy_over_white_ground_returns_all_the_energy(oracle):
    """r + t = 1 leaves, rho = 1 ground, no atmosphere: nothing absorbs, so the radiosity leaving the
    (padded, quasi-infinite) canopy top equals the incoming irradiance E cos(sza) (distantflux)."""
    sza = 35.0
    _, d = _desc(atmosphere=None, integrator="path", rr_depth=50, sza=sza,
                 surface={"type": "diffuse", "reflectance": 1.0},
                 canopy={"lai": 1.5, "radius": 0.1, "size": (3.0, 3.0, 1.0), "padding": 4,
                         "reflectance": 0.55, "transmittance": 0.45, "seed": 4},
                 sensor={"type": "distantflux", "film_resolution": (1, 1)})
    _, _, _, wl = _render(oracle, d, 40000)
    flux = wl.sum()
    assert abs(flux / (E0 * np.cos(np.radians(sza))) - 1.0) < 0.03


def test_perspective_sensor_ray_convention(oracle):
    """perspective.cpp:200-236 + sensor.h:234-269: the film centre looks along +z of the camera,
    film x grows towards -x ('left' of look_at), rays start on the near plane."""
    kd = scenes.atmosphere_scene(geometry="plane_parallel", atmosphere=None, integrator="path",
                                 sensor={"type": "perspective", "origin": [1.0, -8.0, 6.0], "look_at": [1.0, 0.0, 0.0],
                                         "fov": 60.0, "film_resolution": (8, 4)})
    sc = mi_load_dict(kd)
    d = sc.flat.build_desc()
    fs = np.array([[0.5, 0.5], [0.0, 0.5], [1.0, 0.5], [0.5, 0.0], [0.5, 1.0]])
    o, dirs, w = oracle.sensor_ray(d, 0, fs, np.zeros((5, 2)))
    fwd = np.array([0.0, 8.0, -6.0]) / 10.0
    assert np.allclose(dirs[0], fwd, atol=1e-12) and np.allclose(w, 1.0)
    assert np.allclose(o[0], np.array([1.0, -8.0, 6.0]) + 1e-2 * fwd, atol=1e-12)
    # horizontal half angle = fov / 2 at the film edges; vertical one follows the aspect ratio
    assert np.isclose(np.degrees(np.arccos(dirs[1] @ fwd)), 30.0) and np.isclose(np.degrees(np.arccos(dirs[2] @ fwd)), 30.0)
    half_v = np.degrees(np.arctan(np.tan(np.radians(30.0)) / 2.0))
    assert np.isclose(np.degrees(np.arccos(dirs[3] @ fwd)), half_v) and np.isclose(np.degrees(np.arccos(dirs[4] @ fwd)), half_v)
    # look_at's left = up x dir; film x = 0 maps to +left
    left = np.cross([0.0, 0.0, 1.0], fwd)
    left /= np.linalg.norm(left)
    assert dirs[1] @ left > 0 > dirs[2] @ left
    assert dirs[3][2] > dirs[4][2]  # film y = 0 is the top row


def test_central_patch_surface_switches_bsdf_on_the_patch(oracle):
    """CentralPatchSurface = blendbsdf with the 0/1 central-patch mask (blendbsdf.cpp:108-165): seen from
    infinity without an atmosphere, pixels on the patch show rho_1 E cos / pi, pixels around it rho_0 E cos / pi."""
    sza, rho0, rho1 = 40.0, 0.1, 0.6
    sc = mi_load_dict(scenes.atmosphere_scene(
        geometry="plane_parallel", atmosphere=None, integrator="path", sza=sza,
        surface={"type": "diffuse", "reflectance": rho0},
        central_patch={"edges": (4.0, 2.0), "bsdf": {"type": "diffuse", "reflectance": rho1}},
        sensor={"type": "mpdistant", "vza": 30.0, "vaa": 70.0, "film_resolution": (4, 4),
                # an 8 m x 8 m image around the origin: 2 m pixels; the patch covers |x| <= 2, |y| <= 1
                "target": {"type": "rectangle", "to_world": scenes.ScalarTransform4f().scale([4.0, 4.0, 1.0])}}))
    d = sc.flat.build_desc()
    assert d.has_patch == 1 and np.allclose(list(d.patch_rect), [0.0, 0.0, 2.0, 1.0])
    mean, err, _, _ = _render(oracle, d, 20000)
    img, err = mean.reshape(4, 4), err.reshape(4, 4)  # [row = y, column = x], pixels span [-4,-2], [-2,0], [0,2], [2,4]
    e = E0 * np.cos(np.radians(sza)) / np.pi
    assert np.allclose(img[[0, 3], :], rho0 * e, rtol=1e-6)      # |y| > 2: background
    assert np.allclose(img[:, [0, 3]], rho0 * e, rtol=1e-6)      # |x| > 2: background
    # the four central pixels: |x| <= 2 is on the patch, |y| <= 1 is half of each pixel's height
    assert np.all(np.abs(img[1:3, 1:3] - 0.5 * (rho0 + rho1) * e) < 5.0 * err[1:3, 1:3] + 1e-9)


# ------------------------------------------------------------------ tree trunks (cylinder + cap, diffuse)
def _tree_scene(sensor, **kw):
    trees = dict(positions=((0.0, 0.0),), n_leaves=1, crown_radius=0.01, trunk_height=2.0, trunk_radius=0.25,
                 trunk_reflectance=0.4, reflectance=0.0, transmittance=0.0)
    trees.update(kw)
    return mi_load_dict(scenes.atmosphere_scene(
        geometry="plane_parallel", atmosphere=None, integrator="path", sza=50.0, saa=0.0, max_depth=2,
        surface={"type": "diffuse", "reflectance": 0.0}, canopy={"trees": trees, "size": (2.0, 2.0, 2.1)}, sensor=sensor))


def _brute_force_cylinder(c, o, d, tmax):
    """MI/src/shapes/cylinder.cpp:560-615 in numpy (open tube, near root first, far root otherwise)."""
    ax = c[3:6] - c[:3]
    L = np.linalg.norm(ax)
    u = ax / L
    w = o - c[:3]
    du, wu = d @ u, w @ u
    dp, wp = d - du[:, None] * u, w - wu[:, None] * u
    A, B, C_ = (dp * dp).sum(1), 2 * (dp * wp).sum(1), (wp * wp).sum(1) - c[6] ** 2
    disc = B * B - 4 * A * C_
    with np.errstate(invalid="ignore", divide="ignore"):
        sq = np.sqrt(np.maximum(disc, 0))
        t0, t1 = (-B - sq) / (2 * A), (-B + sq) / (2 * A)
    tn, tf = np.minimum(t0, t1), np.maximum(t0, t1)
    zn, zf = wu + du * tn, wu + du * tf
    ok = (disc >= 0) & (tn <= tmax) & (tf >= 0) & ~((tn < 0) & (tf > tmax))
    near = ok & (zn >= 0) & (zn <= L) & (tn >= 0)
    far = ok & ~near & (zf >= 0) & (zf <= L) & (tf <= tmax)
    return np.where(near, tn, np.where(far, tf, np.inf))


def test_cylinder_ray_caster_equals_brute_force(oracle):
    sc = _tree_scene({"type": "mdistant", "vza": [0.0], "vaa": 0.0}, positions=((0.0, 0.0), (1.5, -0.5)))
    d = sc.flat.build_desc()
    g = sc.flat.leaf_groups[0]
    rng = np.random.default_rng(4)
    n = 4000
    o = np.stack([rng.uniform(-2, 3, n), rng.uniform(-2, 2, n), rng.uniform(-0.05, 3.0, n)], axis=1)
    dirs = rng.normal(size=(n, 3))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    inside = rng.random(n) < 0.15  # some origins inside the tube: only the far wall can be hit
    o[inside] = np.stack([rng.uniform(-0.15, 0.15, inside.sum()), rng.uniform(-0.15, 0.15, inside.sum()),
                          rng.uniform(0.1, 1.9, inside.sum())], axis=1)
    tmax = np.where(rng.random(n) < 0.3, rng.uniform(0.05, 2.0, n), np.inf)
    t, nrm, grp = oracle.canopy_intersect(d, o, dirs, tmax)
    want = np.full(n, np.inf)
    for _, off in sc.flat.instances:
        cyl = g.cylinders[0].astype(np.float32).astype(np.float64).copy()
        cyl[:3] += off
        cyl[3:6] += off
        want = np.minimum(want, _brute_force_cylinder(cyl, o, dirs, tmax))
        want = np.minimum(want, _brute_force(np.vstack([g.disks, g.trunk_disks]).astype(np.float32).astype(np.float64),
                                             [off], o, dirs, tmax))
    assert np.array_equal(np.isfinite(t), np.isfinite(want))
    hit = np.isfinite(t)
    assert hit.sum() > 300 and (hit & inside).sum() > 50
    assert np.allclose(t[hit], want[hit], rtol=1e-9, atol=1e-9)


def test_trunk_is_a_one_sided_lambertian_tube(oracle):
    """Rays aimed at known points of a lone trunk (black ground, no atmosphere, direct light only):
    L = rho E max(n . s, 0) / pi with n the outward normal of the tube, or +z on its cap."""
    sza, rho, r = 50.0, 0.4, 0.25
    s_dir = np.array([np.sin(np.radians(sza)), 0.0, np.cos(np.radians(sza))])  # towards the sun (saa = 0 -> +x)
    phis = np.radians([0.0, 60.0, 120.0, 200.0])
    pts = np.stack([r * np.cos(phis), r * np.sin(phis), [0.5, 1.0, 1.5, 0.7]], axis=1)
    nrm = np.stack([np.cos(phis), np.sin(phis), np.zeros(4)], axis=1)
    org = pts + 3.0 * nrm + np.array([0.0, 0.0, 0.4])        # look at the wall from outside, slightly from above
    pts = np.vstack([pts, [[0.05, -0.1, 2.0]]])             # and at the cap from straight above
    nrm = np.vstack([nrm, [[0.0, 0.0, 1.0]]])
    org = np.vstack([org, [[0.05, -0.1, 6.0]]])
    dirs = pts - org
    sc = _tree_scene({"type": "mradiancemeter", "origins": org, "directions": dirs}, trunk_reflectance=rho, trunk_radius=r)
    mean, err, st, _ = _render(oracle, sc.flat.build_desc(), 64)
    want = rho * E0 * np.maximum(nrm @ s_dir, 0.0) / np.pi
    assert np.allclose(mean, want, rtol=1e-6, atol=1e-12), (mean, want)
    assert want[2] == 0.0 and want[3] == 0.0 and want[0] > 0 and want[4] > 0
