"""
Host-side logic of the drop-in boundary (CPU only): dict flattening, parameter
traversal/update protocol, bitmap facade, seeds, error behaviour.  Mirrors the
reference's tests/01_unit/kernel/test_render.py and test_kernel_dict.py.
"""

import ctypes as C
import warnings

import numpy as np
import pytest

from eradiate_b200 import _abi, scenes
from eradiate_b200.kernel import (
    Bitmap, KernelContext, KernelSceneParameterMap, Medium, SceneParameter, SearchSceneParameter,
    SeedState, develop, mi_load_dict, mi_traverse,
)
from eradiate_b200.kernel._scene import unflatten
from eradiate_b200.kernel._types import ScalarTransform4f


def test_unflatten_dotted_keys():
    d = {"type": "mdistant", "film.type": "hdrfilm", "film.width": 3, "film.rfilter.type": "box"}
    assert unflatten(d) == {"type": "mdistant", "film": {"type": "hdrfilm", "width": 3, "rfilter": {"type": "box"}}}


def test_flatten_c2_geometry():
    sc = mi_load_dict(scenes.config_c2(spp=16))
    f = sc.flat
    assert f.geometry == _abi.GEOM_SPHERICAL_SHELL
    assert np.isclose(f.surface_z, scenes.EARTH_RADIUS)
    assert np.isclose(f.medium_top, scenes.EARTH_RADIUS + scenes.TOA)
    assert np.isclose(f.medium_bottom, scenes.EARTH_RADIUS, rtol=1e-12)
    assert f.n_layers() == 1200
    # scene bbox = cube around the TOA sphere -> bounding sphere radius sqrt(3) * R_toa
    assert np.isclose(f.bsphere_radius, np.sqrt(3) * (scenes.EARTH_RADIUS + scenes.TOA))
    d = f.build_desc()
    assert d.n_phase == 1 and d.phase[0].type == _abi.PHASE_RAYLEIGH
    assert d.bsdf_type == _abi.BSDF_RPV
    assert np.allclose(list(d.bsdf_params)[:4], [0.027685, 0.95, -0.1, 0.027685])
    assert d.sensors[0].n_directions == 32 and d.sensors[0].target_type == _abi.TARGET_POINT
    sun = -np.array(list(d.emitter_direction))
    assert np.allclose(sun, scenes.angles_to_direction(30.0, 0.0))
    assert d.max_depth == -1 and d.rr_depth == 5 and d.integrator == _abi.INTEGRATOR_VOLPATH
    assert sc.integrator().moment


def test_flatten_c1_plane_parallel_homogeneous():
    sc = mi_load_dict(scenes.config_c1())
    f = sc.flat
    d = f.build_desc()
    assert f.geometry == _abi.GEOM_PLANE_PARALLEL and d.homogeneous == 1 and d.n_layers == 1
    assert np.isclose(f.surface_z, 0.0) and np.isclose(f.medium_top, scenes.TOA)
    assert np.isclose(d.sigma_t[0], 1.16e-5) and d.albedo[0] == 1.0
    assert d.bsdf_type == _abi.BSDF_DIFFUSE and d.bsdf_params[0] == 0.5


def test_blendphase_tree_flattening():
    """Nested blendphase (scenes/phase/_blend.py:188) -> leaf probabilities per layer."""
    n = 8
    w_outer = np.linspace(0.0, 1.0, n)
    w_inner = np.full(n, 0.25)
    d = scenes.atmosphere_scene(geometry="plane_parallel", atmosphere="afgl", n_layers=n)
    vol = lambda w: scenes._volume(w, False, scenes.EARTH_RADIUS, scenes.TOA, 1e9)  # noqa: E731
    d["phase_atmosphere"] = {
        "type": "blendphase", "id": "phase_atmosphere",
        "phase_0": {"type": "rayleigh"},
        "phase_1": {"type": "blendphase", "weight": vol(w_inner),
                    "phase_0": {"type": "hg", "g": 0.5}, "phase_1": {"type": "isotropic"}},
        "weight": vol(w_outer),
    }
    sc = mi_load_dict(d)
    desc = sc.flat.build_desc()
    assert desc.n_phase == 3
    assert [desc.phase[i].type for i in range(3)] == [_abi.PHASE_RAYLEIGH, _abi.PHASE_HG, _abi.PHASE_ISOTROPIC]
    pw = np.ctypeslib.as_array(desc.phase_weight, shape=(3, n))
    assert np.allclose(pw[0], 1 - w_outer, atol=1e-7)
    assert np.allclose(pw[1], w_outer * 0.75, atol=1e-7)
    assert np.allclose(pw[2], w_outer * 0.25, atol=1e-7)
    assert np.allclose(pw.sum(axis=0), 1.0, atol=1e-6)


def _set_integrator(ty):
    def mutate(d):
        d["integrator"] = {"type": ty}
    return mutate


@pytest.mark.parametrize("mutate,match", [
    (_set_integrator("path"), "ignores participating media"),
    (lambda d: d["surface_bsdf"].update({"type": "dielectric"}), "unsupported plugin type 'dielectric'"),
    (lambda d: d["measure"].update({"type": "perspective"}), "field of view"),
    (lambda d: d.update({"mesh": {"type": "ply"}}), "inside a shapegroup only"),
    (lambda d: d["measure"]["film"].update({"width": 5}), "Film size"),
    (lambda d: d["measure"]["sampler"].update({"type": "stratified"}), "sampler"),
    (lambda d: d["illumination"].update({"type": "constant"}), "unsupported"),
    (lambda d: d["medium_atmosphere"]["sigma_t"]["volume"].update({"filter_type": "trilinear"}), "nearest"),
])
def test_unsupported_plugins_raise_runtime_error(mutate, match):
    # experiments/_core.py:670-671: load errors surface as RuntimeError
    d = scenes.config_c2(spp=4)
    mutate(d)
    with pytest.raises(RuntimeError, match=match):
        mi_load_dict(d)


def test_glint_family_bsdfs_load_with_reference_defaults_and_keys():
    """ocean_mishchenko.cpp:95-123, ocean_grasp.cpp:120-153, maignan.cpp:92-111: constructor defaults, the
    parameters traverse() publishes (scalars by name, textures as `<name>.value`), and what they flatten to."""
    def bsdf_of(surface):
        d = scenes.atmosphere_scene(surface=surface, geometry="plane_parallel", n_layers=4)
        sc = mi_load_dict(d)
        keys = {k.split("surface_bsdf.")[-1] for k in mi_traverse(sc).parameters.keys() if k.startswith("surface_bsdf.")}
        desc = sc.flat.build_desc()
        return desc.bsdf_type, np.array(desc.bsdf_params[:8]), keys

    ty, p, keys = bsdf_of({"type": "ocean_mishchenko"})
    assert ty == _abi.BSDF_OCEAN_MISHCHENKO and np.allclose(p[:4], [0.1, 1.33, 0.0, 1.000277])
    assert keys == {"wind_speed", "eta.value", "k.value", "ext_ior.value"}
    ty, p, keys = bsdf_of({"type": "ocean_grasp", "wavelength": 865.0})
    assert ty == _abi.BSDF_OCEAN_GRASP and np.allclose(p[:7], [865.0, 0.1, 1.33, 0.0, 1.000277, 0.0, 0.0])
    assert keys == {"wavelength", "wind_speed.value", "eta.value", "k.value", "ext_ior.value",
                    "water_body_reflectance.value"}
    ty, p, keys = bsdf_of({"type": "maignan"})
    assert ty == _abi.BSDF_MAIGNAN and np.allclose(p[:5], [0.1, 0.0, 1.5, 0.0, 1.000277])
    assert keys == {"C.value", "ndvi.value", "refr_re.value", "refr_im.value", "ext_ior.value"}
    with pytest.raises(RuntimeError, match="wavelength"):
        bsdf_of({"type": "ocean_grasp"})
    with pytest.raises(RuntimeError, match="component"):
        bsdf_of({"type": "ocean_grasp", "wavelength": 550.0, "component": 2})


def test_multiphase_flattens_like_the_equivalent_blend_tree():
    """ERP/phase/multiphase.cpp:123-207: component i is drawn with probability w_i / sum(w) and evaluated as the
    normalised mixture, i.e. the leaves of the equivalent nested `blendphase` tree; weights need not sum to 1."""
    from tests.scene_battery import multiphase_scene

    a = mi_load_dict(multiphase_scene()).flat.build_desc()
    b = mi_load_dict(multiphase_scene(nested_blend=True)).flat.build_desc()
    assert a.n_phase == b.n_phase == 3
    assert [a.phase[i].type for i in range(3)] == [_abi.PHASE_RAYLEIGH, _abi.PHASE_HG, _abi.PHASE_ISOTROPIC]
    wa = np.ctypeslib.as_array(a.phase_weight, shape=(3, a.n_layers))
    wb = np.ctypeslib.as_array(b.phase_weight, shape=(3, b.n_layers))
    assert np.allclose(wa, wb, atol=2e-7) and np.allclose(wa.sum(axis=0), 1.0, atol=1e-6)
    # traverse() publishes phase<i> / weight<i> (multiphase.cpp:114-119)
    keys = {k.split("phase_atmosphere.")[-1] for k in mi_traverse(mi_load_dict(multiphase_scene())).parameters.keys()
            if k.startswith("phase_atmosphere.")}
    assert {"weight0.data", "weight1.data", "weight2.data", "phase1.g"} <= keys, keys


@pytest.mark.parametrize("phase,match", [
    ({"type": "multiphase", "phase0": {"type": "isotropic"}, "weight0": 1.0}, "At least 2 child phase functions"),
    ({"type": "multiphase", "phase0": {"type": "isotropic"}, "weight0": 1.0, "phase1": {"type": "hg"}}, "weight1"),
    # the mixture weight is only carried for a non-nested multiphase node
    ({"type": "blendphase", "weight": 0.5, "phase_0": {"type": "isotropic"},
      "phase_1": {"type": "multiphase", "phase0": {"type": "isotropic"}, "weight0": 1.0,
                  "phase1": {"type": "rayleigh", "depolarization": 0.03}, "weight1": 1.0}}, "use_mis=False"),
])
def test_multiphase_load_errors(phase, match):
    with pytest.raises(RuntimeError, match=match):
        mi_load_dict(scenes.atmosphere_scene(phase=phase, geometry="plane_parallel", n_layers=4)).flat.build_desc()


def test_multiphase_mis_flag():
    """`phase_mis` is set only when the mixture weight differs from the drawn component's own weight."""
    mk = lambda ph: mi_load_dict(scenes.atmosphere_scene(  # noqa: E731
        phase=ph, geometry="plane_parallel", n_layers=4)).flat.build_desc()
    base = {"type": "multiphase", "phase0": {"type": "hg", "g": 0.3}, "weight0": 1.0}
    assert mk({**base, "phase1": {"type": "rayleigh"}, "weight1": 1.0}).phase_mis == 0
    assert mk({**base, "phase1": {"type": "rayleigh", "depolarization": 0.03}, "weight1": 1.0}).phase_mis == 1
    assert mk({**base, "use_mis": False, "phase1": {"type": "rayleigh", "depolarization": 0.03}, "weight1": 1.0}).phase_mis == 0


def test_multiphase_without_mis_accepts_any_component():
    phase = {"type": "multiphase", "use_mis": False, "phase0": {"type": "hg", "g": 0.3}, "weight0": 0.2,
             "phase1": {"type": "rayleigh", "depolarization": 0.03}, "weight1": 0.6}
    d = mi_load_dict(scenes.atmosphere_scene(phase=phase, geometry="plane_parallel", n_layers=4)).flat.build_desc()
    w = np.ctypeslib.as_array(d.phase_weight, shape=(2, 4))
    assert np.allclose(w[0], 0.25) and np.allclose(w[1], 0.75)


def test_astroobject_emitter_conventions():
    """ERP/emitters/astroobject.cpp:54-108: `direction` / to_world * z points TOWARDS the object (the opposite of
    `directional`), default diameter 0.5358 deg, diameter validated in ]0, 180[."""
    sun = scenes.angles_to_direction(30.0, 40.0)
    a = mi_load_dict(scenes.atmosphere_scene(sza=30.0, saa=40.0, angular_diameter=1.5, n_layers=4)).flat.build_desc()
    b = mi_load_dict(scenes.atmosphere_scene(sza=30.0, saa=40.0, n_layers=4)).flat.build_desc()
    assert np.allclose(list(a.emitter_direction), -sun, atol=1e-12)  # the direction light travels, as `directional`
    assert np.allclose(list(a.emitter_direction), list(b.emitter_direction), atol=1e-12)
    assert a.emitter_angular_diameter == 1.5 and b.emitter_angular_diameter == 0.0
    d = scenes.atmosphere_scene(n_layers=4, angular_diameter=1.0)
    d["illumination"] = {"type": "astroobject", "direction": list(sun)}
    c = mi_load_dict(d).flat.build_desc()
    assert np.allclose(list(c.emitter_direction), -sun, atol=1e-12) and c.emitter_angular_diameter == 0.5358
    for bad in (0.0, 180.0):
        d["illumination"]["angular_diameter"] = bad
        with pytest.raises(RuntimeError, match="Invalid angular diameter"):
            mi_load_dict(d)
    d["illumination"] = {"type": "astroobject", "direction": list(sun), "to_world": np.eye(4)}
    with pytest.raises(RuntimeError, match="Only one of the parameters"):
        mi_load_dict(d)
    # `hide_emitters` (integrator.cpp:29; volpath.cpp:103, :329-330) travels in the descriptor, also under `moment`
    d = scenes.atmosphere_scene(n_layers=4, angular_diameter=1.0, moment=False)
    assert mi_load_dict(d).flat.build_desc().hide_emitters == 0
    d["integrator"]["hide_emitters"] = True
    assert mi_load_dict(d).flat.build_desc().hide_emitters == 1
    assert mi_load_dict(scenes.atmosphere_scene(n_layers=4, angular_diameter=1.0, hide_emitters=True)).flat.build_desc().hide_emitters == 1
    d["illumination"]["type"] = "directional"  # irrelevant for a delta emitter (never hit)
    del d["illumination"]["angular_diameter"]
    mi_load_dict(d)


def test_piecewise_volpath_needs_a_piecewise_medium():
    # only ERP/media/piecewise.cpp overrides the *_real interface (medium.cpp:99-118); the scene is
    # flattened at load time here, so the reference's render-time error surfaces from mi_load_dict
    d = scenes.config_c2(spp=4)
    _set_integrator("piecewise_volpath")(d)
    with pytest.raises(RuntimeError, match=r"HeterogeneousMedium::sample_interaction_real\(\): not implemented!"):
        mi_load_dict(d)


def test_traverse_parameter_keys_and_search():
    sc = mi_load_dict(scenes.config_c2(spp=4))
    umap = scenes.spectral_update_map(1200, spherical=True)
    umap["surface.rho_0"] = SceneParameter(lambda ctx: 0.2, parameter_id="surface_bsdf.rho_0.value")
    w = mi_traverse(sc, umap)
    keys = set(w.parameters.keys())
    # same dotted paths Mitsuba's traversal publishes
    for k in (
        "medium_atmosphere.sigma_t.volume.data",
        "medium_atmosphere.albedo.volume.data",
        "medium_atmosphere.scale",
        "illumination.irradiance.value",
        "surface_bsdf.rho_0.value",
        "surface_bsdf.k.value",
        "surface_bsdf.g.value",
    ):
        assert k in keys, k
    # SearchSceneParameter lookups were resolved (kernel/_render.py:314-321)
    assert w.umap_template["medium_atmosphere.sigma_t"].parameter_id == \
        "medium_atmosphere.sigma_t.volume.data"
    assert w.umap_template["illumination.irradiance.value"].parameter_id == "illumination.irradiance.value"
    # drop_parameters keeps only what the update map touches (kernel/_render.py:122-140)
    w.drop_parameters()
    assert len(w.parameters) == 4
    # update protocol
    ctx = KernelContext(w=440.0)
    w.parameters.update(w.umap_template.render(ctx))
    flat = sc.flat
    st = flat._profile(flat.medium.children["sigma_t"], 1200, "sigma_t")
    assert np.allclose(st, scenes.afgl_like_profile(1200, scenes.TOA, 440.0)[1])
    assert np.isclose(flat.emitter.children["irradiance"].values["value"], 1.8 * 550 / 440)
    assert np.isclose(flat.bsdf_params()[0], 0.2)
    assert flat.medium.children["sigma_t"].children["volume"].dirty


def test_traverse_warns_on_unsuccessful_lookup():
    sc = mi_load_dict(scenes.config_c1())
    umap = KernelSceneParameterMap({"x": SceneParameter(lambda ctx: 1.0,
                                    search=SearchSceneParameter(Medium, "does_not_exist", "scale"))})
    with pytest.warns(UserWarning, match="unsuccessful"):
        mi_traverse(sc, umap)


def test_update_size_mismatch_raises():
    sc = mi_load_dict(scenes.config_c2(spp=4))
    w = mi_traverse(sc)
    with pytest.raises(RuntimeError, match="size mismatch"):
        w.parameters.update({"medium_atmosphere.sigma_t.volume.data": np.zeros(7)})
    with pytest.raises(KeyError):
        w.parameters.update({"nope.value": 1.0})


def test_bitmap_split_matches_experiment_process_protocol():
    # experiments/_core.py:714-744
    sc = mi_load_dict(scenes.config_c2(spp=4, n_vza=5))
    spp = 10
    s1 = np.arange(5, dtype=float)
    bmp = develop(sc, 0, s1 * 2, s1, s1 * s1, spp)
    splits = dict(bmp.split())
    # sorted by name as Bitmap::split does (bitmap.cpp:692-696; order recorded from the reference in
    # tests/golden/reference_boundary.json)
    assert [k for k, _ in bmp.split()] == ["<root>", "m2_nested", "nested"]
    assert splits["<root>"].pixel_format() == Bitmap.PixelFormat.Y
    assert splits["m2_nested"].pixel_format() == Bitmap.PixelFormat.XYZ
    assert np.array(bmp).shape == (1, 5, 7)
    assert np.allclose(np.array(splits["<root>"])[0, :, 0], s1 * 2 / spp)
    assert np.allclose(np.array(splits["m2_nested"])[0, :, 0], s1 * s1 / spp)
    copy = Bitmap(bmp)
    copy._data[:] = 0
    assert np.array(bmp).max() > 0  # deep copy (kernel/_render.py:466)


def test_seed_state_matches_numpy_seed_sequence():
    # src/eradiate/rng.py: SeedSequence.spawn(1)[0].generate_state(1)
    ss = SeedState(0)
    ref = np.random.SeedSequence(0)
    for _ in range(3):
        assert ss.next()[0] == ref.spawn(1)[0].generate_state(1)[0]


def test_desc_struct_layout_is_stable():
    assert C.sizeof(_abi.PhaseDesc) == 80
    assert C.sizeof(_abi.RenderStats) == 56
    assert C.sizeof(_abi.SensorDesc) == 360
    assert C.sizeof(_abi.LeafGroupDesc) == 88
    assert C.sizeof(_abi.SceneDesc) == 752
    assert _abi.SceneDesc.patch_rect.offset + 32 == _abi.SceneDesc.bsdf_table.offset == 712
    assert _abi.SceneDesc.bsdf_table_res.offset == 720


# ------------------------------------------------------------------ canopy / 3D scenes (host side)
def _canopy_scene(**kw):
    kw.setdefault("canopy", {"lai": 1.0, "radius": 0.2, "size": (2.0, 2.0, 1.0), "padding": 1, "seed": 3})
    kw.setdefault("geometry", "plane_parallel")
    kw.setdefault("n_layers", 10)
    return scenes.atmosphere_scene(**kw)


def test_flatten_canopy_groups_instances_and_leaf_optics():
    """shapegroup + instances exactly as _core.py:266-296 / _leaf_cloud.py:1150-1175 emit them."""
    kd = _canopy_scene(canopy={"lai": 1.0, "radius": 0.2, "size": (2.0, 2.0, 1.0), "padding": 1, "seed": 3,
                               "reflectance": 0.3, "transmittance": 0.2})
    sc = mi_load_dict(kd)
    f = sc.flat
    n = int(round(1.0 * 4.0 / (np.pi * 0.04)))
    assert len(f.leaf_groups) == 1 and f.leaf_groups[0].disks.shape == (n, 7) and len(f.instances) == 9
    dk = f.leaf_groups[0].disks
    assert np.allclose(np.linalg.norm(dk[:, 3:6], axis=1), 1.0) and np.allclose(dk[:, 6], 0.2)
    offs = sorted(tuple(o) for _, o in f.instances)
    assert offs[0] == (-2.0, -2.0, 0.0) and offs[-1] == (2.0, 2.0, 0.0)
    d = f.build_desc()
    assert d.n_leaf_groups == 1 and d.n_instances == 9 and d.leaf_groups[0].n_disks == n
    assert np.isclose(d.leaf_groups[0].reflectance, 0.3) and np.isclose(d.leaf_groups[0].transmittance, 0.2)
    # the distant sensor targets the top of the unit cell (experiments/_canopy_atmosphere.py:200-210)
    assert d.sensors[0].target_type == _abi.TARGET_RECTANGLE
    # leaf optics are scene parameters; the leaves themselves are not exposed one by one
    keys = list(mi_traverse(sc).parameters.keys())
    assert "bsdf_leaf_cloud.reflectance.value" in keys and "bsdf_leaf_cloud.transmittance.value" in keys
    assert not any("leaf_cloud_leaf_" in k for k in keys)


def test_perspective_sensor_fov_axis_and_medium():
    cam = {"type": "perspective", "origin": [0.0, -10.0, 5.0], "look_at": [0.0, 0.0, 0.0], "fov": 40.0,
           "film_resolution": (8, 4)}
    kd = _canopy_scene(sensor=dict(cam, medium={"type": "ref", "id": "medium_atmosphere"}))
    d = mi_load_dict(kd).flat.build_desc()
    s = d.sensors[0]
    assert s.type == _abi.SENSOR_PERSPECTIVE and s.in_medium == 1 and np.isclose(s.x_fov_deg, 40.0)
    assert np.isclose(s.near_clip, 1e-2) and np.isclose(s.far_clip, 1e4)
    kd["measure"]["fov_axis"] = "y"  # sensor.cpp parse_fov: vertical fov -> horizontal through the aspect ratio
    s = mi_load_dict(kd).flat.build_desc().sensors[0]
    assert np.isclose(s.x_fov_deg, np.degrees(2 * np.arctan(np.tan(np.radians(20.0)) * 2.0)))


@pytest.mark.parametrize("mutate,match", [
    (lambda d: d["leaf_cloud_instance_0"].update(
        {"to_world": ScalarTransform4f().rotate([0, 0, 1], 30.0)}), "only translations"),
    (lambda d: d["leaf_cloud"].update({"trunk": {"type": "cylinder", "bsdf": {"type": "rpv"}}}), "trunks a diffuse one"),
    (lambda d: d["leaf_cloud"].update({"cube": {"type": "cube"}}), "unsupported child shape"),
    (lambda d: d["bsdf_leaf_cloud"].update({"type": "diffuse"}), "bilambertian"),
    (lambda d: d["leaf_cloud_instance_0"].update(
        {"to_world": ScalarTransform4f().translate([0.0, 0.0, 2.0e5])}), "below the top of the atmosphere"),
    (lambda d: d["leaf_cloud_instance_0"].update(
        {"to_world": ScalarTransform4f().translate([0.0, 0.0, -5.0])}), "above the ground"),
    (lambda d: d["leaf_cloud"].update(
        {"leaf_cloud_leaf_0": dict(d["leaf_cloud"]["leaf_cloud_leaf_0"],
                                   to_world=ScalarTransform4f().scale([1.0, 2.0, 1.0]))}), "circular"),
    (lambda d: d["surface_bsdf"].update({"type": "blendbsdf"}), "CentralPatchSurface"),
])
def test_canopy_restrictions_raise_at_load(mutate, match):
    kd = _canopy_scene()
    mutate(kd)
    with pytest.raises(RuntimeError, match=match):
        mi_load_dict(kd)


def test_canopy_and_camera_need_the_plane_parallel_geometry():
    kd = scenes.atmosphere_scene(geometry="spherical_shell", n_layers=10)
    kd.update(scenes.disc_canopy(n_leaves=3))
    with pytest.raises(RuntimeError, match="plane-parallel scenes only"):
        mi_load_dict(kd)
    kd = scenes.atmosphere_scene(geometry="spherical_shell", n_layers=10, sensor={
        "type": "perspective", "origin": [0.0, 50.0, scenes.EARTH_RADIUS + 10.0], "look_at": [0.0, 0.0, 0.0]})
    with pytest.raises(RuntimeError, match="plane-parallel scenes only"):
        mi_load_dict(kd)


def test_path_integrator_is_accepted_without_a_medium_only():
    sc = mi_load_dict(_canopy_scene(atmosphere=None, integrator="path"))
    assert sc.flat.build_desc().has_medium == 0
    with pytest.raises(RuntimeError, match="ignores participating media"):
        mi_load_dict(_canopy_scene(integrator="path"))


def test_descriptor_owns_its_buffers():
    """A descriptor stays valid when build_desc() is called again (the device scene builds its own)."""
    sc = mi_load_dict(_canopy_scene())
    d1 = sc.flat.build_desc()
    first = np.ctypeslib.as_array(d1.leaf_groups[0].disks, (d1.leaf_groups[0].n_disks, 7)).copy()
    for _ in range(3):
        sc.flat.build_desc()
    again = np.ctypeslib.as_array(d1.leaf_groups[0].disks, (d1.leaf_groups[0].n_disks, 7))
    assert np.array_equal(first, again)
