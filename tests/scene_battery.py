"""Scene battery shared by the golden-fixture generator and the GPU parity tests."""
import os

import numpy as np

from eradiate_b200 import scenes

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def measured(fname: str, wavelength: float) -> dict:
    """`measured_mono` on one of the synthetic RGL tensor files of tests/golden (tools/make_measured_fixture.py)."""
    return {"type": "measured_mono", "filename": os.path.join(GOLDEN, fname), "wavelength": wavelength}


POMMEROL = dict(w=0.526, theta=13.3, b=0.187, c=(1.0 + 0.273) / 2.0, h=0.083, B_0=1.0)
VZA5 = {"type": "mdistant", "vza": [-70.0, -35.0, 0.0, 35.0, 70.0], "vaa": 0.0}


def polarized_aerosol_scene() -> dict:
    """Molecular rayleigh_polarized + aerosol tabphase_polarized blend (C5-like, without the ocean)."""
    d = scenes.atmosphere_scene(geometry="plane_parallel", aerosol=True, aerosol_phase="hg", n_layers=60,
                                phase={"type": "rayleigh_polarized"}, stokes=True, meridian_align=False,
                                sza=30.0, saa=0.0, surface={"type": "diffuse", "reflectance": 0.05},
                                sensor={"type": "mdistant", "vza": [-50.0, -10.0, 40.0], "vaa": 60.0})
    d["phase_atmosphere"]["phase_1"] = scenes.polarized_aerosol_table()
    return d


MESH_TREE = [
    {"id": "crown", "filename": os.path.join(GOLDEN, "mesh_crown.ply"), "scale": 0.01, "reflectance": 0.45, "transmittance": 0.4},
    {"id": "trunk", "filename": os.path.join(GOLDEN, "mesh_trunk.obj"), "scale": 0.01, "reflectance": 0.3, "transmittance": 0.0},
]
MESH_LEAF = {"id": "curled_leaf", "filename": os.path.join(GOLDEN, "mesh_leaf_normals_ascii.ply"), "scale": 1.0,
             "reflectance": 0.5, "transmittance": 0.3}
CANOPY = {"lai": 2.5, "radius": 0.1, "size": (4.0, 4.0, 1.0), "padding": 1, "seed": 6}


def _with_polarized_aerosol(d: dict) -> dict:
    d["phase_atmosphere"]["phase_1"] = scenes.polarized_aerosol_table()
    return d


def _first_sensor(d: dict) -> dict:
    """Fixtures hold sensor 0 only: drop the others."""
    keep, seen = {}, False
    for k, v in d.items():
        is_sensor = isinstance(v, dict) and v.get("type") in ("mdistant", "hdistant", "distantflux", "perspective")
        if is_sensor and seen:
            continue
        seen = seen or is_sensor
        keep[k] = v
    return keep


def multiphase_scene(nested_blend: bool = False) -> dict:
    """Three phase functions mixed by per-layer `multiphase` weights (not normalised: multiphase.cpp:131-138 does
    it) over the AFGL-shaped profile; `nested_blend`: the equivalent tree of two `blendphase` nodes."""
    n = 80
    zn = (np.arange(n) + 0.5) / n
    w = [np.full(n, 1.0), 3.0 * np.exp(-6.0 * zn), 0.5 * zn]
    vol = lambda v: scenes._volume(v, False, scenes.EARTH_RADIUS, scenes.TOA, 1.0e9)  # noqa: E731
    leaves = [{"type": "rayleigh"}, {"type": "hg", "g": 0.7}, {"type": "isotropic"}]
    if nested_blend:  # blendphase.cpp:100-141: weight = probability of phase_1
        phase = {"type": "blendphase", "phase_0": leaves[0],
                 "phase_1": {"type": "blendphase", "phase_0": leaves[1], "phase_1": leaves[2],
                             "weight": vol(w[2] / (w[1] + w[2]))},
                 "weight": vol((w[1] + w[2]) / (w[0] + w[1] + w[2]))}
    else:
        phase = {"type": "multiphase", "use_mis": True}
        for i in range(3):
            phase[f"phase{i}"] = leaves[i]
            phase[f"weight{i}"] = vol(w[i])
    return scenes.atmosphere_scene(geometry="plane_parallel", n_layers=n, phase=phase, sza=50.0, saa=0.0,
                                   surface={"type": "diffuse", "reflectance": 0.15},
                                   sensor={"type": "mdistant", "vza": [-65.0, -30.0, 0.0, 30.0, 65.0], "vaa": 0.0})


def mq_table(nx=16, ny=25, nz=12):
    """A quasi-diffuse measured BRDF table [cos_theta_i][phi_d][cos_theta_o] for `mqdiffuse`, deliberately not
    symmetric in phi_d -> 2 pi - phi_d (so that the sign of the azimuth difference matters) and with distinct
    first / last phi_d planes (so that the un-wrapped azimuth of mqdiffuse.cpp:125-131 shows)."""
    from eradiate_b200.kernel import VolumeGrid

    co, ph, ci = np.meshgrid(np.linspace(0, 1, nx), np.linspace(0, 2 * np.pi, ny), np.linspace(0, 1, nz), indexing="ij")
    v = 0.25 / np.pi * (1.0 + 0.3 * np.cos(ph) + 0.15 * np.sin(ph) + 0.1 * ph / (2 * np.pi)) * (0.7 + 0.5 * co * ci)
    return VolumeGrid(np.ascontiguousarray(v.transpose(2, 1, 0)).astype(np.float32))


def battery() -> dict:
    """name -> scene dict.  Small films; every plugin of SURVEY 8a appears at least once."""
    S = scenes.atmosphere_scene
    return {
        # BASELINE configs at reduced size
        "c1_homogeneous_lambertian_pp": scenes.config_c1(spp=16),
        "c2_afgl_rpv_spherical": scenes.config_c2(spp=16, n_vza=8),
        "c3_afgl_aerosol_tab_hdistant": scenes.config_c3(spp=16, res=4),
        # the same two at the film sizes BASELINE.json quotes (32 view angles; the 32x32 hemispherical film)
        "c2_full_size_32vza": scenes.config_c2(spp=16),
        "c3_full_film_32x32": scenes.config_c3(spp=16),
        # geometry x medium
        "afgl_rpv_pp": S(geometry="plane_parallel", sensor=VZA5, sza=50.0, saa=30.0),
        "homogeneous_spherical_hg": S(geometry="spherical_shell", atmosphere="homogeneous",
                                      homogeneous_sigma_t=4e-6, homogeneous_albedo=0.9,
                                      phase={"type": "hg", "g": 0.6}, sensor=VZA5),
        "thick_isotropic_pp": S(geometry="plane_parallel", atmosphere="homogeneous",
                                homogeneous_sigma_t=5.0 / scenes.TOA, homogeneous_albedo=0.99,
                                phase={"type": "isotropic"}, sensor=VZA5,
                                surface={"type": "diffuse", "reflectance": 0.2}),
        # BSDFs
        "rtls_pp": S(geometry="plane_parallel", surface={"type": "rtls"}, sensor=VZA5, n_layers=200),
        "rtls_rb_spherical": S(surface={"type": "rtls", "f_iso": 0.3, "f_vol": 0.2, "f_geo": 0.05,
                                        "h": 1.5, "r": 1.2, "b": 0.9}, sensor=VZA5, n_layers=200),
        "hapke_spherical": S(surface={"type": "hapke", **POMMEROL}, sensor=VZA5, n_layers=200, sza=20.0),
        "lambertian_spherical_grazing_sun": S(surface={"type": "diffuse", "reflectance": 0.8},
                                              sensor=VZA5, n_layers=200, sza=85.0, saa=120.0),
        # phase functions
        "aerosol_hg_blend_pp": S(geometry="plane_parallel", aerosol=True, aerosol_phase="hg", sensor=VZA5),
        "aerosol_tab_irregular_spherical": S(aerosol=True, aerosol_phase="tabphase_irregular", sensor=VZA5),
        "rayleigh_depolarized_pp": S(geometry="plane_parallel", n_layers=100, sensor=VZA5,
                                     phase={"type": "rayleigh", "depolarization": 0.0279}),
        # depolarization factor as a per-layer volume (rayleigh.cpp:48,79; scenes/phase/_rayleigh.py:98-131),
        # alone and below a blendphase node
        "rayleigh_depolarization_profile_pp": S(geometry="plane_parallel", n_layers=60, sensor=VZA5, sza=40.0,
                                                phase={"type": "rayleigh",
                                                       "depolarization": np.linspace(0.0, 0.45, 60) ** 2 / 0.45}),
        "aerosol_blend_depolarization_profile_spherical": S(
            aerosol=True, aerosol_phase="hg", n_layers=60, sensor=VZA5, sza=55.0, saa=40.0,
            phase={"type": "rayleigh", "depolarization": 0.4 - np.linspace(0.0, 0.6, 60) ** 2}),
        # sensors
        "hdistant_pp": S(geometry="plane_parallel", n_layers=100,
                         sensor={"type": "hdistant", "film_resolution": (3, 3)}),
        "distantflux_spherical": S(n_layers=100, sensor={"type": "distantflux", "film_resolution": (3, 2)}),
        "distantflux_no_target_spherical": S(n_layers=100,
                                             sensor={"type": "distantflux", "film_resolution": (2, 2),
                                                     "target": None}),
        # ocean_legacy (6SV): glint + whitecaps + underlight, wind direction relative to the frame
        "ocean_pp": S(geometry="plane_parallel", n_layers=100, sza=35.0, saa=20.0,
                      surface={"type": "ocean_legacy", "wavelength": 550.0, "wind_speed": 8.0,
                               "wind_direction": 40.0, "chlorinity": 19.0, "pigmentation": 0.3,
                               "shadowing": True},
                      sensor={"type": "mdistant", "vza": [-60.0, -35.0, 0.0, 35.0, 60.0], "vaa": 20.0}),
        "ocean_spherical_nir": S(n_layers=100, sza=35.0, w_nm=865.0,
                                 surface={"type": "ocean_legacy", "wavelength": 865.0, "wind_speed": 3.0,
                                          "wind_direction": 0.0, "shadowing": False},
                                 sensor={"type": "mdistant", "vza": [-50.0, -35.0, -20.0, 20.0, 50.0],
                                         "vaa": 0.0, "target": [0.0, 3.0e5, 6.3710484e6]}),
        # SURVEY 8f-4 glint family: ocean_mishchenko (glint only), ocean_grasp (whitecaps + water body + glint),
        # maignan (polarized land reflectance; eval without cosine, sample weight C F)
        "ocean_mishchenko_pp": S(geometry="plane_parallel", n_layers=100, sza=35.0, saa=20.0,
                                 surface={"type": "ocean_mishchenko", "wind_speed": 5.0, "eta": 1.34, "k": 0.0},
                                 sensor={"type": "mdistant", "vza": [-60.0, -35.0, -25.0, 0.0, 35.0], "vaa": 20.0}),
        "ocean_grasp_spherical": S(n_layers=100, sza=30.0,
                                   surface={"type": "ocean_grasp", "wavelength": 550.0, "wind_speed": 12.0,
                                            "eta": 1.336, "k": 0.0, "water_body_reflectance": 0.02},
                                   sensor={"type": "mdistant", "vza": [-50.0, -30.0, -15.0, 20.0, 50.0], "vaa": 0.0}),
        "maignan_pp": S(geometry="plane_parallel", n_layers=100, sza=40.0, saa=10.0,
                        surface={"type": "maignan", "C": 5.0, "ndvi": 0.4, "refr_re": 1.5, "refr_im": 0.0},
                        sensor={"type": "mdistant", "vza": [-60.0, -40.0, 0.0, 30.0, 60.0], "vaa": 10.0}),
        "multiphase_three_components_pp": multiphase_scene(),
        # multiphase with use_mis over components whose value is not their pdf: the mixture weight is carried by the
        # kernels (phase_mis)
        "multiphase_mis_depolarized_rayleigh_pp": S(
            geometry="plane_parallel", atmosphere="homogeneous", homogeneous_sigma_t=1.5 / scenes.TOA,
            homogeneous_albedo=0.97, sensor=VZA5, sza=40.0, surface={"type": "diffuse", "reflectance": 0.1},
            phase={"type": "multiphase", "use_mis": True,
                   "phase0": {"type": "rayleigh", "depolarization": 0.25}, "weight0": 2.0,
                   "phase1": {"type": "hg", "g": 0.75}, "weight1": 1.0,
                   "phase2": {"type": "isotropic"}, "weight2": 0.5}),
        "mqdiffuse_pp": S(geometry="plane_parallel", n_layers=100, sza=40.0, saa=25.0,
                          surface={"type": "mqdiffuse", "grid": mq_table()},
                          sensor={"type": "mdistant", "vza": [-60.0, -30.0, 0.0, 30.0, 60.0], "vaa": 70.0}),
        "mqdiffuse_spherical_thick": S(geometry="spherical_shell", atmosphere="homogeneous",
                                       homogeneous_sigma_t=1.0 / scenes.TOA, homogeneous_albedo=0.95,
                                       phase={"type": "hg", "g": 0.5}, sza=30.0, saa=200.0,
                                       surface={"type": "mqdiffuse", "grid": mq_table(9, 13, 7)},
                                       sensor={"type": "mdistant", "vza": [-50.0, -20.0, 20.0, 50.0], "vaa": 140.0,
                                               "target": [2.0e5, 3.0e5, 6.3679007e6]}),
        # measured_mono (measured_mono.cpp): isotropic and anisotropic (reduction = 2) tensor files, between two
        # wavelength nodes; the thin atmosphere keeps the surface term dominant
        "measured_iso_pp": S(geometry="plane_parallel", n_layers=100, sza=40.0, saa=25.0,
                             surface=measured("measured_iso.bsdf", 600.0),
                             sensor={"type": "mdistant", "vza": [-60.0, -30.0, 0.0, 30.0, 60.0], "vaa": 70.0}),
        "measured_aniso_spherical_thick": S(geometry="spherical_shell", atmosphere="homogeneous",
                                            homogeneous_sigma_t=1.0 / scenes.TOA, homogeneous_albedo=0.95,
                                            phase={"type": "hg", "g": 0.5}, sza=30.0, saa=200.0,
                                            surface=measured("measured_aniso.bsdf", 520.0),
                                            sensor={"type": "mdistant", "vza": [-50.0, -20.0, 20.0, 50.0], "vaa": 140.0,
                                                    "target": [2.0e5, 3.0e5, 6.3679007e6]}),
        "astro_measured_aniso_pp": S(geometry="plane_parallel", n_layers=40, sza=55.0, saa=120.0, angular_diameter=4.0,
                                     surface=measured("measured_aniso.bsdf", 725.0),
                                     sensor={"type": "mdistant", "vza": [-70.0, -40.0, -10.0, 20.0, 50.0], "vaa": 20.0}),
        # astroobject: a solar disc instead of the delta directional emitter (next-event directions in a cone, the
        # disc seen directly by unscattered primary rays)
        "astro_wide_disc_afgl_rpv_pp": S(geometry="plane_parallel", n_layers=100, sza=50.0, saa=30.0,
                                         angular_diameter=12.0, sensor=VZA5),
        "astro_sun_aerosol_tab_spherical": S(n_layers=120, sza=35.0, aerosol=True, aerosol_phase="tabphase",
                                             angular_diameter=0.5358,
                                             sensor={"type": "mdistant", "vza": [-60.0, -20.0, 20.0, 60.0], "vaa": 0.0}),
        "astro_piecewise_ocean_grasp_pp": S(geometry="plane_parallel", n_layers=80, integrator="piecewise_volpath",
                                            sza=30.0, angular_diameter=6.0,
                                            surface={"type": "ocean_grasp", "wavelength": 550.0, "wind_speed": 3.0,
                                                     "water_body_reflectance": 0.01},
                                            sensor={"type": "mdistant", "vza": [-45.0, -30.0, -15.0, 30.0], "vaa": 0.0}),
        "astro_volpathmis_thick_hg_pp": S(geometry="plane_parallel", atmosphere="homogeneous", integrator="volpathmis",
                                          homogeneous_sigma_t=2.0 / scenes.TOA, homogeneous_albedo=0.95,
                                          phase={"type": "hg", "g": 0.6}, angular_diameter=8.0, sensor=VZA5,
                                          surface={"type": "diffuse", "reflectance": 0.3}),
        "astro_direct_beam_from_ground_spherical": S(
            n_layers=100, sza=40.0, saa=0.0, angular_diameter=0.5358, surface={"type": "diffuse", "reflectance": 0.2},
            sensor={"type": "mradiancemeter", "medium": {"type": "ref", "id": "medium_atmosphere"},
                    # a sun photometer at the ground: disc centre, inside near the limb, the aureole just outside,
                    # and the zenith sky
                    "origins": [[0.0, 0.0, scenes.EARTH_RADIUS + 1.0]] * 4,
                    "directions": [list(scenes.angles_to_direction(40.0, 0.0)), list(scenes.angles_to_direction(40.2, 0.0)),
                                   list(scenes.angles_to_direction(40.4, 0.0)), [0.0, 0.0, 1.0]]}),
        # the same photometer with `hide_emitters` (integrator.cpp:29, volpath.cpp:329-330): the disc itself is not
        # seen, only the light scattered into the line of sight
        "astro_hide_emitters_from_ground_spherical": S(
            n_layers=100, sza=40.0, saa=0.0, angular_diameter=0.5358, surface={"type": "diffuse", "reflectance": 0.2},
            hide_emitters=True,
            sensor={"type": "mradiancemeter", "medium": {"type": "ref", "id": "medium_atmosphere"},
                    "origins": [[0.0, 0.0, scenes.EARTH_RADIUS + 1.0]] * 4,
                    "directions": [list(scenes.angles_to_direction(40.0, 0.0)), list(scenes.angles_to_direction(40.2, 0.0)),
                                   list(scenes.angles_to_direction(40.4, 0.0)), [0.0, 0.0, 1.0]]}),
        # polarized (Stokes) transport: rayleigh_polarized / tabphase_polarized + stokes integrator
        "polarized_rayleigh_pp": S(geometry="plane_parallel", n_layers=100, sza=40.0, saa=30.0, stokes=True,
                                   phase={"type": "rayleigh_polarized", "depolarization": 0.0279},
                                   surface={"type": "diffuse", "reflectance": 0.1},
                                   sensor={"type": "mdistant", "vza": [-60.0, -30.0, 0.0, 30.0, 60.0], "vaa": 90.0}),
        "polarized_rayleigh_spherical_thick": S(geometry="spherical_shell", atmosphere="homogeneous",
                                                homogeneous_sigma_t=1.0 / scenes.TOA, homogeneous_albedo=0.98,
                                                phase={"type": "rayleigh_polarized"}, stokes=True, sza=50.0,
                                                surface={"type": "rpv", "rho_0": 0.1, "k": 0.9, "g": -0.1},
                                                sensor={"type": "mdistant", "vza": [-70.0, -20.0, 20.0, 70.0], "vaa": 45.0}),
        "polarized_rayleigh_depolarization_profile_spherical": S(
            n_layers=60, sza=45.0, saa=20.0, stokes=True,
            phase={"type": "rayleigh_polarized", "depolarization": np.linspace(0.0, 0.45, 60) ** 2 / 0.45},
            surface={"type": "diffuse", "reflectance": 0.05},
            sensor={"type": "mdistant", "vza": [-65.0, -35.0, 0.0, 30.0, 60.0], "vaa": 60.0}),
        "polarized_aerosol_tab_pp": polarized_aerosol_scene(),
        # C5-like: polarized ocean glint (complex Fresnel Mueller matrix) under a Rayleigh atmosphere
        "polarized_ocean_pp": S(geometry="plane_parallel", n_layers=60, sza=40.0, saa=0.0, stokes=True,
                                phase={"type": "rayleigh_polarized"},
                                surface={"type": "ocean_legacy", "wavelength": 550.0, "wind_speed": 5.0,
                                         "wind_direction": 30.0, "shadowing": True},
                                sensor={"type": "mdistant", "vza": [-60.0, -40.0, -20.0, 20.0, 50.0], "vaa": 25.0}),
        "polarized_mishchenko_pp": S(geometry="plane_parallel", n_layers=60, sza=40.0, saa=0.0, stokes=True,
                                     phase={"type": "rayleigh_polarized"},
                                     surface={"type": "ocean_mishchenko", "wind_speed": 4.0, "eta": 1.33, "k": 0.0,
                                              "ext_ior": 1.0},
                                     sensor={"type": "mdistant", "vza": [-60.0, -40.0, -20.0, 20.0, 50.0], "vaa": 25.0}),
        "polarized_grasp_spherical": S(n_layers=60, sza=35.0, stokes=True, phase={"type": "rayleigh_polarized"},
                                       surface={"type": "ocean_grasp", "wavelength": 670.0, "wind_speed": 10.0,
                                                "eta": 1.331, "k": 0.0, "water_body_reflectance": 0.01},
                                       sensor={"type": "mdistant", "vza": [-55.0, -35.0, -10.0, 30.0, 60.0], "vaa": 15.0}),
        "polarized_maignan_pp": S(geometry="plane_parallel", n_layers=60, sza=45.0, saa=0.0, stokes=True,
                                  phase={"type": "rayleigh_polarized", "depolarization": 0.0279},
                                  surface={"type": "maignan", "C": 6.66, "ndvi": 0.3, "refr_re": 1.5, "refr_im": 0.0},
                                  sensor={"type": "mdistant", "vza": [-65.0, -45.0, -20.0, 25.0, 55.0], "vaa": 30.0}),
        "polarized_mqdiffuse_pp": S(geometry="plane_parallel", n_layers=60, sza=50.0, saa=10.0, stokes=True,
                                    phase={"type": "rayleigh_polarized"},
                                    surface={"type": "mqdiffuse", "grid": mq_table()},
                                    sensor={"type": "mdistant", "vza": [-55.0, -25.0, 15.0, 45.0], "vaa": 100.0}),
        "polarized_measured_iso_pp": S(geometry="plane_parallel", n_layers=60, sza=50.0, saa=10.0, stokes=True,
                                       phase={"type": "rayleigh_polarized"},
                                       surface=measured("measured_iso.bsdf", 450.0),
                                       sensor={"type": "mdistant", "vza": [-55.0, -25.0, 15.0, 45.0], "vaa": 100.0}),
        "polarized_astro_mishchenko_pp": S(geometry="plane_parallel", n_layers=60, sza=40.0, saa=0.0, stokes=True,
                                           phase={"type": "rayleigh_polarized"}, angular_diameter=3.0,
                                           surface={"type": "ocean_mishchenko", "wind_speed": 2.0, "eta": 1.33},
                                           sensor={"type": "mdistant", "vza": [-50.0, -40.0, -30.0, 20.0], "vaa": 0.0}),
        "polarized_multiphase_mis_pp": S(
            geometry="plane_parallel", atmosphere="homogeneous", homogeneous_sigma_t=1.0 / scenes.TOA,
            homogeneous_albedo=0.95, stokes=True, sza=45.0, saa=20.0, surface={"type": "diffuse", "reflectance": 0.05},
            sensor={"type": "mdistant", "vza": [-60.0, -25.0, 10.0, 50.0], "vaa": 60.0},
            phase={"type": "multiphase", "use_mis": True,
                   "phase0": {"type": "rayleigh_polarized", "depolarization": 0.03}, "weight0": 1.0,
                   "phase1": {"type": "hg", "g": 0.6}, "weight1": 1.0}),
        # BASELINE C5 at reduced size: polarized ocean + molecular + polarized aerosol, one band of the sweep
        "c5_polarized_ocean_aerosol_reduced": scenes.config_c5(spp=16, n_vza=4, w_nm=865.0, n_layers=120),
        # integrator options
        "volpathmis_thick": S(geometry="plane_parallel", atmosphere="homogeneous", integrator="volpathmis",
                              homogeneous_sigma_t=3.0 / scenes.TOA, homogeneous_albedo=0.95, sensor=VZA5,
                              surface={"type": "diffuse", "reflectance": 0.3}),
        "max_depth_3_rr_2": S(geometry="plane_parallel", atmosphere="homogeneous", max_depth=3, rr_depth=2,
                              homogeneous_sigma_t=2.0 / scenes.TOA, sensor=VZA5),
        "no_atmosphere_rpv_spherical": S(atmosphere=None, sensor=VZA5),
        # ERP/bsdfs/selectbsdf.cpp with a uniform index texture: the selected BSDF (here the second one) is the surface
        "selectbsdf_uniform_index_pp": S(geometry="plane_parallel", n_layers=50, sensor=VZA5, surface={
            "type": "selectbsdf", "indices": {"type": "uniform", "value": 1.0},
            "bsdf_0": {"type": "diffuse", "reflectance": {"type": "uniform", "value": 0.05}},
            "bsdf_1": {"type": "rtls", "f_iso": {"type": "uniform", "value": 0.25},
                       "f_vol": {"type": "uniform", "value": 0.1}, "f_geo": {"type": "uniform", "value": 0.02}}}),
        # piecewise medium + piecewise_volpath (the default Eradiate picks for plane-parallel atmospheres,
        # experiments/_atmosphere.py:165-184): analytic free flights, exact shadow-ray transmittance
        "piecewise_afgl_rpv_pp": S(geometry="plane_parallel", integrator="piecewise_volpath", sensor=VZA5,
                                   sza=50.0, saa=30.0),
        "piecewise_aerosol_blend_rr_pp": S(geometry="plane_parallel", integrator="piecewise_volpath", aerosol=True,
                                           aerosol_phase="tabphase", n_layers=120, rr_depth=1, sza=65.0, saa=200.0,
                                           surface={"type": "diffuse", "reflectance": 0.6},
                                           sensor={"type": "mdistant", "vza": [-80.0, -35.0, 0.0, 35.0, 80.0],
                                                   "vaa": 45.0}),
        "piecewise_ocean_hdistant_pp": S(geometry="plane_parallel", integrator="piecewise_volpath", n_layers=50,
                                         sza=35.0, saa=20.0, max_depth=6,
                                         surface={"type": "ocean_legacy", "wavelength": 550.0, "wind_speed": 6.0,
                                                  "wind_direction": 10.0},
                                         sensor={"type": "hdistant", "film_resolution": (3, 2)}),
        "piecewise_distantflux_coarse_pp": S(geometry="plane_parallel", integrator="piecewise_volpath", n_layers=3,
                                             sensor={"type": "distantflux", "film_resolution": (2, 2)}),
        # explicit 3D canopies (SURVEY 8f-3 / BASELINE C4): disk leaves in instanced shape groups, bilambertian
        # leaf BSDF, plane-parallel atmosphere around them; rendered by the 3D kernel (ertb_canopy.cuh)
        "canopy_path_no_atmosphere": S(geometry="plane_parallel", atmosphere=None, integrator="path", sza=35.0, saa=40.0,
                                       canopy=dict(CANOPY, reflectance=0.5, transmittance=0.4),
                                       surface={"type": "diffuse", "reflectance": 0.3},
                                       sensor={"type": "mdistant", "vza": [-60.0, -20.0, 0.0, 35.0, 70.0], "vaa": 40.0}),
        "canopy_volpath_afgl_rpv_pp": S(geometry="plane_parallel", n_layers=100, w_nm=670.0, sza=30.0, canopy=CANOPY,
                                        sensor=VZA5),
        # the 3D kernel is general: late-plugin ground BSDFs under a canopy (table lookup / glint in the local frame)
        "canopy_mqdiffuse_ground_pp": S(geometry="plane_parallel", n_layers=60, sza=35.0, saa=40.0, canopy=CANOPY,
                                        surface={"type": "mqdiffuse", "grid": mq_table()}, sensor=VZA5),
        "canopy_ocean_grasp_ground_pp": S(geometry="plane_parallel", n_layers=60, sza=30.0, canopy=dict(CANOPY, seed=5),
                                          surface={"type": "ocean_grasp", "wavelength": 550.0, "wind_speed": 8.0,
                                                   "water_body_reflectance": 0.03},
                                          sensor={"type": "mdistant", "vza": [-45.0, -30.0, 0.0, 30.0], "vaa": 0.0}),
        "canopy_piecewise_aerosol_pp": S(geometry="plane_parallel", n_layers=120, integrator="piecewise_volpath",
                                         aerosol=True, aerosol_phase="hg", sza=50.0, saa=120.0,
                                         canopy=dict(CANOPY, orientation="planophile", seed=9), sensor=VZA5),
        "canopy_perspective_inside_pp": S(geometry="plane_parallel", n_layers=100, sza=40.0, saa=200.0,
                                          canopy=dict(CANOPY, reflectance=0.45, transmittance=0.45),
                                          surface={"type": "diffuse", "reflectance": 0.15},
                                          sensor={"type": "perspective", "origin": [1.0, -9.0, 6.0], "look_at": [0.0, 0.0, 0.5],
                                                  "fov": 45.0, "film_resolution": (3, 2),
                                                  "medium": {"type": "ref", "id": "medium_atmosphere"}}),
        "canopy_perspective_above_toa_homogeneous": S(geometry="plane_parallel", atmosphere="homogeneous", toa=2000.0,
                                                      homogeneous_sigma_t=2e-4, homogeneous_albedo=0.95, sza=20.0,
                                                      canopy=CANOPY, surface={"type": "rtls"},
                                                      sensor={"type": "perspective", "origin": [0.0, -300.0, 2500.0],
                                                              "look_at": [0.0, 0.0, 0.0], "fov": 0.25, "far_clip": 1e5,
                                                              "film_resolution": (2, 2)}),
        "canopy_hdistant_maxdepth_pp": S(geometry="plane_parallel", n_layers=50, max_depth=4, rr_depth=2, sza=25.0,
                                         canopy=dict(CANOPY, reflectance=0.6, transmittance=0.3),
                                         sensor={"type": "hdistant", "film_resolution": (2, 2)}),
        "c4_canopy_afgl_rpv_reduced": _first_sensor(scenes.config_c4(spp=16, lai=2.0, radius=0.1, size=(4.0, 4.0, 1.5),
                                                                     padding=1, n_vza=6, film=(2, 2), n_layers=200)),
        # explicit rays (mradiancemeter) and images from infinity (mpdistant)
        "mradiancemeter_sky_and_nadir_spherical": S(
            n_layers=100, sza=50.0, saa=30.0, surface={"type": "diffuse", "reflectance": 0.3},
            sensor={"type": "mradiancemeter", "medium": {"type": "ref", "id": "medium_atmosphere"},
                    # ground-based sky radiance (looking up, off the sun) and an airborne nadir view at 10 km
                    "origins": [[0.0, 0.0, scenes.EARTH_RADIUS + 1.0]] * 3 + [[0.0, 2.0e4, scenes.EARTH_RADIUS + 1.0e4]],
                    "directions": [[0.0, 0.0, 1.0], [0.5, 0.0, 0.8660254], [-0.6, 0.3, 0.7416198], [0.0, 0.1, -0.9949874]]}),
        "mradiancemeter_piecewise_aerosol_pp": S(
            geometry="plane_parallel", n_layers=120, integrator="piecewise_volpath", aerosol=True, aerosol_phase="hg",
            sza=40.0, sensor={"type": "mradiancemeter", "medium": {"type": "ref", "id": "medium_atmosphere"},
                              "origins": [[0.0, 0.0, 2.0], [10.0, 0.0, 3000.0], [0.0, 5.0, 1.5e4]],
                              "directions": [[0.3, 0.2, 0.9327379], [0.0, 0.6, -0.8], [0.0, 0.0, -1.0]]}),
        "mradiancemeter_from_space_no_medium_flag_pp": S(
            geometry="plane_parallel", n_layers=60,
            sensor={"type": "mradiancemeter", "origins": [[0.0, 0.0, 2.0e5], [1.0e4, 0.0, 1.5e5]],
                    "directions": [[0.0, 0.0, -1.0], [0.5, 0.0, -0.8660254]]}),
        # polarized sky radiance seen from the ground (AERONET-like almucantar points): rayleigh_polarized +
        # polarized aerosol blend in a spherical shell -> GEN + POL + BANDS instance of the pool kernel
        "polarized_sky_from_ground_aerosol_spherical": _with_polarized_aerosol(S(
            n_layers=120, sza=55.0, saa=0.0, aerosol=True, aerosol_phase="hg", stokes=True,
            phase={"type": "rayleigh_polarized", "depolarization": 0.0279},
            surface={"type": "diffuse", "reflectance": 0.1},
            sensor={"type": "mradiancemeter", "medium": {"type": "ref", "id": "medium_atmosphere"},
                    "origins": [[0.0, 0.0, scenes.EARTH_RADIUS + 2.0]] * 3,
                    "directions": [[0.0, 0.0, 1.0], [-0.5735764, 0.0, 0.8191520], [0.4096, 0.7094, 0.5735764]]})),
        # CentralPatchSurface: RPV patch under the canopy cell, Lambertian background around it
        "central_patch_canopy_mpdistant_pp": S(
            geometry="plane_parallel", n_layers=60, sza=35.0, saa=10.0, canopy=dict(CANOPY, lai=1.0, padding=0),
            surface={"type": "diffuse", "reflectance": 0.05},
            central_patch={"edges": (4.0, 4.0), "bsdf": {"type": "rpv", "rho_0": 0.25, "k": 0.8, "g": -0.1}},
            sensor={"type": "mpdistant", "vza": 20.0, "vaa": 45.0, "film_resolution": (3, 3),
                    "target": {"type": "rectangle", "to_world": scenes.ScalarTransform4f().scale([4.0, 4.0, 1.0])}}),
        # AbstractTree instances: spherical leaf clouds on cylinder trunks with a diffuse bark
        "canopy_abstract_trees_pp": S(geometry="plane_parallel", n_layers=60, sza=40.0, saa=30.0,
                                      canopy={"trees": {}, "size": (8.0, 8.0, 4.1)},
                                      surface={"type": "diffuse", "reflectance": 0.2},
                                      sensor={"type": "mdistant", "vza": [-55.0, -20.0, 0.0, 30.0, 65.0], "vaa": 30.0}),
        # MeshTree instances (_tree.py:285-478): a smooth-shaded ellipsoidal crown (binary ply, centimetres) on a
        # hexagonal-prism trunk (obj with quads and hexagons), each element with its own bilambertian BSDF
        "canopy_mesh_trees_pp": S(geometry="plane_parallel", n_layers=60, sza=40.0, saa=30.0,
                                  canopy={"mesh_trees": {"elements": MESH_TREE}, "size": (8.0, 8.0, 4.6)},
                                  surface={"type": "diffuse", "reflectance": 0.2},
                                  sensor={"type": "mdistant", "vza": [-55.0, -20.0, 0.0, 30.0, 65.0], "vaa": 30.0}),
        # the same trees without an atmosphere, flat-shaded, plus disc leaves and a mesh leaf with its own vertex
        # normals in the group; seen by a camera
        "canopy_mesh_faceted_path_perspective": S(
            geometry="plane_parallel", atmosphere=None, integrator="path", sza=25.0, saa=200.0,
            canopy={"mesh_trees": {"elements": [dict(e, face_normals=True) for e in MESH_TREE] + [MESH_LEAF],
                                   "positions": ((0.0, 0.0), (2.5, 1.5)),
                                   "leaves": {"n": 60, "radius": 0.08, "centre": (0.0, 0.0, 3.3), "extent": (1.4, 1.2, 1.0),
                                              "reflectance": 0.4, "transmittance": 0.5}},
                    "size": (6.0, 6.0, 4.6)},
            surface={"type": "rpv", "rho_0": 0.1, "k": 0.8, "g": -0.1},
            sensor={"type": "perspective", "origin": [1.0, -9.0, 6.0], "look_at": [1.0, 0.5, 2.5], "fov": 35.0,
                    "film_resolution": (4, 3)}),
        # astroobject / multiphase-MIS in the 3D kernel's general instances: penumbrae of the leaves under a wide disc;
        # a camera under the canopy looking at the disc through the gaps (direct view, volpath.cpp:328-346); a
        # central patch under the disc; the multiphase mixture weight above a canopy
        "astro_canopy_wide_disc_pp": S(geometry="plane_parallel", n_layers=60, sza=35.0, saa=40.0, angular_diameter=8.0,
                                       canopy=dict(CANOPY, reflectance=0.5, transmittance=0.4),
                                       surface={"type": "diffuse", "reflectance": 0.3}, sensor=VZA5),
        "astro_canopy_camera_at_the_disc_pp": S(
            geometry="plane_parallel", n_layers=60, sza=30.0, saa=0.0, angular_diameter=12.0,
            canopy=dict(CANOPY, lai=1.5), surface={"type": "diffuse", "reflectance": 0.2},
            sensor={"type": "perspective", "origin": [0.9, -0.7, 0.02],
                    "look_at": [float(v) for v in np.array([0.9, -0.7, 0.02]) + scenes.angles_to_direction(30.0, 0.0)],
                    "fov": 16.0, "far_clip": 1e7, "film_resolution": (2, 2),
                    "medium": {"type": "ref", "id": "medium_atmosphere"}}),
        # ... and with the default far clip of 10 km: the clipped ray ends inside the medium, where the reference gives
        # it no throughput (transmittance_eval_pdf over an infinite distance, volpath.cpp:229-233): no disc
        # (homogeneous medium: under a heterogeneous one the reference measures the clip distance from the last NULL
        # collision instead of the camera -- ray.maxt is not shortened at :255-258 -- which the kernel does not imitate)
        "astro_camera_far_clip_in_medium_pp": S(
            geometry="plane_parallel", atmosphere="homogeneous", homogeneous_sigma_t=2e-5, homogeneous_albedo=0.95,
            sza=30.0, saa=0.0, angular_diameter=12.0,
            surface={"type": "diffuse", "reflectance": 0.2},
            sensor={"type": "perspective", "origin": [0.0, 0.0, 0.05],
                    "look_at": [float(v) for v in np.array([0.0, 0.0, 0.05]) + scenes.angles_to_direction(30.0, 0.0)],
                    "fov": 16.0, "film_resolution": (2, 2), "medium": {"type": "ref", "id": "medium_atmosphere"}}),
        "astro_central_patch_pp": S(
            geometry="plane_parallel", n_layers=60, sza=50.0, saa=10.0, angular_diameter=5.0,
            surface={"type": "diffuse", "reflectance": 0.05},
            central_patch={"edges": (4.0, 4.0), "bsdf": {"type": "rpv", "rho_0": 0.25, "k": 0.8, "g": -0.1}},
            sensor={"type": "mpdistant", "vza": 20.0, "vaa": 45.0, "film_resolution": (3, 3),
                    "target": {"type": "rectangle", "to_world": scenes.ScalarTransform4f().scale([6.0, 6.0, 1.0])}}),
        "canopy_multiphase_mis_pp": S(
            geometry="plane_parallel", atmosphere="homogeneous", homogeneous_sigma_t=0.8 / scenes.TOA,
            homogeneous_albedo=0.97, sensor=VZA5, sza=40.0, surface={"type": "diffuse", "reflectance": 0.1},
            canopy=dict(CANOPY, lai=1.0),
            phase={"type": "multiphase", "use_mis": True,
                   "phase0": {"type": "rayleigh", "depolarization": 0.25}, "weight0": 2.0,
                   "phase1": {"type": "hg", "g": 0.75}, "weight1": 1.0}),
        "mpdistant_canopy_image_pp": S(geometry="plane_parallel", n_layers=60, sza=30.0, canopy=CANOPY,
                                       sensor={"type": "mpdistant", "vza": 25.0, "vaa": 60.0, "film_resolution": (3, 2)}),
        "mpdistant_spherical": S(n_layers=100, sza=60.0, sensor={"type": "mpdistant", "vza": 40.0, "vaa": 0.0,
                                                                "film_resolution": (2, 2)}),
        "polarized_piecewise_rayleigh_pp": S(geometry="plane_parallel", integrator="piecewise_volpath",
                                             n_layers=100, sza=40.0, saa=30.0, stokes=True,
                                             phase={"type": "rayleigh_polarized", "depolarization": 0.0279},
                                             surface={"type": "diffuse", "reflectance": 0.1},
                                             sensor={"type": "mdistant", "vza": [-60.0, -30.0, 0.0, 30.0, 60.0],
                                                     "vaa": 90.0}),
    }
