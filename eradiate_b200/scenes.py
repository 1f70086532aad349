"""
Synthetic scene dictionaries restating what Eradiate's scene compiler emits for
its 1D atmosphere experiments (reference templates:
``src/eradiate/scenes/atmosphere/_core.py:640-724``, ``shapes/_sphere.py``,
``shapes/_cuboid.py:223-290``, ``shapes/_rectangle.py``, ``bsdfs/_rpv.py:104-125``,
``illumination/_directional.py``, ``measure/_core.py:218-245``,
``measure/_multi_distant.py:655-667``, ``integrators/_path_tracers.py:53-80``).

Eradiate itself cannot be imported in the build environment (no pint / xarray /
joseki / absorption databases), so the AFGL-1986-shaped radiative profile below is
synthetic and deterministic.  The *dict layout* is the
reference's; the kernel consumes it through ``mi_load_dict`` unchanged.
"""

from __future__ import annotations

import numpy as np

from .kernel._kernel_dict import (
    KernelSceneParameterFlags,
    KernelSceneParameterMap,
    SceneParameter,
    SearchSceneParameter,
)
from .kernel._scene import Emitter, Medium
from .kernel._types import ScalarTransform4f, VolumeGrid, map_cube, map_unit_cube

EARTH_RADIUS = 6378.1e3  # m, src/eradiate/constants.py:8
TOA = 120.0e3  # m, default z-grid top (src/eradiate/scenes/geometry.py:69-81)


# ------------------------------------------------------------------------------
#                         synthetic radiative profiles
# ------------------------------------------------------------------------------


def afgl_like_profile(n_layers: int = 1200, toa: float = TOA, w_nm: float = 550.0):
    """
    AFGL-1986-shaped molecular profile: Rayleigh scattering with an
    8 km scale height, lambda^-4 scaling around 550 nm (magnitude from
    ``src/eradiate/radprops/rayleigh.py:77-140`` at standard density) and an
    ozone-like Chappuis absorber centred at 22 km.  Returned float32, exactly as
    the reference stores its grids (``atmosphere/_core.py:659,674``).
    """
    dz = toa / n_layers
    z = (np.arange(n_layers) + 0.5) * dz
    sigma_s = 1.16e-5 * (550.0 / w_nm) ** 4 * np.exp(-z / 8000.0)
    chappuis = np.exp(-(((w_nm - 600.0) / 120.0) ** 2)) / np.exp(-((50.0 / 120.0) ** 2))
    sigma_a = 5.0e-7 * chappuis * np.exp(-(((z - 22000.0) / 5000.0) ** 2))
    sigma_t = sigma_s + sigma_a
    albedo = sigma_s / sigma_t
    return z, sigma_t.astype(np.float32), albedo.astype(np.float32)


def aerosol_layer(z, bottom=1000.0, top=2000.0, tau_ref=0.5, ssa=0.9):
    """Uniform particle layer (``test_cases/atmospheres.py:71-77`` shape)."""
    inside = (z >= bottom) & (z < top)
    sigma_t = np.where(inside, tau_ref / (top - bottom), 0.0)
    return sigma_t.astype(np.float64), np.full_like(sigma_t, ssa)


def hg_table(g: float = 0.7, n: int = 181):
    """HG phase function sampled on a regular cos(theta) grid, physics convention."""
    mu = np.linspace(-1.0, 1.0, n)
    p = (1.0 - g * g) / (4.0 * np.pi * (1.0 + g * g - 2.0 * g * mu) ** 1.5)
    return mu, p


# ------------------------------------------------------------------------------
#                                direction helpers
# ------------------------------------------------------------------------------


def angles_to_direction(zenith_deg, azimuth_deg):
    """Unit vector pointing *towards* (zenith, azimuth); +Z is local up."""
    th, ph = np.deg2rad(zenith_deg), np.deg2rad(azimuth_deg)
    return np.stack(
        [np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)], axis=-1
    )


def _look_at_direction(d):
    """``to_world`` mapping +Z to ``d`` (illumination/_directional.py)."""
    d = np.asarray(d, dtype=np.float64)
    up = np.array([0.0, 0.0, 1.0]) if abs(d[2]) < 0.999 else np.array([1.0, 0.0, 0.0])
    return ScalarTransform4f().look_at([0, 0, 0], d, up)


# ------------------------------------------------------------------------------
#                                  scene builder
# ------------------------------------------------------------------------------


def atmosphere_scene(
    geometry: str = "spherical_shell",
    atmosphere: str | None = "afgl",
    n_layers: int = 1200,
    aerosol: bool = False,
    aerosol_phase: str = "tabphase",
    surface: dict | None = None,
    sza: float = 30.0,
    saa: float = 0.0,
    irradiance: float = 1.8,
    sensor: dict | None = None,
    spp: int = 1024,
    integrator: str = "volpath",
    moment: bool = True,
    max_depth: int | None = None,
    rr_depth: int | None = None,
    w_nm: float = 550.0,
    planet_radius: float = EARTH_RADIUS,
    toa: float = TOA,
    homogeneous_sigma_t: float = 1.16e-5,
    homogeneous_albedo: float = 1.0,
    phase: dict | None = None,
    stokes: bool = False,
    meridian_align: bool = True,
    force_majorant: bool | None = None,
    canopy: dict | None = None,
    extra_sensors: list | None = None,
    central_patch: dict | None = None,
    angular_diameter: float | None = None,
    width: float = 1.0e9,
    hide_emitters: bool = False,
) -> dict:
    """Build the nested scene dict an ``AtmosphereExperiment`` would emit (with ``canopy``: a
    ``CanopyAtmosphereExperiment``, see :func:`disc_canopy`).

    ``force_majorant`` mirrors ``scenes/atmosphere/_core.py:354,650``: a plane-parallel
    atmosphere is emitted as a ``piecewise`` medium unless it is set (default: set for every
    integrator except ``piecewise_volpath``, so that the two media can be compared).
    """
    if force_majorant is None:
        force_majorant = integrator != "piecewise_volpath"
    if surface is None:
        surface = {"type": "rpv", "rho_0": 0.027685, "k": 0.95, "g": -0.1}  # atmospheres.py:100
    surface = _spectrumify(dict(surface))
    surface["id"] = "surface_bsdf"
    scene: dict = {"type": "scene"}

    integ: dict = {"type": integrator}
    if max_depth is not None:
        integ["max_depth"] = max_depth
    if rr_depth is not None:
        integ["rr_depth"] = rr_depth
    if hide_emitters:  # MI/src/render/integrator.cpp:29
        integ["hide_emitters"] = True
    scene["integrator"] = {"type": "moment", "nested": integ} if moment else integ
    if stokes:  # integrators/_path_tracers.py:70-78: the stokes wrapper comes last
        scene["integrator"] = {"type": "stokes", "integrator": scene["integrator"],
                               "meridian_align": meridian_align}

    sun = angles_to_direction(sza, saa)
    scene["illumination"] = {
        "type": "directional",
        "to_world": _look_at_direction(-sun),
        "irradiance": {"type": "uniform", "value": irradiance},
    }
    if angular_diameter is not None:
        # illumination/_astro_object.py:57-77: same look_at, but towards the object (direction with flip=False)
        scene["illumination"] = {
            "type": "astroobject",
            "to_world": _look_at_direction(sun),
            "angular_diameter": angular_diameter,
            "irradiance": {"type": "uniform", "value": irradiance},
        }

    spherical = geometry == "spherical_shell"
    if not spherical and geometry != "plane_parallel":
        raise ValueError(f"unknown geometry '{geometry}'")

    scene["surface_bsdf"] = surface
    # `width`: PlaneParallelGeometry.width, default 1e6 km (geometry.py:182-189)
    if spherical:
        scene["surface_shape"] = {
            "type": "sphere",
            "center": [0.0, 0.0, 0.0],
            "radius": planet_radius,
            "bsdf": {"type": "ref", "id": "surface_bsdf"},
        }
        target = [0.0, 0.0, planet_radius]
    else:
        scene["surface_shape"] = {
            "type": "rectangle",
            "to_world": ScalarTransform4f().scale([0.5 * width, 0.5 * width, 1.0]),
            "bsdf": {"type": "ref", "id": "surface_bsdf"},
        }
        target = [0.0, 0.0, 0.0]

    if central_patch is not None:
        # CentralPatchSurface (scenes/surface/_central_patch.py:185-215): `surface` is the background, the
        # patch BSDF covers `edges` around the origin; blendbsdf weighted by the 3x3 central-patch mask
        if spherical:
            raise ValueError("the central patch needs the plane-parallel geometry")
        ex, ey = central_patch["edges"]
        sx, sy = width / (3.0 * ex), width / (3.0 * ey)
        background = dict(surface)
        background.pop("id", None)
        scene["surface_bsdf"] = {
            "type": "blendbsdf", "id": "surface_bsdf",
            "bsdf_0": background,
            "bsdf_1": _spectrumify(dict(central_patch["bsdf"])),
            "weight": {
                "type": "bitmap", "filename": "texture/central_patch_surface_mask.bmp", "filter_type": "nearest",
                "to_uv": ScalarTransform4f().scale([sx, sy, 1.0]).translate([-0.5 + 0.5 / sx, -0.5 + 0.5 / sy, 0.0]),
                "wrap_mode": "clamp",
            },
        }

    if atmosphere is not None:
        if atmosphere == "afgl":
            z, sigma_t, albedo = afgl_like_profile(n_layers, toa, w_nm)
            sigma_t = sigma_t.astype(np.float64)
            albedo = albedo.astype(np.float64)
            weight = None
            if aerosol:
                st_a, al_a = aerosol_layer(z)
                ss_m, ss_a = sigma_t * albedo, st_a * al_a
                tot = sigma_t + st_a
                albedo = (ss_m + ss_a) / tot
                sigma_t = tot
                with np.errstate(invalid="ignore", divide="ignore"):
                    weight = np.where(ss_m + ss_a > 0, ss_a / (ss_m + ss_a), 0.0)
            medium_type = "heterogeneous"
        elif atmosphere == "homogeneous":
            sigma_t = np.array([homogeneous_sigma_t])
            albedo = np.array([homogeneous_albedo])
            weight = None
            medium_type = "homogeneous"
        else:
            raise ValueError(f"unknown atmosphere '{atmosphere}'")

        if phase is None:
            phase = {"type": "rayleigh"}
        if isinstance(phase.get("depolarization"), (list, tuple, np.ndarray)):
            # scenes/phase/_rayleigh.py:98-131: a per-layer depolarization factor is a volume on the medium's grid
            phase = dict(phase)
            phase["depolarization"] = _volume(np.asarray(phase["depolarization"], dtype=np.float64), spherical,
                                              planet_radius, toa, width)
        if aerosol and weight is not None:
            if aerosol_phase == "tabphase":
                _, p = hg_table()
                aer_phase = {"type": "tabphase", "values": ",".join(map(str, p))}
            elif aerosol_phase == "tabphase_irregular":
                mu = np.concatenate([np.linspace(-1, 0.5, 40), np.linspace(0.5, 1.0, 121)[1:]])
                g = 0.7
                p = (1.0 - g * g) / (4.0 * np.pi * (1.0 + g * g - 2.0 * g * mu) ** 1.5)
                aer_phase = {
                    "type": "tabphase_irregular",
                    "values": ",".join(map(str, p)),
                    "nodes": ",".join(map(str, mu)),
                }
            else:
                aer_phase = {"type": "hg", "g": 0.7}
            phase_dict = {
                "type": "blendphase",
                "phase_0": phase,
                "phase_1": aer_phase,
                "weight": _volume(weight, spherical, planet_radius, toa, width),
            }
        else:
            phase_dict = phase
        phase_dict = dict(phase_dict)
        phase_dict["id"] = "phase_atmosphere"
        scene["phase_atmosphere"] = phase_dict

        if medium_type == "homogeneous":
            medium = {
                "type": "homogeneous",
                "sigma_t": float(sigma_t[0]),
                "albedo": float(albedo[0]),
            }
        else:
            medium = {
                "type": "heterogeneous" if (spherical or force_majorant) else "piecewise",
                "sigma_t": _volume(sigma_t, spherical, planet_radius, toa, width),
                "albedo": _volume(albedo, spherical, planet_radius, toa, width),
            }
        medium["phase"] = {"type": "ref", "id": "phase_atmosphere"}
        medium["id"] = "medium_atmosphere"
        scene["medium_atmosphere"] = medium

        if spherical:
            scene["shape_atmosphere"] = {
                "type": "sphere",
                "center": [0.0, 0.0, 0.0],
                "radius": planet_radius + toa,
                "bsdf": {"type": "null"},
                "interior": {"type": "ref", "id": "medium_atmosphere"},
            }
        else:
            bottom = -0.01 * toa  # shapes/_cuboid.py:288 (1 % below the ground)
            scene["shape_atmosphere"] = {
                "type": "cube",
                "to_world": map_cube(-0.5 * width, 0.5 * width, -0.5 * width, 0.5 * width, bottom, toa),
                "bsdf": {"type": "null"},
                "interior": {"type": "ref", "id": "medium_atmosphere"},
            }

    if canopy is not None:
        if spherical:
            raise ValueError("canopies need the plane-parallel geometry")
        if "trees" in canopy:  # {"trees": {...abstract_tree_canopy arguments...}, "size": (lx, ly, lz)}
            scene.update(abstract_tree_canopy(**canopy["trees"]))
        elif "mesh_trees" in canopy:  # {"mesh_trees": {...mesh_tree_canopy arguments...}, "size": (lx, ly, lz)}
            scene.update(mesh_tree_canopy(**canopy["mesh_trees"]))
        else:
            scene.update(disc_canopy(**canopy))
        # experiments/_canopy_atmosphere.py:200-210: distant measures target the top of the unit cell
        lx, ly, lz = canopy.get("size", (10.0, 10.0, 2.0))
        target = {
            "type": "rectangle",
            "to_world": ScalarTransform4f().translate([0.0, 0.0, lz]).scale([0.5 * lx, 0.5 * ly, 1.0]),
        }
    if sensor is None:
        sensor = {"type": "mdistant", "vza": np.linspace(-75.0, 75.0, 32), "vaa": 0.0}
    scene["measure"] = _sensor_dict(dict(sensor), target, spp)
    for k, extra in enumerate(extra_sensors or []):
        extra = dict(extra)
        extra.setdefault("id", f"measure_{k + 2}")
        if extra.get("type") == "perspective" and extra.pop("inside_atmosphere", False) and atmosphere is not None:
            extra["medium"] = {"type": "ref", "id": "medium_atmosphere"}
        sid = extra["id"]
        scene[sid] = _sensor_dict(extra, target, spp)
    return scene


def leaf_normals(n: int, orientation, rng) -> np.ndarray:
    """Leaf normals: 'uniform' (spherical leaf angle distribution), 'planophile' (all +z) or a vector."""
    if isinstance(orientation, str) and orientation == "uniform":
        mu = rng.uniform(-1.0, 1.0, n)
        phi = rng.uniform(0.0, 2.0 * np.pi, n)
        st = np.sqrt(1.0 - mu * mu)
        return np.stack([st * np.cos(phi), st * np.sin(phi), mu], axis=1)
    if isinstance(orientation, str) and orientation == "planophile":
        return np.tile([0.0, 0.0, 1.0], (n, 1))
    v = np.asarray(orientation, dtype=np.float64)
    return np.tile(v / np.linalg.norm(v), (n, 1))


def disc_canopy(
    lai: float = 3.0,
    radius: float = 0.1,
    size=(10.0, 10.0, 2.0),
    padding: int = 0,
    orientation="uniform",
    reflectance: float = 0.5,
    transmittance: float = 0.4,
    seed: int = 1,
    n_leaves: int | None = None,
    z_bottom: float = 0.0,
) -> dict:
    """
    Homogeneous disc canopy (RAMI "HOM" scenes) exactly as ``DiscreteCanopy.homogeneous`` +
    ``InstancedCanopyElement`` emit it (``scenes/biosphere/_leaf_cloud.py:1150-1175``,
    ``_core.py:266-296``, ``_discrete.py:139-200``): one ``bilambertian`` BSDF, a ``shapegroup`` of
    ``disk`` leaves (to_world = look_at x scale), and (2*padding+1)^2 translated ``instance``s.
    The leaf count follows the leaf area index: n = LAI * lx * ly / (pi r^2).
    """
    lx, ly, lz = (float(v) for v in size)
    if n_leaves is None:
        n_leaves = max(1, int(round(lai * lx * ly / (np.pi * radius * radius))))
    rng = np.random.default_rng(seed)
    pos = np.stack([rng.uniform(-0.5 * lx, 0.5 * lx, n_leaves), rng.uniform(-0.5 * ly, 0.5 * ly, n_leaves),
                    z_bottom + rng.uniform(0.0, lz, n_leaves)], axis=1)
    nrm = leaf_normals(n_leaves, orientation, rng)
    out: dict = {
        "bsdf_leaf_cloud": {
            "type": "bilambertian",
            "reflectance": {"type": "uniform", "value": float(reflectance)},
            "transmittance": {"type": "uniform", "value": float(transmittance)},
        }
    }
    group: dict = {"type": "shapegroup"}
    for i in range(n_leaves):
        n = nrm[i]
        up = np.array([1.0, 0.0, 0.0]) if abs(n[2]) > 0.9 else np.array([0.0, 0.0, 1.0])
        group[f"leaf_cloud_leaf_{i}"] = {
            "type": "disk",
            "bsdf": {"type": "ref", "id": "bsdf_leaf_cloud"},
            "to_world": ScalarTransform4f().look_at(origin=pos[i], target=pos[i] + n, up=up).scale(radius),
        }
    out["leaf_cloud"] = group
    k = 0
    for ix in range(-padding, padding + 1):
        for iy in range(-padding, padding + 1):
            out[f"leaf_cloud_instance_{k}"] = {
                "type": "instance",
                "group": {"type": "ref", "id": "leaf_cloud"},
                "to_world": ScalarTransform4f().translate([ix * lx, iy * ly, 0.0]),
            }
            k += 1
    return out


def _spectrumify(bsdf: dict) -> dict:
    """Wrap scalar BSDF parameters as uniform spectra, as Eradiate does."""
    spectral = {
        "diffuse": ("reflectance",),
        "rpv": ("rho_0", "k", "g", "rho_c"),
        "rtls": ("f_iso", "f_vol", "f_geo"),
        "hapke": ("w", "b", "c", "theta", "B_0", "h"),
    }.get(bsdf.get("type"), ())
    for k in spectral:
        if k in bsdf and not isinstance(bsdf[k], dict):
            bsdf[k] = {"type": "uniform", "value": float(bsdf[k])}
    return bsdf


def _volume(values, spherical: bool, planet_radius: float, toa: float, width: float) -> dict:
    values = np.asarray(values, dtype=np.float64)
    if spherical:
        rtoa = planet_radius + toa
        return {
            "type": "sphericalcoordsvolume",
            "volume": {
                "type": "gridvolume",
                "grid": VolumeGrid(np.reshape(values, (1, 1, -1)).astype(np.float32)),
                "filter_type": "nearest",
            },
            "to_world": map_cube(-rtoa, rtoa, -rtoa, rtoa, -rtoa, rtoa),
            "rmin": planet_radius / rtoa,
        }
    return {
        "type": "gridvolume",
        "grid": VolumeGrid(np.reshape(values, (-1, 1, 1)).astype(np.float32)),
        "to_world": map_unit_cube(-0.5 * width, 0.5 * width, -0.5 * width, 0.5 * width, 0.0, toa),
        "filter_type": "nearest",
    }


def _sensor_dict(sensor: dict, target, spp: int) -> dict:
    ty = sensor.pop("type")
    out: dict = {"type": ty, "id": sensor.pop("id", "measure")}
    if ty == "mdistant":
        vza = np.atleast_1d(np.asarray(sensor.pop("vza"), dtype=np.float64))
        vaa = np.broadcast_to(np.asarray(sensor.pop("vaa", 0.0), dtype=np.float64), vza.shape)
        # hplane layout: negative zeniths point to azimuth + 180 deg
        az = np.where(vza < 0, vaa + 180.0, vaa)
        view = angles_to_direction(np.abs(vza), az)
        out["directions"] = ",".join(map(str, (-view).ravel(order="C")))
        width, height = vza.size, 1
    elif ty == "mradiancemeter":  # scenes/measure/_multi_radiancemeter.py:87-96
        org = np.asarray(sensor.pop("origins"), dtype=np.float64).reshape(-1, 3)
        dirs = np.asarray(sensor.pop("directions"), dtype=np.float64).reshape(-1, 3)
        out["origins"] = ",".join(map(str, org.ravel(order="C")))
        out["directions"] = ",".join(map(str, dirs.ravel(order="C")))
        if "medium" in sensor:
            out["medium"] = sensor.pop("medium")
        width, height = org.shape[0], 1
        target = None
    elif ty == "mpdistant":  # scenes/measure/_distant.py (MultiPixelDistantMeasure): one direction, an image
        res = sensor.pop("film_resolution", (8, 8))
        width, height = int(res[0]), int(res[1])
        view = angles_to_direction(float(sensor.pop("vza", 0.0)), float(sensor.pop("vaa", 0.0)))
        out["direction"] = [float(v) for v in -np.asarray(view).ravel()]
    elif ty == "perspective":  # scenes/measure/_perspective.py:150-165
        res = sensor.pop("film_resolution", (32, 32))
        width, height = int(res[0]), int(res[1])
        out["fov"] = float(sensor.pop("fov", 50.0))
        out["far_clip"] = float(sensor.pop("far_clip", 1e4))
        out["to_world"] = ScalarTransform4f().look_at(
            origin=sensor.pop("origin"), target=sensor.pop("look_at", [0.0, 0.0, 0.0]),
            up=sensor.pop("up", [0.0, 0.0, 1.0]))
        if "medium" in sensor:
            out["medium"] = sensor.pop("medium")
        target = None
    else:
        res = sensor.pop("film_resolution", (32, 32))
        width, height = int(res[0]), int(res[1])
        if "to_world" in sensor:
            out["to_world"] = sensor.pop("to_world")
    tgt = sensor.pop("target", target)
    if tgt is not None and ty not in ("perspective", "mradiancemeter"):
        out["target"] = tgt
    if "ray_offset" in sensor:
        out["ray_offset"] = sensor.pop("ray_offset")
    if "medium" in sensor:
        out["medium"] = sensor.pop("medium")
    out["film"] = {
        "type": "hdrfilm",
        "width": width,
        "height": height,
        "pixel_format": "luminance",
        "component_format": "float32",
        "rfilter": {"type": "box"},
    }
    out["sampler"] = {"type": "independent", "sample_count": int(spp)}
    return out


# ------------------------------------------------------------------------------
#                         BASELINE.json configurations
# ------------------------------------------------------------------------------


def config_c1(spp: int = 4096) -> dict:
    """C1: homogeneous molecular atmosphere, Lambertian rho=0.5, plane-parallel, 1 angle."""
    return atmosphere_scene(
        geometry="plane_parallel",
        atmosphere="homogeneous",
        surface={"type": "diffuse", "reflectance": 0.5},
        sensor={"type": "mdistant", "vza": [0.0], "vaa": 0.0},
        spp=spp,
    )


def config_c2(spp: int = 1 << 20, n_vza: int = 32) -> dict:
    """C2: AFGL1986-shaped molecular atmosphere + RPV, spherical shell, mdistant 32 VZA."""
    return atmosphere_scene(
        geometry="spherical_shell",
        atmosphere="afgl",
        sensor={"type": "mdistant", "vza": np.linspace(-75.0, 75.0, n_vza), "vaa": 0.0},
        spp=spp,
    )


def config_c3(spp: int = 1 << 22, res: int = 32, w_nm: float = 865.0) -> dict:
    """C3: AFGL + aerosol layer (tab_phase), hdistant hemispherical film."""
    return atmosphere_scene(
        geometry="spherical_shell",
        atmosphere="afgl",
        aerosol=True,
        w_nm=w_nm,
        sensor={"type": "hdistant", "film_resolution": (res, res)},
        spp=spp,
    )


def abstract_tree_canopy(
    positions=((0.0, 0.0), (3.0, 1.0), (-2.0, 2.5)),
    trunk_height: float = 2.0,
    trunk_radius: float = 0.1,
    crown_radius: float = 1.0,
    n_leaves: int = 300,
    leaf_radius: float = 0.06,
    reflectance: float = 0.45,
    transmittance: float = 0.45,
    trunk_reflectance: float = 0.3,
    seed: int = 5,
) -> dict:
    """
    Instanced `AbstractTree`s (``scenes/biosphere/_tree.py:150-180``): a shape group made of a spherical leaf
    cloud (`disk` leaves, `bilambertian`), a trunk `cylinder` from z = -0.1 to the crown base and its cap `disk`
    (one `diffuse` BSDF), placed by translated `instance`s (``_core.py:266-296``).
    """
    rng = np.random.default_rng(seed)
    v = rng.normal(size=(n_leaves, 3))
    v *= (crown_radius * rng.uniform(0.0, 1.0, (n_leaves, 1)) ** (1.0 / 3.0)) / np.linalg.norm(v, axis=1, keepdims=True)
    pos = v + np.array([0.0, 0.0, trunk_height + crown_radius])
    nrm = leaf_normals(n_leaves, "uniform", rng)
    out: dict = {
        "bsdf_leaf_cloud": {"type": "bilambertian", "reflectance": {"type": "uniform", "value": float(reflectance)},
                            "transmittance": {"type": "uniform", "value": float(transmittance)}},
        "bsdf_tree": {"type": "diffuse", "reflectance": {"type": "uniform", "value": float(trunk_reflectance)}},
    }
    group: dict = {"type": "shapegroup"}
    for i in range(n_leaves):
        n = nrm[i]
        up = np.array([1.0, 0.0, 0.0]) if abs(n[2]) > 0.9 else np.array([0.0, 0.0, 1.0])
        group[f"leaf_cloud_leaf_{i}"] = {
            "type": "disk", "bsdf": {"type": "ref", "id": "bsdf_leaf_cloud"},
            "to_world": ScalarTransform4f().look_at(origin=pos[i], target=pos[i] + n, up=up).scale(leaf_radius),
        }
    group["trunk_cyl_tree"] = {"type": "cylinder", "bsdf": {"type": "ref", "id": "bsdf_tree"}, "radius": trunk_radius,
                               "p0": [0.0, 0.0, -0.1], "p1": [0.0, 0.0, trunk_height]}
    group["trunk_cap_tree"] = {"type": "disk", "bsdf": {"type": "ref", "id": "bsdf_tree"},
                               "to_world": ScalarTransform4f().scale(trunk_radius).translate([0.0, 0.0, trunk_height / trunk_radius])}
    out["tree"] = group
    for k, (x, y) in enumerate(positions):
        out[f"tree_instance_{k}"] = {"type": "instance", "group": {"type": "ref", "id": "tree"},
                                     "to_world": ScalarTransform4f().translate([float(x), float(y), 0.0])}
    return out


def mesh_tree_canopy(
    elements,
    positions=((0.0, 0.0), (3.0, 1.0), (-2.0, 2.5)),
    leaves: dict | None = None,
) -> dict:
    """
    Instanced `MeshTree`s (``scenes/biosphere/_tree.py:285-478, :600-700``): a shape group of triangle meshes,
    each a `ply` / `obj` file with its own `bilambertian` BSDF and a scaling `to_world` (mesh units -> metres),
    placed by translated `instance`s.  `elements`: [{"id", "filename", "scale", "reflectance", "transmittance",
    optional "face_normals"}]; `leaves`: optional disc leaves in the same group ({"n", "radius", "centre", "extent",
    "reflectance", "transmittance", "seed"}).
    """
    out: dict = {}
    group: dict = {"type": "shapegroup"}
    for e in elements:
        ext = str(e["filename"]).rsplit(".", 1)[-1].lower()
        if ext not in ("ply", "obj"):
            raise ValueError(f"unsupported file extension '.{ext}'")
        out[f"bsdf_{e['id']}"] = {"type": "bilambertian",
                                  "reflectance": {"type": "uniform", "value": float(e["reflectance"])},
                                  "transmittance": {"type": "uniform", "value": float(e["transmittance"])}}
        shape = {"type": ext, "bsdf": {"type": "ref", "id": f"bsdf_{e['id']}"}, "filename": str(e["filename"]),
                 "to_world": ScalarTransform4f().scale(float(e.get("scale", 1.0)))}
        if "face_normals" in e:
            shape["face_normals"] = bool(e["face_normals"])
        group[e["id"]] = shape
    if leaves:
        rng = np.random.default_rng(leaves.get("seed", 3))
        n = int(leaves["n"])
        pos = np.asarray(leaves["centre"]) + (rng.uniform(-1.0, 1.0, (n, 3)) * np.asarray(leaves["extent"]))
        nrm = leaf_normals(n, "uniform", rng)
        out["bsdf_leaf_cloud"] = {"type": "bilambertian",
                                  "reflectance": {"type": "uniform", "value": float(leaves["reflectance"])},
                                  "transmittance": {"type": "uniform", "value": float(leaves["transmittance"])}}
        for i in range(n):
            up = np.array([1.0, 0.0, 0.0]) if abs(nrm[i][2]) > 0.9 else np.array([0.0, 0.0, 1.0])
            group[f"leaf_cloud_leaf_{i}"] = {
                "type": "disk", "bsdf": {"type": "ref", "id": "bsdf_leaf_cloud"},
                "to_world": ScalarTransform4f().look_at(origin=pos[i], target=pos[i] + nrm[i], up=up).scale(float(leaves["radius"])),
            }
    out["mesh_tree"] = group
    for k, (x, y) in enumerate(positions):
        out[f"mesh_tree_instance_{k}"] = {"type": "instance", "group": {"type": "ref", "id": "mesh_tree"},
                                          "to_world": ScalarTransform4f().translate([float(x), float(y), 0.0])}
    return out


def config_c4(spp: int = 1 << 22, lai: float = 3.0, radius: float = 0.1, size=(25.0, 25.0, 2.0),
              padding: int = 2, n_vza: int = 32, film=(64, 64), n_layers: int = 1200) -> dict:
    """C4: RAMI-style homogeneous disc canopy under the AFGL1986-shaped atmosphere, RPV floor,
    mono 670 nm, plane-parallel; mdistant (principal plane) + a perspective camera above the canopy."""
    return atmosphere_scene(
        geometry="plane_parallel", atmosphere="afgl", n_layers=n_layers, w_nm=670.0,
        canopy={"lai": lai, "radius": radius, "size": size, "padding": padding,
                "reflectance": 0.0546, "transmittance": 0.0149},  # RAMI HOM red-band leaf optics
        sensor={"type": "mdistant", "vza": np.linspace(-75.0, 75.0, n_vza), "vaa": 0.0},
        extra_sensors=[{"type": "perspective", "origin": [0.0, -1.5 * size[1], 0.75 * size[1]],
                        "look_at": [0.0, 0.0, 0.5 * size[2]], "fov": 40.0, "film_resolution": film,
                        "inside_atmosphere": True}],
        spp=spp,
    )


def polarized_aerosol_table(g: float = 0.65, n_back: int = 33, n_fwd: int = 41) -> dict:
    """Synthetic `tabphase_polarized` aerosol (HG-shaped m11 on irregular nodes, Rayleigh-like
    linear polarisation, damped; ERP/phase/tabphase_polarized.cpp:228-296) as comma-joined strings."""
    mu = np.concatenate([np.linspace(-1, 0.6, n_back), np.linspace(0.6, 1.0, n_fwd)[1:]])
    m11 = (1.0 - g * g) / (4.0 * np.pi * (1.0 + g * g - 2.0 * g * mu) ** 1.5)
    pol = -0.4 * (1 - mu**2) / (1 + mu**2)
    fmt = lambda a: ",".join(map(str, a))  # noqa: E731
    return {
        "type": "tabphase_polarized", "nodes": fmt(mu), "m11": fmt(m11), "m12": fmt(pol * m11),
        "m22": fmt(0.9 * m11), "m33": fmt(0.8 * mu * m11), "m34": fmt(0.1 * (1 - mu**2) * m11),
        "m44": fmt(0.7 * mu * m11),
    }


def config_c5(spp: int = 1 << 20, n_vza: int = 1, w_nm: float = 550.0, n_layers: int = 1200,
              geometry: str = "spherical_shell") -> dict:
    """C5: polarized (Stokes) transport, `ocean_legacy` surface under the AFGL1986-shaped molecular
    atmosphere (`rayleigh_polarized`) + aerosol layer (`tabphase_polarized`); one context of the
    400-1000 nm sweep (see :func:`spectral_update_map_c5`)."""
    vza = np.linspace(-60.0, 60.0, n_vza) if n_vza > 1 else np.array([30.0])
    sensor = {"type": "mdistant", "vza": vza, "vaa": 40.0}
    if geometry == "spherical_shell":
        # Keep the target away from (0, 0, R): that is the pole of the sphere's (u, v) parametrisation,
        # where dp_du = (-y, x, 0) vanishes (MI/src/shapes/sphere.cpp:703-720) and the shading frame an
        # anisotropic BSDF (the ocean's wind direction) refers to is decided by rounding noise.
        y = 3.0e5
        sensor["target"] = [0.0, y, float(np.sqrt(EARTH_RADIUS**2 - y * y))]
    d = atmosphere_scene(
        geometry=geometry, atmosphere="afgl", n_layers=n_layers, aerosol=True, aerosol_phase="hg", w_nm=w_nm,
        phase={"type": "rayleigh_polarized", "depolarization": 0.0279}, stokes=True, sza=35.0, saa=0.0,
        surface={"type": "ocean_legacy", "wavelength": w_nm, "wind_speed": 5.0, "wind_direction": 30.0,
                 "chlorinity": 19.0, "pigmentation": 0.3, "shadowing": True},
        sensor=sensor, spp=spp)
    d["phase_atmosphere"]["phase_1"] = polarized_aerosol_table()
    return d


def spectral_update_map_c5(n_layers: int = 1200, spherical: bool = True) -> KernelSceneParameterMap:
    """C5 sweep: per-context sigma_t / albedo / blend weight / irradiance and the ocean BSDF wavelength
    (`scenes/bsdfs/_ocean_legacy.py`: `wavelength` is a spectral scene parameter)."""
    from .kernel._scene import BSDF, PhaseFunction

    rel = "volume.data" if spherical else "data"

    def mix(ctx):
        z, st_m, al_m = afgl_like_profile(n_layers, TOA, ctx.si.w)
        st_a, al_a = aerosol_layer(z)
        ss_m, ss_a = st_m.astype(np.float64) * al_m, st_a * al_a
        tot = st_m + st_a
        with np.errstate(invalid="ignore", divide="ignore"):
            w = np.where(ss_m + ss_a > 0, ss_a / (ss_m + ss_a), 0.0)
        return tot.astype(np.float32), ((ss_m + ss_a) / tot).astype(np.float32), w.astype(np.float32)

    umap = spectral_update_map(n_layers, spherical)
    umap.data["medium_atmosphere.sigma_t"] = SceneParameter(
        lambda ctx: mix(ctx)[0], KernelSceneParameterFlags.SPECTRAL,
        search=SearchSceneParameter(Medium, "medium_atmosphere", f"sigma_t.{rel}"))
    umap.data["medium_atmosphere.albedo"] = SceneParameter(
        lambda ctx: mix(ctx)[1], KernelSceneParameterFlags.SPECTRAL,
        search=SearchSceneParameter(Medium, "medium_atmosphere", f"albedo.{rel}"))
    umap.data["phase_atmosphere.weight"] = SceneParameter(
        lambda ctx: mix(ctx)[2], KernelSceneParameterFlags.SPECTRAL,
        search=SearchSceneParameter(PhaseFunction, "phase_atmosphere", f"weight.{rel}"))
    umap.data["surface_bsdf.wavelength"] = SceneParameter(
        lambda ctx: float(ctx.si.w), KernelSceneParameterFlags.SPECTRAL,
        search=SearchSceneParameter(BSDF, "surface_bsdf", "wavelength"))
    return umap


def spectral_update_map(n_layers: int = 1200, spherical: bool = True) -> KernelSceneParameterMap:
    """
    Update map equivalent to ``atmosphere/_core.py:777-805`` +
    ``illumination/_directional.py``: per-context sigma_t / albedo / irradiance.
    """
    rel = "volume.data" if spherical else "data"

    def sigma_t(ctx):
        return afgl_like_profile(n_layers, TOA, ctx.si.w)[1]

    def albedo(ctx):
        return afgl_like_profile(n_layers, TOA, ctx.si.w)[2]

    return KernelSceneParameterMap(
        {
            "medium_atmosphere.sigma_t": SceneParameter(
                sigma_t,
                KernelSceneParameterFlags.SPECTRAL,
                search=SearchSceneParameter(Medium, "medium_atmosphere", f"sigma_t.{rel}"),
            ),
            "medium_atmosphere.albedo": SceneParameter(
                albedo,
                KernelSceneParameterFlags.SPECTRAL,
                search=SearchSceneParameter(Medium, "medium_atmosphere", f"albedo.{rel}"),
            ),
            "illumination.irradiance.value": SceneParameter(
                lambda ctx: 1.8 * (550.0 / ctx.si.w),
                KernelSceneParameterFlags.SPECTRAL,
                search=SearchSceneParameter(Emitter, "illumination", "irradiance.value"),
            ),
        }
    )
