"""
ctypes mirror of ``include/eradiate_b200.h`` (the C ABI of the CUDA library).

Field order and types must match the header exactly; ``tests/test_abi.py``
checks ``sizeof`` and ``ERTB_ABI_VERSION`` against the compiled library.
"""

from __future__ import annotations

import ctypes as C

ABI_VERSION = 16
MAX_PHASE = 4
MAX_BSDF_PARAMS = 16
MAX_LAYERS = 4096
MAX_PHASE_NODES = 2048

# enum ertb_geometry
GEOM_PLANE_PARALLEL = 0
GEOM_SPHERICAL_SHELL = 1

# enum ertb_bsdf_type
BSDF_DIFFUSE = 0
BSDF_RPV = 1
BSDF_RTLS = 2
BSDF_HAPKE = 3
BSDF_OCEAN_LEGACY = 4
BSDF_BLACK = 5
BSDF_OCEAN_MISHCHENKO = 6
BSDF_OCEAN_GRASP = 7
BSDF_MAIGNAN = 8
BSDF_MQDIFFUSE = 9
BSDF_MEASURED_MONO = 10

# enum ertb_phase_type
PHASE_ISOTROPIC = 0
PHASE_RAYLEIGH = 1
PHASE_HG = 2
PHASE_TABULATED = 3
PHASE_TABULATED_IRREGULAR = 4
PHASE_RAYLEIGH_POLARIZED = 5
PHASE_TABULATED_POLARIZED = 6

# enum ertb_sensor_type
SENSOR_MDISTANT = 0
SENSOR_HDISTANT = 1
SENSOR_DISTANTFLUX = 2
SENSOR_PERSPECTIVE = 3
SENSOR_MPDISTANT = 4
SENSOR_MRADIANCEMETER = 5

# enum ertb_target_type
TARGET_NONE = 0
TARGET_POINT = 1
TARGET_RECTANGLE = 2
TARGET_DISK = 3

# enum ertb_integrator_type
INTEGRATOR_VOLPATH = 0
INTEGRATOR_VOLPATHMIS = 1
INTEGRATOR_PIECEWISE_VOLPATH = 2

# enum ertb_param
PARAM_SIGMA_T = 0
PARAM_ALBEDO = 1
PARAM_PHASE_WEIGHT = 2
PARAM_PHASE_VALUES = 3
PARAM_BSDF_PARAMS = 4
PARAM_IRRADIANCE = 5
PARAM_PHASE_PARAMS = 6
PARAM_PHASE_MUELLER = 7
PARAM_LEAF_BSDF = 8
PARAM_PATCH_BSDF_PARAMS = 9
PARAM_TRUNK_BSDF = 10
PARAM_MESH_BSDF = 11

c_float_p = C.POINTER(C.c_float)
c_double_p = C.POINTER(C.c_double)


class PhaseDesc(C.Structure):
    _fields_ = [
        ("type", C.c_int32),
        ("n_nodes", C.c_int32),
        ("params", C.c_float * 4),
        ("values", c_float_p),
        ("nodes", c_float_p),
        ("mueller", c_float_p * 5),
    ]


class SensorDesc(C.Structure):
    _fields_ = [
        ("type", C.c_int32),
        ("width", C.c_int32),
        ("height", C.c_int32),
        ("n_directions", C.c_int32),
        ("directions", c_double_p),
        ("to_world", C.c_double * 16),
        ("target_type", C.c_int32),
        ("_pad0", C.c_int32),
        ("target", C.c_double * 3),
        ("target_to_world", C.c_double * 16),
        ("ray_offset", C.c_double),
        ("x_fov_deg", C.c_double),
        ("near_clip", C.c_double),
        ("far_clip", C.c_double),
        ("in_medium", C.c_int32),
        ("_pad1", C.c_int32),
        ("origins", c_double_p),
    ]


class LeafGroupDesc(C.Structure):
    _fields_ = [
        ("n_disks", C.c_int32),
        ("reflectance", C.c_float),
        ("transmittance", C.c_float),
        ("_pad", C.c_int32),
        ("disks", c_float_p),
        ("n_cylinders", C.c_int32),
        ("n_trunk_disks", C.c_int32),
        ("cylinders", c_float_p),
        ("trunk_disks", c_float_p),
        ("trunk_reflectance", C.c_float),
        ("n_triangles", C.c_int32),
        ("n_mesh_bsdfs", C.c_int32),
        ("_pad2", C.c_int32),
        ("triangles", c_float_p),
        ("triangle_bsdf", C.POINTER(C.c_int32)),
        ("mesh_bsdfs", c_float_p),
    ]


class SceneDesc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("geometry", C.c_int32),
        ("surface_z", C.c_double),
        ("medium_bottom", C.c_double),
        ("medium_top", C.c_double),
        ("bsphere_center", C.c_double * 3),
        ("bsphere_radius", C.c_double),
        ("has_medium", C.c_int32),
        ("n_layers", C.c_int32),
        ("sigma_t", c_float_p),
        ("albedo", c_float_p),
        ("sigma_t_scale", C.c_float),
        ("homogeneous", C.c_int32),
        ("n_phase", C.c_int32),
        ("_pad1", C.c_int32),
        ("phase", PhaseDesc * MAX_PHASE),
        ("phase_weight", c_float_p),
        ("bsdf_type", C.c_int32),
        ("_pad2", C.c_int32),
        ("bsdf_params", C.c_float * MAX_BSDF_PARAMS),
        ("emitter_direction", C.c_double * 3),
        ("irradiance", C.c_float),
        ("_pad3", C.c_int32),
        ("integrator", C.c_int32),
        ("rr_depth", C.c_int32),
        ("max_depth", C.c_int64),
        ("polarized", C.c_int32),
        ("meridian_align", C.c_int32),
        ("n_sensors", C.c_int32),
        ("_pad4", C.c_int32),
        ("sensors", C.POINTER(SensorDesc)),
        ("n_leaf_groups", C.c_int32),
        ("n_instances", C.c_int32),
        ("leaf_groups", C.POINTER(LeafGroupDesc)),
        ("instance_group", C.POINTER(C.c_int32)),
        ("instance_offset", c_double_p),
        ("has_patch", C.c_int32),
        ("patch_bsdf_type", C.c_int32),
        ("patch_bsdf_params", C.c_float * MAX_BSDF_PARAMS),
        ("patch_rect", C.c_double * 4),
        ("bsdf_table", c_float_p),
        ("bsdf_table_res", C.c_int32 * 3),
        ("phase_mis", C.c_int32),
        ("emitter_angular_diameter", C.c_double),
        ("hide_emitters", C.c_int32),
        ("_pad_tail", C.c_int32),
    ]


class RenderStats(C.Structure):
    _fields_ = [
        ("n_paths", C.c_uint64),
        ("trips_main", C.c_uint64),
        ("trips_nee", C.c_uint64),
        ("n_scatter", C.c_uint64),
        ("n_surface", C.c_uint64),
        ("device_ms", C.c_double),
        ("n_launches", C.c_int32),
        ("n_bands", C.c_int32),
    ]

    def as_dict(self) -> dict:
        return {
            k: getattr(self, k)
            for k in (
                "n_paths",
                "trips_main",
                "trips_nee",
                "n_scatter",
                "n_surface",
                "device_ms",
                "n_launches",
                "n_bands",
            )
        }


#: every symbol ``include/eradiate_b200.h`` declares; checked by tests/test_abi.py
EXPORTED_SYMBOLS = (
    "ertb_abi_version",
    "ertb_last_error",
    "ertb_device_count",
    "ertb_scene_create",
    "ertb_scene_destroy",
    "ertb_scene_update",
    "ertb_render",
    "ertb_render_device",
    "ertb_render_stokes",
    "ertb_sensor_pixel_count",
    "ertb_batch_begin",
    "ertb_batch_push",
    "ertb_batch_end",
    "ertb_kat_bsdf_eval",
    "ertb_kat_bsdf_sample",
    "ertb_kat_phase_eval",
    "ertb_kat_phase_sample",
    "ertb_kat_phase_mueller",
    "ertb_kat_bsdf_mueller",
    "ertb_kat_piecewise_sample",
    "ertb_kat_piecewise_transmittance",
    "ertb_kat_sensor_ray",
    "ertb_kat_canopy_intersect",
    "ertb_kat_leaf_bsdf_eval",
    "ertb_kat_leaf_bsdf_sample",
)
