"""
ctypes loader of the CUDA library (``eradiate_b200/csrc/libertb_cuda.so``).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).
There is no CPU fallback: a missing library or a machine without a usable CUDA
device raises ``RuntimeError`` as soon as a scene is created.
"""

from __future__ import annotations

import ctypes as C
import os

from . import _abi

_LIB_PATH = os.environ.get(  # ERTB_LIB: developer knob for A/B runs of two builds of the same ABI
    "ERTB_LIB", os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libertb_cuda.so"))
_lib = None


def lib_path() -> str:
    return _LIB_PATH


def load() -> C.CDLL:
    """Load (once) and annotate the C ABI. Raises RuntimeError if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise RuntimeError(
            f"CUDA library not found at {_LIB_PATH}; build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback)"
        )
    lib = C.CDLL(_LIB_PATH)
    vp, u64, i32 = C.c_void_p, C.c_uint64, C.c_int
    fp, dp = _abi.c_float_p, _abi.c_double_p
    lib.ertb_abi_version.restype = i32
    lib.ertb_last_error.restype = C.c_char_p
    lib.ertb_device_count.restype = i32
    lib.ertb_scene_create.argtypes = [C.POINTER(_abi.SceneDesc), i32, C.POINTER(vp)]
    lib.ertb_scene_destroy.argtypes = [vp]
    lib.ertb_scene_destroy.restype = None
    lib.ertb_scene_update.argtypes = [vp, i32, i32, fp, C.c_size_t]
    lib.ertb_render.argtypes = [vp, i32, u64, u64, u64, dp, dp, dp, C.POINTER(_abi.RenderStats)]
    lib.ertb_render_device.argtypes = [vp, i32, u64, u64, u64, vp, vp, vp]
    lib.ertb_render_stokes.argtypes = [vp, i32, u64, u64, u64, dp, dp, dp, dp, C.POINTER(_abi.RenderStats)]
    lib.ertb_kat_phase_mueller.argtypes = [vp, i32, C.c_size_t, fp, fp, fp, fp]
    lib.ertb_kat_bsdf_mueller.argtypes = [vp, C.c_size_t, fp, fp, fp]
    lib.ertb_sensor_pixel_count.argtypes = [vp, i32]
    lib.ertb_batch_begin.argtypes = [vp, i32, C.POINTER(i32), i32]
    lib.ertb_batch_push.argtypes = [vp, i32, u64, u64, u64]
    lib.ertb_batch_end.argtypes = [vp, dp, C.c_size_t, C.POINTER(_abi.RenderStats), dp]
    lib.ertb_kat_bsdf_eval.argtypes = [vp, C.c_size_t, fp, fp, fp]
    lib.ertb_kat_bsdf_sample.argtypes = [vp, C.c_size_t, fp, fp, fp, fp]
    lib.ertb_kat_phase_eval.argtypes = [vp, i32, C.c_size_t, fp, fp]
    lib.ertb_kat_phase_sample.argtypes = [vp, i32, C.c_size_t, fp, fp, fp, fp]
    lib.ertb_kat_sensor_ray.argtypes = [vp, i32, C.c_size_t, fp, fp, dp, dp, fp]
    lib.ertb_kat_piecewise_sample.argtypes = [vp, C.c_size_t, fp, fp, fp, fp, C.POINTER(i32)]
    lib.ertb_kat_piecewise_transmittance.argtypes = [vp, C.c_size_t, fp, fp, fp]
    lib.ertb_kat_canopy_intersect.argtypes = [vp, C.c_size_t, dp, fp, fp, dp, fp, C.POINTER(i32)]
    lib.ertb_kat_leaf_bsdf_eval.argtypes = [vp, i32, C.c_size_t, fp, fp, fp]
    lib.ertb_kat_leaf_bsdf_sample.argtypes = [vp, i32, C.c_size_t, fp, fp, fp, fp]
    for name in _abi.EXPORTED_SYMBOLS:
        fn = getattr(lib, name)
        if fn.restype is C.c_int and name not in ("ertb_abi_version", "ertb_device_count"):
            fn.restype = i32
    if lib.ertb_abi_version() != _abi.ABI_VERSION:
        raise RuntimeError(
            f"ABI mismatch: library {lib.ertb_abi_version()} != python {_abi.ABI_VERSION}; rebuild"
        )
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != 0:
        msg = load().ertb_last_error()
        raise RuntimeError(msg.decode("utf-8", "replace") if msg else f"ertb error {status}")
