"""
Loader of THE REFERENCE ITSELF: the Mitsuba 3 fork + Dr.Jit + Eradiate plugins compiled from
/root/reference/ext/mitsuba by ``oracle/build_ref.sh`` into ``oracle/_ref`` (git-ignored binaries).

TEST INFRASTRUCTURE ONLY, like the rest of ``oracle/``: imported by ``tools/make_reference_golden.py``
(fixtures), by tests that compare against the reference when it is present, and by the
``--impl reference`` / ``cpu_baseline`` legs of ``bench.py``.  Nothing under ``eradiate_b200/`` imports it.

The binaries keep the absolute rpaths of their build tree, so the shared libraries are pre-loaded by
path (RTLD_GLOBAL) in dependency order before ``drjit`` / ``mitsuba`` are imported: the dynamic linker
then resolves every DT_NEEDED entry against the sonames already in the process.
"""

from __future__ import annotations

import ctypes
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
_PRELOAD = [
    "libnanothread.so", "libdrjit-core.so", "libdrjit-extra.so",
    "libHalf-mitsuba.so", "libIex-mitsuba.so", "libIexMath-mitsuba.so", "libIlmThread-mitsuba.so",
    "libImath-mitsuba.so", "libIlmImf-mitsuba.so", "libpng-mitsuba.so", "libjpeg-mitsuba.so",
    "libpugixml.so", "libasmjit-mitsuba.so", "libmitsuba.so",
]
_mi = None


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "libmitsuba.so")) and os.path.isdir(os.path.join(REF_DIR, "python", "mitsuba"))


def describe() -> str:
    try:
        return open(os.path.join(REF_DIR, "BUILD_INFO.txt")).read().strip().replace("\n", "; ")
    except OSError:
        return "oracle/_ref (no BUILD_INFO.txt)"


def mitsuba(variant: str = "scalar_mono_double"):
    """Import the reference's ``mitsuba`` module (once) and select ``variant``."""
    global _mi
    if _mi is None:
        if not available():
            raise RuntimeError("the reference build is absent: run oracle/build_ref.sh (needs /root/reference)")
        for name in _PRELOAD:
            ctypes.CDLL(os.path.join(REF_DIR, name), mode=ctypes.RTLD_GLOBAL)
        ctypes.CDLL(os.path.join(REF_DIR, "python", "mitsuba", "libnanobind-drjit.so"), mode=ctypes.RTLD_GLOBAL)
        sys.path.insert(0, os.path.join(REF_DIR, "python"))
        import mitsuba as mi  # noqa: E402

        _mi = mi
    if _mi.variant() != variant:
        _mi.set_variant(variant)
    return _mi


def llvm_runtime() -> str | None:
    """Path of a loadable libLLVM (what drjit-core dlopens, drjit-core/src/llvm_api.cpp:84), or None."""
    import glob

    cands = [os.environ.get("DRJIT_LIBLLVM_PATH")] if os.environ.get("DRJIT_LIBLLVM_PATH") else []
    for pat in ("/usr/lib/x86_64-linux-gnu/libLLVM*.so*", "/usr/lib/llvm-*/lib/libLLVM*.so*", "/usr/lib64/libLLVM*.so*",
                "/usr/local/lib/libLLVM*.so*"):
        cands += sorted(glob.glob(pat), reverse=True)
    for c in cands:
        try:
            lib = ctypes.CDLL(c)
            lib.LLVMCreateTargetMachine  # an LLVM-C symbol drjit-core needs (llvm_api.cpp:102-165)
            return c
        except (OSError, AttributeError):
            continue
    return None


def is_polarized(d: dict) -> bool:
    integ = d.get("integrator", {})
    while isinstance(integ, dict):
        if integ.get("type") == "stokes":
            return True
        integ = integ.get("nested", integ.get("integrator"))
    return False


def to_mitsuba(mi, value):
    """The scene dict with this repo's stand-in value types replaced by the real Mitsuba ones."""
    from eradiate_b200.kernel._types import ScalarTransform4f, VolumeGrid

    if isinstance(value, dict):
        if value.get("type") == "bitmap" and str(value.get("filename", "")).endswith("central_patch_surface_mask.bmp"):
            # Eradiate ships this 3x3 mask (white centre) with its data; an equivalent file lives next to the fixtures
            value = dict(value, filename=os.path.join(_HERE, "..", "tests", "golden", "central_patch_surface_mask.bmp"))
        return {k: to_mitsuba(mi, v) for k, v in value.items()}
    if isinstance(value, ScalarTransform4f):
        return mi.ScalarTransform4f(np.asarray(value.matrix, dtype=np.float64))
    if isinstance(value, VolumeGrid):
        return mi.VolumeGrid(np.ascontiguousarray(value.data, dtype=np.float32))
    if isinstance(value, np.ndarray):
        return value.tolist() if value.ndim == 1 and value.size <= 4 else value
    if isinstance(value, (np.floating, np.integer)):
        return value.item()
    if isinstance(value, tuple):
        return list(value)
    return value


def film_channels(mi, sensor) -> dict:
    """{channel name: float64 array [H, W]} of the developed film."""
    bmp = sensor.film().bitmap()
    out = {}
    for name, sub in bmp.split():
        a = np.array(sub, dtype=np.float64)
        if a.ndim == 2:
            a = a[..., None]
        root = "" if name == "<root>" else name + "."
        for k, ch in enumerate(sub.struct_()):
            cname = ch.name
            if root and cname.startswith(root):
                cname = cname[len(root):]
            out[root + cname] = a[..., k]
    return out
