"""
The C-ABI library loads and exports every symbol include/eradiate_b200.h declares
(no compute calls without a GPU), and the product path fails loudly -- no CPU
fallback -- when no CUDA device is present.
"""

import ctypes as C
import os
import re

import pytest
import torch

from eradiate_b200 import _abi, _lib, scenes
from eradiate_b200.kernel import mi_load_dict, render

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported():
    header = open(os.path.join(ROOT, "include", "eradiate_b200.h")).read()
    declared = set(re.findall(r"\b(ertb_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_abi.EXPORTED_SYMBOLS)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name


def test_abi_version_and_enums_match_header():
    header = open(os.path.join(ROOT, "include", "eradiate_b200.h")).read()
    assert int(re.search(r"#define ERTB_ABI_VERSION (\d+)", header).group(1)) == _abi.ABI_VERSION
    assert _lib.load().ertb_abi_version() == _abi.ABI_VERSION
    for name, val in re.findall(r"ERTB_([A-Z_]+) = (\d+)", header):
        py = name.replace("GEOM_", "GEOM_")
        assert getattr(_abi, py) == int(val), name
    for macro in ("MAX_PHASE", "MAX_BSDF_PARAMS", "MAX_LAYERS", "MAX_PHASE_NODES"):
        assert int(re.search(rf"#define ERTB_{macro} (\d+)", header).group(1)) == getattr(_abi, macro)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under eradiate_b200/ may reference it."""
    pkg = os.path.join(ROOT, "eradiate_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "ertbo_" not in src and "libertb_oracle" not in src, f


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_gpu_fails_loudly():
    lib = _lib.load()
    assert lib.ertb_device_count() == -1
    assert b"no usable CUDA device" in lib.ertb_last_error()
    sc = mi_load_dict(scenes.config_c1())
    with pytest.raises(RuntimeError, match="no usable CUDA device"):
        render(sc, 0, 1, 16)


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "_LIB_PATH", "/nonexistent/libertb_cuda.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()
