"""
Loader of THE REFERENCE'S OWN `eradiate.kernel` boundary module (src/eradiate/kernel/_render.py: mi_load_dict,
mi_traverse, SearchSceneParameter, mi_render; _kernel_dict.py: KernelSceneParameterMap, SceneParameter) on top of the
reference Mitsuba compiled into oracle/_ref.  TEST INFRASTRUCTURE ONLY (see oracle/ref.py).

The Eradiate package as a whole cannot be imported in this image (pint, xarray, pinttrs, dessinemoi, joseki ... are
absent), but the kernel boundary needs none of them at run time: the modules are executed from their files under
/root/reference/src/eradiate with the rest of the package stubbed --
  * `eradiate.config`   -> a settings object with `progress = 0` (no progress bars),
  * `eradiate.contexts` -> this repo's KernelContext stand-in (same interface: .si.as_hashable, .active_sensors,
                           .index_formatted, .kwargs),
  * `pint`, `xarray`    -> empty modules (only named in annotations of util/misc.py),
everything else (`attrs.py`, `rng.py`, `typing.py`, `util/misc.py`, `util/numpydoc.py`, `kernel/_kernel_dict.py`,
`kernel/_render.py`) is the reference's code, unmodified.  Nothing is copied: the files are read where they lie.
"""

from __future__ import annotations

import importlib.util
import os
import sys
import types

SRC = "/root/reference/src/eradiate"
_mod = None


def available() -> bool:
    from . import ref

    return ref.available() and os.path.exists(os.path.join(SRC, "kernel", "_render.py"))


def kernel(variant: str = "scalar_mono_double"):
    """(reference `eradiate.kernel._render` module, `eradiate.kernel._kernel_dict` module, mitsuba module)."""
    global _mod
    from . import ref

    mi = ref.mitsuba(variant)
    if _mod is not None:
        return _mod[0], _mod[1], mi
    from eradiate_b200.kernel._kernel_dict import KernelContext

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    def load(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(SRC, rel))
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
        return m

    if "eradiate" in sys.modules:
        raise RuntimeError("an `eradiate` package is already imported")
    stub("eradiate").__path__ = [SRC]
    for ext in ("pint", "xarray"):
        if ext not in sys.modules:
            stub(ext, Quantity=object, Unit=object, Dataset=object, DataArray=object)
    stub("eradiate.config", settings=types.SimpleNamespace(progress=0),
         ProgressLevel=types.SimpleNamespace(NONE=0, SPECTRAL_LOOP=1, KERNEL=2))
    stub("eradiate.util").__path__ = [SRC + "/util"]
    load("eradiate.util.numpydoc", "util/numpydoc.py")
    load("eradiate.attrs", "attrs.py")
    load("eradiate.typing", "typing.py")
    load("eradiate.util.misc", "util/misc.py")
    load("eradiate.rng", "rng.py")
    stub("eradiate.contexts", KernelContext=KernelContext)
    stub("eradiate.kernel").__path__ = [SRC + "/kernel"]
    kd = load("eradiate.kernel._kernel_dict", "kernel/_kernel_dict.py")
    rd = load("eradiate.kernel._render", "kernel/_render.py")
    _mod = (rd, kd)
    return rd, kd, mi


def translate_umap(umap, variant: str = "scalar_mono_double"):
    """This repo's KernelSceneParameterMap -> the reference's, entry by entry: same functions and flags, node types
    replaced by the Mitsuba classes of the same name (what Eradiate's scene elements pass, e.g.
    scenes/atmosphere/_core.py:788 `SearchSceneParameter(node_type=mi.Medium, ...)`)."""
    rd, kd, mi = kernel(variant)
    out = {}
    for key, p in umap.data.items():
        search = None
        if p.search is not None:
            search = rd.SearchSceneParameter(node_type=getattr(mi, p.search.node_type.__name__), node_id=p.search.node_id,
                                             parameter_relpath=p.search.parameter_relpath)
        flags = kd.KernelSceneParameterFlags(p.flags.value)
        out[key] = kd.SceneParameter(p.func, flags, search=search)
    return kd.KernelSceneParameterMap(out)
