"""
Mirror of the small part of ``src/eradiate/kernel/_kernel_dict.py`` and
``src/eradiate/kernel/_render.py:27-66`` that the hot-path boundary needs:
scene-parameter update maps rendered per kernel context.

Real Eradiate objects (``KernelSceneParameterMap``, ``SceneParameter``,
``SearchSceneParameter``, ``KernelContext``) are duck-type compatible with these:
``mi_traverse`` / ``mi_render`` only use ``.data``, ``.items()``, ``.render(ctx)``,
``.parameter_id``, ``.search(node, path)``, ``ctx.si.as_hashable``,
``ctx.active_sensors`` and ``ctx.index_formatted``.
"""

from __future__ import annotations

import enum
import typing as t
from collections import UserDict


class KernelSceneParameterFlags(enum.Flag):
    """``_kernel_dict.py:43-52``."""

    NONE = 0
    SPECTRAL = enum.auto()
    GEOMETRIC = enum.auto()
    ALL = SPECTRAL | GEOMETRIC


class _Unused:
    def __repr__(self):
        return "UNUSED"


class SearchSceneParameter:
    """``_render.py:27-66``: find a parameter by node type + node id + relative path."""

    def __init__(self, node_type: type, node_id: str, parameter_relpath: str):
        if not isinstance(node_type, type):
            raise TypeError("node_type must be a type")
        self.node_type = node_type
        self.node_id = node_id
        self.parameter_relpath = parameter_relpath

    def __call__(self, node, node_path: str | None = None) -> str | None:
        if isinstance(node, self.node_type) and node.id() == self.node_id:
            prefix = f"{node_path}." if node_path is not None else ""
            return f"{prefix}{self.parameter_relpath}"
        return None


class SceneParameter:
    """``_kernel_dict.py:55-103``: a context-dependent scene parameter value."""

    UNUSED = _Unused()

    def __init__(
        self,
        func: t.Callable,
        flags: KernelSceneParameterFlags = KernelSceneParameterFlags.ALL,
        search: SearchSceneParameter | None = None,
        parameter_id: str | None = None,
    ):
        self.func = func
        self.flags = flags
        self.search = search
        self.parameter_id = parameter_id

    def __call__(self, ctx) -> t.Any:
        return self.func(ctx)


class KernelSceneParameterMap(UserDict):
    """``_kernel_dict.py:242-314``."""

    def __init__(self, data: dict | None = None):
        super().__init__()
        if data:
            self.data.update(data)

    def render(self, ctx, flags=KernelSceneParameterFlags.ALL, drop: bool = False) -> dict:
        unused, result = [], {}
        for k in list(self.keys()):
            v = self[k]
            if isinstance(v, SceneParameter) or callable(v):
                key = k if getattr(v, "parameter_id", None) is None else v.parameter_id
                vflags = getattr(v, "flags", KernelSceneParameterFlags.ALL)
                if vflags & flags:
                    result[key] = v(ctx)
                else:
                    unused.append(k)
                    if not drop:
                        result[key] = SceneParameter.UNUSED
            else:
                raise ValueError(f"value for key '{k}' is not a SceneParameter")
        if not drop and unused:
            raise ValueError(f"Unevaluated parameters: {unused}")
        return result


class _SpectralIndex:
    def __init__(self, w: float, g: float | None = None):
        self.w, self.g = float(w), g

    @property
    def as_hashable(self):
        return self.w if self.g is None else (self.w, self.g)

    def __repr__(self):
        return f"SpectralIndex(w={self.w}, g={self.g})"


class KernelContext:
    """Minimal ``eradiate.contexts.KernelContext`` stand-in (one spectral index)."""

    def __init__(self, w: float = 550.0, g: float | None = None, active_sensors=None, **kwargs):
        self.si = _SpectralIndex(w, g)
        self.active_sensors = active_sensors
        self.kwargs = kwargs

    @property
    def index_formatted(self) -> str:
        return f"{self.si.w:g} nm" if self.si.g is None else f"{self.si.w:g} nm:{self.si.g:g}"
