"""
``mi.Bitmap`` facade returned by ``mi_render``.

``Experiment.process`` (``src/eradiate/experiments/_core.py:710-744``) only uses
``bitmap.split()`` (pairs ``("<root>", img)``, ``("nested", img)``,
``("m2_nested", img)``), ``img.pixel_format()`` and ``np.array(img)`` (``[H, W, C]``).
Channel layout follows ``MI/src/films/hdrfilm.cpp:304-405`` for a luminance film
plus the moment integrator's AOVs (``MI/src/integrators/moment.cpp:55-66``).
"""

from __future__ import annotations

import enum

import numpy as np


class PixelFormat(enum.Enum):
    Y = "Y"
    XYZ = "XYZ"
    RGB = "RGB"
    MultiChannel = "MultiChannel"


class Bitmap:
    PixelFormat = PixelFormat

    def __init__(self, data, pixel_format: PixelFormat | None = None, channel_names=None, raw=None):
        if isinstance(data, Bitmap):  # deep copy, like mi.Bitmap(film.bitmap())
            self._data = data._data.copy()
            self._pixel_format = data._pixel_format
            self._channel_names = list(data._channel_names)
            self.raw = None if data.raw is None else {
                k: (v.copy() if hasattr(v, "copy") else v) for k, v in data.raw.items()
            }
            self.stats = getattr(data, "stats", None)
            return
        a = np.array(data, dtype=np.float32)
        if a.ndim == 2:
            a = a[..., None]
        self._data = a
        if pixel_format is None:
            pixel_format = PixelFormat.Y if a.shape[2] == 1 else PixelFormat.MultiChannel
        self._pixel_format = pixel_format
        self._channel_names = list(channel_names) if channel_names is not None else (
            ["Y"] if a.shape[2] == 1 else [f"ch{i}" for i in range(a.shape[2])]
        )
        #: float64 per-pixel sums {"sum_wl", "sum_l", "sum_l2", "spp"} (not part of mi.Bitmap)
        self.raw = raw
        self.stats = None

    def pixel_format(self) -> PixelFormat:
        return self._pixel_format

    def channel_count(self) -> int:
        return self._data.shape[2]

    def channel_names(self) -> list[str]:
        return list(self._channel_names)

    def width(self) -> int:
        return self._data.shape[1]

    def height(self) -> int:
        return self._data.shape[0]

    def size(self):
        return (self.width(), self.height())

    def __array__(self, dtype=None, copy=None):
        return self._data if dtype is None else self._data.astype(dtype)

    def split(self) -> list[tuple[str, "Bitmap"]]:
        """Group channels by prefix (``Bitmap::split``, MI/src/core/bitmap.cpp:589-698): the layers come back SORTED
        by name (byte order, :692-696: ``<root>``, ``S0`` .. ``S3``, ``m2_nested``, ``nested``)."""
        groups: dict[str, list[int]] = {}
        for i, name in enumerate(self._channel_names):
            prefix = name.rsplit(".", 1)[0] if "." in name else "<root>"
            groups.setdefault(prefix, []).append(i)
        out = []
        for prefix, idx in groups.items():
            names = [self._channel_names[i].rsplit(".", 1)[-1] for i in idx]
            if names == ["Y"]:
                fmt = PixelFormat.Y
            elif names == ["X", "Y", "Z"]:
                fmt = PixelFormat.XYZ
            elif names == ["R", "G", "B"]:  # the S0 .. S3 layers of the stokes integrator (stokes.cpp:190-196)
                fmt = PixelFormat.RGB
            else:
                fmt = PixelFormat.MultiChannel
            out.append((prefix, Bitmap(self._data[:, :, idx], fmt, names)))
        out.sort(key=lambda kv: kv[0].encode())
        return out

    def __repr__(self):
        return (
            f"Bitmap[{self.width()}x{self.height()}, {self._pixel_format.name}, "
            f"channels={self._channel_names}]"
        )
