"""
Minimal stand-ins for the Mitsuba value types that appear in the scene
dictionaries Eradiate hands to ``mi_load_dict`` (``mi.ScalarTransform4f``,
``mi.VolumeGrid``, ``mi.ScalarPoint3f``).  The flattener accepts either the
real Mitsuba objects (anything exposing ``.matrix`` / convertible through
``np.array``) or these.

Reference: ``src/eradiate/kernel/transform.py`` (``map_unit_cube``, ``map_cube``),
``MI/include/mitsuba/core/transform.h`` (``look_at``, ``translate``, ``scale``).
"""

from __future__ import annotations

import numpy as np


class ScalarTransform4f:
    """Row-major 4x4 affine transform with Mitsuba's chaining API."""

    __slots__ = ("matrix",)

    def __init__(self, matrix=None):
        if matrix is None:
            self.matrix = np.eye(4, dtype=np.float64)
        else:
            m = np.array(getattr(matrix, "matrix", matrix), dtype=np.float64)
            if m.shape != (4, 4):
                raise ValueError(f"expected a 4x4 matrix, got shape {m.shape}")
            self.matrix = m

    # -- chaining constructors (T().translate(v) == T @ translate(v)) ---------
    def translate(self, v) -> "ScalarTransform4f":
        m = np.eye(4)
        m[:3, 3] = np.asarray(v, dtype=np.float64)
        return ScalarTransform4f(self.matrix @ m)

    def scale(self, v) -> "ScalarTransform4f":
        v = np.asarray(v, dtype=np.float64)
        if v.ndim == 0:
            v = np.full(3, float(v))
        m = np.diag([v[0], v[1], v[2], 1.0])
        return ScalarTransform4f(self.matrix @ m)

    def rotate(self, axis, angle) -> "ScalarTransform4f":
        """Rotation of ``angle`` degrees around ``axis`` (Rodrigues)."""
        a = np.asarray(axis, dtype=np.float64)
        a = a / np.linalg.norm(a)
        t = np.deg2rad(angle)
        c, s = np.cos(t), np.sin(t)
        K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
        R = np.eye(3) * c + s * K + (1 - c) * np.outer(a, a)
        m = np.eye(4)
        m[:3, :3] = R
        return ScalarTransform4f(self.matrix @ m)

    def look_at(self, origin, target, up) -> "ScalarTransform4f":
        """Camera-to-world transform: +Z maps to ``normalize(target - origin)``."""
        origin = np.asarray(origin, dtype=np.float64)
        target = np.asarray(target, dtype=np.float64)
        up = np.asarray(up, dtype=np.float64)
        d = target - origin
        d = d / np.linalg.norm(d)
        left = np.cross(up, d)
        nl = np.linalg.norm(left)
        if nl == 0:
            raise ValueError("look_at(): 'up' is parallel to the viewing direction")
        left /= nl
        new_up = np.cross(d, left)
        m = np.eye(4)
        m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = left, new_up, d, origin
        return ScalarTransform4f(self.matrix @ m)

    def inverse(self) -> "ScalarTransform4f":
        return ScalarTransform4f(np.linalg.inv(self.matrix))

    def __matmul__(self, other):
        if isinstance(other, ScalarTransform4f) or hasattr(other, "matrix"):
            return ScalarTransform4f(self.matrix @ np.array(other.matrix, dtype=np.float64))
        v = np.asarray(other, dtype=np.float64)
        return self.transform_affine(v)

    def transform_affine(self, p):
        p = np.asarray(p, dtype=np.float64)
        return self.matrix[:3, :3] @ p + self.matrix[:3, 3]

    def transform_vector(self, v):
        return self.matrix[:3, :3] @ np.asarray(v, dtype=np.float64)

    def __array__(self, dtype=None, copy=None):
        return self.matrix if dtype is None else self.matrix.astype(dtype)

    def __repr__(self):
        return f"ScalarTransform4f({self.matrix.tolist()})"


class VolumeGrid:
    """``mi.VolumeGrid`` stand-in: float32 data laid out ``[z, y, x]`` or ``[z, y, x, c]``."""

    __slots__ = ("data",)

    def __init__(self, data):
        a = np.array(getattr(data, "data", data), dtype=np.float32)
        if a.ndim == 3:
            a = a[..., None]
        if a.ndim != 4:
            raise ValueError(f"VolumeGrid expects 3 or 4 dimensions, got {a.ndim}")
        self.data = a

    def __array__(self, dtype=None, copy=None):
        return self.data if dtype is None else self.data.astype(dtype)


def to_matrix(value) -> np.ndarray:
    """Coerce a transform-like value (ours, Mitsuba's, ndarray, nested list) to 4x4 float64."""
    if value is None:
        return np.eye(4)
    m = getattr(value, "matrix", value)
    m = np.array(m, dtype=np.float64)
    if m.shape != (4, 4):
        raise RuntimeError(f"unsupported transform value of shape {m.shape}")
    return m


def to_grid(value) -> np.ndarray:
    """Coerce a volume-grid-like value to a float32 ``[z, y, x, c]`` array."""
    if isinstance(value, VolumeGrid):
        return value.data
    a = np.array(value, dtype=np.float32)
    if a.ndim == 3:
        a = a[..., None]
    if a.ndim != 4:
        raise RuntimeError(f"unsupported volume grid of shape {a.shape}")
    return a


def map_unit_cube(xmin, xmax, ymin, ymax, zmin, zmax) -> ScalarTransform4f:
    """``src/eradiate/kernel/transform.py:11-52``: map [0,1]^3 to the given box."""
    return ScalarTransform4f().translate([xmin, ymin, zmin]) @ ScalarTransform4f().scale(
        [xmax - xmin, ymax - ymin, zmax - zmin]
    )


def map_cube(xmin, xmax, ymin, ymax, zmin, zmax) -> ScalarTransform4f:
    """``src/eradiate/kernel/transform.py:55-100``: map [-1,1]^3 to the given box."""
    half = [0.5 * (xmax - xmin), 0.5 * (ymax - ymin), 0.5 * (zmax - zmin)]
    centre = [0.5 * (xmax + xmin), 0.5 * (ymax + ymin), 0.5 * (zmax + zmin)]
    return ScalarTransform4f().translate(centre) @ ScalarTransform4f().scale(half)
