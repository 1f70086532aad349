"""
Point-wise evaluation of the device plugin implementations through the C ABI's
``ertb_kat_*`` entry points (known-answer tests; one GPU thread per query).
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi, _lib
from .kernel._render import _device_scene


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _fp(a):
    return a.ctypes.data_as(_abi.c_float_p)


def _dp(a):
    return a.ctypes.data_as(_abi.c_double_p)


def bsdf_eval(scene, wi, wo):
    dev = _device_scene(scene)
    dev.sync()
    wi, wo = _f(wi).reshape(-1, 3), _f(wo).reshape(-1, 3)
    out = np.zeros(wi.shape[0], dtype=np.float32)
    _lib.check(dev.lib.ertb_kat_bsdf_eval(dev.handle, wi.shape[0], _fp(wi), _fp(wo), _fp(out)))
    return out


def bsdf_mueller(scene, wi, wo):
    """BSDF::eval of a polarized scene as 4x4 Mueller matrices (local frame, implicit Stokes bases)."""
    dev = _device_scene(scene)
    dev.sync()
    wi, wo = _f(wi).reshape(-1, 3), _f(wo).reshape(-1, 3)
    M = np.zeros((wi.shape[0], 4, 4), dtype=np.float32)
    _lib.check(dev.lib.ertb_kat_bsdf_mueller(dev.handle, wi.shape[0], _fp(wi), _fp(wo), _fp(M)))
    return M


def bsdf_sample(scene, wi, u):
    dev = _device_scene(scene)
    dev.sync()
    wi, u = _f(wi).reshape(-1, 3), _f(u).reshape(-1, 3)
    wo = np.zeros_like(wi)
    w = np.zeros(wi.shape[0], dtype=np.float32)
    _lib.check(dev.lib.ertb_kat_bsdf_sample(dev.handle, wi.shape[0], _fp(wi), _fp(u), _fp(wo), _fp(w)))
    return wo, w


def phase_eval(scene, leaf, cos_theta):
    dev = _device_scene(scene)
    dev.sync()
    c = _f(cos_theta).reshape(-1)
    out = np.zeros_like(c)
    _lib.check(dev.lib.ertb_kat_phase_eval(dev.handle, leaf, c.size, _fp(c), _fp(out)))
    return out


def phase_sample(scene, leaf, u):
    dev = _device_scene(scene)
    dev.sync()
    u = _f(u).reshape(-1, 2)
    ct, w, pdf = (np.zeros(u.shape[0], dtype=np.float32) for _ in range(3))
    _lib.check(dev.lib.ertb_kat_phase_sample(dev.handle, leaf, u.shape[0], _fp(u), _fp(ct), _fp(w), _fp(pdf)))
    return ct, w, pdf


def sensor_ray(scene, sensor, film_sample, aperture_sample):
    dev = _device_scene(scene)
    dev.sync()
    fs, ap = _f(film_sample).reshape(-1, 2), _f(aperture_sample).reshape(-1, 2)
    n = fs.shape[0]
    o, d = np.zeros((n, 3)), np.zeros((n, 3))
    w = np.zeros(n, dtype=np.float32)
    _lib.check(dev.lib.ertb_kat_sensor_ray(dev.handle, sensor, n, _fp(fs), _fp(ap), _dp(o), _dp(d), _fp(w)))
    return o, d, w


def phase_mueller(scene, leaf, wi, wo):
    """Mueller matrices (n, 4, 4) and pdf (n) of one phase-function leaf, world-space directions."""
    dev = _device_scene(scene)
    dev.sync()
    wi, wo = _f(wi).reshape(-1, 3), _f(wo).reshape(-1, 3)
    n = wi.shape[0]
    M = np.zeros((n, 4, 4), dtype=np.float32)
    pdf = np.zeros(n, dtype=np.float32)
    _lib.check(dev.lib.ertb_kat_phase_mueller(dev.handle, leaf, n, _fp(wi), _fp(wo), _fp(M), _fp(pdf)))
    return M, pdf


def piecewise_sample(scene, altitude, mu, u):
    """Analytic free flight of the piecewise medium: (distance, kind) per query; kind 0 = collision,
    1 = reached the ground, 2 = left through the top of the atmosphere."""
    dev = _device_scene(scene)
    dev.sync()
    z, mu, u = (np.ascontiguousarray(np.broadcast_arrays(_f(altitude), _f(mu), _f(u))[i]) for i in range(3))
    n = z.size
    t = np.zeros(n, dtype=np.float32)
    kind = np.zeros(n, dtype=np.int32)
    _lib.check(dev.lib.ertb_kat_piecewise_sample(dev.handle, n, _fp(z), _fp(mu), _fp(u), _fp(t),
                                                 kind.ctypes.data_as(C.POINTER(C.c_int32))))
    return t, kind


def piecewise_transmittance(scene, altitude, mu):
    """exp(-optical depth) from each altitude to the top of the atmosphere along cosine mu > 0."""
    dev = _device_scene(scene)
    dev.sync()
    z, mu = (np.ascontiguousarray(np.broadcast_arrays(_f(altitude), _f(mu))[i]) for i in range(2))
    tr = np.zeros(z.size, dtype=np.float32)
    _lib.check(dev.lib.ertb_kat_piecewise_transmittance(dev.handle, z.size, _fp(z), _fp(mu), _fp(tr)))
    return tr


def canopy_intersect(scene, origin, direction, tmax=None):
    """Nearest leaf along world-space rays: (t, normal, group); t = inf for a miss."""
    dev = _device_scene(scene)
    o = np.ascontiguousarray(origin, dtype=np.float64).reshape(-1, 3)
    d = _f(direction).reshape(-1, 3)
    n = o.shape[0]
    tm = np.full(n, 1e30, np.float32) if tmax is None else _f(tmax).reshape(-1)
    t = np.zeros(n, np.float64)
    nrm = np.zeros((n, 3), np.float32)
    grp = np.zeros(n, np.int32)
    _lib.check(dev.lib.ertb_kat_canopy_intersect(dev.handle, n, _dp(o), _fp(d), _fp(tm), _dp(t), _fp(nrm),
                                                 grp.ctypes.data_as(C.POINTER(C.c_int32))))
    return t, nrm, grp


def leaf_bsdf_eval(scene, group, cos_i, cos_o):
    dev = _device_scene(scene)
    ci, co = _f(cos_i).reshape(-1), _f(cos_o).reshape(-1)
    out = np.zeros(ci.size, np.float32)
    _lib.check(dev.lib.ertb_kat_leaf_bsdf_eval(dev.handle, group, ci.size, _fp(ci), _fp(co), _fp(out)))
    return out


def leaf_bsdf_sample(scene, group, cos_i, u):
    dev = _device_scene(scene)
    ci, u = _f(cos_i).reshape(-1), _f(u).reshape(-1, 3)
    wo = np.zeros((ci.size, 3), np.float32)
    w = np.zeros(ci.size, np.float32)
    _lib.check(dev.lib.ertb_kat_leaf_bsdf_sample(dev.handle, group, ci.size, _fp(ci), _fp(u), _fp(wo), _fp(w)))
    return wo, w
