"""
ctypes wrapper of the CPU oracle (``oracle/libertb_oracle.so``).

TEST INFRASTRUCTURE ONLY -- imported by ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s cpu_baseline / ``--impl reference`` legs.  The product package
``eradiate_b200`` never imports this module.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from eradiate_b200 import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libertb_oracle.so")
_lib = None

dp = _abi.c_double_p
fp = _abi.c_float_p


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc, OpenMP)."""
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    srcs.append(os.path.join(_HERE, "..", "include", "eradiate_b200.h"))
    stale = force or not os.path.exists(_SO) or any(
        os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs
    )
    if stale:
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)
    return _SO


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        lib = C.CDLL(_SO)
        lib.ertbo_last_error.restype = C.c_char_p
        _lib = lib
    return _lib


def _check(status: int):
    if status != 0:
        raise RuntimeError(load().ertbo_last_error().decode())


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def render(desc: _abi.SceneDesc, sensor: int = 0, seed: int = 0, spp: int = 1024,
           sample_offset: int = 0, n_threads: int = 0):
    """Returns (sum_wl, sum_l, sum_l2, stats dict)."""
    lib = load()
    sd = desc.sensors[sensor]
    npix = sd.width * sd.height
    out = [np.zeros(npix, dtype=np.float64) for _ in range(3)]
    stats = _abi.RenderStats()
    _check(
        lib.ertbo_render(
            C.byref(desc), C.c_int(sensor), C.c_uint64(seed), C.c_uint64(spp),
            C.c_uint64(sample_offset), *[o.ctypes.data_as(dp) for o in out], C.byref(stats),
            C.c_int(n_threads),
        )
    )
    return out[0], out[1], out[2], _with_flights(stats.as_dict())


def _with_flights(st: dict) -> dict:
    """Adds the free-flight counts of the last render (what a kernel without stencil-crossing
    iterations and with culled zero-weight shadow rays counts as its loop trips)."""
    f = (C.c_uint64 * 2)()
    load().ertbo_last_flights(f)
    st["flights_main"], st["flights_nee"] = int(f[0]), int(f[1])
    return st


def render_stokes(desc: _abi.SceneDesc, sensor: int = 0, seed: int = 0, spp: int = 1024,
                  sample_offset: int = 0, n_threads: int = 0):
    """Returns (sum_wl, sum_l, sum_l2, sum_stokes[4, npix], stats dict)."""
    lib = load()
    sd = desc.sensors[sensor]
    npix = sd.width * sd.height
    out = [np.zeros(npix, dtype=np.float64) for _ in range(3)]
    st = np.zeros((4, npix), dtype=np.float64)
    stats = _abi.RenderStats()
    _check(
        lib.ertbo_render_stokes(
            C.byref(desc), C.c_int(sensor), C.c_uint64(seed), C.c_uint64(spp),
            C.c_uint64(sample_offset), *[o.ctypes.data_as(dp) for o in out], st.ctypes.data_as(dp),
            C.byref(stats), C.c_int(n_threads),
        )
    )
    return out[0], out[1], out[2], st, _with_flights(stats.as_dict())


def phase_mueller(desc, leaf, wi, wo):
    wi, wo = _d(wi).reshape(-1, 3), _d(wo).reshape(-1, 3)
    n = wi.shape[0]
    M, pdf = np.zeros((n, 4, 4)), np.zeros(n)
    _check(load().ertbo_phase_mueller(C.byref(desc), C.c_int(leaf), C.c_size_t(n), wi.ctypes.data_as(dp),
                                      wo.ctypes.data_as(dp), M.ctypes.data_as(dp), pdf.ctypes.data_as(dp)))
    return M, pdf


def piecewise_sample(desc, o, d, sample, si_t=None, half_width=0.0):
    """piecewise.cpp sample_interaction_real -> (t, tr, pdf)."""
    o, d = _d(o).reshape(-1, 3), _d(d).reshape(-1, 3)
    n = o.shape[0]
    u = _d(np.broadcast_to(sample, (n,)))
    st = _d(np.full(n, np.inf) if si_t is None else np.broadcast_to(si_t, (n,)))
    t, tr, pdf = np.zeros(n), np.zeros(n), np.zeros(n)
    _check(load().ertbo_piecewise_sample(C.byref(desc), C.c_double(half_width), C.c_size_t(n),
                                         o.ctypes.data_as(dp), d.ctypes.data_as(dp), u.ctypes.data_as(dp),
                                         st.ctypes.data_as(dp), t.ctypes.data_as(dp), tr.ctypes.data_as(dp),
                                         pdf.ctypes.data_as(dp)))
    return t, tr, pdf


def piecewise_eval(desc, o, d, si_t=None, half_width=0.0):
    """piecewise.cpp eval_transmittance_pdf_real -> (tr, pdf, escaped)."""
    o, d = _d(o).reshape(-1, 3), _d(d).reshape(-1, 3)
    n = o.shape[0]
    st = _d(np.full(n, np.inf) if si_t is None else np.broadcast_to(si_t, (n,)))
    tr, pdf = np.zeros(n), np.zeros(n)
    esc = np.zeros(n, dtype=np.int32)
    _check(load().ertbo_piecewise_eval(C.byref(desc), C.c_double(half_width), C.c_size_t(n),
                                       o.ctypes.data_as(dp), d.ctypes.data_as(dp), st.ctypes.data_as(dp),
                                       tr.ctypes.data_as(dp), pdf.ctypes.data_as(dp),
                                       esc.ctypes.data_as(C.POINTER(C.c_int))))
    return tr, pdf, esc.astype(bool)


def bsdf_mueller(desc, wi, wo):
    wi, wo = _d(wi).reshape(-1, 3), _d(wo).reshape(-1, 3)
    M = np.zeros((wi.shape[0], 4, 4))
    _check(load().ertbo_bsdf_mueller(C.byref(desc), C.c_size_t(wi.shape[0]), wi.ctypes.data_as(dp),
                                     wo.ctypes.data_as(dp), M.ctypes.data_as(dp)))
    return M


def bsdf_eval(desc, wi, wo):
    wi, wo = _d(wi).reshape(-1, 3), _d(wo).reshape(-1, 3)
    out = np.zeros(wi.shape[0])
    _check(load().ertbo_bsdf_eval(C.byref(desc), C.c_size_t(wi.shape[0]), wi.ctypes.data_as(dp),
                                  wo.ctypes.data_as(dp), out.ctypes.data_as(dp)))
    return out


def bsdf_pdf(desc, wi, wo):
    wi, wo = _d(wi).reshape(-1, 3), _d(wo).reshape(-1, 3)
    out = np.zeros(wi.shape[0])
    _check(load().ertbo_bsdf_pdf(C.byref(desc), C.c_size_t(wi.shape[0]), wi.ctypes.data_as(dp),
                                 wo.ctypes.data_as(dp), out.ctypes.data_as(dp)))
    return out


def bsdf_sample(desc, wi, u):
    wi, u = _d(wi).reshape(-1, 3), _d(u).reshape(-1, 3)
    wo, w = np.zeros_like(wi), np.zeros(wi.shape[0])
    _check(load().ertbo_bsdf_sample(C.byref(desc), C.c_size_t(wi.shape[0]), wi.ctypes.data_as(dp),
                                    u.ctypes.data_as(dp), wo.ctypes.data_as(dp), w.ctypes.data_as(dp)))
    return wo, w


def phase_eval(desc, leaf, cos_theta):
    c = _d(cos_theta).reshape(-1)
    out = np.zeros_like(c)
    _check(load().ertbo_phase_eval(C.byref(desc), C.c_int(leaf), C.c_size_t(c.size),
                                   c.ctypes.data_as(dp), out.ctypes.data_as(dp)))
    return out


def phase_sample(desc, leaf, u):
    u = _d(u).reshape(-1, 2)
    ct, w, pdf = (np.zeros(u.shape[0]) for _ in range(3))
    _check(load().ertbo_phase_sample(C.byref(desc), C.c_int(leaf), C.c_size_t(u.shape[0]),
                                     u.ctypes.data_as(dp), ct.ctypes.data_as(dp),
                                     w.ctypes.data_as(dp), pdf.ctypes.data_as(dp)))
    return ct, w, pdf


def sensor_ray(desc, sensor, film_sample, aperture_sample):
    fs, ap = _d(film_sample).reshape(-1, 2), _d(aperture_sample).reshape(-1, 2)
    n = fs.shape[0]
    o, d, w = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros(n)
    _check(load().ertbo_sensor_ray(C.byref(desc), C.c_int(sensor), C.c_size_t(n),
                                   fs.ctypes.data_as(dp), ap.ctypes.data_as(dp),
                                   o.ctypes.data_as(dp), d.ctypes.data_as(dp), w.ctypes.data_as(dp)))
    return o, d, w


def distr_regular(pdf, u=None, xq=None):
    pdf = np.ascontiguousarray(pdf, dtype=np.float32)
    nq = max(0 if u is None else len(u), 0 if xq is None else len(xq))
    u_ = _d(np.zeros(nq) if u is None else u)
    xq_ = _d(np.zeros(nq) if xq is None else xq)
    xs, pe, integ = np.zeros(nq), np.zeros(nq), C.c_double()
    _check(load().ertbo_distr_regular(pdf.ctypes.data_as(fp), C.c_int(pdf.size), C.c_size_t(nq),
                                      u_.ctypes.data_as(dp), xs.ctypes.data_as(dp),
                                      xq_.ctypes.data_as(dp), pe.ctypes.data_as(dp), C.byref(integ)))
    return xs, pe, integ.value


def distr_irregular(nodes, pdf, u=None, xq=None):
    nodes = np.ascontiguousarray(nodes, dtype=np.float32)
    pdf = np.ascontiguousarray(pdf, dtype=np.float32)
    nq = max(0 if u is None else len(u), 0 if xq is None else len(xq))
    u_ = _d(np.zeros(nq) if u is None else u)
    xq_ = _d(np.zeros(nq) if xq is None else xq)
    xs, pe, integ = np.zeros(nq), np.zeros(nq), C.c_double()
    _check(load().ertbo_distr_irregular(nodes.ctypes.data_as(fp), pdf.ctypes.data_as(fp),
                                        C.c_int(pdf.size), C.c_size_t(nq), u_.ctypes.data_as(dp),
                                        xs.ctypes.data_as(dp), xq_.ctypes.data_as(dp),
                                        pe.ctypes.data_as(dp), C.byref(integ)))
    return xs, pe, integ.value


def warp(name: str, u, v):
    lib = load()
    u, v = np.atleast_1d(_d(u)), np.atleast_1d(_d(v))
    if name == "uniform_disk_concentric":
        out = np.zeros((u.size, 2))
        x, y = C.c_double(), C.c_double()
        for i in range(u.size):
            lib.ertbo_square_to_uniform_disk_concentric(C.c_double(u[i]), C.c_double(v[i]),
                                                        C.byref(x), C.byref(y))
            out[i] = (x.value, y.value)
        return out
    fn = {"cosine_hemisphere": lib.ertbo_square_to_cosine_hemisphere,
          "uniform_hemisphere": lib.ertbo_square_to_uniform_hemisphere}[name]
    out = np.zeros((u.size, 3))
    buf = (C.c_double * 3)()
    for i in range(u.size):
        fn(C.c_double(u[i]), C.c_double(v[i]), buf)
        out[i] = list(buf)
    return out


def medium_lookup(desc, points):
    p = _d(points).reshape(-1, 3)
    st, al = np.zeros(p.shape[0]), np.zeros(p.shape[0])
    _check(load().ertbo_medium_lookup(C.byref(desc), C.c_size_t(p.shape[0]), p.ctypes.data_as(dp),
                                      st.ctypes.data_as(dp), al.ctypes.data_as(dp)))
    return st, al


def canopy_intersect(desc, origin, direction, tmax=None):
    """Nearest leaf along world-space rays (uniform-grid DDA): (t, normal, group)."""
    o, d = _d(origin).reshape(-1, 3), _d(direction).reshape(-1, 3)
    n = o.shape[0]
    tm = _d(np.full(n, np.inf) if tmax is None else np.broadcast_to(tmax, (n,)))
    t, nrm = np.zeros(n), np.zeros((n, 3))
    grp = np.zeros(n, dtype=np.int32)
    _check(load().ertbo_canopy_intersect(C.byref(desc), C.c_size_t(n), o.ctypes.data_as(dp), d.ctypes.data_as(dp),
                                         tm.ctypes.data_as(dp), t.ctypes.data_as(dp), nrm.ctypes.data_as(dp),
                                         grp.ctypes.data_as(C.POINTER(C.c_int))))
    return t, nrm, grp


def leaf_bsdf(desc, group, mode, wi, wo=None, u=None):
    """bilambertian eval ('eval'), pdf ('pdf') or sample ('sample' -> (wo, weight)) in the leaf frame."""
    wi = _d(wi).reshape(-1, 3)
    n = wi.shape[0]
    out = np.zeros(n)
    m = {"eval": 0, "pdf": 1, "sample": 2}[mode]
    wo = np.zeros((n, 3)) if wo is None else _d(wo).reshape(-1, 3).copy()
    uu = np.zeros((n, 3)) if u is None else _d(u).reshape(-1, 3)
    _check(load().ertbo_leaf_bsdf(C.byref(desc), C.c_int(group), C.c_int(m), C.c_size_t(n), wi.ctypes.data_as(dp),
                                  wo.ctypes.data_as(dp), uu.ctypes.data_as(dp), out.ctypes.data_as(dp)))
    return (wo, out) if mode == "sample" else out
