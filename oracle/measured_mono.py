"""
numpy restatement of the `measured_mono` BSDF -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.py).

Follows, in float64 and one query at a time:
  * the tensor-file container          MI/src/core/tensor.cpp:12-57
  * Marginal2D<Dim, Continuous = true>  MI/include/mitsuba/core/distr_2d.h:868-1480 (constructor :907-1022,
    eval :1058-1090, sample_continuous :1288-1377, invert_continuous :1379-1453, sample_segment / invert_segment
    :1455-1470, parameter interpolation :255-292, lookup :1117-1138)
  * MeasuredMono::{sample, eval, pdf}  ERP/bsdfs/measured_mono.cpp:234-447

Pinned on values computed by the compiled reference itself on two synthetic tensor files
(tests/golden/measured_mono_reference.json, tools/make_measured_fixture.py).
"""

from __future__ import annotations

import struct

import numpy as np

_NP = {1: np.uint8, 2: np.int8, 3: np.uint16, 4: np.int16, 5: np.uint32, 6: np.int32, 7: np.uint64, 8: np.int64,
       9: np.float16, 10: np.float32, 11: np.float64}


def read_tensor_file(path: str) -> dict:
    """{field name: ndarray} (tensor.cpp:12-57)."""
    raw = open(path, "rb").read()
    if len(raw) < 18 or raw[:12] != b"tensor_file\0":
        raise RuntimeError("Invalid tensor file: invalid header.")
    (n_fields,) = struct.unpack_from("<I", raw, 14)
    pos, out = 18, {}
    for _ in range(n_fields):
        (nl,) = struct.unpack_from("<H", raw, pos)
        name = raw[pos + 2:pos + 2 + nl].decode()
        pos += 2 + nl
        ndim, dtype, offset = struct.unpack_from("<HBQ", raw, pos)
        pos += 11
        shape = struct.unpack_from("<" + "Q" * ndim, raw, pos)
        pos += 8 * ndim
        if dtype not in _NP:
            raise RuntimeError("Invalid tensor file: unknown type.")
        n = int(np.prod(shape)) if ndim else 1
        out[name] = np.frombuffer(raw, dtype=_NP[dtype], count=n, offset=offset).reshape(shape).copy()
    return out


class Marginal2D:
    """Marginal2D<Dim, true>: `data` has shape param_res + (h, w)."""

    def __init__(self, data, param_values=(), normalize=True, enable_sampling=True):
        data = np.asarray(data, dtype=np.float64)
        self.dim = len(param_values)
        self.param_values = [np.asarray(p, dtype=np.float64) for p in param_values]
        self.h, self.w = data.shape[-2:]
        self.normalized = normalize
        self.slices = int(np.prod(data.shape[:-2])) if self.dim else 1
        # distr_2d.h:244-251: strides in slices, 0 for a parameter of resolution 1
        self.strides, s = [0] * self.dim, 1
        for i in range(self.dim - 1, -1, -1):
            self.strides[i] = s if len(self.param_values[i]) > 1 else 0
            s *= len(self.param_values[i])
        d = data.reshape(self.slices, self.h, self.w).copy()
        w, h = self.w, self.h
        self.marg = np.zeros((self.slices, h - 1))
        self.cond = np.zeros((self.slices, h, w - 1))
        for k in range(self.slices):
            norm = 1.0
            if enable_sampling:  # :928-955 (continuous)
                self.cond[k] = np.cumsum(0.5 / (w - 1) * (d[k, :, :-1] + d[k, :, 1:]), axis=1)
                row_sum = self.cond[k, :, -1]
                self.marg[k] = np.cumsum(0.5 / (h - 1) * (row_sum[:-1] + row_sum[1:]))
                if normalize:
                    norm = 1.0 / self.marg[k, -1]
            elif normalize:  # :996-1013
                ssum = (d[k, :-1, :-1] + d[k, :-1, 1:] + d[k, 1:, :-1] + d[k, 1:, 1:]).sum()
                norm = 1.0 / (0.5 / (w - 1) * 0.5 / (h - 1) * ssum)
            self.cond[k] *= norm
            self.marg[k] *= norm
            d[k] *= norm
        self.data = d

    # -- parameter interpolation (:255-292) and the recursive lookup (:1117-1138) --------------------------
    def _weights(self, param):
        wts, off = [], 0
        for dim in range(self.dim):
            pv = self.param_values[dim]
            if len(pv) == 1:
                wts.append((1.0, 0.0))
                continue
            # math::find_interval: the largest index with pv[i] < param, clamped to [0, n - 2]
            idx = int(np.clip(np.searchsorted(pv, param[dim], side="left") - 1, 0, len(pv) - 2))
            w1 = float(np.clip((param[dim] - pv[idx]) / (pv[idx + 1] - pv[idx]), 0.0, 1.0))
            wts.append((1.0 - w1, w1))
            off += self.strides[dim] * idx
        return off, wts

    def _lookup(self, table, slice_off, wts, index, dim=None):
        """table: [slices, n]; weighted sum over the 2^Dim neighbouring slices."""
        dim = self.dim if dim is None else dim
        if dim == 0:
            return table[slice_off, index] if index >= 0 else 0.0
        w0, w1 = wts[dim - 1]
        v0 = self._lookup(table, slice_off, wts, index, dim - 1)
        v1 = self._lookup(table, slice_off + self.strides[dim - 1], wts, index, dim - 1) if w1 != 0.0 or True else 0.0
        return v0 * w0 + v1 * w1

    def _tables(self):
        return (self.data.reshape(self.slices, -1), self.marg.reshape(self.slices, -1),
                self.cond.reshape(self.slices, -1))

    # -- eval (:1058-1090) ----------------------------------------------------------------------------------
    def eval(self, pos, param=()):
        off, wts = self._weights(param)
        data, _, _ = self._tables()
        x = min(max(pos[0], 0.0), 1.0) * (self.w - 1)
        y = min(max(pos[1], 0.0), 1.0) * (self.h - 1)
        ox, oy = min(int(x), self.w - 2), min(int(y), self.h - 2)
        x, y = x - ox, y - oy
        i = ox + oy * self.w
        v00, v10 = self._lookup(data, off, wts, i), self._lookup(data, off, wts, i + 1)
        v01, v11 = self._lookup(data, off, wts, i + self.w), self._lookup(data, off, wts, i + self.w + 1)
        return (v00 * (1 - x) + v10 * x) * (1 - y) + (v01 * (1 - x) + v11 * x) * y

    @staticmethod
    def _sample_segment(s, inv_width, v0, v1):  # :1455-1464
        non_const = abs(v0 - v1) > 1e-4 * (v0 + v1)
        divisor = (v0 - v1) if non_const else (v0 + v1)
        s *= 2.0 * inv_width
        if non_const:
            s = v0 - np.sqrt(max(v0 * v0 + s * (v1 - v0), 0.0))
        if divisor != 0.0:
            s /= divisor
        return s

    # -- sample_continuous (:1288-1377) ------------------------------------------------------------------------
    def sample(self, sample, param=()):
        off, wts = self._weights(param)
        data, marg, cond = self._tables()
        w, h = self.w, self.h
        eps = np.finfo(np.float64).eps / 2
        sx = min(max(sample[0], eps), 1.0 - eps)
        sy = min(max(sample[1], eps), 1.0 - eps)
        fm = lambda idx: self._lookup(marg, off, wts, idx)  # noqa: E731
        if not self.normalized:
            sy *= fm(h - 2)
        # dr::binary_search(0, n_marg - 1, pred): first index in [0, n_marg - 1] where pred is false
        row = 0
        while row < h - 2 and fm(row) < sy:
            row += 1
        sy -= fm(row - 1) if row > 0 else 0.0
        base = row * (w - 1)
        r0 = self._lookup(cond, off, wts, base + (w - 1) - 1)
        r1 = self._lookup(cond, off, wts, base + 2 * (w - 1) - 1)
        sy = self._sample_segment(sy, h - 1, r0, r1)
        sx *= r0 + (r1 - r0) * sy

        def fc(idx):
            if idx < 0:
                return 0.0
            v0 = self._lookup(cond, off, wts, base + idx)
            v1 = self._lookup(cond, off, wts, base + idx + (w - 1))
            return v0 + (v1 - v0) * sy

        col = 0
        while col < w - 1 and fc(col) < sx:  # binary_search(0, w - 1, ...)
            col += 1
        col = min(col, w - 2)
        sx -= fc(col - 1) if col > 0 else 0.0
        i = row * w + col
        v00, v10 = self._lookup(data, off, wts, i), self._lookup(data, off, wts, i + 1)
        v01, v11 = self._lookup(data, off, wts, i + w), self._lookup(data, off, wts, i + w + 1)
        c0, c1 = v00 + (v01 - v00) * sy, v10 + (v11 - v10) * sy
        sx = self._sample_segment(sx, w - 1, c0, c1)
        return ((col + sx) / (w - 1), (row + sy) / (h - 1)), c0 + (c1 - c0) * sx

    # -- invert_continuous (:1379-1453) ------------------------------------------------------------------------
    def invert(self, sample, param=()):
        off, wts = self._weights(param)
        data, marg, cond = self._tables()
        w, h = self.w, self.h
        x = min(max(sample[0], 0.0), 1.0) * (w - 1)
        y = min(max(sample[1], 0.0), 1.0) * (h - 1)
        px, py = min(int(x), w - 2), min(int(y), h - 2)
        x, y = x - px, y - py
        i = py * w + px
        v00, v10 = self._lookup(data, off, wts, i), self._lookup(data, off, wts, i + 1)
        v01, v11 = self._lookup(data, off, wts, i + w), self._lookup(data, off, wts, i + w + 1)
        c0, c1 = v00 + (v01 - v00) * y, v10 + (v11 - v10) * y
        pdf = c0 + (c1 - c0) * x
        x = x * (c0 + (c1 - c0) * 0.5 * x) / (w - 1)  # invert_segment
        base = py * (w - 1)

        def fc(idx):
            if idx < 0:
                return 0.0
            v0 = self._lookup(cond, off, wts, base + idx)
            v1 = self._lookup(cond, off, wts, base + idx + (w - 1))
            return v0 + (v1 - v0) * y

        x += fc(px - 1) if px > 0 else 0.0
        r0 = self._lookup(cond, off, wts, base + (w - 1) - 1)
        r1 = self._lookup(cond, off, wts, base + 2 * (w - 1) - 1)
        x /= r0 + (r1 - r0) * y
        y = y * (r0 + (r1 - r0) * 0.5 * y) / (h - 1)
        y += self._lookup(marg, off, wts, py - 1) if py > 0 else 0.0
        if not self.normalized:
            y /= self._lookup(marg, off, wts, h - 2)
        return (x, y), pdf


class MeasuredMono:
    """ERP/bsdfs/measured_mono.cpp."""

    def __init__(self, path: str, wavelength: float = 550.0):
        tf = read_tensor_file(path)
        if "wavelengths" not in tf:
            raise RuntimeError("Measurements in RGB format cannot be used with the measured_mono plugin")
        self.wavelength = float(wavelength)
        phi_i, theta_i, wav = tf["phi_i"], tf["theta_i"], tf["wavelengths"]
        self.isotropic = phi_i.shape[0] <= 2
        self.jacobian = bool(tf["jacobian"][0])
        self.reduction = 0 if self.isotropic else int(np.rint(2 * np.pi / (float(phi_i[-1]) - float(phi_i[0]))))
        self.ndf = Marginal2D(tf["ndf"], (), False, False)
        self.sigma = Marginal2D(tf["sigma"], (), False, False)
        self.vndf = Marginal2D(tf["vndf"], (phi_i, theta_i))
        self.luminance = Marginal2D(tf["luminance"], (phi_i, theta_i))
        self.spectra = Marginal2D(tf["spectra"], (phi_i, theta_i, wav), False, False)

    # :496-510
    @staticmethod
    def _u2theta(u):
        return u * u * (np.pi / 2)

    @staticmethod
    def _u2phi(u):
        return (2 * u - 1) * np.pi

    @staticmethod
    def _theta2u(t):
        return np.sqrt(t * (2 / np.pi))

    @staticmethod
    def _phi2u(p):
        return (p + np.pi) / (2 * np.pi)

    @staticmethod
    def _elevation(d):  # :226-232
        dist = np.sqrt(d[0] ** 2 + d[1] ** 2 + (d[2] - 1.0) ** 2)
        return 2.0 * np.arcsin(min(max(0.5 * dist, -1.0), 1.0))

    @staticmethod
    def _mulsign_neg(x, s):  # x * -sign(s)
        return -x if not np.signbit(s) else x

    def _reduce(self, wi, wo=None):
        sx = sy = -1.0
        wi = np.array(wi, dtype=np.float64)
        wo = None if wo is None else np.array(wo, dtype=np.float64)
        if self.reduction >= 2:
            sy = wi[1]
            sx = wi[0] if self.reduction == 4 else sy
            wi[0], wi[1] = self._mulsign_neg(wi[0], sx), self._mulsign_neg(wi[1], sy)
            if wo is not None:
                wo[0], wo[1] = self._mulsign_neg(wo[0], sx), self._mulsign_neg(wo[1], sy)
        return wi, wo, sx, sy

    def _common(self, wi, wo):
        m = wo + wi
        m /= np.linalg.norm(m)
        theta_i, phi_i = self._elevation(wi), np.arctan2(wi[1], wi[0])
        theta_m, phi_m = self._elevation(m), np.arctan2(m[1], m[0])
        u_wi = (self._theta2u(theta_i), self._phi2u(phi_i))
        um1 = self._phi2u(phi_m - phi_i if self.isotropic else phi_m)
        u_m = (self._theta2u(theta_m), um1 - np.floor(um1))
        return m, (phi_i, theta_i), u_wi, u_m

    def eval(self, wi, wo) -> float:  # :339-393 (value, no cosine factor: the tables hold f * cos already)
        if not (wi[2] > 0 and wo[2] > 0):
            return 0.0
        wi, wo, _, _ = self._reduce(wi, wo)
        m, params, u_wi, u_m = self._common(wi, wo)
        sample, _ = self.vndf.invert(u_m, params)
        spec = self.spectra.eval(sample, params + (self.wavelength,))
        if self.jacobian:
            spec *= self.ndf.eval(u_m) / (4 * self.sigma.eval(u_wi))
        return float(spec)

    def pdf(self, wi, wo) -> float:  # :395-447
        if not (wi[2] > 0 and wo[2] > 0):
            return 0.0
        wi, wo, _, _ = self._reduce(wi, wo)
        m, params, u_wi, u_m = self._common(wi, wo)
        sample, vndf_pdf = self.vndf.invert(u_m, params)
        pdf = self.luminance.eval(sample, params)
        sin_theta_m = np.sqrt(max(1.0 - m[2] * m[2], 0.0))
        jac = max(2 * np.pi**2 * u_m[0] * sin_theta_m, 1e-6) * 4 * float(np.dot(wi, m))
        return float(vndf_pdf * pdf / jac)

    def sample(self, wi, sample2):  # :234-337 -> (wo, weight, pdf)
        if not wi[2] > 0:
            return np.zeros(3), 0.0, 0.0
        wi, _, sx, sy = self._reduce(wi)
        theta_i, phi_i = self._elevation(wi), np.arctan2(wi[1], wi[0])
        params = (phi_i, theta_i)
        u_wi = (self._theta2u(theta_i), self._phi2u(phi_i))
        s, lum_pdf = self.luminance.sample((sample2[1], sample2[0]), params)
        u_m, ndf_pdf = self.vndf.sample(s, params)
        phi_m, theta_m = self._u2phi(u_m[1]), self._u2theta(u_m[0])
        if self.isotropic:
            phi_m += phi_i
        m = np.array([np.cos(phi_m) * np.sin(theta_m), np.sin(phi_m) * np.sin(theta_m), np.cos(theta_m)])
        jac = max(2 * np.pi**2 * u_m[0] * np.sin(theta_m), 1e-6) * 4 * float(np.dot(wi, m))
        wo = m * 2 * float(np.dot(m, wi)) - wi
        pdf = ndf_pdf * lum_pdf / jac
        spec = self.spectra.eval(s, params + (self.wavelength,))
        if self.jacobian:
            spec *= self.ndf.eval(u_m) / (4 * self.sigma.eval(u_wi))
        wo[0], wo[1] = self._mulsign_neg(wo[0], sx), self._mulsign_neg(wo[1], sy)
        if not wo[2] > 0:
            return wo, 0.0, float(pdf)
        return wo, float(spec / pdf), float(pdf)
