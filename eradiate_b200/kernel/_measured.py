"""
Host side of the `measured_mono` BSDF (``ERP/bsdfs/measured_mono.cpp``): reads the RGL tensor file and
flattens the five interpolants the plugin builds into ONE float32 table for the device.

* container: ``MI/src/core/tensor.cpp:12-57`` ("tensor_file", version, fields: name, ndim, dtype, offset, shape);
* the plugin's structure checks: ``measured_mono.cpp:86-125``;
* ``Marginal2D<Dim, Continuous = true>`` tables (``MI/include/mitsuba/core/distr_2d.h:907-1022``): per slice the
  data, the conditional CDF of every row (trapezoid rule, accumulated in double) and the marginal CDF over rows,
  all scaled by 1 / (last marginal entry) when the interpolant is normalised (`vndf`, `luminance`); `ndf`, `sigma`
  and `spectra` are evaluated only and keep their values (``normalize = false, enable_sampling = false``).

In a mono variant the wavelength is a scene parameter, and `spectra` is only ever *evaluated*: its dependence on
the wavelength is the linear blend of the two neighbouring slices (``distr_2d.h:255-292, :1117-1138``), so the
blend is done here once per wavelength and the device sees a two-parameter interpolant like the others.

Table layout (float32; integers are stored as floats, all < 2^24):
    [0] n_phi  [1] n_theta  [2] isotropic  [3] jacobian  [4] reduction
    [5] ndf w  [6] ndf h  [7] ndf data        [8] sigma w  [9] sigma h  [10] sigma data
    [11] vndf w [12] vndf h [13] data [14] marg [15] cond
    [16] lum w  [17] lum h  [18] data [19] marg [20] cond
    [21] spectra w [22] spectra h [23] data
    [24] phi_i values [25] theta_i values [26] slice stride of phi_i [27] slice stride of theta_i
(entries 7, 10, 13-15, 18-20, 23-25 are offsets into the table; the header is 32 floats long).
"""

from __future__ import annotations

import struct

import numpy as np

_NP = {1: np.uint8, 2: np.int8, 3: np.uint16, 4: np.int16, 5: np.uint32, 6: np.int32, 7: np.uint64, 8: np.int64,
       9: np.float16, 10: np.float32, 11: np.float64}
HEADER = 32


def read_tensor_file(path: str) -> dict:
    try:
        raw = open(path, "rb").read()
    except OSError as e:
        raise RuntimeError(f"measured_mono: cannot open '{path}': {e}") from e
    if len(raw) < 12 + 2 + 4:
        raise RuntimeError("Invalid tensor file: too small, truncated?")
    if raw[:12] != b"tensor_file\0":
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
        out[name] = np.frombuffer(raw, dtype=_NP[dtype], count=n, offset=offset).reshape(shape)
    return out


def _check(tf: dict) -> None:
    """measured_mono.cpp:74-125."""
    for name in ("theta_i", "phi_i", "ndf", "sigma", "vndf", "luminance", "description", "jacobian"):
        if name not in tf:
            raise RuntimeError(f'TensorFile: field "{name}" not found!')
    if "wavelengths" not in tf:
        raise RuntimeError("Measurements in RGB format cannot be used with the measured_mono plugin")
    if "spectra" not in tf:
        raise RuntimeError('TensorFile: field "spectra" not found!')
    f32 = np.float32
    t, p, w = tf["theta_i"], tf["phi_i"], tf["wavelengths"]
    ok = (tf["description"].ndim == 1 and tf["description"].dtype == np.uint8
          and t.ndim == 1 and t.dtype == f32 and p.ndim == 1 and p.dtype == f32 and w.ndim == 1 and w.dtype == f32
          and tf["ndf"].ndim == 2 and tf["ndf"].dtype == f32 and tf["sigma"].ndim == 2 and tf["sigma"].dtype == f32
          and tf["vndf"].ndim == 4 and tf["vndf"].dtype == f32
          and tf["vndf"].shape[0] == p.shape[0] and tf["vndf"].shape[1] == t.shape[0]
          and tf["luminance"].ndim == 4 and tf["luminance"].dtype == f32
          and tf["luminance"].shape[0] == p.shape[0] and tf["luminance"].shape[1] == t.shape[0]
          and tf["luminance"].shape[2:] == tf["vndf"].shape[2:]
          and tf["spectra"].ndim == 5 and tf["spectra"].dtype == f32
          and tf["spectra"].shape[0] == p.shape[0] and tf["spectra"].shape[1] == t.shape[0]
          and tf["spectra"].shape[2] == w.shape[0] and tf["spectra"].shape[3:] == tf["luminance"].shape[2:]
          and tf["jacobian"].shape == (1,) and tf["jacobian"].dtype == np.uint8)
    if not ok:
        raise RuntimeError("Invalid file structure: measured_mono expects the RGL material database layout")
    for a in (tf["ndf"], tf["sigma"], tf["vndf"], tf["luminance"], tf["spectra"]):
        if a.shape[-1] < 2 or a.shape[-2] < 2:
            raise RuntimeError("Distribution2D(): input array resolution must be >= 2!")


def _marginal_tables(data: np.ndarray, normalize: bool, sampling: bool):
    """[slices, h, w] float64 -> (data, marg [slices, h-1], cond [slices, h, w-1]), scaled as the constructor does."""
    d = np.array(data, dtype=np.float64)
    s, h, w = d.shape
    marg, cond = np.zeros((s, h - 1)), np.zeros((s, h, w - 1))
    for k in range(s):
        norm = 1.0
        if sampling:
            cond[k] = np.cumsum(0.5 / (w - 1) * (d[k, :, :-1] + d[k, :, 1:]), axis=1)
            rows = cond[k, :, -1]
            marg[k] = np.cumsum(0.5 / (h - 1) * (rows[:-1] + rows[1:]))
            if normalize:
                norm = 1.0 / marg[k, -1]
        elif normalize:
            ssum = (d[k, :-1, :-1] + d[k, :-1, 1:] + d[k, 1:, :-1] + d[k, 1:, 1:]).sum()
            norm = 1.0 / (0.5 / (w - 1) * 0.5 / (h - 1) * ssum)
        cond[k] *= norm
        marg[k] *= norm
        d[k] *= norm
    return d, marg, cond


class MeasuredData:
    """The file, kept on the host so that a `wavelength` update only re-blends the spectral slices."""

    def __init__(self, path: str):
        self.tf = read_tensor_file(path)
        _check(self.tf)
        tf = self.tf
        self.phi_i = tf["phi_i"].astype(np.float64)
        self.theta_i = tf["theta_i"].astype(np.float64)
        self.wavelengths = tf["wavelengths"].astype(np.float64)
        self.isotropic = self.phi_i.shape[0] <= 2
        self.jacobian = int(tf["jacobian"][0])
        self.reduction = 0 if self.isotropic else int(np.rint(2.0 * np.pi / (self.phi_i[-1] - self.phi_i[0])))
        n_phi, n_theta = self.phi_i.size, self.theta_i.size
        self._fixed = {
            "ndf": _marginal_tables(tf["ndf"][None], False, False),
            "sigma": _marginal_tables(tf["sigma"][None], False, False),
            "vndf": _marginal_tables(tf["vndf"].reshape((n_phi * n_theta,) + tf["vndf"].shape[2:]), True, True),
            "luminance": _marginal_tables(tf["luminance"].reshape((n_phi * n_theta,) + tf["luminance"].shape[2:]), True, True),
        }

    def spectra_at(self, wavelength: float) -> np.ndarray:
        """[n_phi * n_theta, h, w]: the slices of `spectra` blended at `wavelength` (clamped weights)."""
        sp, wv = self.tf["spectra"].astype(np.float64), self.wavelengths
        if wv.size == 1:
            out = sp[:, :, 0]
        else:
            i = int(np.clip(np.searchsorted(wv, wavelength, side="left") - 1, 0, wv.size - 2))
            w1 = float(np.clip((wavelength - wv[i]) / (wv[i + 1] - wv[i]), 0.0, 1.0))
            out = sp[:, :, i] * (1.0 - w1) + sp[:, :, i + 1] * w1
        return out.reshape((-1,) + sp.shape[3:])

    def table(self, wavelength: float) -> np.ndarray:
        n_phi, n_theta = self.phi_i.size, self.theta_i.size
        spec = self.spectra_at(wavelength)
        parts, off = [], HEADER
        H = np.zeros(HEADER, dtype=np.float64)

        def add(a):
            nonlocal off
            a = np.asarray(a, dtype=np.float64).ravel()
            parts.append(a)
            start = off
            off += a.size
            return start

        H[0:5] = [n_phi, n_theta, int(self.isotropic), self.jacobian, self.reduction]
        d, _, _ = self._fixed["ndf"]
        H[5:8] = [d.shape[2], d.shape[1], add(d)]
        d, _, _ = self._fixed["sigma"]
        H[8:11] = [d.shape[2], d.shape[1], add(d)]
        d, m, c = self._fixed["vndf"]
        H[11:16] = [d.shape[2], d.shape[1], add(d), add(m), add(c)]
        d, m, c = self._fixed["luminance"]
        H[16:21] = [d.shape[2], d.shape[1], add(d), add(m), add(c)]
        H[21:24] = [spec.shape[2], spec.shape[1], add(spec)]
        H[24], H[25] = add(self.phi_i), add(self.theta_i)
        # distr_2d.h:244-251: stride (in slices) of each parameter, 0 when its resolution is 1
        H[26] = n_theta if n_phi > 1 else 0
        H[27] = 1 if n_theta > 1 else 0
        if off >= (1 << 24):
            raise RuntimeError("measured_mono: tables too large for the device layout (>= 2^24 floats)")
        return np.concatenate([H] + parts).astype(np.float32)
