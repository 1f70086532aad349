#!/usr/bin/env python
"""
Fixtures for the `measured_mono` BSDF (ERP/bsdfs/measured_mono.cpp): the reference's tests need a file of the RGL
material database that cannot be fetched here (test_measured_mono.py:19-32 skips without it), so

  1. two small SYNTHETIC tensor files in the RGL layout are written (tests/golden/measured_iso.bsdf: isotropic,
     one phi_i; measured_aniso.bsdf: four phi_i over [0, pi] -> reduction 2), smooth positive random tables;
  2. the compiled reference (oracle/_ref, scalar_mono_double) loads them through its own TensorFile / Marginal2D
     code and evaluates eval / pdf / sample on fixed directions and samples at two wavelengths;
  3. the numbers go to tests/golden/measured_mono_reference.json.

The numpy restatement (oracle/measured_mono.py) and the device code are pinned on these values.
"""
import json
import os
import struct
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
DTYPES = {np.dtype("uint8"): 1, np.dtype("float32"): 10}


def write_tensor_file(path, fields):
    """MI/src/core/tensor.cpp:12-57: 'tensor_file\\0', version (1, 0), n_fields, then per field
    name_len u16 | name | ndim u16 | dtype u8 | offset u64 | shape u64[ndim]; data blocks follow."""
    header = bytearray(b"tensor_file\0") + bytes([1, 0]) + struct.pack("<I", len(fields))
    size = len(header) + sum(2 + len(n.encode()) + 2 + 1 + 8 + 8 * a.ndim for n, a in fields.items())
    blocks, recs = [], bytearray()
    offset = (size + 7) // 8 * 8
    for name, a in fields.items():
        a = np.ascontiguousarray(a)
        nb = name.encode()
        recs += struct.pack("<H", len(nb)) + nb + struct.pack("<HBQ", a.ndim, DTYPES[a.dtype], offset)
        recs += b"".join(struct.pack("<Q", s) for s in a.shape)
        blocks.append((offset, a.tobytes()))
        offset = (offset + a.nbytes + 7) // 8 * 8
    with open(path, "wb") as f:
        f.write(header + recs)
        for off, b in blocks:
            f.seek(off)
            f.write(b)


def synthetic(n_phi, seed, res=8, n_theta=6, n_wav=5):
    rng = np.random.default_rng(seed)

    def smooth(*shape):
        a = rng.uniform(0.3, 1.0, shape)
        yy, xx = np.meshgrid(np.linspace(0, 1, shape[-2]), np.linspace(0, 1, shape[-1]), indexing="ij")
        return (a * (0.4 + np.exp(-3.0 * ((xx - 0.4) ** 2 + (yy - 0.5) ** 2)))).astype(np.float32)

    phi_i = np.array([0.0], np.float32) if n_phi == 1 else np.linspace(0.0, np.pi, n_phi).astype(np.float32)
    theta_i = np.linspace(0.0, 1.45, n_theta).astype(np.float32)
    return {
        "description": np.frombuffer(b"synthetic material (eradiate_b200 test fixture)", dtype=np.uint8).copy(),
        "theta_i": theta_i, "phi_i": phi_i,
        "wavelengths": np.linspace(400.0, 800.0, n_wav).astype(np.float32),
        "ndf": smooth(res, res), "sigma": (0.5 + smooth(res, res)).astype(np.float32),
        "vndf": smooth(n_phi, n_theta, res, res), "luminance": smooth(n_phi, n_theta, res, res),
        "spectra": (0.2 * smooth(n_phi, n_theta, n_wav, res, res)).astype(np.float32),
        "jacobian": np.array([1], np.uint8),
    }


def directions(n, seed):
    rng = np.random.default_rng(seed)
    th, ph = rng.uniform(0.02, 1.4, n), rng.uniform(-np.pi, np.pi, n)
    return np.stack([np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)], axis=1)


def main():
    from oracle import ref

    files = {"measured_iso.bsdf": synthetic(1, 11), "measured_aniso.bsdf": synthetic(4, 12)}
    for name, fields in files.items():
        write_tensor_file(os.path.join(GOLD, name), fields)
    mi = ref.mitsuba("scalar_mono_double")
    import drjit as dr

    out = {"generator": "tools/make_measured_fixture.py", "reference": ref.describe(), "cases": []}
    n = 48
    wi, wo = directions(n, 1), directions(n, 2)
    u = np.random.default_rng(3).uniform(0.02, 0.98, (n, 2))
    for fname in files:
        for w in (450.0, 725.0):
            bsdf = mi.load_dict({"type": "measured_mono", "filename": os.path.join(GOLD, fname), "wavelength": w})
            ctx = mi.BSDFContext()
            ev, pdf, swo, sw, spdf = [], [], [], [], []
            for k in range(n):
                si = dr.zeros(mi.SurfaceInteraction3f)
                si.wi = mi.Vector3f(*wi[k])
                ev.append(float(bsdf.eval(ctx, si, mi.Vector3f(*wo[k]))[0]))
                pdf.append(float(bsdf.pdf(ctx, si, mi.Vector3f(*wo[k]))))
                bs, wgt = bsdf.sample(ctx, si, 0.5, mi.Point2f(*u[k]))
                swo.append([float(bs.wo.x), float(bs.wo.y), float(bs.wo.z)])
                sw.append(float(wgt[0]))
                spdf.append(float(bs.pdf))
            out["cases"].append({"file": fname, "wavelength": w, "wi": wi.tolist(), "wo": wo.tolist(), "u": u.tolist(),
                                 "eval": ev, "pdf": pdf, "sample_wo": swo, "sample_weight": sw, "sample_pdf": spdf})
            print(fname, w, "eval[0:3]", ev[:3], "pdf[0:3]", pdf[:3])
    with open(os.path.join(GOLD, "measured_mono_reference.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote measured_mono_reference.json")


if __name__ == "__main__":
    main()
