// ertb_measured.cuh -- the `measured_mono` BSDF (ERP/bsdfs/measured_mono.cpp:234-447) on the device.
//
// The plugin evaluates five piecewise-bilinear interpolants of the RGL material format (Dupuy & Jakob 2018):
// ndf and sigma (no parameter, evaluated only), vndf and luminance (parameters phi_i, theta_i; normalised, with
// marginal / conditional CDFs for sample warping and its inverse) and spectra (phi_i, theta_i, wavelength;
// evaluated only -- the host blends its wavelength slices, eradiate_b200/kernel/_measured.py, so the device sees
// two parameters there too).  The functions below are Marginal2D<Dim, Continuous = true>
// (MI/include/mitsuba/core/distr_2d.h): eval :1058-1090, sample_continuous :1288-1377, invert_continuous
// :1379-1453, sample_segment / invert_segment :1455-1470, parameter weights :255-292, lookup :1117-1138.
// All tables live in ONE float array in global memory (L2-resident), P.ocean_tables; layout in _measured.py.
#pragma once

#include "ertb_device.cuh"

struct MmWeights { // the <= 4 slices a (phi_i, theta_i) pair interpolates between, and their weights
    int slice[4];
    float w[4];
    int n;
};

__device__ __forceinline__ int mm_i(const float *T, int k) { return (int) __ldg(T + k); }

// Distribution2D::interpolate_weights for the two parameters (phi_i, theta_i)
__device__ __forceinline__ MmWeights mm_weights(const float *T, float phi, float theta) {
    const int n_par[2] = { mm_i(T, 0), mm_i(T, 1) }, off[2] = { mm_i(T, 24), mm_i(T, 25) }, stride[2] = { mm_i(T, 26), mm_i(T, 27) };
    const float par[2] = { phi, theta };
    int base = 0, d_slice[2] = { 0, 0 };
    float w1[2] = { 0.f, 0.f };
#pragma unroll
    for (int d = 0; d < 2; ++d) {
        if (n_par[d] == 1) continue;
        // math::find_interval: the last index whose value is < param, clamped to [0, n - 2]
        int lo = 0, hi = n_par[d] - 1;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(T + off[d] + mid) < par[d]) lo = mid; else hi = mid;
        }
        const float p0 = __ldg(T + off[d] + lo), p1 = __ldg(T + off[d] + lo + 1);
        w1[d] = fminf(fmaxf((par[d] - p0) / (p1 - p0), 0.f), 1.f);
        base += stride[d] * lo;
        d_slice[d] = stride[d];
    }
    MmWeights W;
    W.n = 4;
    W.slice[0] = base;                         W.w[0] = (1.f - w1[0]) * (1.f - w1[1]);
    W.slice[1] = base + d_slice[1];            W.w[1] = (1.f - w1[0]) * w1[1];
    W.slice[2] = base + d_slice[0];            W.w[2] = w1[0] * (1.f - w1[1]);
    W.slice[3] = base + d_slice[0] + d_slice[1]; W.w[3] = w1[0] * w1[1];
    return W;
}
__device__ __forceinline__ MmWeights mm_weights_none() {
    MmWeights W;
    W.n = 1; W.slice[0] = 0; W.w[0] = 1.f;
    W.slice[1] = W.slice[2] = W.slice[3] = 0; W.w[1] = W.w[2] = W.w[3] = 0.f;
    return W;
}
// Marginal2D::lookup: weighted sum over the neighbouring slices; `n` = entries per slice
__device__ __forceinline__ float mm_lookup(const float *tab, int n, const MmWeights &W, int idx) {
    float v = 0.f;
    for (int k = 0; k < W.n; ++k)
        if (W.w[k] != 0.f) v = fmaf(W.w[k], __ldg(tab + (size_t) W.slice[k] * n + idx), v);
    return v;
}

__device__ __forceinline__ float mm_eval2d(const float *data, int w, int h, const MmWeights &W, float px, float py) {
    px = fminf(fmaxf(px, 0.f), 1.f) * (float) (w - 1);
    py = fminf(fmaxf(py, 0.f), 1.f) * (float) (h - 1);
    const int ox = min((int) px, w - 2), oy = min((int) py, h - 2);
    px -= (float) ox; py -= (float) oy;
    const int i = ox + oy * w, n = w * h;
    const float v00 = mm_lookup(data, n, W, i), v10 = mm_lookup(data, n, W, i + 1);
    const float v01 = mm_lookup(data, n, W, i + w), v11 = mm_lookup(data, n, W, i + w + 1);
    const float a = fmaf(px, v10 - v00, v00), b = fmaf(px, v11 - v01, v01);
    return fmaf(py, b - a, a);
}

__device__ __forceinline__ float mm_sample_segment(float s, float inv_width, float v0, float v1) {
    const bool non_const = fabsf(v0 - v1) > 1e-4f * (v0 + v1);
    const float divisor = non_const ? v0 - v1 : v0 + v1;
    s *= 2.f * inv_width;
    if (non_const) s = v0 - sqrtf(fmaxf(fmaf(s, v1 - v0, v0 * v0), 0.f));
    if (divisor != 0.f) s /= divisor;
    return s;
}

// sample_continuous: (sx, sy) uniform -> position in [0,1]^2 and its density
__device__ __forceinline__ float mm_sample2d(const float *data, const float *marg, const float *cond, int w, int h,
                                             const MmWeights &W, float &sx, float &sy) {
    const int n_cond = h * (w - 1), n_marg = h - 1, n_data = w * h;
    sx = fminf(fmaxf(sx, 5.9604645e-8f), 1.f - 5.9604645e-8f);
    sy = fminf(fmaxf(sy, 5.9604645e-8f), 1.f - 5.9604645e-8f);
    // first row whose marginal CDF is >= sy, in [0, n_marg - 1]
    int lo = 0, hi = n_marg - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (mm_lookup(marg, n_marg, W, mid) < sy) lo = mid + 1; else hi = mid;
    }
    const int row = lo;
    if (row > 0) sy -= mm_lookup(marg, n_marg, W, row - 1);
    const int base = row * (w - 1);
    const float r0 = mm_lookup(cond, n_cond, W, base + (w - 1) - 1), r1 = mm_lookup(cond, n_cond, W, base + 2 * (w - 1) - 1);
    sy = mm_sample_segment(sy, (float) (h - 1), r0, r1);
    sx *= fmaf(sy, r1 - r0, r0);
    // first column whose (row-interpolated) conditional CDF is >= sx, in [0, w - 1] (clamped to a valid patch)
    lo = 0; hi = w - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const float v0 = mm_lookup(cond, n_cond, W, base + mid), v1 = mm_lookup(cond, n_cond, W, base + mid + (w - 1));
        if (fmaf(sy, v1 - v0, v0) < sx) lo = mid + 1; else hi = mid;
    }
    const int col = min(lo, w - 2);
    if (col > 0) {
        const float v0 = mm_lookup(cond, n_cond, W, base + col - 1), v1 = mm_lookup(cond, n_cond, W, base + col - 1 + (w - 1));
        sx -= fmaf(sy, v1 - v0, v0);
    }
    const int i = row * w + col;
    const float v00 = mm_lookup(data, n_data, W, i), v10 = mm_lookup(data, n_data, W, i + 1);
    const float v01 = mm_lookup(data, n_data, W, i + w), v11 = mm_lookup(data, n_data, W, i + w + 1);
    const float c0 = fmaf(sy, v01 - v00, v00), c1 = fmaf(sy, v11 - v10, v10);
    sx = mm_sample_segment(sx, (float) (w - 1), c0, c1);
    const float pdf = fmaf(sx, c1 - c0, c0);
    sx = ((float) col + sx) / (float) (w - 1);
    sy = ((float) row + sy) / (float) (h - 1);
    return pdf;
}

// invert_continuous: position -> the uniform sample that maps to it, and the density there
__device__ __forceinline__ float mm_invert2d(const float *data, const float *marg, const float *cond, int w, int h,
                                             const MmWeights &W, float &x, float &y) {
    const int n_cond = h * (w - 1), n_marg = h - 1, n_data = w * h;
    x = fminf(fmaxf(x, 0.f), 1.f) * (float) (w - 1);
    y = fminf(fmaxf(y, 0.f), 1.f) * (float) (h - 1);
    const int px = min((int) x, w - 2), py = min((int) y, h - 2);
    x -= (float) px; y -= (float) py;
    const int i = py * w + px;
    const float v00 = mm_lookup(data, n_data, W, i), v10 = mm_lookup(data, n_data, W, i + 1);
    const float v01 = mm_lookup(data, n_data, W, i + w), v11 = mm_lookup(data, n_data, W, i + w + 1);
    const float c0 = fmaf(y, v01 - v00, v00), c1 = fmaf(y, v11 - v10, v10);
    const float pdf = fmaf(x, c1 - c0, c0);
    x = x * fmaf(0.5f * x, c1 - c0, c0) / (float) (w - 1); // invert_segment
    const int base = py * (w - 1);
    if (px > 0) {
        const float v0 = mm_lookup(cond, n_cond, W, base + px - 1), v1 = mm_lookup(cond, n_cond, W, base + px - 1 + (w - 1));
        x += fmaf(y, v1 - v0, v0);
    }
    const float r0 = mm_lookup(cond, n_cond, W, base + (w - 1) - 1), r1 = mm_lookup(cond, n_cond, W, base + 2 * (w - 1) - 1);
    x /= fmaf(y, r1 - r0, r0);
    y = y * fmaf(0.5f * y, r1 - r0, r0) / (float) (h - 1);
    if (py > 0) y += mm_lookup(marg, n_marg, W, py - 1);
    return pdf; // (vndf and luminance are normalised: no division by the last marginal entry)
}

// ---- measured_mono.cpp helpers ------------------------------------------------------------------------------------
__device__ __forceinline__ float mm_elevation(f3 d) { // :226-232, stable acos(d.z)
    const float dist = sqrtf(d.x * d.x + d.y * d.y + (d.z - 1.f) * (d.z - 1.f));
    return 2.f * asinf(fminf(fmaxf(0.5f * dist, -1.f), 1.f));
}
__device__ __forceinline__ float mm_mulsign_neg(float x, float s) { return signbit(s) ? x : -x; } // x * -sign(s)
__device__ __forceinline__ float mm_theta2u(float t) { return sqrtf(t * (2.f * ERTB_INV_PI)); }
__device__ __forceinline__ float mm_phi2u(float p) { return (p + ERTB_PI) * (0.5f * ERTB_INV_PI); }

// BSDF::eval (:339-393).  The tables hold f * cos(theta_o): no cosine factor is applied.
// (not inlined: a rarely used BSDF whose loops would otherwise weigh on the register allocation of every kernel that
// can meet it -- the 3D canopy kernel lost 10 % on C4 with it inline)
__device__ __noinline__ float mm_eval(const ErtbParams &P, f3 wi, f3 wo) {
    if (!(wi.z > 0.f && wo.z > 0.f)) return 0.f;
    const float *T = P.ocean_tables;
    const int reduction = mm_i(T, 4);
    if (reduction >= 2) {
        const float sy = wi.y, sx = reduction == 4 ? wi.x : sy;
        wi.x = mm_mulsign_neg(wi.x, sx); wi.y = mm_mulsign_neg(wi.y, sy);
        wo.x = mm_mulsign_neg(wo.x, sx); wo.y = mm_mulsign_neg(wo.y, sy);
    }
    const f3 m = normalize3(mk3(wo.x + wi.x, wo.y + wi.y, wo.z + wi.z));
    const float theta_i = mm_elevation(wi), phi_i = atan2f(wi.y, wi.x), theta_m = mm_elevation(m), phi_m = atan2f(m.y, m.x);
    const float u_wi_x = mm_theta2u(theta_i), u_wi_y = mm_phi2u(phi_i);
    float um_x = mm_theta2u(theta_m), um_y = mm_phi2u(mm_i(T, 2) ? phi_m - phi_i : phi_m);
    um_y -= floorf(um_y);
    const MmWeights W = mm_weights(T, phi_i, theta_i), W0 = mm_weights_none();
    float sx = um_x, sy = um_y;
    mm_invert2d(T + mm_i(T, 13), T + mm_i(T, 14), T + mm_i(T, 15), mm_i(T, 11), mm_i(T, 12), W, sx, sy);
    float spec = mm_eval2d(T + mm_i(T, 23), mm_i(T, 21), mm_i(T, 22), W, sx, sy);
    if (mm_i(T, 3))
        spec *= mm_eval2d(T + mm_i(T, 7), mm_i(T, 5), mm_i(T, 6), W0, um_x, um_y) /
                (4.f * mm_eval2d(T + mm_i(T, 10), mm_i(T, 8), mm_i(T, 9), W0, u_wi_x, u_wi_y));
    return spec;
}

// BSDF::sample (:234-337): returns the weight spec / pdf; `pdf` is set for callers that want it
__device__ __noinline__ float mm_sample(const ErtbParams &P, f3 wi, float u1, float u2, f3 &wo, float &pdf) {
    wo = mk3(0.f, 0.f, 1.f);
    pdf = 0.f;
    if (!(wi.z > 0.f)) return 0.f;
    const float *T = P.ocean_tables;
    const int reduction = mm_i(T, 4);
    float rsx = -1.f, rsy = -1.f;
    if (reduction >= 2) {
        rsy = wi.y; rsx = reduction == 4 ? wi.x : rsy;
        wi.x = mm_mulsign_neg(wi.x, rsx); wi.y = mm_mulsign_neg(wi.y, rsy);
    }
    const float theta_i = mm_elevation(wi), phi_i = atan2f(wi.y, wi.x);
    const float u_wi_x = mm_theta2u(theta_i), u_wi_y = mm_phi2u(phi_i);
    const MmWeights W = mm_weights(T, phi_i, theta_i), W0 = mm_weights_none();
    float sx = u2, sy = u1; // sample = (sample2.y, sample2.x), :263
    const float lum_pdf = mm_sample2d(T + mm_i(T, 18), T + mm_i(T, 19), T + mm_i(T, 20), mm_i(T, 16), mm_i(T, 17), W, sx, sy);
    float mx = sx, my = sy;
    const float ndf_pdf = mm_sample2d(T + mm_i(T, 13), T + mm_i(T, 14), T + mm_i(T, 15), mm_i(T, 11), mm_i(T, 12), W, mx, my);
    float phi_m = (2.f * my - 1.f) * ERTB_PI;
    const float theta_m = mx * mx * (0.5f * ERTB_PI);
    if (mm_i(T, 2)) phi_m += phi_i;
    float sp, cp, st, ct;
    sincosf(phi_m, &sp, &cp);
    sincosf(theta_m, &st, &ct);
    const f3 m = mk3(cp * st, sp * st, ct);
    const float wim = dot3(wi, m);
    const float jac = fmaxf(2.f * ERTB_PI * ERTB_PI * mx * st, 1e-6f) * 4.f * wim;
    wo = mk3(fmaf(m.x, 2.f * wim, -wi.x), fmaf(m.y, 2.f * wim, -wi.y), fmaf(m.z, 2.f * wim, -wi.z));
    pdf = ndf_pdf * lum_pdf / jac;
    float spec = mm_eval2d(T + mm_i(T, 23), mm_i(T, 21), mm_i(T, 22), W, sx, sy);
    if (mm_i(T, 3))
        spec *= mm_eval2d(T + mm_i(T, 7), mm_i(T, 5), mm_i(T, 6), W0, mx, my) /
                (4.f * mm_eval2d(T + mm_i(T, 10), mm_i(T, 8), mm_i(T, 9), W0, u_wi_x, u_wi_y));
    wo.x = mm_mulsign_neg(wo.x, rsx); wo.y = mm_mulsign_neg(wo.y, rsy);
    if (!(wo.z > 0.f)) return 0.f;
    return spec / pdf;
}
