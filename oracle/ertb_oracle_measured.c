/* ertb_oracle_measured.c -- `measured_mono` BSDF restatement in double precision (TEST INFRASTRUCTURE ONLY).
 *
 * Reads the flattened table the host builds (layout: eradiate_b200/kernel/_measured.py; the descriptor only carries
 * that form) and evaluates it the way the plugin does:
 *   Marginal2D<Dim, Continuous = true>   MI/include/mitsuba/core/distr_2d.h (parameter weights :255-292, lookup
 *     :1117-1138, eval :1058-1090, sample_continuous :1288-1377, invert_continuous :1379-1453, segments :1455-1470)
 *   MeasuredMono::sample / eval          ERP/bsdfs/measured_mono.cpp:234-393
 * Pinned by tests/test_measured_mono.py on values the compiled reference returns for two synthetic tensor files
 * (tests/golden/measured_mono_reference.json) -- together with oracle/measured_mono.py, which restates the same
 * algorithm from the TENSOR FILE (its own table construction), so that the host's flattening is checked as well. */
#include <math.h>
#include <stddef.h>

#include "ertb_oracle_measured.h"

#define MM_PI 3.14159265358979323846

typedef struct { /* one interpolant inside the table */
    const float *data, *marg, *cond;
    int w, h;
    int off[4];     /* slice index of the <= 4 neighbours in parameter space */
    double wt[4];   /* and their weights (0 when unused) */
} m2d_t;

static int hdr(const float *T, int k) { return (int) T[k]; }

/* the slices a (phi_i, theta_i) pair interpolates between */
static void m2d_bind(m2d_t *m, const float *T, int use_param, double phi, double theta) {
    for (int k = 0; k < 4; ++k) { m->off[k] = 0; m->wt[k] = 0.0; }
    m->wt[0] = 1.0;
    if (!use_param) return;
    const int n[2] = { hdr(T, 0), hdr(T, 1) }, stride[2] = { hdr(T, 26), hdr(T, 27) };
    const float *val[2] = { T + hdr(T, 24), T + hdr(T, 25) };
    const double p[2] = { phi, theta };
    int base = 0, step[2] = { 0, 0 };
    double w1[2] = { 0.0, 0.0 };
    for (int d = 0; d < 2; ++d) {
        if (n[d] == 1) continue;
        int idx = 0; /* math::find_interval: last node below the parameter, clamped to [0, n - 2] */
        while (idx < n[d] - 2 && (double) val[d][idx + 1] < p[d]) ++idx;
        double t = (p[d] - val[d][idx]) / ((double) val[d][idx + 1] - val[d][idx]);
        w1[d] = t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t);
        base += stride[d] * idx;
        step[d] = stride[d];
    }
    m->off[0] = base;                     m->wt[0] = (1 - w1[0]) * (1 - w1[1]);
    m->off[1] = base + step[1];           m->wt[1] = (1 - w1[0]) * w1[1];
    m->off[2] = base + step[0];           m->wt[2] = w1[0] * (1 - w1[1]);
    m->off[3] = base + step[0] + step[1]; m->wt[3] = w1[0] * w1[1];
}
static double m2d_at(const m2d_t *m, const float *tab, int per_slice, int idx) {
    double v = 0.0;
    for (int k = 0; k < 4; ++k)
        if (m->wt[k] != 0.0) v += m->wt[k] * (double) tab[(size_t) m->off[k] * per_slice + idx];
    return v;
}
static double clamp01(double x) { return x < 0.0 ? 0.0 : (x > 1.0 ? 1.0 : x); }

static double m2d_eval(const m2d_t *m, double x, double y) {
    const int w = m->w, h = m->h, n = w * h;
    x = clamp01(x) * (w - 1); y = clamp01(y) * (h - 1);
    int ox = (int) x, oy = (int) y;
    if (ox > w - 2) ox = w - 2;
    if (oy > h - 2) oy = h - 2;
    x -= ox; y -= oy;
    int i = ox + oy * w;
    double v00 = m2d_at(m, m->data, n, i), v10 = m2d_at(m, m->data, n, i + 1);
    double v01 = m2d_at(m, m->data, n, i + w), v11 = m2d_at(m, m->data, n, i + w + 1);
    return (v00 * (1 - x) + v10 * x) * (1 - y) + (v01 * (1 - x) + v11 * x) * y;
}
static double seg_sample(double s, double inv_width, double v0, double v1) {
    int non_const = fabs(v0 - v1) > 1e-4 * (v0 + v1);
    double divisor = non_const ? v0 - v1 : v0 + v1;
    s *= 2.0 * inv_width;
    if (non_const) {
        double q = v0 * v0 + s * (v1 - v0);
        s = v0 - sqrt(q > 0.0 ? q : 0.0);
    }
    if (divisor != 0.0) s /= divisor;
    return s;
}
static double m2d_cond(const m2d_t *m, int base, int idx, double y) { /* conditional CDF between two rows */
    if (idx < 0) return 0.0;
    const int n = m->h * (m->w - 1);
    double v0 = m2d_at(m, m->cond, n, base + idx), v1 = m2d_at(m, m->cond, n, base + idx + (m->w - 1));
    return v0 + (v1 - v0) * y;
}
/* (x, y) uniform in, position out; returns the density */
static double m2d_sample(const m2d_t *m, double *x, double *y) {
    const int w = m->w, h = m->h, n_marg = h - 1;
    const double eps = 1.1102230246251565e-16;
    double sx = *x < eps ? eps : (*x > 1 - eps ? 1 - eps : *x), sy = *y < eps ? eps : (*y > 1 - eps ? 1 - eps : *y);
    int row = 0;
    while (row < h - 2 && m2d_at(m, m->marg, n_marg, row) < sy) ++row;
    if (row > 0) sy -= m2d_at(m, m->marg, n_marg, row - 1);
    const int base = row * (w - 1);
    double r0 = m2d_cond(m, base, w - 2, 0.0), r1 = m2d_cond(m, base, w - 2, 1.0);
    sy = seg_sample(sy, h - 1, r0, r1);
    sx *= r0 + (r1 - r0) * sy;
    int col = 0;
    while (col < w - 1 && m2d_cond(m, base, col, sy) < sx) ++col;
    if (col > w - 2) col = w - 2;
    sx -= m2d_cond(m, base, col - 1, sy);
    const int i = row * w + col, n = w * h;
    double v00 = m2d_at(m, m->data, n, i), v10 = m2d_at(m, m->data, n, i + 1);
    double v01 = m2d_at(m, m->data, n, i + w), v11 = m2d_at(m, m->data, n, i + w + 1);
    double c0 = v00 + (v01 - v00) * sy, c1 = v10 + (v11 - v10) * sy;
    sx = seg_sample(sx, w - 1, c0, c1);
    *x = (col + sx) / (w - 1);
    *y = (row + sy) / (h - 1);
    return c0 + (c1 - c0) * sx;
}
/* position in, the uniform sample that maps to it out; returns the density */
static double m2d_invert(const m2d_t *m, double *x, double *y) {
    const int w = m->w, h = m->h, n = w * h;
    double px = clamp01(*x) * (w - 1), py = clamp01(*y) * (h - 1);
    int ix = (int) px, iy = (int) py;
    if (ix > w - 2) ix = w - 2;
    if (iy > h - 2) iy = h - 2;
    px -= ix; py -= iy;
    const int i = iy * w + ix;
    double v00 = m2d_at(m, m->data, n, i), v10 = m2d_at(m, m->data, n, i + 1);
    double v01 = m2d_at(m, m->data, n, i + w), v11 = m2d_at(m, m->data, n, i + w + 1);
    double c0 = v00 + (v01 - v00) * py, c1 = v10 + (v11 - v10) * py;
    double pdf = c0 + (c1 - c0) * px;
    const int base = iy * (w - 1);
    double sx = px * (c0 + 0.5 * px * (c1 - c0)) / (w - 1) + m2d_cond(m, base, ix - 1, py);
    double r0 = m2d_cond(m, base, w - 2, 0.0), r1 = m2d_cond(m, base, w - 2, 1.0);
    sx /= r0 + (r1 - r0) * py;
    double sy = py * (r0 + 0.5 * py * (r1 - r0)) / (h - 1);
    if (iy > 0) sy += m2d_at(m, m->marg, h - 1, iy - 1);
    *x = sx; *y = sy; /* vndf / luminance are normalised: the last marginal entry is 1 */
    return pdf;
}

typedef struct { m2d_t ndf, sigma, vndf, lum, spectra; int isotropic, jacobian, reduction; } mm_t;

static void mm_bind(mm_t *M, const float *T, double phi_i, double theta_i) {
    M->isotropic = hdr(T, 2); M->jacobian = hdr(T, 3); M->reduction = hdr(T, 4);
    M->ndf.w = hdr(T, 5); M->ndf.h = hdr(T, 6); M->ndf.data = T + hdr(T, 7); M->ndf.marg = M->ndf.cond = NULL;
    M->sigma.w = hdr(T, 8); M->sigma.h = hdr(T, 9); M->sigma.data = T + hdr(T, 10); M->sigma.marg = M->sigma.cond = NULL;
    M->vndf.w = hdr(T, 11); M->vndf.h = hdr(T, 12);
    M->vndf.data = T + hdr(T, 13); M->vndf.marg = T + hdr(T, 14); M->vndf.cond = T + hdr(T, 15);
    M->lum.w = hdr(T, 16); M->lum.h = hdr(T, 17);
    M->lum.data = T + hdr(T, 18); M->lum.marg = T + hdr(T, 19); M->lum.cond = T + hdr(T, 20);
    M->spectra.w = hdr(T, 21); M->spectra.h = hdr(T, 22); M->spectra.data = T + hdr(T, 23); M->spectra.marg = M->spectra.cond = NULL;
    m2d_bind(&M->ndf, T, 0, 0, 0);
    m2d_bind(&M->sigma, T, 0, 0, 0);
    m2d_bind(&M->vndf, T, 1, phi_i, theta_i);
    m2d_bind(&M->lum, T, 1, phi_i, theta_i);
    m2d_bind(&M->spectra, T, 1, phi_i, theta_i);
}
static double elevation(const double d[3]) { /* measured_mono.cpp:226-232 */
    double dist = sqrt(d[0] * d[0] + d[1] * d[1] + (d[2] - 1.0) * (d[2] - 1.0)), h = 0.5 * dist;
    return 2.0 * asin(h > 1.0 ? 1.0 : h);
}
static double flip(double x, double s) { return signbit(s) ? x : -x; } /* dr::mulsign_neg: x * -sign(s) */
static double theta2u(double t) { return sqrt(t * (2.0 / MM_PI)); }
static double phi2u(double p) { return (p + MM_PI) / (2.0 * MM_PI); }

double mm_oracle_eval(const float *T, const double wi_in[3], const double wo_in[3]) {
    if (!(wi_in[2] > 0.0 && wo_in[2] > 0.0)) return 0.0;
    double wi[3] = { wi_in[0], wi_in[1], wi_in[2] }, wo[3] = { wo_in[0], wo_in[1], wo_in[2] };
    const int reduction = hdr(T, 4);
    if (reduction >= 2) { /* :363-371 */
        double sy = wi[1], sx = reduction == 4 ? wi[0] : sy;
        wi[0] = flip(wi[0], sx); wi[1] = flip(wi[1], sy);
        wo[0] = flip(wo[0], sx); wo[1] = flip(wo[1], sy);
    }
    double m[3] = { wo[0] + wi[0], wo[1] + wi[1], wo[2] + wi[2] };
    double nm = sqrt(m[0] * m[0] + m[1] * m[1] + m[2] * m[2]);
    for (int k = 0; k < 3; ++k) m[k] /= nm;
    double theta_i = elevation(wi), phi_i = atan2(wi[1], wi[0]), theta_m = elevation(m), phi_m = atan2(m[1], m[0]);
    mm_t M;
    mm_bind(&M, T, phi_i, theta_i);
    double umx = theta2u(theta_m), umy = phi2u(M.isotropic ? phi_m - phi_i : phi_m);
    umy -= floor(umy);
    double sx = umx, sy = umy;
    m2d_invert(&M.vndf, &sx, &sy);
    double spec = m2d_eval(&M.spectra, sx, sy);
    if (M.jacobian) spec *= m2d_eval(&M.ndf, umx, umy) / (4.0 * m2d_eval(&M.sigma, theta2u(theta_i), phi2u(phi_i)));
    return spec;
}

/* MeasuredMono::pdf, :395-447 */
double mm_oracle_pdf(const float *T, const double wi_in[3], const double wo_in[3]) {
    if (!(wi_in[2] > 0.0 && wo_in[2] > 0.0)) return 0.0;
    double wi[3] = { wi_in[0], wi_in[1], wi_in[2] }, wo[3] = { wo_in[0], wo_in[1], wo_in[2] };
    const int reduction = hdr(T, 4);
    if (reduction >= 2) {
        double sy = wi[1], sx = reduction == 4 ? wi[0] : sy;
        wi[0] = flip(wi[0], sx); wi[1] = flip(wi[1], sy);
        wo[0] = flip(wo[0], sx); wo[1] = flip(wo[1], sy);
    }
    double m[3] = { wo[0] + wi[0], wo[1] + wi[1], wo[2] + wi[2] };
    double nm = sqrt(m[0] * m[0] + m[1] * m[1] + m[2] * m[2]);
    for (int k = 0; k < 3; ++k) m[k] /= nm;
    double theta_i = elevation(wi), phi_i = atan2(wi[1], wi[0]), theta_m = elevation(m), phi_m = atan2(m[1], m[0]);
    mm_t M;
    mm_bind(&M, T, phi_i, theta_i);
    double umx = theta2u(theta_m), umy = phi2u(M.isotropic ? phi_m - phi_i : phi_m);
    umy -= floor(umy);
    double sx = umx, sy = umy;
    double vndf_pdf = m2d_invert(&M.vndf, &sx, &sy);
    double lum = m2d_eval(&M.lum, sx, sy);
    double st = 1.0 - m[2] * m[2];
    double jac = 2.0 * MM_PI * MM_PI * umx * sqrt(st > 0.0 ? st : 0.0);
    jac = (jac > 1e-6 ? jac : 1e-6) * 4.0 * (wi[0] * m[0] + wi[1] * m[1] + wi[2] * m[2]);
    return vndf_pdf * lum / jac;
}

double mm_oracle_sample(const float *T, const double wi_in[3], double u1, double u2, double wo[3], double *pdf_out) {
    wo[0] = wo[1] = 0.0; wo[2] = 1.0;
    if (pdf_out) *pdf_out = 0.0;
    if (!(wi_in[2] > 0.0)) return 0.0;
    double wi[3] = { wi_in[0], wi_in[1], wi_in[2] }, fx = -1.0, fy = -1.0;
    const int reduction = hdr(T, 4);
    if (reduction >= 2) {
        fy = wi[1]; fx = reduction == 4 ? wi[0] : fy;
        wi[0] = flip(wi[0], fx); wi[1] = flip(wi[1], fy);
    }
    double theta_i = elevation(wi), phi_i = atan2(wi[1], wi[0]);
    mm_t M;
    mm_bind(&M, T, phi_i, theta_i);
    double sx = u2, sy = u1; /* :263: Point2f(sample2.y(), sample2.x()) */
    double lum_pdf = m2d_sample(&M.lum, &sx, &sy);
    double mx = sx, my = sy;
    double ndf_pdf = m2d_sample(&M.vndf, &mx, &my);
    double phi_m = (2.0 * my - 1.0) * MM_PI, theta_m = mx * mx * (0.5 * MM_PI);
    if (M.isotropic) phi_m += phi_i;
    double m[3] = { cos(phi_m) * sin(theta_m), sin(phi_m) * sin(theta_m), cos(theta_m) };
    double wim = wi[0] * m[0] + wi[1] * m[1] + wi[2] * m[2];
    double jac = 2.0 * MM_PI * MM_PI * mx * sin(theta_m);
    jac = (jac > 1e-6 ? jac : 1e-6) * 4.0 * wim;
    for (int k = 0; k < 3; ++k) wo[k] = 2.0 * wim * m[k] - wi[k];
    double pdf = ndf_pdf * lum_pdf / jac;
    double spec = m2d_eval(&M.spectra, sx, sy);
    if (M.jacobian) spec *= m2d_eval(&M.ndf, mx, my) / (4.0 * m2d_eval(&M.sigma, theta2u(theta_i), phi2u(phi_i)));
    wo[0] = flip(wo[0], fx); wo[1] = flip(wo[1], fy);
    if (pdf_out) *pdf_out = pdf;
    if (!(wo[2] > 0.0)) return 0.0;
    return spec / pdf;
}
