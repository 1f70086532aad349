// ertb_ocean.cuh -- the local-frame BSDFs of the sm_100a path tracer: ocean_legacy (6SV ocean model) first, then
// (end of the file) the isotropic-Beckmann glint family -- ocean_mishchenko, ocean_grasp, maignan -- and mqdiffuse.
//
// Reference: ERP/bsdfs/ocean_legacy.cpp (update :313-372, glint :405-447, underlight
// :449-491, sample :494-559, eval :561-661, pdf :663-713),
// MI/include/mitsuba/eradiate/oceanprops.h (water_ior, fresnel_sunglint_legacy, Cox-Munk /
// Gram-Charlier, r_omega, whitecaps) and MI/include/mitsuba/render/microfacet.h:195-530
// (rotated anisotropic Beckmann distribution, visible-normal sampling, Smith G1, height-
// correlated G).  Host part: per-wavelength scalars from the literature tables (Whitlock et
// al. 1982; Hale & Querry 1973; Morel 1988, as tabulated by 6SV) and a CUDA kernel that
// fills the two 64x64 transmittance tables (64x64 Gauss-Legendre quadrature per entry,
// ocean_legacy.cpp:137-243).  Device part: fp32 eval / pdf / sample in the local frame.
#pragma once

#include "ertb_device.cuh"
#include "ertb_measured.cuh"

#define ERTB_OC_RES 64

// indices into ErtbParams::bsdf for the ocean model (derived, device-side)
enum : int {
    OC_N_REAL = 0, OC_N_IMAG, OC_SIGMA_U, OC_SIGMA_C, OC_WIND_DIR, OC_R_OMEGA, OC_COVERAGE, OC_WHITECAP,
    OC_UNDERLIGHT_ON, OC_SHADOWING, OC_WIND_SPEED, OC_ALPHA_UP, OC_ALPHA_VP, OC_CORR, OC_UL_NORM
};

// ----------------------------------------------------------------------------- host side
namespace ertb_ocean_host {

static const double WC_DATA[39] = {
    0.220, 0.220, 0.220, 0.220, 0.220, 0.220, 0.215, 0.210, 0.200, 0.190, 0.175, 0.155, 0.130,
    0.080, 0.100, 0.105, 0.100, 0.080, 0.045, 0.055, 0.065, 0.060, 0.055, 0.040, 0.0, 0.0,
    0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0 };
static const double IOR_WL[64] = {
    200,  225,  250,  275,  300,  325,  345,  375,  400,  425,  445,  475,  500,  525,  550,  575,
    600,  625,  650,  675,  700,  725,  750,  775,  800,  825,  850,  875,  900,  925,  950,  975,
    1000, 1200, 1400, 1600, 1800, 2000, 2200, 2400, 2600, 2650, 2700, 2750, 2800, 2850, 2900, 2950,
    3000, 3050, 3100, 3150, 3200, 3250, 3300, 3350, 3400, 3450, 3500, 3600, 3700, 3800, 3900, 4000 };
static const double IOR_RE[64] = {
    1.369, 1.373, 1.362, 1.354, 1.349, 1.346, 1.343, 1.341, 1.339, 1.338, 1.337, 1.336, 1.335,
    1.334, 1.333, 1.333, 1.332, 1.332, 1.331, 1.331, 1.331, 1.330, 1.330, 1.330, 1.329, 1.329,
    1.329, 1.328, 1.328, 1.328, 1.327, 1.327, 1.327, 1.324, 1.321, 1.317, 1.312, 1.306, 1.296,
    1.279, 1.242, 1.219, 1.188, 1.157, 1.142, 1.149, 1.201, 1.292, 1.371, 1.426, 1.467, 1.483,
    1.478, 1.467, 1.450, 1.432, 1.420, 1.410, 1.400, 1.385, 1.374, 1.364, 1.357, 1.351 };
static const double IOR_IM[64] = {
    1.10e-07, 4.90e-08, 3.35e-08, 2.35e-08, 1.60e-08, 1.08e-08, 6.50e-09, 3.50e-09, 1.86e-09,
    1.30e-09, 1.02e-09, 9.35e-10, 1.00e-09, 1.32e-09, 1.96e-09, 3.60e-09, 1.09e-08, 1.39e-08,
    1.64e-08, 2.23e-08, 3.35e-08, 9.15e-08, 1.56e-07, 1.48e-07, 1.25e-07, 1.82e-07, 2.93e-07,
    3.91e-07, 4.86e-07, 1.06e-06, 2.93e-06, 3.48e-06, 2.89e-06, 9.89e-06, 1.38e-04, 8.55e-05,
    1.15e-04, 1.10e-03, 2.89e-04, 9.56e-04, 3.17e-03, 6.70e-03, 1.90e-02, 5.90e-02, 1.15e-01,
    1.85e-01, 2.68e-01, 2.98e-01, 2.72e-01, 2.40e-01, 1.92e-01, 1.35e-01, 9.24e-02, 6.10e-02,
    3.68e-02, 2.61e-02, 1.95e-02, 1.32e-02, 9.40e-03, 5.15e-03, 3.60e-03, 3.40e-03, 3.80e-03,
    4.60e-03 };
static const double ATTN_K[61] = {
    0.0209, 0.0200, 0.0196, 0.0189, 0.0183, 0.0182, 0.0171, 0.0170, 0.0168, 0.0166, 0.0168, 0.0170,
    0.0173, 0.0174, 0.0175, 0.0184, 0.0194, 0.0203, 0.0217, 0.0240, 0.0271, 0.0320, 0.0384, 0.0445,
    0.0490, 0.0505, 0.0518, 0.0543, 0.0568, 0.0615, 0.0640, 0.0640, 0.0717, 0.0762, 0.0807, 0.0940,
    0.1070, 0.1280, 0.1570, 0.2000, 0.2530, 0.2790, 0.2960, 0.3030, 0.3100, 0.3150, 0.3200, 0.3250,
    0.3300, 0.3400, 0.3500, 0.3700, 0.4050, 0.4180, 0.4300, 0.4400, 0.4500, 0.4700, 0.5000, 0.5500,
    0.6500 };
static const double ATTN_CHI[61] = {
    0.1100, 0.1110, 0.1125, 0.1135, 0.1126, 0.1104, 0.1078, 0.1065, 0.1041, 0.0996, 0.0971, 0.0939,
    0.0896, 0.0859, 0.0823, 0.0788, 0.0746, 0.0726, 0.0690, 0.0660, 0.0636, 0.0600, 0.0578, 0.0540,
    0.0498, 0.0475, 0.0467, 0.0450, 0.0440, 0.0426, 0.0410, 0.0400, 0.0390, 0.0375, 0.0360, 0.0340,
    0.0330, 0.0328, 0.0325, 0.0330, 0.0340, 0.0350, 0.0360, 0.0375, 0.0385, 0.0400, 0.0420, 0.0430,
    0.0440, 0.0445, 0.0450, 0.0460, 0.0475, 0.0490, 0.0515, 0.0520, 0.0505, 0.0440, 0.0390, 0.0340,
    0.0300 };
static const double ATTN_E[61] = {
    0.668, 0.672, 0.680, 0.687, 0.693, 0.701, 0.707, 0.708, 0.707, 0.704, 0.701, 0.699, 0.700, 0.703,
    0.703, 0.703, 0.703, 0.704, 0.702, 0.700, 0.700, 0.695, 0.690, 0.685, 0.680, 0.675, 0.670, 0.665,
    0.660, 0.655, 0.650, 0.645, 0.640, 0.630, 0.623, 0.615, 0.610, 0.614, 0.618, 0.622, 0.626, 0.630,
    0.634, 0.638, 0.642, 0.647, 0.653, 0.658, 0.663, 0.667, 0.672, 0.677, 0.682, 0.687, 0.695, 0.697,
    0.693, 0.665, 0.640, 0.620, 0.600 };
static const double MOL_6S[61] = {
    0.0076, 0.0072, 0.0068, 0.0064, 0.0061, 0.0058, 0.0055, 0.0052, 0.0049, 0.0047, 0.0045, 0.0043,
    0.0041, 0.0039, 0.0037, 0.0036, 0.0034, 0.0033, 0.0031, 0.0030, 0.0029, 0.0027, 0.0026, 0.0025,
    0.0024, 0.0023, 0.0022, 0.0022, 0.0021, 0.0020, 0.0019, 0.0018, 0.0018, 0.0017, 0.0017, 0.0016,
    0.0016, 0.0015, 0.0015, 0.0014, 0.0014, 0.0013, 0.0013, 0.0012, 0.0012, 0.0011, 0.0011, 0.0010,
    0.0010, 0.0010, 0.0010, 0.0009, 0.0008, 0.0008, 0.0008, 0.0007, 0.0007, 0.0007, 0.0007, 0.0007,
    0.0007 };

inline double lerp_regular(const double *y, int n, double x0, double x1, double x) {
    if (!(x >= x0 && x <= x1)) return 0.0;
    double xs = (x - x0) * (n - 1) / (x1 - x0);
    int i = (int) xs;
    i = i < 0 ? 0 : (i > n - 2 ? n - 2 : i);
    double w = xs - i;
    return (1.0 - w) * y[i] + w * y[i + 1];
}
inline double lerp_irregular(const double *xn, const double *y, int n, double x) {
    if (!(x >= xn[0] && x <= xn[n - 1])) return 0.0;
    int lo = 0, hi = n - 1;
    while (hi - lo > 1) {
        int mid = (lo + hi) / 2;
        if (xn[mid] < x) lo = mid; else hi = mid;
    }
    return y[lo] + (x - xn[lo]) / (xn[lo + 1] - xn[lo]) * (y[lo + 1] - y[lo]);
}

// raw plugin parameters (bsdf_params of the descriptor) -> derived device scalars
inline void derive(const float *raw, float *out, double &n_real, double &n_imag) {
    const double wl = raw[0], ws = raw[1], chl = raw[3], pig = raw[4];
    double wd = -(double) raw[2] + 90.0; // North-left -> East-right (ocean_legacy.cpp:275-280)
    wd = (wd - 360.0 * floor(wd / 360.0)) * M_PI / 180.0;
    n_real = lerp_irregular(IOR_WL, IOR_RE, 64, wl) + 0.00017492711 * (0.03 + 1.805 * chl);
    n_imag = lerp_irregular(IOR_WL, IOR_IM, 64, wl);
    const double su = sqrt(0.00316 * ws), sc = sqrt(0.003 + 0.00192 * ws);
    // r_omega (oceanprops.h:692-740)
    double r_om = 0.0;
    {
        double plog = log(pig) / log(10.0);
        double mol = lerp_regular(MOL_6S, 61, 400.0, 700.0, wl);
        double bb = 0.5 * mol + 0.30 * pow(pig, 0.62) * (0.002 + 0.02 * (0.5 - 0.25 * plog) * (550.0 / wl));
        double attn = lerp_regular(ATTN_K, 61, 400.0, 700.0, wl) +
                      lerp_regular(ATTN_CHI, 61, 400.0, 700.0, wl) * pow(pig, lerp_regular(ATTN_E, 61, 400.0, 700.0, wl));
        if (bb != 0.0 && attn != 0.0) {
            double u = 0.75, r = 0.33 * bb / u / attn;
            for (int it = 0; it < 1000; ++it) {
                u = (0.9 * (1.0 - r)) / (1.0 + 2.25 * r);
                double rn = 0.33 * bb / (u * attn);
                if (fabs((rn - r) / rn) < 0.0001) break;
                r = rn;
            }
            r_om = r;
        }
    }
    double cov = fmin(1.0, fmax(0.0, 2.95e-06 * pow(ws, 3.52)));
    double au = fmax(sqrt(2.0) * su, 1e-4), av = fmax(sqrt(2.0) * sc, 1e-4);
    double s = sin(wd), c = cos(wd);
    for (int i = 0; i < ERTB_MAX_BSDF_PARAMS; ++i) out[i] = 0.f;
    out[OC_N_REAL] = (float) n_real;
    out[OC_N_IMAG] = (float) n_imag;
    out[OC_SIGMA_U] = (float) su;
    out[OC_SIGMA_C] = (float) sc;
    out[OC_WIND_DIR] = (float) wd;
    out[OC_R_OMEGA] = (float) r_om;
    out[OC_COVERAGE] = (float) cov;
    out[OC_WHITECAP] = (float) (cov * lerp_regular(WC_DATA, 39, 200.0, 4000.0, wl));
    out[OC_UNDERLIGHT_ON] = (wl < 400.0 || wl > 700.0) ? 0.f : 1.f;
    out[OC_SHADOWING] = raw[5] != 0.f ? 1.f : 0.f;
    out[OC_WIND_SPEED] = (float) ws;
    out[OC_ALPHA_UP] = (float) sqrt((au * c) * (au * c) + (av * s) * (av * s));
    out[OC_ALPHA_VP] = (float) sqrt((au * s) * (au * s) + (av * c) * (av * c));
    out[OC_CORR] = (float) (2.0 * (au * au - av * av) * c * s);
    out[OC_UL_NORM] = (float) ((1.0 / (n_real * n_real + n_imag * n_imag)) * r_om / (1.0 - 0.485 * r_om));
}

inline void gauss_legendre(int n, double *x, double *w) {
    for (int i = 0; i < n; ++i) {
        double z = cos(M_PI * (i + 0.75) / (n + 0.5)), pp = 0.0;
        for (int it = 0; it < 100; ++it) {
            double p1 = 1.0, p2 = 0.0;
            for (int j = 1; j <= n; ++j) {
                double p3 = p2;
                p2 = p1;
                p1 = ((2.0 * j - 1.0) * z * p2 - (j - 1.0) * p3) / j;
            }
            pp = n * (z * p1 - p2) / (z * z - 1.0);
            double z1 = z;
            z = z1 - p1 / pp;
            if (fabs(z - z1) < 1e-15) break;
        }
        x[n - 1 - i] = z;
        w[n - 1 - i] = 2.0 / ((1.0 - z * z) * pp * pp);
    }
}
// ocean_mishchenko / ocean_grasp / maignan: raw plugin parameters -> the slots the device code reads
// (ocean_mishchenko.cpp:136-142, ocean_grasp.cpp:155-188, oceanprops.h:330-363, :582-584)
inline void derive_glint(int type, const float *raw, float *out) {
    for (int i = 0; i < ERTB_MAX_BSDF_PARAMS; ++i) out[i] = 0.f;
    double ws = 0.0, eta, k, ext;
    if (type == ERTB_BSDF_OCEAN_MISHCHENKO) { ws = raw[0]; eta = raw[1]; k = raw[2]; ext = raw[3]; }
    else if (type == ERTB_BSDF_OCEAN_GRASP) { ws = raw[1]; eta = raw[2]; k = raw[3]; ext = raw[4]; }
    else { eta = raw[2]; k = raw[3]; ext = raw[4]; }
    out[OC_N_REAL] = (float) (eta / ext);
    out[OC_N_IMAG] = (float) (k / ext);
    if (type == ERTB_BSDF_MAIGNAN) {
        out[15] = (float) ((double) raw[0] * exp(-(double) raw[1])); // MG_CEXP
        return;
    }
    const double sigma = sqrt(0.5 * (0.00512 * ws + 0.003)), alpha = fmax(sqrt(2.0) * sigma, 1e-4);
    out[OC_SIGMA_U] = out[OC_SIGMA_C] = (float) sigma;
    out[OC_ALPHA_UP] = out[OC_ALPHA_VP] = (float) alpha;
    out[OC_SHADOWING] = 1.f;
    out[OC_WIND_SPEED] = (float) ws;
    if (type == ERTB_BSDF_OCEAN_GRASP) {
        const double cov = fmin(1.0, fmax(0.0, 2.95e-06 * pow(ws, 3.52))), um = (double) raw[0] * 0.001;
        const double eff = um >= 0.6 ? 0.22 * exp(-1.75 * pow(um - 0.6, 0.99)) : 0.22;
        out[OC_COVERAGE] = (float) cov;
        out[OC_WHITECAP] = (float) (cov * eff);
        out[OC_R_OMEGA] = raw[5]; // water body reflectance
    }
}
} // namespace ertb_ocean_host

// ----------------------------------------------------------------- transmittance tables
__device__ inline double oc_fresnel_d(double nr, double ni, double coschi, double sinchi) {
    double nr2 = nr * nr, ni2 = ni * ni;
    double s = nr2 - ni2 - sinchi * sinchi;
    double a1 = fabs(s), a2 = sqrt(s * s + 4.0 * nr2 * ni2);
    double u = sqrt(0.5 * fabs(a1 + a2)), v = sqrt(0.5 * fabs(a2 - a1));
    double b1 = (nr2 - ni2) * coschi, b2 = 2.0 * nr * ni * coschi;
    double right = ((coschi - u) * (coschi - u) + v * v) / ((coschi + u) * (coschi + u) + v * v);
    double left = ((b1 - u) * (b1 - u) + (b2 + v) * (b2 + v)) / ((b1 + u) * (b1 + u) + (b2 - v) * (b2 - v));
    return 0.5 * (right + left);
}

// One thread per table entry; out[0..4095] = downwelling, out[4096..8191] = upwelling.
// Entry idx = i*64 + j  <->  (zenith j, azimuth i)  (meshgrid(zeniths, azimuths)).
__global__ void ertb_ocean_tables_kernel(double n_real, double n_imag, double wind_speed,
                                         const double *__restrict__ gl /* nodes[64] | weights[64] */,
                                         float *__restrict__ out) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * ERTB_OC_RES * ERTB_OC_RES) return;
    const bool up = t >= ERTB_OC_RES * ERTB_OC_RES;
    int idx = up ? t - ERTB_OC_RES * ERTB_OC_RES : t;
    int i = idx / ERTB_OC_RES, j = idx % ERTB_OC_RES;
    const double PI_D = 3.14159265358979323846;
    double theta = 0.5 * PI_D * j / (ERTB_OC_RES - 1), phi = 2.0 * PI_D * i / (ERTB_OC_RES - 1);
    double nr = n_real, ni = n_imag;
    if (up) {
        theta = asin(sin(theta) / nr);
        nr = 1.0 / nr;
        ni = 0.0;
    }
    const double su = sqrt(0.00316 * wind_speed), sc = sqrt(0.003 + 0.00192 * wind_speed);
    const double c21 = 0.01 - 0.0086 * wind_speed, c03 = 0.04 - 0.033 * wind_speed;
    const double sp = sin(phi), cp = cos(phi);
    const double wix = sin(theta), wiz = cos(theta);
    double td = 0.0, summ = 0.0;
    for (int iy = 0; iy < ERTB_OC_RES; ++iy) {
        double phi_o = gl[iy] * PI_D + PI_D, wy = PI_D * gl[ERTB_OC_RES + iy];
        double so = sin(phi_o), co = cos(phi_o);
        for (int ix = 0; ix < ERTB_OC_RES; ++ix) {
            double theta_o = gl[ix] * 0.25 * PI_D + 0.25 * PI_D, wx = 0.25 * PI_D * gl[ERTB_OC_RES + ix];
            double sz = sin(theta_o), cz = cos(theta_o);
            double gweight = cz * sz * wy * wx;
            double wox = sz * co, woy = sz * so, woz = cz;
            double cti = wiz < 1e-6 ? 1e-6 : wiz, cto = woz < 1e-6 ? 1e-6 : woz;
            double mx = wix + wox, my = woy, mz = wiz + woz;
            double inv = rsqrt(mx * mx + my * my + mz * mz);
            mx *= inv; my *= inv; mz *= inv;
            // cox_munk_anisotropic_distrib (oceanprops.h:585-630) with wind direction = phi
            double px = cp * mx + sp * my, py = -sp * mx + cp * my;
            double xn = px / (su * mz), xe = py / (sc * mz);
            double xe2 = xe * xe, xn2 = xn * xn;
            double coef = 1.0 - (c21 / 2.0) * (xe2 - 1.0) * xn - (c03 / 6.0) * (xn2 - 3.0) * xn;
            coef += (0.40 / 24.0) * (xe2 * xe2 - 6.0 * xe2 + 3.0);
            coef += (0.23 / 24.0) * (xn2 * xn2 - 6.0 * xn2 + 3.0);
            coef += (0.12 / 4.0) * (xe2 - 1.0) * (xn2 - 1.0);
            double D = fmax(coef / (2.0 * PI_D) / (su * sc) * exp(-(xe2 + xn2) * 0.5), 0.0);
            double mz2 = mz * mz;
            D /= mz2 * mz2;
            double cos_chi = fmin(fmax(wox * mx + woy * my + woz * mz, -0.999999999), 0.999999999);
            double sin_chi = fmin(sqrt(1.0 - cos_chi * cos_chi), 0.999999999);
            double glint = D * oc_fresnel_d(nr, ni, cos_chi, sin_chi) * PI_D / (4.0 * cti * cto);
            td += glint * gweight;
            summ += gweight;
        }
    }
    if (td >= summ) td = summ;
    out[t] = (float) (1.0 - td / summ);
}

// -------------------------------------------------------------------------- device side
__device__ __forceinline__ float oc_tex(const float *__restrict__ data, float u, float v) {
    // drjit texture.h:500-530, linear filter, clamp wrap; u -> zenith axis, v -> azimuth axis
    float pu = fmaf(u, (float) ERTB_OC_RES, -0.5f), pv = fmaf(v, (float) ERTB_OC_RES, -0.5f);
    float fu = floorf(pu), fv = floorf(pv);
    int iu = (int) fu, iv = (int) fv;
    float wu1 = pu - fu, wv1 = pv - fv;
    int u0 = min(max(iu, 0), ERTB_OC_RES - 1), u1 = min(max(iu + 1, 0), ERTB_OC_RES - 1);
    int v0 = min(max(iv, 0), ERTB_OC_RES - 1), v1 = min(max(iv + 1, 0), ERTB_OC_RES - 1);
    float a = __ldg(data + v0 * ERTB_OC_RES + u0), b = __ldg(data + v0 * ERTB_OC_RES + u1);
    float c = __ldg(data + v1 * ERTB_OC_RES + u0), d = __ldg(data + v1 * ERTB_OC_RES + u1);
    float r0 = fmaf(wu1, b - a, a), r1 = fmaf(wu1, d - c, c);
    return fmaf(wv1, r1 - r0, r0);
}
__device__ __forceinline__ float oc_transmittance(const ErtbParams &P, const float *data, float cos_theta, float vx, float vy) {
    float u = acosf(clampf(cos_theta, -1.f, 1.f)) * (2.f * ERTB_INV_PI);
    float v = (atan2f(vy, vx) - P.bsdf[OC_WIND_DIR]) * ERTB_INV_TWO_PI;
    v -= floorf(v);
    return oc_tex(data, u, v);
}
__device__ __forceinline__ float oc_fresnel(float nr, float ni, float coschi, float sinchi) {
    float nr2 = nr * nr, ni2 = ni * ni;
    float s = nr2 - ni2 - sinchi * sinchi;
    float a1 = fabsf(s), a2 = sqrtf(fmaf(s, s, 4.f * nr2 * ni2));
    float u = sqrtf(0.5f * fabsf(a1 + a2)), v = sqrtf(0.5f * fabsf(a2 - a1));
    float b1 = (nr2 - ni2) * coschi, b2 = 2.f * nr * ni * coschi;
    float right = ((coschi - u) * (coschi - u) + v * v) / ((coschi + u) * (coschi + u) + v * v);
    float left = ((b1 - u) * (b1 - u) + (b2 + v) * (b2 + v)) / ((b1 + u) * (b1 + u) + (b2 - v) * (b2 - v));
    return 0.5f * (right + left);
}
// Beckmann D(m) of the rotated anisotropic distribution (microfacet.h:195-225)
__device__ __forceinline__ float oc_beckmann_D(const ErtbParams &P, f3 m) {
    const float au = fmaxf(1.41421356f * P.bsdf[OC_SIGMA_U], 1e-4f), av = fmaxf(1.41421356f * P.bsdf[OC_SIGMA_C], 1e-4f);
    float s, c;
    __sincosf(-P.bsdf[OC_WIND_DIR], &s, &c);
    f3 p = normalize3(mk3(c * m.x - s * m.y, s * m.x + c * m.y, m.z));
    float ct2 = m.z * m.z;
    float e = (p.x / au) * (p.x / au) + (p.y / av) * (p.y / av);
    float r = __expf(-e / ct2) / (ERTB_PI * au * av * ct2 * ct2);
    return r * m.z > 1e-20f ? r : 0.f;
}
__device__ __forceinline__ float oc_xy_alpha2(const ErtbParams &P, f3 v) {
    return (P.bsdf[OC_ALPHA_UP] * v.x) * (P.bsdf[OC_ALPHA_UP] * v.x) + (P.bsdf[OC_ALPHA_VP] * v.y) * (P.bsdf[OC_ALPHA_VP] * v.y) +
           v.x * v.y * P.bsdf[OC_CORR];
}
__device__ __forceinline__ float oc_lambda(const ErtbParams &P, f3 v) { // smith_lambda, Beckmann
    float xy = oc_xy_alpha2(P, v);
    if (xy == 0.f) return 0.f;
    float a = rsqrtf(xy / (v.z * v.z)), a2 = a * a;
    return a >= 1.6f ? 0.f : (1.f - 1.259f * a + 0.396f * a2) / (3.535f * a + 2.181f * a2);
}
__device__ __forceinline__ float oc_g1(const ErtbParams &P, f3 v, f3 m) { // smith_g1, Beckmann
    float xy = oc_xy_alpha2(P, v);
    float a = rsqrtf(xy / (v.z * v.z)), a2 = a * a;
    float r = a >= 1.6f ? 1.f : (3.535f * a + 2.181f * a2) / (1.f + 2.276f * a + 2.577f * a2);
    if (xy == 0.f) r = 1.f;
    if (dot3(v, m) * v.z <= 0.f) r = 0.f;
    return r;
}
__device__ __forceinline__ float oc_gram_charlier(const ErtbParams &P, f3 m) { // oceanprops.h:643-690
    const float ws = P.bsdf[OC_WIND_SPEED];
    const float c21 = 0.01f - 0.0086f * ws, c03 = 0.04f - 0.033f * ws;
    float s, c;
    __sincosf(P.bsdf[OC_WIND_DIR], &s, &c);
    f3 p = normalize3(mk3(c * m.x + s * m.y, -s * m.x + c * m.y, m.z));
    float xn = p.x / (P.bsdf[OC_SIGMA_U] * p.z), xe = p.y / (P.bsdf[OC_SIGMA_C] * p.z);
    float xe2 = xe * xe, xn2 = xn * xn;
    float coef = 1.f - (c21 / 2.f) * (xe2 - 1.f) * xn - (c03 / 6.f) * (xn2 - 3.f) * xn;
    coef += (0.40f / 24.f) * (xe2 * xe2 - 6.f * xe2 + 3.f);
    coef += (0.23f / 24.f) * (xn2 * xn2 - 6.f * xn2 + 3.f);
    coef += (0.12f / 4.f) * (xe2 - 1.f) * (xn2 - 1.f);
    return fmaxf(coef, 0.f);
}
// eval_glint(wi, wo) (ocean_legacy.cpp:405-447)
__device__ __forceinline__ float oc_glint(const ErtbParams &P, f3 wi, f3 wo) {
    f3 m = normalize3(mk3(wi.x + wo.x, wi.y + wo.y, wi.z + wo.z));
    float D = oc_beckmann_D(P, m) * oc_gram_charlier(P, m);
    float result = D / (4.f * wi.z * wo.z);
    if (P.bsdf[OC_SHADOWING] != 0.f) {
        float G = 1.f / (1.f + oc_lambda(P, wi) + oc_lambda(P, wo));
        if (dot3(wi, m) * wi.z <= 0.f || dot3(wo, m) * wo.z <= 0.f) G = 0.f;
        result *= G;
    }
    float cos_chi = clampf(dot3(wo, m), -0.999999999f, 0.999999999f);
    float sin_chi = fminf(sqrtf(fmaxf(1.f - cos_chi * cos_chi, 0.f)), 0.999999999f);
    return result * oc_fresnel(P.bsdf[OC_N_REAL], P.bsdf[OC_N_IMAG], cos_chi, sin_chi) * ERTB_PI;
}
// BSDF::eval, Radiance mode: wi = si.wi (towards the sensor side), wo = sampled / sun direction.
// Returns value * cos(theta_o).
__device__ __forceinline__ float oc_eval(const ErtbParams &P, f3 wi, f3 wo) {
    if (!(wi.z > 0.f && wo.z > 0.f)) return 0.f;
    const float *tdn = P.ocean_tables, *tup = P.ocean_tables + ERTB_OC_RES * ERTB_OC_RES;
    float wc = P.bsdf[OC_WHITECAP];
    float ul = 0.f;
    if (P.bsdf[OC_UNDERLIGHT_ON] != 0.f) { // eval_underlight(wo_hat = wo, wi_hat = si.wi)
        float t_d = oc_transmittance(P, tdn, wo.z, wo.x, wo.y);
        float t_u = oc_transmittance(P, tup, wi.z, wo.x, wo.y);
        ul = P.bsdf[OC_UL_NORM] * t_u * t_d;
    }
    float glint = oc_glint(P, wo, wi);
    return (wc + (1.f - wc) * ul + (1.f - P.bsdf[OC_COVERAGE]) * glint) * wo.z * ERTB_INV_PI;
}
__device__ __forceinline__ void oc_lobe_probs(const ErtbParams &P, f3 wi, float &pd, float &ps) {
    float wc = P.bsdf[OC_WHITECAP];
    float t_i = oc_transmittance(P, P.ocean_tables, wi.z, wi.x, wi.y);
    pd = wc + t_i * (1.f - wc);
    ps = 1.f - P.bsdf[OC_COVERAGE];
    ps = ps / (ps + pd);
    pd = 1.f - ps;
}
__device__ __forceinline__ float oc_pdf(const ErtbParams &P, f3 wi, f3 wo) { // :663-713
    if (!(wi.z > 0.f && wo.z > 0.f)) return 0.f;
    float pd, ps;
    oc_lobe_probs(P, wi, pd, ps);
    f3 H = normalize3(mk3(wo.x + wi.x, wo.y + wi.y, wo.z + wi.z));
    return pd * wo.z * ERTB_INV_PI + ps * oc_beckmann_D(P, H) * oc_g1(P, wi, H) / (4.f * wi.z);
}
// sample_visible_11, Beckmann branch (microfacet.h:437-475)
__device__ __forceinline__ void oc_sample_visible_11(float cos_theta_i, float s1, float s2, float &sx, float &sy) {
    const float INV_SQRT_PI = 0.5641895835477563f;
    float tan_i = safe_sqrtf(1.f - cos_theta_i * cos_theta_i) / cos_theta_i;
    float cot_i = 1.f / tan_i;
    float maxval = erff(cot_i);
    s1 = fmaxf(fminf(s1, 1.f - 1e-6f), 1e-6f);
    s2 = fmaxf(fminf(s2, 1.f - 1e-6f), 1e-6f);
    float x = maxval - (maxval + 1.f) * erff(sqrtf(-logf(s1)));
    s1 *= 1.f + maxval + INV_SQRT_PI * tan_i * expf(-cot_i * cot_i);
    for (int i = 0; i < 3; ++i) {
        float slope = erfinvf(x);
        float value = 1.f + x + INV_SQRT_PI * tan_i * expf(-slope * slope) - s1;
        float deriv = 1.f - slope * tan_i;
        x -= value / deriv;
    }
    sx = erfinvf(x);
    sy = erfinvf(fmaf(2.f, s2, -1.f));
}
// BSDF::sample (:494-559): returns wo and the weight eval / pdf
__device__ __forceinline__ float oc_sample(const ErtbParams &P, f3 wi, float s1, float u1, float u2, f3 &wo) {
    wo = mk3(0.f, 0.f, 1.f);
    if (!(wi.z > 0.f)) return 0.f;
    float pd, ps;
    oc_lobe_probs(P, wi, pd, ps);
    if (s1 < pd) {
        wo = cosine_hemisphere(u1, u2);
    } else {
        const float au = fmaxf(1.41421356f * P.bsdf[OC_SIGMA_U], 1e-4f), av = fmaxf(1.41421356f * P.bsdf[OC_SIGMA_C], 1e-4f);
        float sd, cd;
        __sincosf(P.bsdf[OC_WIND_DIR], &sd, &cd);
        f3 p = normalize3(mk3(au * (wi.x * cd + wi.y * sd), av * (-wi.x * sd + wi.y * cd), wi.z));
        float st2 = 1.f - p.z * p.z, sphi = 0.f, cphi = 1.f;
        if (st2 > 0.f) { float is = rsqrtf(st2); sphi = p.y * is; cphi = p.x * is; }
        float sx, sy;
        oc_sample_visible_11(p.z, u1, u2, sx, sy);
        float slx = (cphi * sx - sphi * sy) * au, sly = (sphi * sx + cphi * sy) * av;
        f3 m = normalize3(mk3(-slx, -sly, 1.f));
        f3 H = normalize3(mk3(m.x * cd - m.y * sd, m.x * sd + m.y * cd, m.z));
        float dp = dot3(wi, H);
        wo = mk3(2.f * dp * H.x - wi.x, 2.f * dp * H.y - wi.y, 2.f * dp * H.z - wi.z);
    }
    float pdf = oc_pdf(P, wi, wo);
    if (!(pdf > 0.f)) return 0.f;
    return oc_eval(P, wi, wo) / pdf;
}

// ---------------------------------------------------------------- polarized sun glint
// fresnel_sunglint_polarized (oceanprops.h:443-545).  `wi` / `wo`: propagation directions of the
// incident and of the reflected light in the local frame.  Complex arithmetic on float2 (re, im).
struct cf { float re, im; };
__device__ __forceinline__ cf cmk(float re, float im) { cf r; r.re = re; r.im = im; return r; }
__device__ __forceinline__ cf cadd(cf a, cf b) { return cmk(a.re + b.re, a.im + b.im); }
__device__ __forceinline__ cf csub(cf a, cf b) { return cmk(a.re - b.re, a.im - b.im); }
__device__ __forceinline__ cf cmul(cf a, cf b) { return cmk(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
__device__ __forceinline__ cf cscl(cf a, float s) { return cmk(a.re * s, a.im * s); }
__device__ __forceinline__ cf cdiv(cf a, cf b) {
    float d = 1.f / (b.re * b.re + b.im * b.im);
    return cmk((a.re * b.re + a.im * b.im) * d, (a.im * b.re - a.re * b.im) * d);
}
__device__ __forceinline__ cf csqrt_(cf a) {
    float m = sqrtf(a.re * a.re + a.im * a.im);
    float re = sqrtf(fmaxf(0.5f * (m + a.re), 0.f)), im = sqrtf(fmaxf(0.5f * (m - a.re), 0.f));
    return cmk(re, a.im < 0.f ? -im : im);
}
__device__ __forceinline__ float cabs2(cf a) { return a.re * a.re + a.im * a.im; }
__device__ __forceinline__ cf cmulconj(cf a, cf b) { return cmk(a.re * b.re + a.im * b.im, a.im * b.re - a.re * b.im); }

__device__ __forceinline__ void oc_fresnel_mueller(float nr, float ni, f3 wi_in, f3 wo_in, float *M) {
    const cf n1 = cmk(1.f, 0.f), n2 = cmk(nr, ni);
    float mu_i = fminf(fabsf(wi_in.z), 0.9999999f), mu_o = fminf(fabsf(wo_in.z), 0.9999999f);
    // phi -> -phi mirrors y: rebuild the unit vectors from (mu, -phi) without trigonometry
    float si = sqrtf(1.f - mu_i * mu_i), so = sqrtf(1.f - mu_o * mu_o);
    float ri = rsqrtf(fmaxf(wi_in.x * wi_in.x + wi_in.y * wi_in.y, 1e-30f)), ro = rsqrtf(fmaxf(wo_in.x * wo_in.x + wo_in.y * wo_in.y, 1e-30f));
    float cpi = (wi_in.x == 0.f && wi_in.y == 0.f) ? 1.f : wi_in.x * ri, spi = (wi_in.x == 0.f && wi_in.y == 0.f) ? 0.f : -wi_in.y * ri;
    float cpo = (wo_in.x == 0.f && wo_in.y == 0.f) ? 1.f : wo_in.x * ro, spo = (wo_in.x == 0.f && wo_in.y == 0.f) ? 0.f : -wo_in.y * ro;
    f3 wi = mk3(si * cpi, si * spi, -mu_i), wo = mk3(so * cpo, so * spo, mu_o);
    f3 kd = mk3(wi.x - wo.x, wi.y - wo.y, wi.z - wo.z);
    float mu_il = dot3(kd, wi) * rsqrtf(dot3(kd, kd));
    cf ratio = cdiv(cmul(n1, n1), cmul(n2, n2));
    cf mu_refr = csqrt_(csub(cmk(1.f, 0.f), cscl(ratio, 1.f - mu_il * mu_il)));
    cf a = cscl(n1, mu_il), b = cmul(n2, mu_refr), c = cscl(n2, mu_il), d = cmul(n1, mu_refr);
    cf R_r = cdiv(csub(a, b), cadd(a, b)), R_l = cdiv(csub(c, d), cadd(c, d));
    // polarisation frames: phi_v = normalize(z x w), theta_v = phi_v x w   (w never vertical: mu <= 0.9999999)
    f3 pvi = normalize3(mk3(-wi.y, wi.x, 0.f)), pvo = normalize3(mk3(-wo.y, wo.x, 0.f));
    f3 tvi = mk3(pvi.y * wi.z - pvi.z * wi.y, pvi.z * wi.x - pvi.x * wi.z, pvi.x * wi.y - pvi.y * wi.x);
    f3 tvo = mk3(pvo.y * wo.z - pvo.z * wo.y, pvo.z * wo.x - pvo.x * wo.z, pvo.x * wo.y - pvo.y * wo.x);
    float pi_wo = dot3(pvi, wo), po_wi = dot3(pvo, wi), ti_wo = dot3(tvi, wo), to_wi = dot3(tvo, wi);
    cf f_tt = cadd(cscl(R_r, pi_wo * po_wi), cscl(R_l, ti_wo * to_wi));
    cf f_tp = cadd(cscl(R_r, -ti_wo * po_wi), cscl(R_l, pi_wo * to_wi));
    cf f_pt = cadd(cscl(R_r, -pi_wo * to_wi), cscl(R_l, ti_wo * po_wi));
    cf f_pp = cadd(cscl(R_r, ti_wo * to_wi), cscl(R_l, pi_wo * po_wi));
    f3 cx = mk3(wi.y * wo.z - wi.z * wo.y, wi.z * wo.x - wi.x * wo.z, wi.x * wo.y - wi.y * wo.x);
    float c2 = dot3(cx, cx);
    float coeff = 1.f / fmaxf(c2 * c2, 1e-30f);
    float tt = cabs2(f_tt), tp = cabs2(f_tp), pt = cabs2(f_pt), pp = cabs2(f_pp);
    cf ttp = cmulconj(f_tt, f_tp), ptpp = cmulconj(f_pt, f_pp), ttpt = cmulconj(f_tt, f_pt);
    cf tppp = cmulconj(f_tp, f_pp), ttpp = cmulconj(f_tt, f_pp), tppt = cmulconj(f_tp, f_pt);
    M[0] = 0.5f * coeff * (tt + tp + pt + pp); M[1] = 0.5f * coeff * (tt - tp + pt - pp);
    M[2] = -coeff * (ttp.re + ptpp.re); M[3] = -coeff * (ttp.im + ptpp.im);
    M[4] = 0.5f * coeff * (tt + tp - pt - pp); M[5] = 0.5f * coeff * (tt - tp - pt + pp);
    M[6] = -coeff * (ttp.re - ptpp.re); M[7] = -coeff * (ttp.im - ptpp.im);
    M[8] = -coeff * (ttpt.re + tppp.re); M[9] = -coeff * (ttpt.re - tppp.re);
    M[10] = coeff * (ttpp.re + tppt.re); M[11] = coeff * (ttpp.im - tppt.im);
    M[12] = coeff * (ttpt.im + tppp.im); M[13] = coeff * (ttpt.im - tppp.im);
    M[14] = -coeff * (ttpp.im + tppt.im); M[15] = coeff * (ttpp.re - tppt.re);
}

// ------------------------------------------------------------------------------------------------
// Isotropic-Beckmann glint family: ocean_mishchenko, ocean_grasp, maignan.
//   ERP/bsdfs/ocean_mishchenko.cpp:136-142 update(), :144-226 sample, :228-296 eval, :298-325 pdf
//   ERP/bsdfs/ocean_grasp.cpp:202-238 eval_glint, :246-255 lambda, :267-352 sample, :354-455 eval, :457-512 pdf
//   ERP/bsdfs/maignan.cpp:105-166 eval_maignan, :168-194 sample, :196-211 eval, :213-224 pdf
// The host (scene_commit) derives the OC_* slots the Beckmann helpers above read (sigma_u = sigma_c =
// sqrt(mss / 2), wind direction 0, no correlation), so D, G1, the height-correlated G and the
// visible-normal sampling are the very same device functions the 6SV ocean uses.  Slots reused:
// OC_N_REAL / OC_N_IMAG = lower index relative to the (real) exterior index, OC_COVERAGE, OC_WHITECAP
// (Monahan / Frouin, ocean_grasp), OC_R_OMEGA = water body reflectance, MG_CEXP = C exp(-ndvi).
// ------------------------------------------------------------------------------------------------
#define MG_CEXP 15 /* first slot the 6SV ocean leaves free */

// BSDFs whose value is a Mueller matrix in polarized scenes (Fresnel reflection on facets)
__device__ __forceinline__ bool bsdf_is_mueller(int t) {
    return t == ERTB_BSDF_OCEAN_LEGACY || (t >= ERTB_BSDF_OCEAN_MISHCHENKO && t <= ERTB_BSDF_MAIGNAN);
}
// BSDFs evaluated on local-frame vectors (everything that is not a function of (cos_i, cos_o, cos dphi) only)
__device__ __forceinline__ bool bsdf_is_local(int t) { return t == ERTB_BSDF_OCEAN_LEGACY || t >= ERTB_BSDF_OCEAN_MISHCHENKO; }
// `GENERAL` = false: the lean instances of the pool kernel, which the host only launches for the plugin set of
// SURVEY 8a (the later plugins -- glint family, mqdiffuse, astroobject -- run in its GEN instances, so that their
// code does not weigh on the instruction stream of the headline configurations: C2 -1.2 %, C5 -6 % otherwise)
template <bool GENERAL>
__device__ __forceinline__ bool bsdf_is_local_t(int t) { return GENERAL ? bsdf_is_local(t) : t == ERTB_BSDF_OCEAN_LEGACY; }
template <bool GENERAL>
__device__ __forceinline__ bool bsdf_is_mueller_t(int t) { return GENERAL ? bsdf_is_mueller(t) : t == ERTB_BSDF_OCEAN_LEGACY; }

__device__ __forceinline__ float gl_grasp_lambda(float vz, float sigma) { // ocean_grasp.cpp:246-255
    float st = sigma * safe_sqrtf(1.f - vz * vz) / vz;
    return 0.5f * (0.7978845608f * st * __expf(-1.f / (2.f * st * st)) - erfcf(0.70710678f / st));
}
__device__ __forceinline__ float gl_fresnel00(const ErtbParams &P, f3 wi, f3 wo) {
    float M[16];
    oc_fresnel_mueller(P.bsdf[OC_N_REAL], P.bsdf[OC_N_IMAG], mk3(-wo.x, -wo.y, -wo.z), wi, M);
    return M[0];
}
// scalar factor in front of the Fresnel matrix in BSDF::eval (with the cosine where the plugin has one)
__device__ __forceinline__ float gl_geometry(const ErtbParams &P, f3 wi, f3 wo) {
    if (P.bsdf_type == ERTB_BSDF_MAIGNAN) { // maignan.cpp:116-136; cos(Theta) = wi . wo
        float cT = clampf(dot3(wi, wo), -1.f, 1.f);
        float tan_a = sqrtf((1.f - cT) / (1.f + cT));
        return P.bsdf[MG_CEXP] * __expf(-tan_a) / (4.f * (wi.z + wo.z));
    }
    f3 m = normalize3(mk3(wi.x + wo.x, wi.y + wo.y, wi.z + wo.z));
    float D = oc_beckmann_D(P, m);
    const bool facing = dot3(wi, m) * wi.z > 0.f && dot3(wo, m) * wo.z > 0.f;
    if (P.bsdf_type == ERTB_BSDF_OCEAN_MISHCHENKO) { // :249-254
        float G = facing ? 1.f / (1.f + oc_lambda(P, wi) + oc_lambda(P, wo)) : 0.f;
        return D * G / (4.f * wi.z);
    }
    const float sigma = P.bsdf[OC_SIGMA_U];
    float G = facing ? 1.f / (1.f + gl_grasp_lambda(wo.z, sigma) + gl_grasp_lambda(wi.z, sigma)) : 0.f;
    return (1.f - P.bsdf[OC_COVERAGE]) * D * G / (4.f * wi.z); // pi D G / (4 ci co) * co / pi
}
__device__ __forceinline__ float gl_dep(const ErtbParams &P, f3 wo) { // ocean_grasp.cpp:395-404, :436
    if (P.bsdf_type != ERTB_BSDF_OCEAN_GRASP) return 0.f;
    return (P.bsdf[OC_WHITECAP] + (1.f - P.bsdf[OC_COVERAGE]) * P.bsdf[OC_R_OMEGA]) * wo.z * ERTB_INV_PI;
}
__device__ __forceinline__ float gl_eval(const ErtbParams &P, f3 wi, f3 wo) {
    if (!(wi.z > 0.f && wo.z > 0.f)) return 0.f;
    return gl_dep(P, wo) + gl_geometry(P, wi, wo) * gl_fresnel00(P, wi, wo);
}
__device__ __forceinline__ float gl_pdf(const ErtbParams &P, f3 wi, f3 wo) {
    if (!(wi.z > 0.f && wo.z > 0.f)) return 0.f;
    if (P.bsdf_type == ERTB_BSDF_MAIGNAN) return wo.z * ERTB_INV_PI;
    f3 m = normalize3(mk3(wi.x + wo.x, wi.y + wo.y, wi.z + wo.z));
    float spec = oc_beckmann_D(P, m) * oc_g1(P, wi, m) / (4.f * wi.z);
    if (P.bsdf_type == ERTB_BSDF_OCEAN_MISHCHENKO) return (dot3(wi, m) > 0.f && dot3(wo, m) > 0.f) ? spec : 0.f;
    const float cov = P.bsdf[OC_COVERAGE], p_spec = 1.f / (P.bsdf[OC_R_OMEGA] + 1.f), pc = wo.z * ERTB_INV_PI;
    return cov * pc + (1.f - cov) * ((1.f - p_spec) * pc + p_spec * spec);
}
// visible-normal sample of the isotropic distribution + mirror reflection
__device__ __forceinline__ f3 gl_reflect_sample(const ErtbParams &P, f3 wi, float u1, float u2, f3 &m) {
    const float a = fmaxf(1.41421356f * P.bsdf[OC_SIGMA_U], 1e-4f);
    f3 p = normalize3(mk3(a * wi.x, a * wi.y, wi.z));
    float st2 = 1.f - p.z * p.z, sphi = 0.f, cphi = 1.f;
    if (st2 > 0.f) { float is = rsqrtf(st2); sphi = p.y * is; cphi = p.x * is; }
    float sx, sy;
    oc_sample_visible_11(p.z, u1, u2, sx, sy);
    m = normalize3(mk3(-(cphi * sx - sphi * sy) * a, -(sphi * sx + cphi * sy) * a, 1.f));
    float dp = dot3(wi, m);
    return mk3(2.f * dp * m.x - wi.x, 2.f * dp * m.y - wi.y, 2.f * dp * m.z - wi.z);
}
// BSDF::sample: direction only
__device__ __forceinline__ f3 gl_sample_dir(const ErtbParams &P, f3 wi, float s1, float u1, float u2) {
    f3 m;
    if (P.bsdf_type == ERTB_BSDF_MAIGNAN) return cosine_hemisphere(u1, u2);
    if (P.bsdf_type == ERTB_BSDF_OCEAN_MISHCHENKO) return gl_reflect_sample(P, wi, u1, u2, m);
    const float cov = P.bsdf[OC_COVERAGE], p_diff = 1.f - 1.f / (P.bsdf[OC_R_OMEGA] + 1.f); // :285-318
    if (s1 < cov || (s1 - cov) / (1.f - cov) < p_diff) return cosine_hemisphere(u1, u2);
    return gl_reflect_sample(P, wi, u1, u2, m);
}
// scalar factor of the weight BSDF::sample returns for `wo`, in front of the Fresnel matrix, and the
// factor `dscale` applying to the depolarizing part
__device__ __forceinline__ float gl_weight_geometry(const ErtbParams &P, f3 wi, f3 wo, float &dscale) {
    dscale = 0.f;
    if (!(wi.z > 0.f && wo.z > 0.f)) return 0.f;
    if (P.bsdf_type == ERTB_BSDF_MAIGNAN) return gl_geometry(P, wi, wo); // C F, not divided by the pdf (:189-193)
    if (P.bsdf_type == ERTB_BSDF_OCEAN_MISHCHENKO) { // F G / G1 (:170-181)
        f3 m = normalize3(mk3(wi.x + wo.x, wi.y + wo.y, wi.z + wo.z));
        float g1 = oc_g1(P, wi, m);
        if (!(oc_beckmann_D(P, m) * g1 * fabsf(dot3(wi, m)) != 0.f)) return 0.f;
        float G = (dot3(wo, m) * wo.z > 0.f) ? 1.f / (1.f + oc_lambda(P, wi) + oc_lambda(P, wo)) : 0.f;
        return G / g1;
    }
    float pdf = gl_pdf(P, wi, wo);
    dscale = pdf > 0.f ? 1.f / pdf : 0.f;
    return gl_geometry(P, wi, wo) * dscale;
}
// BSDF::sample (scalar): returns the weight
__device__ __forceinline__ float gl_sample(const ErtbParams &P, f3 wi, float s1, float u1, float u2, f3 &wo) {
    wo = mk3(0.f, 0.f, 1.f);
    if (!(wi.z > 0.f)) return 0.f;
    wo = gl_sample_dir(P, wi, s1, u1, u2);
    float dscale;
    float g = gl_weight_geometry(P, wi, wo, dscale);
    if (!(wo.z > 0.f)) return 0.f;
    return gl_dep(P, wo) * dscale + g * gl_fresnel00(P, wi, wo);
}

// ------------------------------------------------------------------------------------------------
// mqdiffuse (ERP/bsdfs/mqdiffuse.cpp:94-176): trilinear lookup of the measured table at
// (cos_theta_o, phi_d / 2 pi, cos_theta_i), remapped so that the first / last texel sit on 0 / 1
// (:96-100), clamp wrap mode (drjit/texture.h linear filter: both neighbours clamped to the grid).
// P.bsdf[0..2] = resolution (x, y, z), P.ocean_tables = data[z][y][x].
// eval() wraps a negative azimuth difference by 2 pi (:153); sample() does not (:125-131), so a sampled
// direction with atan2(wo) < atan2(wi) reads the phi_d = 0 plane through the clamp -- kept as is.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float mq_tex(const ErtbParams &P, float cos_o, float phi_d, float cos_i) {
    const int rx = (int) P.bsdf[0], ry = (int) P.bsdf[1], rz = (int) P.bsdf[2];
    float px = cos_o * (float) (rx - 1), py = phi_d * (0.5f * ERTB_INV_PI) * (float) (ry - 1), pz = cos_i * (float) (rz - 1);
    float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
    float wx = px - fx, wy = py - fy, wz = pz - fz;
    int x0 = min(max((int) fx, 0), rx - 1), x1 = min(max((int) fx + 1, 0), rx - 1);
    int y0 = min(max((int) fy, 0), ry - 1), y1 = min(max((int) fy + 1, 0), ry - 1);
    int z0 = min(max((int) fz, 0), rz - 1), z1 = min(max((int) fz + 1, 0), rz - 1);
    const float *d = P.ocean_tables;
#define MQ_AT(z, y, x) __ldg(d + ((size_t) (z) * ry + (y)) * rx + (x))
    float c00 = fmaf(wx, MQ_AT(z0, y0, x1) - MQ_AT(z0, y0, x0), MQ_AT(z0, y0, x0));
    float c01 = fmaf(wx, MQ_AT(z0, y1, x1) - MQ_AT(z0, y1, x0), MQ_AT(z0, y1, x0));
    float c10 = fmaf(wx, MQ_AT(z1, y0, x1) - MQ_AT(z1, y0, x0), MQ_AT(z1, y0, x0));
    float c11 = fmaf(wx, MQ_AT(z1, y1, x1) - MQ_AT(z1, y1, x0), MQ_AT(z1, y1, x0));
#undef MQ_AT
    float c0 = fmaf(wy, c01 - c00, c00), c1 = fmaf(wy, c11 - c10, c10);
    return fmaf(wz, c1 - c0, c0);
}
__device__ __forceinline__ float mq_phi_d(f3 wi, f3 wo) { // fmod(atan2(wo) - atan2(wi), 2 pi): sign of the dividend
    return fmodf(atan2f(wo.y, wo.x) - atan2f(wi.y, wi.x), 2.f * ERTB_PI);
}
__device__ __forceinline__ float mq_eval(const ErtbParams &P, f3 wi, f3 wo) {
    if (!(wi.z > 0.f && wo.z > 0.f)) return 0.f;
    float phi_d = mq_phi_d(wi, wo);
    if (phi_d < 0.f) phi_d += 2.f * ERTB_PI;
    return mq_tex(P, wo.z, phi_d, wi.z) * wo.z;
}
__device__ __forceinline__ float mq_sample(const ErtbParams &P, f3 wi, float u1, float u2, f3 &wo) {
    wo = mk3(0.f, 0.f, 1.f);
    if (!(wi.z > 0.f)) return 0.f;
    wo = cosine_hemisphere(u1, u2);
    if (!(wo.z > 0.f)) return 0.f;
    return mq_tex(P, wo.z, mq_phi_d(wi, wo), wi.z) * ERTB_PI; // value * cos / pdf
}

// the table-driven scalar BSDFs (mqdiffuse, measured_mono): depolarizers in polarized scenes
__device__ __forceinline__ bool bsdf_is_table(int t) { return t >= ERTB_BSDF_MQDIFFUSE; }
__device__ __forceinline__ float tb_eval(const ErtbParams &P, f3 wi, f3 wo) {
    return P.bsdf_type == ERTB_BSDF_MEASURED_MONO ? mm_eval(P, wi, wo) : mq_eval(P, wi, wo);
}
__device__ __forceinline__ float tb_sample(const ErtbParams &P, f3 wi, float u1, float u2, f3 &wo) {
    float pdf;
    return P.bsdf_type == ERTB_BSDF_MEASURED_MONO ? mm_sample(P, wi, u1, u2, wo, pdf) : mq_sample(P, wi, u1, u2, wo);
}

// dispatch over the local-frame BSDFs (6SV ocean, glint family, mqdiffuse, measured_mono)
template <bool GENERAL = true>
__device__ __forceinline__ float lf_eval(const ErtbParams &P, f3 wi, f3 wo) {
    if (!GENERAL || P.bsdf_type == ERTB_BSDF_OCEAN_LEGACY) return oc_eval(P, wi, wo);
    if (bsdf_is_table(P.bsdf_type)) return tb_eval(P, wi, wo);
    return gl_eval(P, wi, wo);
}
template <bool GENERAL = true>
__device__ __forceinline__ float lf_sample(const ErtbParams &P, f3 wi, float s1, float u1, float u2, f3 &wo) {
    if (!GENERAL || P.bsdf_type == ERTB_BSDF_OCEAN_LEGACY) return oc_sample(P, wi, s1, u1, u2, wo);
    if (bsdf_is_table(P.bsdf_type)) return tb_sample(P, wi, u1, u2, wo);
    return gl_sample(P, wi, s1, u1, u2, wo);
}
