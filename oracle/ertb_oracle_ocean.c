/*
 * ertb_oracle_ocean.c -- CPU oracle: ocean_legacy BSDF (6SV ocean model); second half of the file: the
 * isotropic-Beckmann glint family (ocean_mishchenko, ocean_grasp, maignan).
 * TEST INFRASTRUCTURE (see ertb_oracle.c).  Restates, in scalar double C:
 *   ERP/bsdfs/ocean_legacy.cpp:137-243  eval_ocean_transmittance (64x64 Gauss-Legendre)
 *   ERP/bsdfs/ocean_legacy.cpp:313-372  update(); :384-393 whitecaps; :405-447 glint;
 *                              :449-491 transmittance lookup / underlight;
 *                              :494-559 sample; :561-661 eval; :663-713 pdf
 *   MI/include/mitsuba/eradiate/oceanprops.h (whitecap coverage :330, water_ior :389,
 *       fresnel_sunglint_legacy :415, cox_munk_* :566-690, r_omega :692)
 *   MI/include/mitsuba/render/microfacet.h:195-530 (anisotropic rotated Beckmann:
 *       eval, visible-normal sampling, smith_g1, G_height_correlated)
 *   MI/ext/drjit/include/drjit/texture.h:500-530 (bilinear lookup, clamp wrap)
 * The spectral tables are literature data (Whitlock et al. 1982; Hale & Querry 1973;
 * Morel 1988) as tabulated by 6SV and by oceanprops.h:32-160.
 */
#include "ertb_oracle_ocean.h"
#include "../include/eradiate_b200.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define PI 3.14159265358979323846
#define OC_RES 64

/* ------------------------------------------------------------- literature tables */
/* Whitecap effective reflectance, 0.2 .. 4.0 um every 0.1 um (Whitlock et al. 1982) */
static const double WC_DATA[39] = {
    0.220, 0.220, 0.220, 0.220, 0.220, 0.220, 0.215, 0.210, 0.200, 0.190, 0.175, 0.155, 0.130,
    0.080, 0.100, 0.105, 0.100, 0.080, 0.045, 0.055, 0.065, 0.060, 0.055, 0.040, 0.000, 0.000,
    0.000, 0.000, 0.000, 0.000, 0.000, 0.000, 0.000, 0.000, 0.000, 0.000, 0.000, 0.000, 0.000 };
/* Complex refractive index of water (Hale & Querry 1973), wavelengths in nm */
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
/* Morel (1988) case-1 water: K_w, chi, e on 400..700 nm every 5 nm */
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
/* Molecular scattering coefficient of sea water as used by 6S, 400..700 nm every 5 nm */
static const double MOL_6S[61] = {
    0.0076, 0.0072, 0.0068, 0.0064, 0.0061, 0.0058, 0.0055, 0.0052, 0.0049, 0.0047, 0.0045, 0.0043,
    0.0041, 0.0039, 0.0037, 0.0036, 0.0034, 0.0033, 0.0031, 0.0030, 0.0029, 0.0027, 0.0026, 0.0025,
    0.0024, 0.0023, 0.0022, 0.0022, 0.0021, 0.0020, 0.0019, 0.0018, 0.0018, 0.0017, 0.0017, 0.0016,
    0.0016, 0.0015, 0.0015, 0.0014, 0.0014, 0.0013, 0.0013, 0.0012, 0.0012, 0.0011, 0.0011, 0.0010,
    0.0010, 0.0010, 0.0010, 0.0009, 0.0008, 0.0008, 0.0008, 0.0007, 0.0007, 0.0007, 0.0007, 0.0007,
    0.0007 };

/* ContinuousDistribution::eval_pdf (distr_1d.h:370-392): linear interpolation, 0 outside */
static double interp_regular(const double *y, int n, double x0, double x1, double x) {
    if (!(x >= x0 && x <= x1)) return 0.0;
    double xs = (x - x0) * (n - 1) / (x1 - x0);
    int i = (int) xs;
    if (i < 0) i = 0;
    if (i > n - 2) i = n - 2;
    double w1 = xs - i;
    return (1.0 - w1) * y[i] + w1 * y[i + 1];
}
/* IrregularContinuousDistribution::eval_pdf (distr_1d.h:712-735) */
static double interp_irregular(const double *xn, const double *y, int n, double x) {
    if (!(x >= xn[0] && x <= xn[n - 1])) return 0.0;
    int lo = 0, hi = n - 1;
    while (hi - lo > 1) {
        int mid = (lo + hi) / 2;
        if (xn[mid] < x) lo = mid; else hi = mid;
    }
    double t = (x - xn[lo]) / (xn[lo + 1] - xn[lo]);
    return y[lo] + t * (y[lo + 1] - y[lo]);
}

/* ------------------------------------------------------------------ oceanprops.h */
static double fresnel_legacy(double nr, double ni, double coschi, double sinchi) { /* :415-440 */
    double nr2 = nr * nr, ni2 = ni * ni;
    double s = nr2 - ni2 - sinchi * sinchi;
    double a1 = fabs(s), a2 = sqrt(s * s + 4.0 * nr2 * ni2);
    double u = sqrt(0.5 * fabs(a1 + a2)), v = sqrt(0.5 * fabs(a2 - a1));
    double b1 = (nr2 - ni2) * coschi, b2 = 2.0 * nr * ni * coschi;
    double right = ((coschi - u) * (coschi - u) + v * v) / ((coschi + u) * (coschi + u) + v * v);
    double left = ((b1 - u) * (b1 - u) + (b2 + v) * (b2 + v)) / ((b1 + u) * (b1 + u) + (b2 - v) * (b2 - v));
    return 0.5 * (right + left);
}

/* Gram-Charlier series factor of the Cox-Munk slope distribution (:643-690) */
static double gram_charlier(double wind_dir, double wind_speed, double su, double sc, const double m[3]) {
    const double c40 = 0.40, c22 = 0.12, c04 = 0.23;
    double c21 = 0.01 - 0.0086 * wind_speed, c03 = 0.04 - 0.033 * wind_speed;
    double sp = sin(wind_dir), cp = cos(wind_dir);
    double mx = cp * m[0] + sp * m[1], my = -sp * m[0] + cp * m[1], mz = m[2];
    double inv = 1.0 / sqrt(mx * mx + my * my + mz * mz);
    mx *= inv; my *= inv; mz *= inv;
    double xn = mx / (su * mz), xe = my / (sc * mz);
    double xe2 = xe * xe, xn2 = xn * xn;
    double coef = 1.0 - (c21 / 2.0) * (xe2 - 1.0) * xn - (c03 / 6.0) * (xn2 - 3.0) * xn;
    coef += (c40 / 24.0) * (xe2 * xe2 - 6.0 * xe2 + 3.0);
    coef += (c04 / 24.0) * (xn2 * xn2 - 6.0 * xn2 + 3.0);
    coef += (c22 / 4.0) * (xe2 - 1.0) * (xn2 - 1.0);
    return coef;
}
/* cox_munk_anisotropic_distrib (:585-630) */
static double cox_munk_distrib(double wind_dir, double wind_speed, double su, double sc, const double m[3]) {
    double sp = sin(wind_dir), cp = cos(wind_dir);
    double mx = cp * m[0] + sp * m[1], my = -sp * m[0] + cp * m[1], mz = m[2];
    double inv = 1.0 / sqrt(mx * mx + my * my + mz * mz);
    mx *= inv; my *= inv; mz *= inv;
    double xn = mx / (su * mz), xe = my / (sc * mz);
    double coef = gram_charlier(wind_dir, wind_speed, su, sc, m);
    double prob = coef / (2.0 * PI) / (su * sc) * exp(-(xe * xe + xn * xn) * 0.5);
    return prob > 0.0 ? prob : 0.0;
}

/* r_omega (:692-740): iterative underlight reflectance */
static double r_omega(double wavelength, double pigmentation) {
    double pigment_log = log(pigmentation) / log(10.0);
    double mol = interp_regular(MOL_6S, 61, 400.0, 700.0, wavelength);
    double scat = 0.30 * pow(pigmentation, 0.62);
    double bratio = 0.002 + 0.02 * (0.5 - 0.25 * pigment_log) * (550.0 / wavelength);
    double bb = 0.5 * mol + scat * bratio;
    double k = interp_regular(ATTN_K, 61, 400.0, 700.0, wavelength);
    double chi = interp_regular(ATTN_CHI, 61, 400.0, 700.0, wavelength);
    double e = interp_regular(ATTN_E, 61, 400.0, 700.0, wavelength);
    double attn = k + chi * pow(pigmentation, e);
    if (bb == 0.0 || attn == 0.0) return 0.0;
    double u = 0.75, r = 0.33 * bb / u / attn;
    for (int it = 0; it < 1000; ++it) {
        u = (0.9 * (1.0 - r)) / (1.0 + 2.25 * r);
        double rn = 0.33 * bb / (u * attn);
        if (fabs((rn - r) / rn) < 0.0001) break;
        r = rn;
    }
    return r;
}

/* Gauss-Legendre nodes/weights on [-1, 1] (quad::gauss_legendre) by Newton iteration */
static void gauss_legendre(int n, double *x, double *w) {
    for (int i = 0; i < n; ++i) {
        double z = cos(PI * (i + 0.75) / (n + 0.5)), pp = 0.0;
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
        x[n - 1 - i] = z; /* ascending */
        w[n - 1 - i] = 2.0 / ((1.0 - z * z) * pp * pp);
    }
}

/* eval_ocean_transmittance (ocean_legacy.cpp:137-243) for one (theta, phi) */
static double transmittance(double theta, double phi, double nr, double ni, double wind_speed, int upwelling,
                            const double *gx, const double *gw) {
    if (upwelling) {
        theta = asin(sin(theta) / nr);
        nr = 1.0 / nr;
        ni = 0.0;
    }
    double su = sqrt(0.00316 * wind_speed), sc = sqrt(0.003 + 0.00192 * wind_speed);
    double wi[3] = { sin(theta), 0.0, cos(theta) };
    double td = 0.0, summ = 0.0;
    /* meshgrid(nodes, nodes): nodes_x (zenith) varies fastest */
    for (int iy = 0; iy < OC_RES; ++iy) {
        double phi_o = gx[iy] * PI + PI, wy = PI * gw[iy];
        for (int ix = 0; ix < OC_RES; ++ix) {
            double theta_o = gx[ix] * 0.25 * PI + 0.25 * PI, wx = 0.25 * PI * gw[ix];
            double sz = sin(theta_o), cz = cos(theta_o);
            double gweight = cz * sz * wy * wx;
            double wo[3] = { sz * cos(phi_o), sz * sin(phi_o), cz };
            double cti = wi[2] < 1e-6 ? 1e-6 : wi[2], cto = wo[2] < 1e-6 ? 1e-6 : wo[2];
            double m[3] = { wi[0] + wo[0], wi[1] + wo[1], wi[2] + wo[2] };
            double inv = 1.0 / sqrt(m[0] * m[0] + m[1] * m[1] + m[2] * m[2]);
            m[0] *= inv; m[1] *= inv; m[2] *= inv;
            double D = cox_munk_distrib(phi, wind_speed, su, sc, m) / pow(m[2], 4.0);
            double cos_chi = wo[0] * m[0] + wo[1] * m[1] + wo[2] * m[2];
            cos_chi = fmax(fmin(cos_chi, 0.999999999), -0.999999999);
            double sin_chi = fmax(fmin(sqrt(1.0 - cos_chi * cos_chi), 0.999999999), -0.999999999);
            double F = fresnel_legacy(nr, ni, cos_chi, sin_chi);
            double glint = D * F * PI / (4.0 * cti * cto);
            if (!(cti > 0.0 && cto > 0.0)) glint = 1.0;
            td += glint * gweight;
            summ += gweight;
        }
    }
    if (td >= summ) td = summ;
    return 1.0 - td / summ;
}

int ocean_init(ocean_state_t *o, const float *p) {
    memset(o, 0, sizeof *o);
    o->wavelength = p[0];
    o->wind_speed = p[1];
    double wd = -(double) p[2] + 90.0; /* North-left -> East-right (ocean_legacy.cpp:275-280) */
    wd = wd - 360.0 * floor(wd / 360.0);
    o->wind_direction = wd * PI / 180.0;
    o->chlorinity = p[3];
    o->pigmentation = p[4];
    o->shadowing = p[5] != 0.f;
    o->n_real = interp_irregular(IOR_WL, IOR_RE, 64, o->wavelength) + 0.00017492711 * (0.03 + 1.805 * o->chlorinity);
    o->n_imag = interp_irregular(IOR_WL, IOR_IM, 64, o->wavelength);
    o->sigma_c2 = 0.003 + 0.00192 * o->wind_speed;
    o->sigma_u2 = 0.00316 * o->wind_speed;
    o->r_omega = r_omega(o->wavelength, o->pigmentation);
    double cov = 2.95e-06 * pow(o->wind_speed, 3.52); /* whitecap_coverage_monahan :330-336 */
    o->whitecap_coverage = fmax(0.0, fmin(1.0, cov));
    o->whitecap_reflectance = o->whitecap_coverage * interp_regular(WC_DATA, 39, 200.0, 4000.0, o->wavelength);
    o->underlight_attn = 0.485;
    o->n_tab = OC_RES;
    o->tr_down = malloc(sizeof(double) * OC_RES * OC_RES);
    o->tr_up = malloc(sizeof(double) * OC_RES * OC_RES);
    double gx[OC_RES], gw[OC_RES];
    gauss_legendre(OC_RES, gx, gw);
    /* meshgrid(zeniths, azimuths): data[i * 64 + j] <-> (zenith j, azimuth i) */
#pragma omp parallel for schedule(dynamic, 4)
    for (int idx = 0; idx < OC_RES * OC_RES; ++idx) {
        int i = idx / OC_RES, j = idx % OC_RES;
        double zen = fmax(0.0, 0.5 * PI * j / (OC_RES - 1)), az = fmax(0.0, 2.0 * PI * i / (OC_RES - 1));
        o->tr_down[idx] = transmittance(zen, az, o->n_real, o->n_imag, o->wind_speed, 0, gx, gw);
        o->tr_up[idx] = transmittance(zen, az, o->n_real, o->n_imag, o->wind_speed, 1, gx, gw);
    }
    o->ready = 1;
    return 0;
}

void ocean_free(ocean_state_t *o) {
    free(o->tr_down);
    free(o->tr_up);
    o->tr_down = o->tr_up = NULL;
}

/* Texture2f linear / clamp lookup (drjit texture.h:500-530); u -> zenith axis, v -> azimuth */
static double tex_lookup(const double *data, double u, double v) {
    double pu = u * OC_RES - 0.5, pv = v * OC_RES - 0.5;
    int iu = (int) floor(pu), iv = (int) floor(pv);
    double wu1 = pu - iu, wv1 = pv - iv, wu0 = 1.0 - wu1, wv0 = 1.0 - wv1;
    int u0 = iu < 0 ? 0 : (iu > OC_RES - 1 ? OC_RES - 1 : iu), u1 = iu + 1 < 0 ? 0 : (iu + 1 > OC_RES - 1 ? OC_RES - 1 : iu + 1);
    int v0 = iv < 0 ? 0 : (iv > OC_RES - 1 ? OC_RES - 1 : iv), v1 = iv + 1 < 0 ? 0 : (iv + 1 > OC_RES - 1 ? OC_RES - 1 : iv + 1);
    return data[v0 * OC_RES + u0] * wu0 * wv0 + data[v0 * OC_RES + u1] * wu1 * wv0 +
           data[v1 * OC_RES + u0] * wu0 * wv1 + data[v1 * OC_RES + u1] * wu1 * wv1;
}
/* eval_transmittance (ocean_legacy.cpp:449-466) */
static double eval_transmittance(const ocean_state_t *o, const double *data, double cos_theta, double vx, double vy) {
    double c = cos_theta < -1.0 ? -1.0 : (cos_theta > 1.0 ? 1.0 : cos_theta);
    double u = acos(c) * (2.0 / PI);
    double v = (atan2(vy, vx) - o->wind_direction) / (2.0 * PI);
    v = v - floor(v);
    return tex_lookup(data, u, v);
}

/* ----- rotated anisotropic Beckmann distribution (microfacet.h) ----- */
typedef struct { double au, av, angle, aup, avp, corr; } beckmann_t;
static beckmann_t beckmann_make(const ocean_state_t *o) {
    beckmann_t B;
    B.au = fmax(sqrt(2.0) * sqrt(o->sigma_u2), 1e-4);
    B.av = fmax(sqrt(2.0) * sqrt(o->sigma_c2), 1e-4);
    B.angle = o->wind_direction;
    double s = sin(B.angle), c = cos(B.angle);
    B.aup = sqrt((B.au * c) * (B.au * c) + (B.av * s) * (B.av * s));
    B.avp = sqrt((B.au * s) * (B.au * s) + (B.av * c) * (B.av * c));
    B.corr = 2.0 * (B.au * B.au - B.av * B.av) * c * s;
    return B;
}
static double beckmann_eval(const beckmann_t *B, const double m[3]) { /* :195-225 */
    double ct = m[2], ct2 = ct * ct;
    double s = sin(-B->angle), c = cos(-B->angle);
    double px = c * m[0] - s * m[1], py = s * m[0] + c * m[1], pz = m[2];
    double inv = 1.0 / sqrt(px * px + py * py + pz * pz);
    px *= inv; py *= inv;
    double r = exp(-((px / B->au) * (px / B->au) + (py / B->av) * (py / B->av)) / ct2) / (PI * B->au * B->av * ct2 * ct2);
    return r * ct > 1e-20 ? r : 0.0;
}
static double beckmann_lambda(const beckmann_t *B, const double v[3]) { /* :400-420 */
    double xy = (B->aup * v[0]) * (B->aup * v[0]) + (B->avp * v[1]) * (B->avp * v[1]) + v[0] * v[1] * B->corr;
    if (xy == 0.0) return 0.0;
    double t2 = xy / (v[2] * v[2]);
    double a = 1.0 / sqrt(t2), a2 = a * a;
    return a >= 1.6 ? 0.0 : (1.0 - 1.259 * a + 0.396 * a2) / (3.535 * a + 2.181 * a2);
}
static double beckmann_g1(const beckmann_t *B, const double v[3], const double m[3]) { /* :375-398 */
    double xy = (B->aup * v[0]) * (B->aup * v[0]) + (B->avp * v[1]) * (B->avp * v[1]) + v[0] * v[1] * B->corr;
    double t2 = xy / (v[2] * v[2]);
    double a = 1.0 / sqrt(t2), a2 = a * a;
    double r = a >= 1.6 ? 1.0 : (3.535 * a + 2.181 * a2) / (1.0 + 2.276 * a + 2.577 * a2);
    if (xy == 0.0) r = 1.0;
    if ((v[0] * m[0] + v[1] * m[1] + v[2] * m[2]) * v[2] <= 0.0) r = 0.0;
    return r;
}
static double beckmann_g_hc(const beckmann_t *B, const double wi[3], const double wo[3], const double m[3]) { /* :356-366 */
    double r = 1.0 / (1.0 + beckmann_lambda(B, wi) + beckmann_lambda(B, wo));
    if ((wi[0] * m[0] + wi[1] * m[1] + wi[2] * m[2]) * wi[2] <= 0.0) r = 0.0;
    if ((wo[0] * m[0] + wo[1] * m[1] + wo[2] * m[2]) * wo[2] <= 0.0) r = 0.0;
    return r;
}
/* erfinv via Newton on erf (libm has no erfinv) */
static double erfinv_d(double y) {
    if (y <= -1.0) return -INFINITY;
    if (y >= 1.0) return INFINITY;
    double w = -log((1.0 - y) * (1.0 + y)), x;
    if (w < 5.0) {
        w -= 2.5;
        x = 2.81022636e-08; x = 3.43273939e-07 + x * w; x = -3.5233877e-06 + x * w;
        x = -4.39150654e-06 + x * w; x = 0.00021858087 + x * w; x = -0.00125372503 + x * w;
        x = -0.00417768164 + x * w; x = 0.246640727 + x * w; x = 1.50140941 + x * w;
    } else {
        w = sqrt(w) - 3.0;
        x = -0.000200214257; x = 0.000100950558 + x * w; x = 0.00134934322 + x * w;
        x = -0.00367342844 + x * w; x = 0.00573950773 + x * w; x = -0.0076224613 + x * w;
        x = 0.00943887047 + x * w; x = 1.00167406 + x * w; x = 2.83297682 + x * w;
    }
    x *= y;
    for (int i = 0; i < 3; ++i) /* polish */
        x -= (erf(x) - y) / (2.0 / sqrt(PI) * exp(-x * x));
    return x;
}
/* sample_visible_11 (:437-475), Beckmann branch */
static void sample_visible_11(double cos_theta_i, double s1, double s2, double *sx, double *sy) {
    double tan_i = sqrt(fmax(0.0, 1.0 - cos_theta_i * cos_theta_i)) / cos_theta_i;
    double cot_i = 1.0 / tan_i;
    double maxval = erf(cot_i);
    s1 = fmax(fmin(s1, 1.0 - 1e-6), 1e-6);
    s2 = fmax(fmin(s2, 1.0 - 1e-6), 1e-6);
    double x = maxval - (maxval + 1.0) * erf(sqrt(-log(s1)));
    s1 *= 1.0 + maxval + (1.0 / sqrt(PI)) * tan_i * exp(-cot_i * cot_i);
    for (int i = 0; i < 3; ++i) {
        double slope = erfinv_d(x);
        double value = 1.0 + x + (1.0 / sqrt(PI)) * tan_i * exp(-slope * slope) - s1;
        double deriv = 1.0 - slope * tan_i;
        x -= value / deriv;
    }
    *sx = erfinv_d(x);
    *sy = erfinv_d(2.0 * s2 - 1.0);
}
/* MicrofacetDistribution::sample, visible-normal branch (:305-348) */
static void beckmann_sample(const beckmann_t *B, const double wi[3], double s1, double s2, double m[3]) {
    double sd = sin(B->angle), cd = cos(B->angle);
    double px = B->au * (wi[0] * cd + wi[1] * sd), py = B->av * (-wi[0] * sd + wi[1] * cd), pz = wi[2];
    double inv = 1.0 / sqrt(px * px + py * py + pz * pz);
    px *= inv; py *= inv; pz *= inv;
    double st2 = 1.0 - pz * pz, sphi = 0.0, cphi = 1.0;
    if (st2 > 0.0) { double is = 1.0 / sqrt(st2); sphi = py * is; cphi = px * is; }
    double sx, sy;
    sample_visible_11(pz, s1, s2, &sx, &sy);
    double slx = (cphi * sx - sphi * sy) * B->au, sly = (sphi * sx + cphi * sy) * B->av;
    double mx = -slx, my = -sly, mz = 1.0;
    inv = 1.0 / sqrt(mx * mx + my * my + mz * mz);
    mx *= inv; my *= inv; mz *= inv;
    double rx = mx * cd - my * sd, ry = mx * sd + my * cd;
    inv = 1.0 / sqrt(rx * rx + ry * ry + mz * mz);
    m[0] = rx * inv; m[1] = ry * inv; m[2] = mz * inv;
}

/* eval_glint(wi, wo) (:405-447) */
static double eval_glint(const ocean_state_t *o, const double wi[3], const double wo[3]) {
    beckmann_t B = beckmann_make(o);
    double m[3] = { wi[0] + wo[0], wi[1] + wo[1], wi[2] + wo[2] };
    double inv = 1.0 / sqrt(m[0] * m[0] + m[1] * m[1] + m[2] * m[2]);
    m[0] *= inv; m[1] *= inv; m[2] *= inv;
    double su = sqrt(o->sigma_u2), sc = sqrt(o->sigma_c2);
    double D = beckmann_eval(&B, m);
    double gc = gram_charlier(o->wind_direction, o->wind_speed, su, sc, m);
    D *= gc > 0.0 ? gc : 0.0;
    double result = D / (4.0 * wi[2] * wo[2]);
    if (o->shadowing) result *= beckmann_g_hc(&B, wi, wo, m);
    double cos_chi = wo[0] * m[0] + wo[1] * m[1] + wo[2] * m[2];
    cos_chi = fmax(fmin(cos_chi, 0.999999999), -0.999999999);
    double sin_chi = fmax(fmin(sqrt(1.0 - cos_chi * cos_chi), 0.999999999), -0.999999999);
    return result * fresnel_legacy(o->n_real, o->n_imag, cos_chi, sin_chi) * PI;
}
/* eval_underlight(wi, wo) (:468-491) */
static double eval_underlight(const ocean_state_t *o, const double wi[3], const double wo[3]) {
    if (o->wavelength < 400.0 || o->wavelength > 700.0) return 0.0;
    double t_d = eval_transmittance(o, o->tr_down, wi[2], wi[0], wi[1]);
    double t_u = eval_transmittance(o, o->tr_up, wo[2], wi[0], wi[1]);
    return (1.0 / (o->n_real * o->n_real + o->n_imag * o->n_imag)) * (o->r_omega * t_u * t_d) /
           (1.0 - o->underlight_attn * o->r_omega);
}

/* BSDF::eval in Radiance mode (:561-661), `si.wi` = wi, returns value * cos(theta_o) */
double ocean_eval(const ocean_state_t *o, double wix, double wiy, double wiz, double wox, double woy, double woz) {
    if (!(wiz > 0.0 && woz > 0.0)) return 0.0;
    double wo_hat[3] = { wox, woy, woz }, wi_hat[3] = { wix, wiy, wiz }; /* wo_hat = wo, wi_hat = si.wi */
    double wc = o->whitecap_reflectance;
    double ul = eval_underlight(o, wo_hat, wi_hat);
    double glint = eval_glint(o, wo_hat, wi_hat);
    double result = wc + (1.0 - wc) * ul + (1.0 - o->whitecap_coverage) * glint;
    return result * woz / PI;
}

/* BSDF::pdf (:663-713) */
double ocean_pdf(const ocean_state_t *o, double wix, double wiy, double wiz, double wox, double woy, double woz) {
    if (!(wiz > 0.0 && woz > 0.0)) return 0.0;
    double wc = o->whitecap_reflectance;
    double t_i = eval_transmittance(o, o->tr_down, wiz, wix, wiy);
    double pd = t_i * (1.0 - wc) + wc, ps = 1.0 - o->whitecap_coverage;
    ps = ps / (ps + pd);
    pd = 1.0 - ps;
    pd *= woz / PI;
    double wi[3] = { wix, wiy, wiz };
    double H[3] = { wox + wix, woy + wiy, woz + wiz };
    double inv = 1.0 / sqrt(H[0] * H[0] + H[1] * H[1] + H[2] * H[2]);
    H[0] *= inv; H[1] *= inv; H[2] *= inv;
    beckmann_t B = beckmann_make(o);
    ps *= beckmann_eval(&B, H) * beckmann_g1(&B, wi, H) / (4.0 * wiz);
    return pd + ps;
}

/* BSDF::sample (:494-559): returns eval / pdf */
double ocean_sample(const ocean_state_t *o, double wix, double wiy, double wiz, double s1, double u1, double u2, double *wo) {
    wo[0] = wo[1] = 0.0; wo[2] = 1.0;
    if (!(wiz > 0.0)) return 0.0;
    double wc = o->whitecap_reflectance;
    double t_i = eval_transmittance(o, o->tr_down, wiz, wix, wiy);
    double pd = wc + t_i * (1.0 - wc), ps = 1.0 - o->whitecap_coverage;
    ps = ps / (ps + pd);
    pd = 1.0 - ps;
    if (s1 < pd) {
        /* cosine hemisphere: reuse the oracle's warp through a local copy (warp.h:412-433) */
        double x = 2.0 * u1 - 1.0, y = 2.0 * u2 - 1.0, r, phi;
        if (x == 0.0 && y == 0.0) { r = 0.0; phi = 0.0; }
        else if (fabs(x) < fabs(y)) { r = y; phi = 0.5 * PI - 0.25 * PI * x / y; }
        else { r = x; phi = 0.25 * PI * y / x; }
        wo[0] = r * cos(phi); wo[1] = r * sin(phi);
        wo[2] = sqrt(fmax(0.0, 1.0 - wo[0] * wo[0] - wo[1] * wo[1]));
    } else {
        beckmann_t B = beckmann_make(o);
        double wi[3] = { wix, wiy, wiz }, H[3];
        beckmann_sample(&B, wi, u1, u2, H);
        double dp = wi[0] * H[0] + wi[1] * H[1] + wi[2] * H[2];
        wo[0] = 2.0 * dp * H[0] - wi[0]; wo[1] = 2.0 * dp * H[1] - wi[1]; wo[2] = 2.0 * dp * H[2] - wi[2];
    }
    double pdf = ocean_pdf(o, wix, wiy, wiz, wo[0], wo[1], wo[2]);
    if (!(pdf > 0.0)) return 0.0;
    return ocean_eval(o, wix, wiy, wiz, wo[0], wo[1], wo[2]) / pdf;
}

/* ------------------------------------------------------------------ polarized glint */
typedef struct { double re, im; } cplx;
static cplx c_make(double re, double im) { cplx r = { re, im }; return r; }
static cplx c_add(cplx a, cplx b) { return c_make(a.re + b.re, a.im + b.im); }
static cplx c_sub(cplx a, cplx b) { return c_make(a.re - b.re, a.im - b.im); }
static cplx c_mul(cplx a, cplx b) { return c_make(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
static cplx c_scale(cplx a, double s) { return c_make(a.re * s, a.im * s); }
static cplx c_div(cplx a, cplx b) { double d = b.re * b.re + b.im * b.im; return c_make((a.re * b.re + a.im * b.im) / d, (a.im * b.re - a.re * b.im) / d); }
static cplx c_conj(cplx a) { return c_make(a.re, -a.im); }
static cplx c_sqrt(cplx a) {
    double m = sqrt(a.re * a.re + a.im * a.im);
    double re = sqrt(0.5 * (m + a.re)), im = sqrt(fmax(0.0, 0.5 * (m - a.re)));
    return c_make(re, a.im < 0.0 ? -im : im);
}
static double c_abs2(cplx a) { return a.re * a.re + a.im * a.im; }

/* fresnel_sunglint_polarized (oceanprops.h:443-545): Mueller matrix of the reflection on a facet,
 * `wi` = propagation direction of the incident light, `wo` = of the reflected light (local frame). */
static void fresnel_polarized(double nr, double ni, const double wi_in[3], const double wo_in[3], double M[16]) {
    cplx n1 = c_make(1.0, 0.0), n2 = c_make(nr, ni);
    double mu_i = fabs(wi_in[2]), mu_o = fabs(wo_in[2]);
    double phi_i = -atan2(wi_in[1], wi_in[0]), phi_o = -atan2(wo_in[1], wo_in[0]);
    if (mu_i > 0.9999999) mu_i = 0.9999999;
    if (mu_o > 0.9999999) mu_o = 0.9999999;
    double si = sqrt(1.0 - mu_i * mu_i), so = sqrt(1.0 - mu_o * mu_o);
    double wi[3] = { si * cos(phi_i), si * sin(phi_i), -mu_i }, wo[3] = { so * cos(phi_o), so * sin(phi_o), mu_o };
    double kd[3] = { wi[0] - wo[0], wi[1] - wo[1], wi[2] - wo[2] };
    double kd2 = kd[0] * kd[0] + kd[1] * kd[1] + kd[2] * kd[2];
    double mu_il = (kd[0] * wi[0] + kd[1] * wi[1] + kd[2] * wi[2]) / sqrt(kd2);
    cplx ratio = c_div(c_mul(n1, n1), c_mul(n2, n2));
    cplx mu_refr = c_sqrt(c_sub(c_make(1.0, 0.0), c_scale(ratio, 1.0 - mu_il * mu_il)));
    cplx a = c_scale(n1, mu_il), b = c_mul(n2, mu_refr), c = c_scale(n2, mu_il), d = c_mul(n1, mu_refr);
    cplx R_r = c_div(c_sub(a, b), c_add(a, b)), R_l = c_div(c_sub(c, d), c_add(c, d));
    /* polarisation frames */
    double pvi[3], tvi[3], pvo[3], tvo[3];
    if (wi[0] == 0.0 && wi[1] == 0.0 && wi[2] == -1.0) { pvi[0] = 0; pvi[1] = 1; pvi[2] = 0; }
    else { double x = -wi[1], y = wi[0], n = sqrt(x * x + y * y); pvi[0] = x / n; pvi[1] = y / n; pvi[2] = 0; } /* z x wi */
    tvi[0] = pvi[1] * wi[2] - pvi[2] * wi[1]; tvi[1] = pvi[2] * wi[0] - pvi[0] * wi[2]; tvi[2] = pvi[0] * wi[1] - pvi[1] * wi[0];
    if (wo[0] == 0.0 && wo[1] == 0.0 && wo[2] == 1.0) { pvo[0] = 0; pvo[1] = 1; pvo[2] = 0; }
    else { double x = -wo[1], y = wo[0], n = sqrt(x * x + y * y); pvo[0] = x / n; pvo[1] = y / n; pvo[2] = 0; }
    tvo[0] = pvo[1] * wo[2] - pvo[2] * wo[1]; tvo[1] = pvo[2] * wo[0] - pvo[0] * wo[2]; tvo[2] = pvo[0] * wo[1] - pvo[1] * wo[0];
    double pi_wo = pvi[0] * wo[0] + pvi[1] * wo[1] + pvi[2] * wo[2], po_wi = pvo[0] * wi[0] + pvo[1] * wi[1] + pvo[2] * wi[2];
    double ti_wo = tvi[0] * wo[0] + tvi[1] * wo[1] + tvi[2] * wo[2], to_wi = tvo[0] * wi[0] + tvo[1] * wi[1] + tvo[2] * wi[2];
    cplx f_tt = c_add(c_scale(R_r, pi_wo * po_wi), c_scale(R_l, ti_wo * to_wi));
    cplx f_tp = c_add(c_scale(R_r, -ti_wo * po_wi), c_scale(R_l, pi_wo * to_wi));
    cplx f_pt = c_add(c_scale(R_r, -pi_wo * to_wi), c_scale(R_l, ti_wo * po_wi));
    cplx f_pp = c_add(c_scale(R_r, ti_wo * to_wi), c_scale(R_l, pi_wo * po_wi));
    double cx[3] = { wi[1] * wo[2] - wi[2] * wo[1], wi[2] * wo[0] - wi[0] * wo[2], wi[0] * wo[1] - wi[1] * wo[0] };
    double c2 = cx[0] * cx[0] + cx[1] * cx[1] + cx[2] * cx[2];
    int collinear = wo[0] == -wi[0] && wo[1] == -wi[1] && wo[2] == -wi[2];
    double coeff = 1.0 / (collinear ? 0.000001 : c2 * c2);
    double tt = c_abs2(f_tt), tp = c_abs2(f_tp), pt = c_abs2(f_pt), pp = c_abs2(f_pp);
    cplx ttp = c_mul(f_tt, c_conj(f_tp)), ptpp = c_mul(f_pt, c_conj(f_pp));
    cplx ttpt = c_mul(f_tt, c_conj(f_pt)), tppp = c_mul(f_tp, c_conj(f_pp));
    cplx ttpp = c_mul(f_tt, c_conj(f_pp)), tppt = c_mul(f_tp, c_conj(f_pt));
    M[0] = 0.5 * coeff * (tt + tp + pt + pp); M[1] = 0.5 * coeff * (tt - tp + pt - pp);
    M[2] = -coeff * (ttp.re + ptpp.re); M[3] = -coeff * (ttp.im + ptpp.im);
    M[4] = 0.5 * coeff * (tt + tp - pt - pp); M[5] = 0.5 * coeff * (tt - tp - pt + pp);
    M[6] = -coeff * (ttp.re - ptpp.re); M[7] = -coeff * (ttp.im - ptpp.im);
    M[8] = -coeff * (ttpt.re + tppp.re); M[9] = -coeff * (ttpt.re - tppp.re);
    M[10] = coeff * (ttpp.re + tppt.re); M[11] = coeff * (ttpp.im - tppt.im);
    M[12] = coeff * (ttpt.im + tppp.im); M[13] = coeff * (ttpt.im - tppp.im);
    M[14] = -coeff * (ttpp.im + tppt.im); M[15] = coeff * (ttpp.re - tppt.re);
}

/* Scalar factor of the glint lobe without the Fresnel term (eval_glint, :405-420) */
static double glint_geometry(const ocean_state_t *o, const double wi[3], const double wo[3]) {
    beckmann_t B = beckmann_make(o);
    double m[3] = { wi[0] + wo[0], wi[1] + wo[1], wi[2] + wo[2] };
    double inv = 1.0 / sqrt(m[0] * m[0] + m[1] * m[1] + m[2] * m[2]);
    m[0] *= inv; m[1] *= inv; m[2] *= inv;
    double gc = gram_charlier(o->wind_direction, o->wind_speed, sqrt(o->sigma_u2), sqrt(o->sigma_c2), m);
    double result = beckmann_eval(&B, m) * (gc > 0.0 ? gc : 0.0) / (4.0 * wi[2] * wo[2]);
    if (o->shadowing) result *= beckmann_g_hc(&B, wi, wo, m);
    return result * PI;
}

/* Polarized BSDF::eval in Radiance mode (:561-661) BEFORE the basis rotations: returns the
 * depolarizing part `dep` (whitecaps + underlight, times cos/pi), and the glint Mueller matrix in
 * the meridian-plane bases (times (1 - coverage) cos/pi). */
void ocean_eval_polarized(const ocean_state_t *o, const double wi_si[3], const double wo[3], double *dep, double glint[16]) {
    for (int i = 0; i < 16; ++i) glint[i] = 0.0;
    *dep = 0.0;
    if (!(wi_si[2] > 0.0 && wo[2] > 0.0)) return;
    const double *wo_hat = wo, *wi_hat = wi_si;
    double wc = o->whitecap_reflectance;
    double ul = eval_underlight(o, wo_hat, wi_hat);
    double scale = wo[2] / PI;
    *dep = (wc + (1.0 - wc) * ul) * scale;
    double g = glint_geometry(o, wo_hat, wi_hat);
    double minus_wi[3] = { -wo_hat[0], -wo_hat[1], -wo_hat[2] }; /* eval_glint(wi := wo_hat, wo := wi_hat): F(-wi, wo) */
    double F[16];
    fresnel_polarized(o->n_real, o->n_imag, minus_wi, wi_hat, F);
    for (int i = 0; i < 16; ++i) glint[i] = F[i] * g * (1.0 - o->whitecap_coverage) * scale;
}

/* ====================================================================================================
 * Isotropic-Beckmann glint family (TEST INFRASTRUCTURE, as above).  Restates
 *   ERP/bsdfs/ocean_mishchenko.cpp:136-142 update(), :144-226 sample, :228-296 eval, :298-325 pdf
 *   ERP/bsdfs/ocean_grasp.cpp:155-165 parameters_changed, :172-174 eval_sigma, :186-188 whitecaps,
 *       :202-238 eval_glint, :246-255 lambda, :267-352 sample, :354-455 eval, :457-512 pdf
 *   ERP/bsdfs/maignan.cpp:105-166 eval_maignan, :168-194 sample, :196-211 eval, :213-224 pdf
 *   oceanprops.h:330-336 whitecap_coverage_monahan, :350-363 whitecap_reflectance_frouin,
 *       :443-545 fresnel_sunglint_polarized, :582-584 cox_munk_msslope_squared
 * The exterior index is real, so only n_lower / n_ext enters the Fresnel coefficients.
 * ==================================================================================================== */
void glint_init(glint_state_t *g, int type, const float *p) {
    memset(g, 0, sizeof *g);
    g->type = type;
    if (type == ERTB_BSDF_OCEAN_MISHCHENKO) { /* params: wind_speed, eta, k, ext_ior */
        g->sigma = sqrt(0.5 * (0.00512 * (double) p[0] + 0.003));
        g->nr = (double) p[1] / (double) p[3]; g->ni = (double) p[2] / (double) p[3];
    } else if (type == ERTB_BSDF_OCEAN_GRASP) { /* params: wavelength, wind_speed, eta, k, ext_ior, wbr */
        double wl = p[0], ws = p[1];
        g->sigma = sqrt(0.5 * (0.00512 * ws + 0.003));
        g->nr = (double) p[2] / (double) p[4]; g->ni = (double) p[3] / (double) p[4];
        g->coverage = fmax(0.0, fmin(1.0, 2.95e-06 * pow(ws, 3.52)));
        double um = wl * 0.001;
        double eff = um >= 0.6 ? 0.22 * exp(-1.75 * pow(um - 0.6, 0.99)) : 0.22;
        g->whitecap = g->coverage * eff;
        g->wbr = p[5];
    } else { /* maignan: C, ndvi, refr_re, refr_im, ext_ior */
        g->cexp = (double) p[0] * exp(-(double) p[1]);
        g->nr = (double) p[2] / (double) p[4]; g->ni = (double) p[3] / (double) p[4];
    }
}
static beckmann_t glint_beckmann(const glint_state_t *g) {
    beckmann_t B;
    B.au = B.av = fmax(sqrt(2.0) * g->sigma, 1e-4); /* MicrofacetDistribution(Beckmann, alpha, sample_visible) */
    B.angle = 0.0;
    B.aup = B.avp = B.au;
    B.corr = 0.0;
    return B;
}
static void half_vector(const double a[3], const double b[3], double m[3]) {
    m[0] = a[0] + b[0]; m[1] = a[1] + b[1]; m[2] = a[2] + b[2];
    double inv = 1.0 / sqrt(m[0] * m[0] + m[1] * m[1] + m[2] * m[2]);
    m[0] *= inv; m[1] *= inv; m[2] *= inv;
}
static double dot3d(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
/* ocean_grasp.cpp:246-255 */
static double grasp_lambda(const double v[3], double sigma) {
    double st = sigma * sqrt(1.0 - v[2] * v[2]) / v[2];
    return 0.5 * (sqrt(2.0 / PI) * st * exp(-1.0 / (2.0 * st * st)) - (1.0 - erf(1.0 / (sqrt(2.0) * st))));
}
/* scalar geometry factor in front of the Fresnel matrix in BSDF::eval (cosine included where the plugin has it) */
static double glint_geometry_iso(const glint_state_t *g, const double wi[3], const double wo[3]) {
    beckmann_t B = glint_beckmann(g);
    double m[3];
    half_vector(wi, wo, m);
    double D = beckmann_eval(&B, m);
    if (g->type == ERTB_BSDF_OCEAN_MISHCHENKO) { /* :249-254 */
        if (D == 0.0) return 0.0;
        return D * beckmann_g_hc(&B, wi, wo, m) / (4.0 * wi[2]);
    }
    /* ocean_grasp eval_glint(wi := wo_hat = wo, wo := wi_hat = si.wi) :202-238, then (1 - coverage) cos / pi */
    double G = 1.0 / (1.0 + grasp_lambda(wo, g->sigma) + grasp_lambda(wi, g->sigma));
    if (dot3d(wo, m) * wo[2] <= 0.0) G = 0.0;
    if (dot3d(wi, m) * wi[2] <= 0.0) G = 0.0;
    double value = PI * D * G / (4.0 * wo[2] * wi[2]);
    return (1.0 - g->coverage) * value * wo[2] / PI;
}
static double maignan_c(const glint_state_t *g, const double wi[3], const double wo[3]) { /* :116-136 */
    double ci = wi[2], co = wo[2];
    double si = sqrt(fmax(0.0, 1.0 - ci * ci)), so = sqrt(fmax(0.0, 1.0 - co * co));
    double spi = 0, cpi = 1, spo = 0, cpo = 1; /* Frame3f::sincos_phi: (0, 1) along the normal */
    double ni2 = wi[0] * wi[0] + wi[1] * wi[1], no2 = wo[0] * wo[0] + wo[1] * wo[1];
    if (ni2 > 0.0) { double inv = 1.0 / sqrt(ni2); spi = wi[1] * inv; cpi = wi[0] * inv; }
    if (no2 > 0.0) { double inv = 1.0 / sqrt(no2); spo = wo[1] * inv; cpo = wo[0] * inv; }
    double cdphi = cpi * cpo + spi * spo;
    double cT = ci * co + si * so * cdphi;
    double tan_a = sqrt((1.0 - cT) / (1.0 + cT));
    return g->cexp * exp(-tan_a) / (4.0 * (ci + co));
}
static double glint_dep(const glint_state_t *g, const double wo[3]) {
    if (g->type != ERTB_BSDF_OCEAN_GRASP) return 0.0;
    return (g->whitecap + (1.0 - g->coverage) * g->wbr) * wo[2] / PI; /* :395-404, :436 */
}

double glint_eval(const glint_state_t *g, const double wi[3], const double wo[3]) {
    if (!(wi[2] > 0.0 && wo[2] > 0.0)) return 0.0;
    double in_fwd[3] = { -wo[0], -wo[1], -wo[2] }, F[16];
    fresnel_polarized(g->nr, g->ni, in_fwd, wi, F); /* F(-wo_hat, wi_hat), wo_hat = wo, wi_hat = si.wi */
    if (g->type == ERTB_BSDF_MAIGNAN) return maignan_c(g, wi, wo) * F[0];
    return glint_dep(g, wo) + glint_geometry_iso(g, wi, wo) * F[0];
}
double glint_pdf(const glint_state_t *g, const double wi[3], const double wo[3]) {
    if (!(wi[2] > 0.0 && wo[2] > 0.0)) return 0.0;
    if (g->type == ERTB_BSDF_MAIGNAN) return wo[2] / PI;
    beckmann_t B = glint_beckmann(g);
    double m[3];
    half_vector(wi, wo, m);
    double spec = beckmann_eval(&B, m) * beckmann_g1(&B, wi, m) / (4.0 * wi[2]);
    if (g->type == ERTB_BSDF_OCEAN_MISHCHENKO)
        return (dot3d(wi, m) > 0.0 && dot3d(wo, m) > 0.0) ? spec : 0.0;
    double p_spec = 1.0 / (g->wbr + 1.0), p_diff = 1.0 - p_spec, pc = wo[2] / PI;
    return g->coverage * pc + (1.0 - g->coverage) * (p_diff * pc + p_spec * spec);
}
static void cosine_hemisphere_d(double u1, double u2, double wo[3]) { /* warp.h:412-433 */
    double x = 2.0 * u1 - 1.0, y = 2.0 * u2 - 1.0, r, phi;
    if (x == 0.0 && y == 0.0) { r = 0.0; phi = 0.0; }
    else if (fabs(x) < fabs(y)) { r = y; phi = 0.5 * PI - 0.25 * PI * x / y; }
    else { r = x; phi = 0.25 * PI * y / x; }
    wo[0] = r * cos(phi); wo[1] = r * sin(phi);
    wo[2] = sqrt(fmax(0.0, 1.0 - wo[0] * wo[0] - wo[1] * wo[1]));
}
/* validity of a Mishchenko sample (:170-178): distr.sample's pdf != 0 and the reflected direction is up */
static int mishchenko_valid(const beckmann_t *B, const double wi[3], const double wo[3], const double m[3]) {
    double pdf_m = beckmann_eval(B, m) * beckmann_g1(B, wi, m) * fabs(dot3d(wi, m)) / wi[2];
    return pdf_m != 0.0 && wo[2] > 0.0;
}
double glint_sample(const glint_state_t *g, const double wi[3], double s1, double u1, double u2, double wo[3]) {
    wo[0] = wo[1] = 0.0; wo[2] = 1.0;
    if (!(wi[2] > 0.0)) return 0.0;
    beckmann_t B = glint_beckmann(g);
    if (g->type == ERTB_BSDF_MAIGNAN) { /* :183-193: C * F, NOT divided by the pdf */
        cosine_hemisphere_d(u1, u2, wo);
        if (!(wo[2] / PI > 0.0)) return 0.0;
        double in_fwd[3] = { -wo[0], -wo[1], -wo[2] }, F[16];
        fresnel_polarized(g->nr, g->ni, in_fwd, wi, F);
        return maignan_c(g, wi, wo) * F[0];
    }
    if (g->type == ERTB_BSDF_OCEAN_MISHCHENKO) {
        double m[3];
        beckmann_sample(&B, wi, u1, u2, m);
        double dp = dot3d(wi, m);
        wo[0] = 2.0 * dp * m[0] - wi[0]; wo[1] = 2.0 * dp * m[1] - wi[1]; wo[2] = 2.0 * dp * m[2] - wi[2];
        if (!mishchenko_valid(&B, wi, wo, m)) return 0.0;
        double in_fwd[3] = { -wo[0], -wo[1], -wo[2] }, F[16];
        fresnel_polarized(g->nr, g->ni, in_fwd, wi, F);
        return F[0] * beckmann_g_hc(&B, wi, wo, m) / beckmann_g1(&B, wi, m);
    }
    /* ocean_grasp :285-351 */
    double p_spec = 1.0 / (g->wbr + 1.0), p_diff = 1.0 - p_spec;
    int foam = s1 < g->coverage;
    double s1p = (s1 - g->coverage) / (1.0 - g->coverage);
    if (foam || s1p < p_diff) {
        cosine_hemisphere_d(u1, u2, wo);
    } else {
        double m[3];
        beckmann_sample(&B, wi, u1, u2, m);
        double dp = dot3d(wi, m);
        wo[0] = 2.0 * dp * m[0] - wi[0]; wo[1] = 2.0 * dp * m[1] - wi[1]; wo[2] = 2.0 * dp * m[2] - wi[2];
    }
    double pdf = glint_pdf(g, wi, wo);
    if (!(pdf > 0.0)) return 0.0;
    return glint_eval(g, wi, wo) / pdf;
}
void glint_polarized(const glint_state_t *g, int weight, const double wi[3], const double wo[3], double *dep, double M[16]) {
    for (int i = 0; i < 16; ++i) M[i] = 0.0;
    *dep = 0.0;
    if (!(wi[2] > 0.0 && wo[2] > 0.0)) return;
    double in_fwd[3] = { -wo[0], -wo[1], -wo[2] }, F[16], scale, dscale = 1.0;
    fresnel_polarized(g->nr, g->ni, in_fwd, wi, F);
    if (g->type == ERTB_BSDF_MAIGNAN) {
        scale = maignan_c(g, wi, wo); /* eval and sample weight coincide (maignan.cpp:189-193, :207-210) */
    } else if (g->type == ERTB_BSDF_OCEAN_MISHCHENKO && weight) {
        beckmann_t B = glint_beckmann(g);
        double m[3];
        half_vector(wi, wo, m);
        if (!mishchenko_valid(&B, wi, wo, m)) return;
        scale = beckmann_g_hc(&B, wi, wo, m) / beckmann_g1(&B, wi, m);
    } else {
        scale = glint_geometry_iso(g, wi, wo);
        *dep = glint_dep(g, wo);
        if (weight) { /* ocean_grasp: eval / pdf */
            double pdf = glint_pdf(g, wi, wo);
            dscale = pdf > 0.0 ? 1.0 / pdf : 0.0;
        }
    }
    *dep *= dscale;
    for (int i = 0; i < 16; ++i) M[i] = F[i] * scale * dscale;
}
