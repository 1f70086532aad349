/* ertb_oracle_ocean.h -- ocean_legacy BSDF restatement (oracle side). See ertb_oracle_ocean.c */
#ifndef ERTB_ORACLE_OCEAN_H
#define ERTB_ORACLE_OCEAN_H
typedef struct ocean_state {
    int ready;
    double wavelength, wind_speed, wind_direction, chlorinity, pigmentation;
    int shadowing;
    double n_real, n_imag;           /* water index of refraction */
    double r_omega;                  /* underlight reflectance */
    double whitecap_coverage, whitecap_reflectance, underlight_attn;
    double sigma2, sigma_c2, sigma_u2; /* Cox-Munk */
    double *tr_down, *tr_up;         /* 64-node transmittance tables */
    int n_tab;
} ocean_state_t;
int ocean_init(ocean_state_t *o, const float *params);
void ocean_free(ocean_state_t *o);
double ocean_eval(const ocean_state_t *o, double wix, double wiy, double wiz, double wox, double woy, double woz);
double ocean_sample(const ocean_state_t *o, double wix, double wiy, double wiz, double s1, double u1, double u2, double *wo);
double ocean_pdf(const ocean_state_t *o, double wix, double wiy, double wiz, double wox, double woy, double woz);
void ocean_eval_polarized(const ocean_state_t *o, const double wi_si[3], const double wo[3], double *dep, double glint[16]);

/* Isotropic-Beckmann glint family: ERP/bsdfs/ocean_mishchenko.cpp, ocean_grasp.cpp, maignan.cpp */
typedef struct glint_state {
    int type;                 /* enum ertb_bsdf_type */
    double nr, ni;            /* index of the lower medium relative to the (real) exterior index */
    double sigma;             /* sqrt(0.5 * Cox-Munk mean square slope); Beckmann alpha = sqrt(2) * sigma */
    double coverage, whitecap, wbr; /* ocean_grasp: Monahan coverage, Frouin whitecap reflectance, water body */
    double cexp;              /* maignan: C * exp(-ndvi) */
} glint_state_t;
void glint_init(glint_state_t *g, int type, const float *params);
/* BSDF::eval / pdf / sample exactly as the plugins return them (Radiance mode, wi = si.wi) */
double glint_eval(const glint_state_t *g, const double wi[3], const double wo[3]);
double glint_pdf(const glint_state_t *g, const double wi[3], const double wo[3]);
double glint_sample(const glint_state_t *g, const double wi[3], double s1, double u1, double u2, double wo[3]);
/* Polarized variants BEFORE the basis rotations: depolarizing part + Mueller matrix in the meridian-plane
 * bases.  `weight` = 0: BSDF::eval(wi, wo); 1: the weight BSDF::sample returns for the sampled `wo`. */
void glint_polarized(const glint_state_t *g, int weight, const double wi[3], const double wo[3], double *dep, double M[16]);
#endif
