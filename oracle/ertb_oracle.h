/*
 * ertb_oracle.h -- CPU restatement (double precision, plain C) of the reference's
 * null-collision volumetric path tracer.  TEST INFRASTRUCTURE ONLY: it is the
 * checker for the CUDA path (tests/, __graft_entry__.smoke(), bench.py's
 * cpu_baseline / --impl reference legs).  Nothing under eradiate_b200/ may
 * import, link or call it.
 *
 * It consumes the same flat scene descriptor as the CUDA library
 * (include/eradiate_b200.h) but follows the reference's own formulation:
 * world-space 3D rays, shape intersections, medium AABB, PCG32 sampler.
 */
#ifndef ERTB_ORACLE_H
#define ERTB_ORACLE_H

#include "../include/eradiate_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

const char *ertbo_last_error(void);

/* MI/src/render/integrator.cpp:398-520 + volpath.cpp:93-572, n_threads OpenMP threads
 * (0 = all).  Same outputs as ertb_render. */
int ertbo_render(const ertb_scene_desc *desc, int sensor, uint64_t seed, uint64_t spp,
                 uint64_t sample_offset, double *sum_wl, double *sum_l, double *sum_l2,
                 ertb_render_stats *stats, int n_threads);

int ertbo_render_stokes(const ertb_scene_desc *desc, int sensor, uint64_t seed, uint64_t spp,
                        uint64_t sample_offset, double *sum_wl, double *sum_l, double *sum_l2,
                        double *sum_stokes, ertb_render_stats *stats, int n_threads);
/* Free flights sampled by the last ertbo_render / ertbo_render_stokes call: out[0] = main walk
 * (Medium::sample_interaction calls of volpath.cpp:220), out[1] = shadow rays that can contribute
 * (volpath.cpp:479; rays ending on an opaque surface or carrying a zero weight excluded). */
void ertbo_last_flights(uint64_t out[2]);
int ertbo_piecewise_sample(const ertb_scene_desc *desc, double half_width, size_t n, const double *o,
                           const double *d, const double *sample, const double *si_t, double *t,
                           double *tr, double *pdf);
int ertbo_piecewise_eval(const ertb_scene_desc *desc, double half_width, size_t n, const double *o,
                         const double *d, const double *si_t, double *tr, double *pdf, int *escaped);
int ertbo_phase_mueller(const ertb_scene_desc *desc, int leaf, size_t n, const double *wi, const double *wo,
                        double *mueller, double *pdf);

/* BSDF::eval of a polarized variant: 4x4 Mueller matrices (16*n, row-major) in the implicit Stokes bases
 * of -wo and wi, local frame (z = normal). */
int ertbo_bsdf_mueller(const ertb_scene_desc *desc, size_t n, const double *wi, const double *wo, double *mueller);

/* Point-wise plugin evaluations (double). Same conventions as ertb_kat_*. */
/* BSDF::pdf of the ground BSDF (solid-angle density of BSDF::sample) */
int ertbo_bsdf_pdf(const ertb_scene_desc *desc, size_t n, const double *wi, const double *wo, double *out);
int ertbo_bsdf_eval(const ertb_scene_desc *desc, size_t n, const double *wi, const double *wo,
                    double *out);
int ertbo_bsdf_sample(const ertb_scene_desc *desc, size_t n, const double *wi, const double *u,
                      double *wo, double *weight);
int ertbo_phase_eval(const ertb_scene_desc *desc, int leaf, size_t n, const double *cos_theta,
                     double *out);
int ertbo_phase_sample(const ertb_scene_desc *desc, int leaf, size_t n, const double *u,
                       double *cos_theta, double *weight, double *pdf);
int ertbo_sensor_ray(const ertb_scene_desc *desc, int sensor, size_t n, const double *film_sample,
                     const double *aperture_sample, double *origin, double *dir, double *weight);

/* distr_1d.h restatements exposed for the golden-vector tests. */
int ertbo_distr_regular(const float *pdf, int n, size_t nq, const double *u, double *x_sampled,
                        const double *xq, double *pdf_eval, double *integral);
int ertbo_distr_irregular(const float *nodes, const float *pdf, int n, size_t nq, const double *u,
                          double *x_sampled, const double *xq, double *pdf_eval, double *integral);

/* Warps (MI/include/mitsuba/core/warp.h) for KATs */
void ertbo_square_to_uniform_disk_concentric(double u, double v, double *x, double *y);
void ertbo_square_to_cosine_hemisphere(double u, double v, double *out3);
void ertbo_square_to_uniform_hemisphere(double u, double v, double *out3);

/* Analytic helpers used by the tests: vertical optical thickness etc. */
int ertbo_medium_lookup(const ertb_scene_desc *desc, size_t n, const double *p_xyz, double *sigma_t,
                        double *albedo);

#ifdef __cplusplus
}
#endif
/* canopy KATs (ertb_oracle_canopy.c): nearest leaf along world-space rays; bilambertian
 * eval (mode 0) / pdf (1) / sample (2) of leaf group `group` in the leaf's local frame */
int ertbo_canopy_intersect(const ertb_scene_desc *desc, size_t n, const double *o, const double *d,
                           const double *tmax, double *t, double *normal, int *group);
int ertbo_leaf_bsdf(const ertb_scene_desc *desc, int group, int mode, size_t n, const double *wi, double *wo,
                    const double *u, double *out);

#endif
