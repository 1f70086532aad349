/*
 * ertb_oracle.c -- CPU oracle: double-precision restatement of the reference's
 * null-collision volumetric path tracer (Eradiate 1.1.0 / eradiate-mitsuba 0.4.3).
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.  The product
 * (eradiate_b200/) never does.
 *
 * Parity pinning: the plugin-level functions below are checked against every
 * golden vector / numpy reference implementation the reference's own tests hold
 * for this path (tests/test_oracle_golden.py lists them with file:line).  The
 * reference itself (Mitsuba + Dr.Jit, cmake + generated config headers) cannot be
 * built or imported in this environment, so the END-TO-END render of the oracle
 * is pinned only through the reference's analytic system tests (BRF == rho,
 * L = rho E / pi, RPV(k=1,g=0,rho_c=1) == Lambertian, transmittance KATs).
 *
 * "MI"  = /root/reference/ext/mitsuba, "ERP" = MI/src/eradiate_plugins.
 * Nothing here is copied from the reference; each function restates the cited
 * lines' arithmetic in plain scalar C.
 */
#include "ertb_oracle.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "ertb_oracle_ocean.h"
#include "ertb_oracle_canopy.h"
#include "ertb_oracle_measured.h"

#define PI 3.14159265358979323846
#define INV_PI (1.0 / PI)
#define INV_TWO_PI (0.5 / PI)
#define INV_FOUR_PI (0.25 / PI)

/* MI/include/mitsuba/core/math.h:18-23 (double variant) */
#define RAY_EPS (DBL_EPSILON * 0.5 * 1500.0) /* dr::Epsilon<double> = 2^-53 */
#define SHADOW_EPS (RAY_EPS * 10.0)

static __thread char g_err[512];
const char *ertbo_last_error(void) { return g_err; }
static int fail(const char *msg) {
    snprintf(g_err, sizeof g_err, "%s", msg);
    return 1;
}

/* ------------------------------------------------------------------ vectors */
typedef struct { double x, y, z; } v3;
static inline v3 V(double x, double y, double z) { v3 r = { x, y, z }; return r; }
static inline v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vmul(v3 a, double s) { return V(a.x * s, a.y * s, a.z * s); }
static inline v3 vfma(v3 a, double s, v3 b) { return V(a.x * s + b.x, a.y * s + b.y, a.z * s + b.z); }
static inline double vdot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline double vnorm(v3 a) { return sqrt(vdot(a, a)); }
static inline v3 vnormalize(v3 a) { return vmul(a, 1.0 / vnorm(a)); }
static inline v3 vneg(v3 a) { return V(-a.x, -a.y, -a.z); }
static inline double sqr(double x) { return x * x; }
static inline double safe_sqrt(double x) { return sqrt(x > 0.0 ? x : 0.0); }
static inline double safe_acos(double x) { return acos(x < -1.0 ? -1.0 : (x > 1.0 ? 1.0 : x)); }

/* MI/include/mitsuba/core/vector.h:118-140  coordinate_system (Duff et al.) */
static void coordinate_system(v3 n, v3 *s, v3 *t) {
    double sign = copysign(1.0, n.z);
    double a = -1.0 / (sign + n.z);
    double b = n.x * n.y * a;
    *s = V(sign * sqr(n.x) * a + 1.0, sign * b, -sign * n.x);
    *t = V(b, sqr(n.y) * a + sign, -n.y);
}
typedef struct { v3 s, t, n; } frame_t;
static frame_t make_frame(v3 n) { frame_t f; f.n = n; coordinate_system(n, &f.s, &f.t); return f; }
static v3 to_local(const frame_t *f, v3 v) { return V(vdot(v, f->s), vdot(v, f->t), vdot(v, f->n)); }
static v3 to_world(const frame_t *f, v3 v) {
    return vadd(vadd(vmul(f->s, v.x), vmul(f->t, v.y)), vmul(f->n, v.z));
}

/* --------------------------------------------------------------------- RNG */
/* MI/ext/drjit/include/drjit/random.h:108-195 : PCG32 (O'Neill).  One stream per
 * path; the reference seeds one stream per pixel (integrator.cpp:418-421) -- same
 * distribution, different realisation (parity is statistical, SURVEY 8b "Seeds"). */
typedef struct { uint64_t state, inc; } pcg32;
#define PCG_MULT 6364136223846793005ULL
static inline uint32_t pcg_next(pcg32 *r) {
    uint64_t old = r->state;
    r->state = old * PCG_MULT + r->inc;
    uint32_t xs = (uint32_t) (((old >> 18u) ^ old) >> 27u);
    uint32_t rot = (uint32_t) (old >> 59u);
    return (xs >> rot) | (xs << ((-rot) & 31));
}
static inline void pcg_seed(pcg32 *r, uint64_t initstate, uint64_t initseq) {
    r->state = 0;
    r->inc = (initseq << 1u) | 1u;
    pcg_next(r);
    r->state += initstate;
    pcg_next(r);
}
/* MI/src/samplers/independent.cpp:77-86 next_1d (double variant: 53 random bits) */
static inline double next_1d(pcg32 *r) {
    uint64_t a = pcg_next(r) >> 5, b = pcg_next(r) >> 6;
    return (double) (a * 67108864ULL + b) * (1.0 / 9007199254740992.0);
}
static inline uint64_t mix64(uint64_t z) { /* splitmix64 finaliser */
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

/* ------------------------------------------------------------------- warps */
/* MI/include/mitsuba/core/warp.h:54-90 */
void ertbo_square_to_uniform_disk_concentric(double u, double v, double *ox, double *oy) {
    double x = 2.0 * u - 1.0, y = 2.0 * v - 1.0;
    int is_zero = (x == 0.0 && y == 0.0);
    int q13 = fabs(x) < fabs(y);
    double r = q13 ? y : x, rp = q13 ? x : y;
    double phi = 0.25 * PI * rp / r;
    if (q13) phi = 0.5 * PI - phi;
    if (is_zero) phi = 0.0;
    *ox = r * cos(phi);
    *oy = r * sin(phi);
}
/* warp.h:412-433 */
void ertbo_square_to_cosine_hemisphere(double u, double v, double *o) {
    double x, y;
    ertbo_square_to_uniform_disk_concentric(u, v, &x, &y);
    o[0] = x; o[1] = y; o[2] = safe_sqrt(1.0 - x * x - y * y);
}
/* warp.h:374-388 */
void ertbo_square_to_uniform_hemisphere(double u, double v, double *o) {
    double x, y;
    ertbo_square_to_uniform_disk_concentric(u, v, &x, &y);
    double z = 1.0 - (x * x + y * y);
    double s = sqrt(z + 1.0);
    o[0] = x * s; o[1] = y * s; o[2] = z;
}

/* ----------------------------------------------------- 1D distributions */
/* MI/include/mitsuba/core/distr_1d.h: ContinuousDistribution (regular, :300-620)
 * and IrregularContinuousDistribution (:628-1000).  Storage is float (the
 * reference stores ScalarFloat pdf/cdf even though the CDF is accumulated in
 * double, :548-600); in the double variant ScalarFloat = double, so the oracle
 * keeps double storage and parity tests state a tolerance. */
typedef struct {
    int n;            /* nodes */
    int irregular;
    double *nodes;    /* irregular only */
    double *pdf, *cdf;
    double integral, normalization, interval_size, inv_interval_size, x0, x1;
    int valid0, valid1;
} distr_t;

static void distr_free(distr_t *d) { free(d->nodes); free(d->pdf); free(d->cdf); memset(d, 0, sizeof *d); }

static int distr_init(distr_t *d, const float *nodes, const float *pdf, int n) {
    memset(d, 0, sizeof *d);
    if (n < 2) return fail("distribution needs at least two entries");
    d->n = n;
    d->irregular = nodes != NULL;
    d->pdf = malloc(sizeof(double) * n);
    d->cdf = malloc(sizeof(double) * n);
    if (nodes) d->nodes = malloc(sizeof(double) * n);
    for (int i = 0; i < n; ++i) {
        d->pdf[i] = pdf[i];
        if (nodes) d->nodes[i] = nodes[i];
    }
    d->x0 = nodes ? nodes[0] : -1.0;
    d->x1 = nodes ? nodes[n - 1] : 1.0;
    d->interval_size = (d->x1 - d->x0) / (n - 1);
    d->inv_interval_size = 1.0 / d->interval_size;
    d->valid0 = d->valid1 = -1;
    double integral = 0.0;
    for (int i = 0; i < n - 1; ++i) {
        double w = nodes ? ((double) nodes[i + 1] - (double) nodes[i]) : d->interval_size;
        double y0 = pdf[i], y1 = pdf[i + 1];
        if (nodes && !(w > 0.0)) return fail("node positions must be strictly increasing");
        if (y0 < 0.0 || y1 < 0.0) return fail("entries must be non-negative");
        double value = 0.5 * w * (y0 + y1);
        integral += value;
        d->cdf[i] = integral;
        if (value > 0.0) {
            if (d->valid0 < 0) d->valid0 = i;
            d->valid1 = i;
        }
    }
    if (d->valid0 < 0) return fail("no probability mass found");
    d->integral = d->cdf[d->valid1];
    d->normalization = 1.0 / d->integral;
    return 0;
}

/* MI/ext/drjit/include/drjit/util.h:136-187 binary_search (scalar branch) */
static int bsearch_pred_cdf(const distr_t *d, int start, int end, double sample) {
    int iterations = 0;
    if (start < end) { /* log2i(end-start)+1 */
        unsigned v = (unsigned) (end - start);
        int l = 0;
        while (v >>= 1) ++l;
        iterations = l + 1;
    }
    for (int i = 0; i < iterations; ++i) {
        int middle = (start + end) >> 1;
        if (d->cdf[middle] < sample) start = (middle + 1 < end) ? middle + 1 : end;
        else end = middle;
    }
    return start;
}
static int bsearch_pred_nodes(const distr_t *d, double x) {
    int start = 0, end = d->n, iterations = 0;
    unsigned v = (unsigned) (end - start);
    int l = 0;
    while (v >>= 1) ++l;
    iterations = l + 1;
    for (int i = 0; i < iterations; ++i) {
        int middle = (start + end) >> 1;
        int cond = (middle < d->n) ? (d->nodes[middle] < x) : 0;
        if (cond) start = (middle + 1 < end) ? middle + 1 : end;
        else end = middle;
    }
    return start;
}

/* distr_1d.h:429-459 (regular) / :790-822 (irregular): sample */
static double distr_sample(const distr_t *d, double sample) {
    sample *= d->integral;
    int index = bsearch_pred_cdf(d, d->valid0, d->valid1, sample);
    double y0 = d->pdf[index], y1 = d->pdf[index + 1];
    double c0 = index > 0 ? d->cdf[index - 1] : 0.0;
    double w = d->irregular ? d->nodes[index + 1] - d->nodes[index] : d->interval_size;
    sample = (sample - c0) / w;
    double t_linear = (y0 - safe_sqrt(y0 * y0 + 2.0 * sample * (y1 - y0))) / (y0 - y1);
    double t_const = sample / y0;
    double t = (y0 == y1) ? t_const : t_linear;
    if (d->irregular) return t * w + d->nodes[index];
    return ((double) index + t) * d->interval_size + d->x0;
}
/* distr_1d.h:370-392 (regular) / :712-735 (irregular): eval_pdf (unnormalised) */
static double distr_eval_pdf(const distr_t *d, double x) {
    if (!(x >= d->x0 && x <= d->x1)) return 0.0;
    if (d->irregular) {
        int index = bsearch_pred_nodes(d, x);
        if (index > d->n - 1) index = d->n - 1;
        if (index < 1) index = 1;
        index -= 1;
        double xa = d->nodes[index], xb = d->nodes[index + 1];
        double t = (x - xa) / (xb - xa);
        return t * (d->pdf[index + 1] - d->pdf[index]) + d->pdf[index];
    }
    double xs = (x - d->x0) * d->inv_interval_size;
    int index = (int) xs;
    if (index < 0) index = 0;
    if (index > d->n - 2) index = d->n - 2;
    double w1 = xs - index, w0 = 1.0 - w1;
    return w0 * d->pdf[index] + w1 * d->pdf[index + 1];
}

int ertbo_distr_regular(const float *pdf, int n, size_t nq, const double *u, double *xs,
                        const double *xq, double *pe, double *integral) {
    distr_t d;
    if (distr_init(&d, NULL, pdf, n)) return 1;
    for (size_t i = 0; i < nq; ++i) {
        if (u && xs) xs[i] = distr_sample(&d, u[i]);
        if (xq && pe) pe[i] = distr_eval_pdf(&d, xq[i]) * d.normalization;
    }
    if (integral) *integral = d.integral;
    distr_free(&d);
    return 0;
}
int ertbo_distr_irregular(const float *nodes, const float *pdf, int n, size_t nq, const double *u,
                          double *xs, const double *xq, double *pe, double *integral) {
    distr_t d;
    if (distr_init(&d, nodes, pdf, n)) return 1;
    for (size_t i = 0; i < nq; ++i) {
        if (u && xs) xs[i] = distr_sample(&d, u[i]);
        if (xq && pe) pe[i] = distr_eval_pdf(&d, xq[i]) * d.normalization;
    }
    if (integral) *integral = d.integral;
    distr_free(&d);
    return 0;
}

/* ------------------------------------------------------------ scene state */
typedef struct {
    const ertb_scene_desc *desc;
    int spherical;
    double majorant;            /* heterogeneous.cpp:163 m_scale * max(sigma_t) */
    distr_t distr[ERTB_MAX_PHASE];
    int has_distr[ERTB_MAX_PHASE];
    v3 emitter_d;               /* normalised propagation direction */
    ocean_state_t ocean;        /* ocean_legacy precomputed tables */
    glint_state_t glint;        /* ocean_mishchenko / ocean_grasp / maignan */
    int astro;                  /* astroobject.cpp: finite disc instead of the delta directional emitter */
    double astro_cos, astro_omega; /* cos(angular radius), solid angle 2 pi (1 - cos) */
    double *pw_cum, *pw_rcum;   /* piecewise.cpp m_cum_opt_thickness / m_reverse_cum_opt_thickness */
    double pp_half_width;       /* > 0: finite slab bbox in x, y (known-answer tests only) */
    canopy_t canopy;            /* explicit disk-leaf canopy (plane-parallel scenes) */
} scene_t;

static int piecewise_init(scene_t *S);
static int is_glint_family(int type) {
    return type == ERTB_BSDF_OCEAN_MISHCHENKO || type == ERTB_BSDF_OCEAN_GRASP || type == ERTB_BSDF_MAIGNAN;
}

static void scene_free(scene_t *S) {
    for (int i = 0; i < ERTB_MAX_PHASE; ++i)
        if (S->has_distr[i]) distr_free(&S->distr[i]);
    ocean_free(&S->ocean);
    free(S->pw_cum); free(S->pw_rcum);
    canopy_free(&S->canopy);
}

static int scene_init(scene_t *S, const ertb_scene_desc *d) {
    memset(S, 0, sizeof *S);
    if (d->abi_version != ERTB_ABI_VERSION) return fail("ABI version mismatch");
    S->desc = d;
    S->spherical = d->geometry == ERTB_GEOM_SPHERICAL_SHELL;
    if (d->has_medium) {
        if (d->n_layers < 1 || !d->sigma_t || !d->albedo) return fail("medium arrays missing");
        float m = d->sigma_t[0];
        for (int i = 1; i < d->n_layers; ++i) if (d->sigma_t[i] > m) m = d->sigma_t[i];
        /* scalar_mono_double: Float = double, grid data float32 */
        S->majorant = (double) d->sigma_t_scale * (double) m;
        if (d->n_phase < 1 || d->n_phase > ERTB_MAX_PHASE) return fail("invalid n_phase");
        for (int i = 0; i < d->n_phase; ++i) {
            const ertb_phase_desc *p = &d->phase[i];
            if (p->type == ERTB_PHASE_TABULATED || p->type == ERTB_PHASE_TABULATED_IRREGULAR ||
                p->type == ERTB_PHASE_TABULATED_POLARIZED) {
                const float *nodes = p->type != ERTB_PHASE_TABULATED ? p->nodes : NULL;
                if (p->type != ERTB_PHASE_TABULATED && !nodes) return fail("nodes missing");
                if (distr_init(&S->distr[i], nodes, p->values, p->n_nodes)) return 1;
                S->has_distr[i] = 1;
            }
        }
        if (d->integrator == ERTB_INTEGRATOR_PIECEWISE_VOLPATH) {
            /* medium.cpp:99-118: only the piecewise medium implements the *_real interface;
             * src/eradiate/experiments/_helpers.py:127-165: plane-parallel geometry only */
            if (d->homogeneous || S->spherical) return fail("sample_interaction_real is not implemented for this medium");
            if (piecewise_init(S)) return 1;
        }
    }
    S->emitter_d = vnormalize(V(d->emitter_direction[0], d->emitter_direction[1], d->emitter_direction[2]));
    if (d->bsdf_type == ERTB_BSDF_OCEAN_LEGACY)
        if (ocean_init(&S->ocean, d->bsdf_params)) return fail("ocean_legacy init failed");
    S->astro = d->emitter_angular_diameter > 0.0;
    if (S->astro) { /* astroobject.cpp:75-80 */
        if (!(d->emitter_angular_diameter < 180.0)) return fail("Invalid angular diameter specified! (must be in ]0, 180[)");
        S->astro_cos = cos(0.5 * d->emitter_angular_diameter * PI / 180.0);
        S->astro_omega = 2.0 * PI * (1.0 - S->astro_cos);
    }
    if (d->bsdf_type == ERTB_BSDF_MEASURED_MONO && (!d->bsdf_table || d->bsdf_table_res[0] < 32))
        return fail("measured_mono: missing table");
    if (d->bsdf_type == ERTB_BSDF_MQDIFFUSE &&
        (!d->bsdf_table || d->bsdf_table_res[0] < 1 || d->bsdf_table_res[1] < 1 || d->bsdf_table_res[2] < 1))
        return fail("mqdiffuse: missing table");
    if (is_glint_family(d->bsdf_type)) {
        if (d->bsdf_type == ERTB_BSDF_OCEAN_GRASP && d->bsdf_params[6] != 0.f)
            return fail("ocean_grasp: only component=0 (full BRDF) is supported");
        glint_init(&S->glint, d->bsdf_type, d->bsdf_params);
    }
    if (d->n_instances > 0) {
        if (S->spherical || d->polarized) return fail("canopies: plane-parallel, unpolarized scenes only");
        if (canopy_init(&S->canopy, d)) return fail("canopy init failed");
    }
    return 0;
}

/* ------------------------------------------------------------- geometry */
typedef struct { v3 o, d; double maxt; } ray_t;
enum { SHAPE_GROUND = 0, SHAPE_TOA = 1, SHAPE_LEAF = 2 };
typedef struct { double t; v3 p, n; int shape; int group; int trunk; /* t = INFINITY when invalid */
                 int mesh; v3 sh_n; double mesh_r, mesh_t; /* mesh triangle: shading normal and its bilambertian */ } si_t;

static inline v3 ray_at(const ray_t *r, double t) { return vfma(r->d, t, r->o); }

/* MI/src/shapes/sphere.cpp:519-580: plane-shifted quadratic in double */
static double sphere_intersect(const ray_t *ray, double radius) {
    v3 l = ray->o; /* center = 0 */
    v3 d = ray->d;
    double plane_t = vdot(vneg(l), d) / vnorm(d);
    v3 o = ray_at(ray, plane_t);
    double A = vdot(d, d), B = 2.0 * vdot(o, d), C = vdot(o, o) - radius * radius;
    /* math::solve_quadratic (MI/include/mitsuba/core/math.h) */
    double disc = B * B - 4.0 * A * C;
    if (disc < 0.0) return INFINITY;
    double temp = -0.5 * (B + copysign(sqrt(disc), B));
    double x0 = temp / A, x1 = C / temp;
    if (temp == 0.0) { x0 = x1 = 0.0; }
    double near_t = fmin(x0, x1) + plane_t, far_t = fmax(x0, x1) + plane_t;
    if (!(near_t <= ray->maxt && far_t >= 0.0)) return INFINITY;
    if (near_t < 0.0 && far_t > ray->maxt) return INFINITY;
    return near_t < 0.0 ? far_t : near_t;
}

/* Scene::ray_intersect over the two stencil surfaces Eradiate emits
 * (src/eradiate/scenes/geometry.py:191-266; ERP/shapes/arectangle.cpp; the slab
 * `cube` top face; horizontal extent treated as unbounded: default width 1e6 km) */
static si_t scene_intersect(const scene_t *S, const ray_t *ray) {
    const ertb_scene_desc *d = S->desc;
    si_t best; memset(&best, 0, sizeof best); best.t = INFINITY; best.shape = -1; best.group = -1;
    double tg = INFINITY, tt = INFINITY;
    if (S->spherical) {
        tg = sphere_intersect(ray, d->surface_z);
        if (d->has_medium) tt = sphere_intersect(ray, d->medium_top);
    } else {
        if (ray->d.z != 0.0) {
            double t = (d->surface_z - ray->o.z) / ray->d.z;
            if (t > 0.0 && t <= ray->maxt) tg = t;
            if (d->has_medium) {
                t = (d->medium_top - ray->o.z) / ray->d.z;
                if (t > 0.0 && t <= ray->maxt) tt = t;
            }
        }
    }
    if (tg < tt) { best.t = tg; best.shape = SHAPE_GROUND; }
    else if (tt < INFINITY) { best.t = tt; best.shape = SHAPE_TOA; }
    if (best.t < INFINITY) {
        best.p = ray_at(ray, best.t);
        if (S->spherical) {
            /* sphere.cpp:684-688: "Re-project onto the sphere to improve accuracy" */
            best.n = vnormalize(best.p);
            best.p = vmul(best.n, best.shape == SHAPE_GROUND ? d->surface_z : d->medium_top);
        } else {
            /* arectangle.cpp:536-540 / rectangle.cpp: "Re-project onto the rectangle" (the
             * slab's top face is a planar mesh face, same treatment).  Without it, distant-sensor
             * rays (t ~ 1e9 m) land ~1e-7 m off the plane, more than the spawn offset. */
            best.n = V(0, 0, 1);
            best.p.z = best.shape == SHAPE_GROUND ? d->surface_z : d->medium_top;
        }
    }
    if (S->canopy.n_instances > 0) { /* leaves: disks in instanced shape groups */
        double o[3] = { ray->o.x, ray->o.y, ray->o.z }, dd[3] = { ray->d.x, ray->d.y, ray->d.z };
        canopy_hit_t h = canopy_intersect(&S->canopy, o, dd, fmin(ray->maxt, best.t));
        if (h.t < best.t) {
            best.t = h.t; best.shape = SHAPE_LEAF; best.group = h.group; best.trunk = h.kind == CANOPY_TRUNK;
            best.p = V(h.p[0], h.p[1], h.p[2]);
            best.n = V(h.n[0], h.n[1], h.n[2]);
            best.mesh = h.kind == CANOPY_MESH;
            best.sh_n = V(h.sh_n[0], h.sh_n[1], h.sh_n[2]);
            best.mesh_r = h.mesh_r; best.mesh_t = h.mesh_t;
        }
    }
    return best;
}

/* MI/include/mitsuba/render/interaction.h:140-169 */
static v3 offset_p(v3 p, v3 n, v3 d) {
    double m = fmax(fabs(p.x), fmax(fabs(p.y), fabs(p.z)));
    double mag = (1.0 + m) * RAY_EPS;
    double dn = vdot(n, d);
    if (dn < 0.0) mag = -mag; /* mulsign */
    return vfma(n, mag, p);
}
static ray_t spawn_ray(v3 p, v3 n, v3 d) {
    ray_t r; r.o = offset_p(p, n, d); r.d = d; r.maxt = DBL_MAX; return r;
}
static ray_t spawn_ray_to(v3 p, v3 n, v3 t) {
    ray_t r;
    r.o = offset_p(p, n, vsub(t, p));
    v3 d = vsub(t, r.o);
    double dist = vnorm(d);
    r.d = vmul(d, 1.0 / dist);
    r.maxt = dist * (1.0 - SHADOW_EPS);
    return r;
}

/* --------------------------------------------------------------- medium */
/* Layer index at a world-space point: GridVolume nearest lookup
 * (MI/src/volumes/grid.cpp:548-562, MI/ext/drjit/include/drjit/texture.h:493-498,
 * clamp wrap mode) behind either the slab to_world (plane-parallel,
 * src/eradiate/scenes/geometry.py:197-210) or the radial remap of
 * ERP/volumes/sphericalcoords.cpp:102-123.  Returns -1 for fill regions. */
static int layer_index(const scene_t *S, v3 p) {
    const ertb_scene_desc *d = S->desc;
    double u;
    if (S->spherical) {
        /* to_local scales the TOA sphere to the unit sphere */
        double r = vnorm(p) / d->medium_top;
        double rmin = d->medium_bottom / d->medium_top;
        if (r < rmin || r > 1.0) return -1; /* fillmin / fillmax = 0 */
        u = (r - rmin) / (1.0 - rmin);
    } else {
        u = (p.z - d->medium_bottom) / (d->medium_top - d->medium_bottom);
    }
    int idx = (int) floor(u * d->n_layers);
    if (idx < 0) idx = 0;
    if (idx > d->n_layers - 1) idx = d->n_layers - 1;
    return idx;
}

/* BoundingBox::ray_intersect of the sigma_t volume's bbox
 * (heterogeneous.cpp:199; sphericalcoords.cpp update_bbox_sphere) */
static int medium_aabb(const scene_t *S, const ray_t *ray, double *mint, double *maxt) {
    const ertb_scene_desc *d = S->desc;
    if (d->homogeneous) { *mint = 0.0; *maxt = INFINITY; return 1; } /* homogeneous.cpp:176-181 */
    double lo[3], hi[3];
    if (S->spherical) {
        for (int i = 0; i < 3; ++i) { lo[i] = -d->medium_top; hi[i] = d->medium_top; }
    } else {
        lo[0] = lo[1] = -INFINITY; hi[0] = hi[1] = INFINITY;
        if (S->pp_half_width > 0.0) { lo[0] = lo[1] = -S->pp_half_width; hi[0] = hi[1] = S->pp_half_width; }
        lo[2] = d->medium_bottom; hi[2] = d->medium_top;
    }
    double o[3] = { ray->o.x, ray->o.y, ray->o.z }, dd[3] = { ray->d.x, ray->d.y, ray->d.z };
    double tn = -INFINITY, tf = INFINITY;
    int active = 1;
    for (int i = 0; i < 3; ++i) {
        if (!isfinite(lo[i])) continue;
        if (dd[i] == 0.0) {
            if (!(o[i] > lo[i] && o[i] < hi[i])) active = 0;
            continue;
        }
        double t1 = (lo[i] - o[i]) / dd[i], t2 = (hi[i] - o[i]) / dd[i];
        double a = fmin(t1, t2), b = fmax(t1, t2);
        if (a > tn) tn = a;
        if (b < tf) tf = b;
    }
    active = active && (tf >= tn);
    *mint = tn; *maxt = tf;
    return active;
}

typedef struct {
    double t, mint, sigma_s, sigma_n, sigma_t, combined;
    v3 p, wi;
    int layer;
} mei_t;

/* MI/src/render/medium.cpp:42-82 + heterogeneous.cpp:185-196 / homogeneous.cpp:157-170 */
static mei_t sample_interaction(const scene_t *S, const ray_t *ray, double sample) {
    const ertb_scene_desc *d = S->desc;
    mei_t mei; memset(&mei, 0, sizeof mei);
    mei.wi = vneg(ray->d);
    double mint, maxt;
    int active = medium_aabb(S, ray, &mint, &maxt);
    active = active && (isfinite(mint) || isfinite(maxt));
    if (!active) { mint = 0.0; maxt = INFINITY; }
    mint = fmax(0.0, mint);
    maxt = fmin(ray->maxt, maxt);
    double m = S->majorant;
    double sampled_t = mint + (-log(1.0 - sample) / m);
    int valid = active && (sampled_t <= maxt);
    mei.t = valid ? sampled_t : INFINITY;
    mei.p = ray_at(ray, sampled_t);
    mei.mint = mint;
    mei.combined = m;
    mei.layer = -1;
    if (valid) {
        int l = layer_index(S, mei.p);
        mei.layer = l;
        double st = l >= 0 ? (double) d->sigma_t_scale * (double) d->sigma_t[l] : 0.0;
        double al = l >= 0 ? (double) d->albedo[l] : 0.0;
        mei.sigma_t = st;
        mei.sigma_s = st * al;
        mei.sigma_n = m - st;
    }
    return mei;
}

/* ------------------------------------------------------ piecewise medium */
/* ERP/media/piecewise.cpp:445-507 precompute_optical_thickness: running sums of the layer
 * extinctions (not yet multiplied by the layer thickness), bottom-up and top-down. */
static int piecewise_init(scene_t *S) {
    const ertb_scene_desc *d = S->desc;
    int n = d->n_layers;
    S->pw_cum = (double *) malloc(sizeof(double) * (size_t) (n > 0 ? n : 1));
    S->pw_rcum = (double *) malloc(sizeof(double) * (size_t) (n > 0 ? n : 1));
    if (!S->pw_cum || !S->pw_rcum) return fail("out of memory");
    double c = 0.0;
    for (int i = 0; i < n; ++i) { c += (double) d->sigma_t_scale * (double) d->sigma_t[i]; S->pw_cum[i] = c; }
    c = 0.0;
    for (int i = n - 1; i >= 0; --i) { c += (double) d->sigma_t_scale * (double) d->sigma_t[i]; S->pw_rcum[n - 1 - i] = c; }
    return 0;
}

static double pw_sigma_t(const scene_t *S, int l) {
    return (double) S->desc->sigma_t_scale * (double) S->desc->sigma_t[l];
}

static int pw_cell(const scene_t *S, double z) {
    const ertb_scene_desc *d = S->desc;
    double voxel = (d->medium_top - d->medium_bottom) / d->n_layers;
    double f = floor((z - d->medium_bottom) / voxel);
    if (!(f > 0.0)) return 0; /* also NaN / -inf */
    return f > (double) (d->n_layers - 1) ? d->n_layers - 1 : (int) f;
}

/* drjit util.h:135-181 binary_search: first index of [start, end) where pred is false, else end */
static int pw_bsearch(const double *table, int start, int end, double a, double offset) {
    int iterations = start < end ? (31 - __builtin_clz((unsigned) (end - start))) + 1 : 0;
    for (int i = 0; i < iterations; ++i) {
        int middle = (start + end) >> 1;
        int cond = a > (table[middle] - offset);
        if (cond) start = middle + 1 < end ? middle + 1 : end;
        else end = middle;
    }
    return start;
}

/* piecewise.cpp:183-332 sample_interaction_real: analytic free-flight sampling through the
 * stack of homogeneous layers.  `aabb` = (active, mint, maxt) of the medium bbox. */
static mei_t pw_sample_interaction_real(const scene_t *S, const ray_t *ray, double si_t, double sample,
                                        int aabb_hit, double mint, double maxt, double *tr_out, double *pdf_out) {
    const ertb_scene_desc *d = S->desc;
    const int n = d->n_layers;
    const double voxel = (d->medium_top - d->medium_bottom) / n, inv_voxel = 1.0 / voxel;
    mei_t mei; memset(&mei, 0, sizeof mei);
    mei.wi = vneg(ray->d);
    mei.layer = -1;
    int active = aabb_hit && (isfinite(mint) || isfinite(maxt));
    if (!active) { mint = 0.0; maxt = INFINITY; }
    mint = fmax(0.0, mint);
    maxt = fmin(si_t, fmin(ray->maxt, maxt));
    int escaped = !active;
    mei.mint = mint;
    mei.t = mint;
    double cum_opt_thick = 0.0, sampled_t = INFINITY, tr = 0.0, pdf = 0.0;
    double n_dot_d = fabs(vnormalize(ray->d).z);
    double delta = n_dot_d == 0.0 ? INFINITY : voxel / n_dot_d;
    double idelta = 1.0 / delta;
    int going_up = ray->d.z >= 0.0;
    int start_idx = pw_cell(S, ray_at(ray, mint).z);
    int end_idx = pw_cell(S, ray_at(ray, maxt).z);
    int same_cell = start_idx == end_idx;
    int opt_start = going_up ? start_idx : n - 1 - start_idx;
    int opt_end = going_up ? end_idx : n - 1 - end_idx;
    int index = opt_start;
    double start_height = (ray_at(ray, mint + RAY_EPS).z - d->medium_bottom) * inv_voxel - (double) start_idx;
    if (going_up) start_height = 1.0 - start_height;
    double sigma_t = pw_sigma_t(S, pw_cell(S, ray_at(ray, mint).z));
    const double *table = going_up ? S->pw_cum : S->pw_rcum;
    double offset = active ? table[opt_start] : 0.0;
    offset -= start_height * sigma_t;
    double log_sample = log(1.0 - sample);
    int search = active && !same_cell;
    if (search) index = pw_bsearch(table, opt_start, opt_end, -log_sample * idelta, offset);
    same_cell |= index == opt_start;
    search = search && !same_cell;
    if (search) {
        mei.t += (start_height + (double) (index - opt_start - 1)) * delta;
        cum_opt_thick = (table[index - 1] - offset) * delta;
    }
    int cell = going_up ? index : n - 1 - index;
    if (!same_cell) sigma_t = pw_sigma_t(S, cell);
    escaped |= mei.t > maxt;
    int sampled = !escaped;
    if (sampled) sampled_t = -(1.0 / sigma_t) * (log_sample + cum_opt_thick) + mei.t;
    escaped |= sampled && sampled_t > maxt;
    sampled |= escaped;
    if (sampled) {
        if (escaped) sampled_t = maxt;
        tr = exp(-(sampled_t - mei.t) * sigma_t - cum_opt_thick);
        pdf = sampled_t == maxt ? tr : tr * sigma_t;
    }
    mei.t = !escaped ? sampled_t : INFINITY;
    if (!escaped) {
        mei.p = ray_at(ray, mei.t);
        mei.layer = cell;
        mei.sigma_t = sigma_t;
        mei.sigma_s = sigma_t * (double) d->albedo[cell];
        mei.sigma_n = 0.0;
        mei.combined = sigma_t;
    }
    *tr_out = tr; *pdf_out = pdf;
    return mei;
}

/* piecewise.cpp:335-429 eval_transmittance_pdf_real: exact transmittance of the ray segment
 * clipped to the medium bbox.  Returns the `escaped` mask. */
static int pw_eval_transmittance_pdf_real(const scene_t *S, const ray_t *ray, double si_t, int aabb_hit,
                                          double mint, double maxt, double *tr_out, double *pdf_out) {
    const ertb_scene_desc *d = S->desc;
    const int n = d->n_layers;
    const double voxel = (d->medium_top - d->medium_bottom) / n, inv_voxel = 1.0 / voxel;
    int active = aabb_hit && (isfinite(mint) || isfinite(maxt));
    mint = fmax(0.0, mint);
    int escaped = active && ((maxt >= ray->maxt) || (maxt >= si_t));
    maxt = active ? fmin(ray->maxt, fmin(maxt, si_t)) : INFINITY;
    maxt = fmax(0.0, maxt);
    double n_dot_d = fabs(vnormalize(ray->d).z);
    double delta = n_dot_d == 0.0 ? INFINITY : voxel / n_dot_d;
    int going_up = ray->d.z >= 0.0;
    double zs = ray_at(ray, mint).z, ze = ray_at(ray, maxt).z;
    int start_idx = pw_cell(S, zs), end_idx = pw_cell(S, ze);
    int same_cell = start_idx == end_idx;
    double start_height = (zs - d->medium_bottom) * inv_voxel - (double) start_idx;
    double end_height = (ze - d->medium_bottom) * inv_voxel - (double) end_idx;
    if (going_up) start_height = 1.0 - start_height;
    else end_height = 1.0 - end_height;
    double s_sigma_t = pw_sigma_t(S, start_idx), e_sigma_t = pw_sigma_t(S, end_idx);
    int hi = start_idx > end_idx ? start_idx : end_idx, lo = start_idx < end_idx ? start_idx : end_idx;
    int max_idx = active ? (hi - 1 > 0 ? hi - 1 : 0) : 0;
    int min_idx = active ? (lo > 0 ? lo : 0) : 0;
    int use_precomputed = active && max_idx > min_idx;
    double cum = use_precomputed ? S->pw_cum[max_idx] - S->pw_cum[min_idx] : 0.0;
    cum += s_sigma_t * start_height + e_sigma_t * end_height;
    cum *= delta;
    double opt_thick = active ? (same_cell ? (maxt - mint) * s_sigma_t : cum) : 0.0;
    double tr = active ? exp(-opt_thick) : 0.0;
    double pdf = active ? ((si_t < maxt || ray->maxt < maxt) ? tr : tr * e_sigma_t) : 0.0;
    *tr_out = tr; *pdf_out = pdf;
    return escaped;
}

/* ---------------------------------------------------------------- phase */
static double eval_rayleigh(double c, double rho) { /* rayleigh.cpp:61-67 */
    double r1 = (1.0 - rho) / (1.0 + rho / 2.0), r2 = (1.0 + rho) / (1.0 - rho);
    return (3.0 / 16.0) * INV_PI * r1 * (r2 + c * c);
}
static double eval_rayleigh_pdf(double c) { return (3.0 / 16.0) * INV_PI * (1.0 + c * c); }
static double eval_hg(double g, double c) { /* hg.cpp:64-68 */
    double temp = 1.0 + g * g + 2.0 * g * c;
    return INV_FOUR_PI * (1.0 - g * g) / (temp * sqrt(temp));
}

/* eval_pdf of one leaf; `c` = dot(wo, wi) with wi = -ray.d (graphics convention) */
static void leaf_eval_pdf(const scene_t *S, int leaf, double c, double *val, double *pdf) {
    const ertb_phase_desc *p = &S->desc->phase[leaf];
    switch (p->type) {
        case ERTB_PHASE_ISOTROPIC: /* isotropic.cpp:39-60 */
            *val = *pdf = INV_FOUR_PI; break;
        case ERTB_PHASE_RAYLEIGH:  /* rayleigh.cpp:97-107 */
        case ERTB_PHASE_RAYLEIGH_POLARIZED: /* rayleigh_polarized.cpp:122-127 (unpolarized branch) */
            *val = eval_rayleigh(c, p->params[0]); *pdf = eval_rayleigh_pdf(c); break;
        case ERTB_PHASE_HG:        /* hg.cpp:92-99 */
            *val = *pdf = eval_hg(p->params[0], c); break;
        default: {                 /* tabphase.cpp:107-118: data in physics convention */
            double v = distr_eval_pdf(&S->distr[leaf], -c) * S->distr[leaf].normalization * INV_TWO_PI;
            *val = *pdf = v; break;
        }
    }
}

/* sample of one leaf: returns local direction in the frame of wi (z = wi). */
static void leaf_sample(const scene_t *S, int leaf, double u1, double u2, v3 *wo_local,
                        double *weight, double *pdf) {
    const ertb_phase_desc *p = &S->desc->phase[leaf];
    double sphi = sin(2.0 * PI * u2), cphi = cos(2.0 * PI * u2);
    switch (p->type) {
        case ERTB_PHASE_ISOTROPIC: { /* warp::square_to_uniform_sphere */
            double z = 1.0 - 2.0 * u2;
            double r = safe_sqrt(1.0 - z * z);
            double s2 = sin(2.0 * PI * u1), c2 = cos(2.0 * PI * u1);
            *wo_local = V(r * c2, r * s2, z);
            *weight = 1.0; *pdf = INV_FOUR_PI;
            break;
        }
        case ERTB_PHASE_RAYLEIGH_POLARIZED: /* rayleigh_polarized.cpp:133-160: same inversion */
        case ERTB_PHASE_RAYLEIGH: { /* rayleigh.cpp:75-95 */
            double z = 2.0 * (2.0 * u1 - 1.0);
            double tmp = sqrt(z * z + 1.0);
            double A = cbrt(z + tmp), B = cbrt(z - tmp);
            double ct = A + B, st = safe_sqrt(1.0 - ct * ct);
            *wo_local = V(st * cphi, st * sphi, ct);
            *pdf = eval_rayleigh_pdf(-ct);
            *weight = eval_rayleigh(-ct, p->params[0]) / *pdf;
            break;
        }
        case ERTB_PHASE_HG: { /* hg.cpp:70-90 */
            double g = p->params[0];
            double sq = (1.0 - g * g) / (1.0 - g + 2.0 * g * u1);
            double ct = (1.0 + g * g - sq * sq) / (2.0 * g);
            if (fabs(g) < DBL_EPSILON * 0.5) ct = 1.0 - 2.0 * u1;
            double st = safe_sqrt(1.0 - ct * ct);
            *wo_local = V(st * cphi, st * sphi, -ct);
            *weight = 1.0; *pdf = eval_hg(g, -ct);
            break;
        }
        default: { /* tabphase.cpp:77-105: wo = -to_world(physics-convention dir) */
            const distr_t *D = &S->distr[leaf];
            double ctp = distr_sample(D, u1), stp = safe_sqrt(1.0 - ctp * ctp);
            *wo_local = V(-stp * cphi, -stp * sphi, -ctp);
            *pdf = distr_eval_pdf(D, ctp) * D->normalization * INV_TWO_PI;
            *weight = 1.0;
            break;
        }
    }
}

static double leaf_prob(const scene_t *S, int leaf, int layer) {
    const ertb_scene_desc *d = S->desc;
    if (d->n_phase == 1 || !d->phase_weight) return leaf == 0 ? 1.0 : 0.0;
    if (layer < 0) return leaf == 0 ? 1.0 : 0.0;
    return (double) d->phase_weight[(size_t) leaf * d->n_layers + layer];
}

/* blendphase.cpp:172-190 (nested lerp == sum of leaf probabilities) */
static void phase_eval_pdf(const scene_t *S, int layer, v3 wi, v3 wo, double *val, double *pdf) {
    double c = vdot(wo, wi), v = 0.0, p = 0.0;
    for (int i = 0; i < S->desc->n_phase; ++i) {
        double w = leaf_prob(S, i, layer);
        if (w == 0.0) continue;
        double vi, pi_;
        leaf_eval_pdf(S, i, c, &vi, &pi_);
        v += w * vi; p += w * pi_;
    }
    *val = v; *pdf = p;
}

/* blendphase.cpp:100-141: pick a component with sample1 (rescaled and forwarded
 * as the nested sample1, unused by every leaf), return the component's own
 * weight and pdf (no mixture MIS). */
static void phase_sample(const scene_t *S, int layer, v3 wi, double s1, double u1, double u2,
                         v3 *wo, double *weight, double *pdf) {
    int n = S->desc->n_phase, leaf = 0;
    if (n > 1) {
        double acc = 0.0;
        leaf = n - 1;
        for (int i = 0; i < n; ++i) {
            acc += leaf_prob(S, i, layer);
            if (s1 < acc) { leaf = i; break; }
        }
    }
    v3 wl;
    leaf_sample(S, leaf, u1, u2, &wl, weight, pdf);
    frame_t f = make_frame(wi); /* medium.cpp:51 sh_frame = Frame3f(wi) */
    *wo = to_world(&f, wl);
    if (S->desc->phase_mis && n > 1) {
        /* multiphase.cpp:176-200: value and pdf of every component at the sampled direction;
         * weight = sum_j w_j value_j / sum_j w_j pdf_j, pdf = the mixture pdf */
        double v, p;
        phase_eval_pdf(S, layer, wi, *wo, &v, &p);
        *weight = p > 1e-8 ? v / p : 0.0;
        *pdf = p;
    }
}


/* ------------------------------------------------- polarization (Mueller) */
/* MI/include/mitsuba/render/mueller.h; matrices are row-major double[16]. */
typedef struct { double m[16]; } mueller_t;
static mueller_t mu_zero(void) { mueller_t r; memset(&r, 0, sizeof r); return r; }
static mueller_t mu_identity(double v) { mueller_t r = mu_zero(); r.m[0] = r.m[5] = r.m[10] = r.m[15] = v; return r; }
static mueller_t mu_mul(const mueller_t *a, const mueller_t *b) {
    mueller_t r = mu_zero();
    for (int i = 0; i < 4; ++i) for (int k = 0; k < 4; ++k) { double aik = a->m[4 * i + k];
        if (aik != 0.0) for (int j = 0; j < 4; ++j) r.m[4 * i + j] += aik * b->m[4 * k + j]; }
    return r;
}
static mueller_t mu_scale(const mueller_t *a, double s) { mueller_t r; for (int i = 0; i < 16; ++i) r.m[i] = a->m[i] * s; return r; }
static mueller_t mu_transpose(const mueller_t *a) { mueller_t r; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r.m[4 * i + j] = a->m[4 * j + i]; return r; }
static mueller_t mu_rotator(double theta) { /* mueller.h:164-173 */
    double s = sin(2.0 * theta), c = cos(2.0 * theta);
    mueller_t r = mu_zero();
    r.m[0] = 1; r.m[5] = c; r.m[6] = s; r.m[9] = -s; r.m[10] = c; r.m[15] = 1;
    return r;
}
static v3 vcross(v3 a, v3 b) { return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static v3 stokes_basis(v3 forward) { v3 s, t; coordinate_system(forward, &s, &t); return s; } /* mueller.h:286 */
/* drjit/sphere.h:53-62 unit_angle */
static double unit_angle(v3 a, v3 b) {
    double d = vdot(a, b);
    v3 diff = d >= 0.0 ? vsub(b, a) : vadd(b, a);
    double temp = 2.0 * asin(fmin(1.0, 0.5 * vnorm(diff)));
    return d >= 0.0 ? temp : PI - temp;
}
/* mueller.h:316-324 */
static mueller_t rotate_stokes_basis(v3 forward, v3 basis_current, v3 basis_target) {
    double theta = unit_angle(vnormalize(basis_current), vnormalize(basis_target));
    if (vdot(forward, vcross(basis_current, basis_target)) < 0.0) theta = -theta;
    return mu_rotator(theta);
}
/* mueller.h:362-372: R_out * M * R_in^T */
static mueller_t rotate_mueller_basis(const mueller_t *M, v3 in_fwd, v3 in_cur, v3 in_tgt, v3 out_fwd, v3 out_cur, v3 out_tgt) {
    mueller_t Rin = rotate_stokes_basis(in_fwd, in_cur, in_tgt), Rout = rotate_stokes_basis(out_fwd, out_cur, out_tgt);
    mueller_t RinT = mu_transpose(&Rin), t = mu_mul(M, &RinT);
    return mu_mul(&Rout, &t);
}
static int mu_has_nan(const mueller_t *a) { for (int i = 0; i < 16; ++i) if (isnan(a->m[i])) return 1; return 0; }

/* Mueller-valued eval_pdf of one leaf in world space (Radiance mode: light arrives along -wo and
 * leaves along +wi).  rayleigh_polarized.cpp:55-127, tabphase_polarized.cpp:318-368; the scalar
 * plugins return Spectrum(value) = value * Identity in a polarized variant. */
static void leaf_eval_mueller(const scene_t *S, int leaf, v3 wi, v3 wo, mueller_t *M, double *pdf) {
    const ertb_phase_desc *p = &S->desc->phase[leaf];
    double ct = -vdot(wo, wi); /* physics convention */
    int polarized_leaf = 0;
    if (p->type == ERTB_PHASE_RAYLEIGH_POLARIZED) {
        double rho = p->params[0];
        double r1 = (1.0 - rho) / (1.0 + rho / 2.0), r2 = (1.0 + rho) / (1.0 - rho), r3 = (1.0 - 2.0 * rho) / (1.0 - rho);
        double a = r2 + ct * ct, b = ct * ct + 1.0, c = ct * ct - 1.0, d = 2.0 * ct;
        double k = (3.0 / 16.0) * INV_PI * r1;
        *M = mu_zero();
        M->m[0] = k * a; M->m[1] = k * c; M->m[4] = k * c; M->m[5] = k * b; M->m[10] = k * d; M->m[15] = k * d * r3;
        *pdf = eval_rayleigh_pdf(ct);
        polarized_leaf = 1;
    } else if (p->type == ERTB_PHASE_TABULATED_POLARIZED) {
        const distr_t *D = &S->distr[leaf];
        double m11 = distr_eval_pdf(D, ct), norm = D->normalization * INV_TWO_PI;
        double ms[5] = { 0, 0, 0, 0, 0 };
        if (ct >= D->x0 && ct <= D->x1) { /* IrregularInterpolant::eval_data (tabphase_polarized.cpp:178-206) */
            int index = bsearch_pred_nodes(D, ct);
            if (index > D->n - 1) index = D->n - 1;
            if (index < 1) index = 1;
            index -= 1;
            double t = (ct - D->nodes[index]) / (D->nodes[index + 1] - D->nodes[index]);
            for (int k = 0; k < 5; ++k)
                if (p->mueller[k]) ms[k] = (double) p->mueller[k][index] + t * ((double) p->mueller[k][index + 1] - (double) p->mueller[k][index]);
        }
        *M = mu_zero();
        M->m[0] = m11; M->m[1] = ms[0]; M->m[4] = ms[0]; M->m[5] = ms[1];
        M->m[10] = ms[2]; M->m[11] = ms[3]; M->m[14] = -ms[3]; M->m[15] = ms[4];
        *M = mu_scale(M, norm);
        *pdf = m11 * norm;
        polarized_leaf = 1;
    } else {
        double v, pp;
        leaf_eval_pdf(S, leaf, vdot(wo, wi), &v, &pp);
        *M = mu_identity(v);
        *pdf = pp;
    }
    if (polarized_leaf) {
        v3 wo_hat = wo, wi_hat = wi;
        v3 x_hat = vnormalize(vcross(vneg(wo_hat), wi_hat));
        v3 p_in = vnormalize(vcross(x_hat, vneg(wo_hat))), p_out = vnormalize(vcross(x_hat, wi_hat));
        *M = rotate_mueller_basis(M, vneg(wo_hat), p_in, stokes_basis(vneg(wo_hat)), wi_hat, p_out, stokes_basis(wi_hat));
        if (mu_has_nan(M)) *M = mu_zero();
    }
}

/* blendphase.cpp:172-190 with Mueller-valued components */
static void phase_eval_mueller_pdf(const scene_t *S, int layer, v3 wi, v3 wo, mueller_t *M, double *pdf) {
    *pdf = 0.0;
    *M = mu_zero();
    for (int i = 0; i < S->desc->n_phase; ++i) {
        double w = leaf_prob(S, i, layer);
        if (w == 0.0) continue;
        mueller_t Mi; double pi_;
        leaf_eval_mueller(S, i, wi, wo, &Mi, &pi_);
        for (int k = 0; k < 16; ++k) M->m[k] += w * Mi.m[k];
        *pdf += w * pi_;
    }
}
/* sample(): direction from the leaf's scalar sampler, weight = Mueller value / pdf
 * (rayleigh_polarized.cpp:133-160, tabphase_polarized.cpp:296-316) */
static void phase_sample_mueller(const scene_t *S, int layer, v3 wi, double s1, double u1, double u2,
                                 v3 *wo, mueller_t *W, double *pdf) {
    int n = S->desc->n_phase, leaf = 0;
    if (n > 1) {
        double acc = 0.0;
        leaf = n - 1;
        for (int i = 0; i < n; ++i) { acc += leaf_prob(S, i, layer); if (s1 < acc) { leaf = i; break; } }
    }
    v3 wl; double w_scalar, p_scalar;
    const ertb_phase_desc *p = &S->desc->phase[leaf];
    if (p->type == ERTB_PHASE_TABULATED_POLARIZED) { /* m11 drives the sampling (:300-309) */
        const distr_t *D = &S->distr[leaf];
        double ctp = distr_sample(D, u1), stp = safe_sqrt(1.0 - ctp * ctp);
        wl = V(-stp * cos(2.0 * PI * u2), -stp * sin(2.0 * PI * u2), -ctp);
        w_scalar = 1.0; p_scalar = 0.0;
    } else {
        leaf_sample(S, leaf, u1, u2, &wl, &w_scalar, &p_scalar);
    }
    frame_t f = make_frame(wi);
    *wo = to_world(&f, wl);
    if (p->type == ERTB_PHASE_RAYLEIGH_POLARIZED || p->type == ERTB_PHASE_TABULATED_POLARIZED) {
        mueller_t M; double pp;
        leaf_eval_mueller(S, leaf, wi, *wo, &M, &pp);
        if (p->type == ERTB_PHASE_RAYLEIGH_POLARIZED) pp = p_scalar; /* pdf from the sampled cosine */
        *pdf = pp;
        *W = pp > 0.0 ? mu_scale(&M, 1.0 / pp) : mu_zero();
    } else {
        *pdf = p_scalar;
        *W = mu_identity(w_scalar);
    }
    if (S->desc->phase_mis && n > 1) { /* multiphase.cpp:176-200 with Spectrum = Mueller matrix */
        mueller_t M; double pm;
        phase_eval_mueller_pdf(S, layer, wi, *wo, &M, &pm);
        *W = pm > 1e-8 ? mu_scale(&M, 1.0 / pm) : mu_zero();
        *pdf = pm;
    }
}

/* ----------------------------------------------------------------- BSDFs */
static double cos_theta(v3 v) { return v.z; }
static double sin_theta(v3 v) { return safe_sqrt(1.0 - v.z * v.z); }
static double tan_theta(v3 v) { return safe_sqrt(1.0 - v.z * v.z) / v.z; }
/* Frame3f::sincos_phi (MI/include/mitsuba/core/frame.h) */
static void sincos_phi(v3 v, double *s, double *c) {
    double st2 = 1.0 - v.z * v.z;
    if (st2 <= 0.0) { *s = 0.0; *c = 1.0; return; }
    double inv = 1.0 / sqrt(st2);
    *s = v.y * inv; *c = v.x * inv;
    if (*s > 1.0) *s = 1.0; if (*s < -1.0) *s = -1.0;
    if (*c > 1.0) *c = 1.0; if (*c < -1.0) *c = -1.0;
}

/* ERP/bsdfs/rpv.cpp:128-167 */
static double eval_rpv(const float *P, v3 wi, v3 wo) {
    double rho_0 = P[0], k = P[1], g = P[2], rho_c = P[3];
    double spi, cpi, spo, cpo;
    sincos_phi(wi, &spi, &cpi);
    sincos_phi(wo, &spo, &cpo);
    double cdphi = cpi * cpo + spi * spo;
    double sti = sin_theta(wi), cti = cos_theta(wi), tti = tan_theta(wi);
    double sto = sin_theta(wo), cto = cos_theta(wo), tto = tan_theta(wo);
    double cT = cti * cto + sti * sto * cdphi;
    double F = (1.0 - g * g) / pow(1.0 + g * g + 2.0 * g * cT, 1.5);
    double G = safe_sqrt(tti * tti + tto * tto - 2.0 * tti * tto * cdphi);
    double H = 1.0 + (1.0 - rho_c) / (1.0 + G);
    double M = pow(cti * cto * (cti + cto), k - 1.0);
    return rho_0 * M * F * H * INV_PI;
}

/* ERP/bsdfs/rtls.cpp:116-243 */
static double eval_rtls(const float *P, v3 wi, v3 wo) {
    double f_iso = P[0], f_vol = P[1], f_geo = P[2], h = P[3], r = P[4], b = P[5];
    double spi, cpi, spo, cpo;
    sincos_phi(wi, &spi, &cpi);
    sincos_phi(wo, &spo, &cpo);
    double sti = sin_theta(wi), cti = cos_theta(wi), tti = tan_theta(wi);
    double sto = sin_theta(wo), cto = cos_theta(wo), tto = tan_theta(wo);
    double cdphi = cpi * cpo + spi * spo, sdphi = spi * cpo - cpi * spo;
    double cpsi = cti * cto + sti * sto * cdphi;
    double spsi = sqrt(1.0 - cpsi * cpsi), psi = acos(cpsi);
    double K_vol = ((PI / 2.0 - psi) * cpsi + spsi) / (cti + cto) - PI / 4.0;
    double ci = cti, co = cto, ti = tti, to = tto, cp = cpsi;
    if (fabs(r - b) > FLT_EPSILON * 0.5) { /* dr::Epsilon<ScalarFloat>; rtls.cpp:197-214 */
        ti = b / r * tti; to = b / r * tto;
        double thi = atan(ti), tho = atan(to);
        ci = cos(thi); co = cos(tho);
        cp = ci * co + sin(thi) * sin(tho) * cdphi;
    }
    double sec_i = 1.0 / ci, sec_o = 1.0 / co, sec_sum = sec_i + sec_o;
    double D = sqrt(ti * ti + to * to - 2.0 * ti * to * cdphi);
    double tsp = ti * to * sdphi;
    double cos_t = (h / b) * sqrt(D * D + tsp * tsp) / sec_sum;
    cos_t = fmax(fmin(cos_t, 1.0), -1.0);
    double t = acos(cos_t), sin_t = sin(t);
    double O = INV_PI * (t - sin_t * cos_t) * sec_sum;
    double K_geo = O - sec_sum + 0.5 * (1.0 + cp) * sec_i * sec_o;
    return (f_iso + f_vol * K_vol + f_geo * K_geo) * INV_PI;
}

/* ERP/bsdfs/hapke.cpp:120-332 */
static double hapke_H(double w, double x) {
    double gamma = sqrt(1.0 - w), ro = (1.0 - gamma) / (1.0 + gamma);
    return 1.0 / (1.0 - w * x * (ro + (1.0 - 2.0 * ro * x) * 0.5 * log((1.0 + x) / x)));
}
static double hapke_E1(double tt, double x) { return exp(-2.0 * INV_PI / tt / tan(x)); }
static double hapke_E2(double tt, double x) { return exp(-INV_PI / (tt * tt) / sqr(tan(x))); }
static double hapke_mu(double tt, double e, double i, double cos_x, double sin_x, double phi,
                       double opt_cos_phi, double sign) {
    double chi = 1.0 / sqrt(1.0 + PI * tt * tt);
    double E1e = hapke_E1(tt, e), E1i = hapke_E1(tt, i), E2e = hapke_E2(tt, e), E2i = hapke_E2(tt, i);
    double s2 = sin(phi * 0.5);
    return chi * (cos_x + sin_x * tt * (opt_cos_phi * E2e + sign * s2 * s2 * E2i) /
                              (2.0 - E1e - phi * INV_PI * E1i));
}
static double eval_hapke(const float *P, v3 wi, v3 wo) {
    double w = P[0], b = P[1], c = P[2], theta = P[3] * PI / 180.0, B0 = P[4], h = P[5];
    double tt = tan(theta);
    double spe, cpe, spi, cpi;
    sincos_phi(wo, &spe, &cpe);
    sincos_phi(wi, &spi, &cpi);
    double cos_phi = cpe * cpi + spe * spi;
    double sin_e = sin_theta(wo), mu = cos_theta(wo), tan_e = tan_theta(wo);
    double sin_i = sin_theta(wi), mu_0 = cos_theta(wi), tan_i = tan_theta(wi);
    double i = atan(tan_i), e = atan(tan_e);
    double fr_phi = safe_acos(cos_phi);
    double phi = fabs(fr_phi > PI ? 2.0 * PI - fr_phi : fr_phi);
    /* eval_mu_0eG / eval_mu_eG (:207-235) */
    double a_ = e <= i ? i : e, b_ = e <= i ? e : i;
    double mu_0eG = hapke_mu(tt, a_, b_, cos(i), sin(i), phi, e <= i ? 1.0 : cos_phi, e <= i ? -1.0 : 1.0);
    double mu_eG = hapke_mu(tt, a_, b_, cos(e), sin(e), phi, e <= i ? cos_phi : 1.0, e <= i ? 1.0 : -1.0);
    double mu_ratio = mu_0eG / (mu_0eG + mu_eG) / mu_0;
    double cos_g = mu_0 * mu + sin_i * sin_e * cos_phi;
    double g = safe_acos(cos_g);
    double num = 1.0 - b * b;
    double Pf = (1.0 - c) * num / pow(1.0 + 2.0 * b * cos_g + b * b, 1.5) +
                c * num / pow(1.0 - 2.0 * b * cos_g + b * b, 1.5);
    double B = B0 / (1.0 + 1.0 / h * tan(g / 2.0));
    double M = hapke_H(w, mu_0eG) * hapke_H(w, mu_eG) - 1.0;
    /* eval_f (:134-138): clip uses dr::Epsilon<Float> of the variant */
    double half = phi / 2.0, lim = PI / 2.0 - DBL_EPSILON * 0.5;
    if (half < 0.0) half = 0.0; if (half > lim) half = lim;
    double f = exp(-2.0 * tan(half));
    double chi = 1.0 / sqrt(1.0 + PI * tt * tt);
    double E1e = hapke_E1(tt, e), E1i = hapke_E1(tt, i), E2e = hapke_E2(tt, e), E2i = hapke_E2(tt, i);
    double eta_0e = chi * (mu_0 + sin_i * tt * E2i / (2.0 - E1i));
    double eta_e = chi * (mu + sin_e * tt * E2e / (2.0 - E1e));
    double opt_mu = e < i ? mu : mu_0, opt_eta = e < i ? eta_e : eta_0e;
    double Sf = (mu_eG * mu_0 * chi) / (eta_e * eta_0e * (1.0 - f + f * chi * opt_mu / opt_eta));
    return w * 0.25 * INV_PI * mu_ratio * (Pf * (1.0 + B) + M) * Sf;
}

/* mqdiffuse.cpp:94-107 eval_texture: p = (cos_theta_o, phi_d / 2 pi, cos_theta_i) remapped by (1 - 1/res) + 0.5/res,
 * then the Dr.Jit linear filter with clamp wrap mode (drjit/texture.h: pos = p * res - 0.5, both neighbours of
 * every axis clamped to [0, res - 1]).  data[z][y][x]. */
static double mq_texture(const ertb_scene_desc *d, double cos_o, double phi_d, double cos_i) {
    const int res[3] = { d->bsdf_table_res[0], d->bsdf_table_res[1], d->bsdf_table_res[2] };
    double p[3] = { cos_o, phi_d / (2.0 * PI), cos_i }, w[3];
    int i0[3], i1[3];
    for (int a = 0; a < 3; ++a) {
        double ps = 1.0 / res[a];
        double q = p[a] * (1.0 - ps) + 0.5 * ps;
        double pos = q * res[a] - 0.5, fl = floor(pos);
        w[a] = pos - fl;
        int i = (int) fl;
        i0[a] = i < 0 ? 0 : (i > res[a] - 1 ? res[a] - 1 : i);
        i1[a] = i + 1 < 0 ? 0 : (i + 1 > res[a] - 1 ? res[a] - 1 : i + 1);
    }
#define MQ_AT(z, y, x) ((double) d->bsdf_table[((size_t) (z) * res[1] + (y)) * res[0] + (x)])
    double c00 = MQ_AT(i0[2], i0[1], i0[0]) * (1 - w[0]) + MQ_AT(i0[2], i0[1], i1[0]) * w[0];
    double c01 = MQ_AT(i0[2], i1[1], i0[0]) * (1 - w[0]) + MQ_AT(i0[2], i1[1], i1[0]) * w[0];
    double c10 = MQ_AT(i1[2], i0[1], i0[0]) * (1 - w[0]) + MQ_AT(i1[2], i0[1], i1[0]) * w[0];
    double c11 = MQ_AT(i1[2], i1[1], i0[0]) * (1 - w[0]) + MQ_AT(i1[2], i1[1], i1[0]) * w[0];
#undef MQ_AT
    double c0 = c00 * (1 - w[1]) + c01 * w[1], c1 = c10 * (1 - w[1]) + c11 * w[1];
    return c0 * (1 - w[2]) + c1 * w[2];
}

/* BSDF::eval (value * cos_theta_o), local frame */
static double bsdf_eval_tp(const scene_t *S, int type, const float *P, v3 wi, v3 wo) {
    double cti = wi.z, cto = wo.z;
    if (!(cti > 0.0 && cto > 0.0)) return 0.0;
    switch (type) {
        case ERTB_BSDF_DIFFUSE: return (double) P[0] * INV_PI * cto;  /* diffuse.cpp:127-143 */
        case ERTB_BSDF_RPV: return eval_rpv(P, wi, wo) * fabs(cto);   /* rpv.cpp:169-181 */
        case ERTB_BSDF_RTLS: return eval_rtls(P, wi, wo) * fabs(cto); /* rtls.cpp:245-257 */
        case ERTB_BSDF_HAPKE: return eval_hapke(P, wi, wo) * fabs(cto);
        case ERTB_BSDF_OCEAN_LEGACY: return ocean_eval(&S->ocean, wi.x, wi.y, wi.z, wo.x, wo.y, wo.z);
        case ERTB_BSDF_OCEAN_MISHCHENKO: case ERTB_BSDF_OCEAN_GRASP: case ERTB_BSDF_MAIGNAN: {
            double a[3] = { wi.x, wi.y, wi.z }, b[3] = { wo.x, wo.y, wo.z };
            return glint_eval(&S->glint, a, b);
        }
        case ERTB_BSDF_MQDIFFUSE: { /* mqdiffuse.cpp:139-161 */
            double phi_d = fmod(atan2(wo.y, wo.x) - atan2(wi.y, wi.x), 2.0 * PI);
            if (phi_d < 0.0) phi_d += 2.0 * PI;
            return mq_texture(S->desc, cto, phi_d, cti) * cto;
        }
        case ERTB_BSDF_MEASURED_MONO: { /* measured_mono.cpp:339-393: the tables already hold f * cos */
            double a[3] = { wi.x, wi.y, wi.z }, b[3] = { wo.x, wo.y, wo.z };
            return mm_oracle_eval(S->desc->bsdf_table, a, b);
        }
        default: return 0.0;
    }
}
static double bsdf_eval(const scene_t *S, v3 wi, v3 wo) {
    return bsdf_eval_tp(S, S->desc->bsdf_type, S->desc->bsdf_params, wi, wo);
}
/* BSDF::sample -> wo (local), weight = eval / pdf */
static double bsdf_sample_tp(const scene_t *S, int type, const float *P, v3 wi, double s1, double u1, double u2, v3 *wo) {
    *wo = V(0, 0, 1);
    if (!(wi.z > 0.0)) return 0.0;
    if (type == ERTB_BSDF_OCEAN_LEGACY) {
        double o[3], w = ocean_sample(&S->ocean, wi.x, wi.y, wi.z, s1, u1, u2, o);
        *wo = V(o[0], o[1], o[2]);
        return w;
    }
    if (is_glint_family(type)) {
        double a[3] = { wi.x, wi.y, wi.z }, o[3], w = glint_sample(&S->glint, a, s1, u1, u2, o);
        *wo = V(o[0], o[1], o[2]);
        return w;
    }
    if (type == ERTB_BSDF_MEASURED_MONO) { /* measured_mono.cpp:234-337 */
        double a[3] = { wi.x, wi.y, wi.z }, o[3], w = mm_oracle_sample(S->desc->bsdf_table, a, u1, u2, o, NULL);
        *wo = V(o[0], o[1], o[2]);
        return w;
    }
    double o[3];
    ertbo_square_to_cosine_hemisphere(u1, u2, o); /* rpv.cpp:113 etc. */
    *wo = V(o[0], o[1], o[2]);
    double pdf = INV_PI * o[2];
    if (!(pdf > 0.0)) return 0.0;
    switch (type) {
        case ERTB_BSDF_DIFFUSE: return (double) P[0];               /* diffuse.cpp:121-123 */
        case ERTB_BSDF_RPV: return eval_rpv(P, wi, *wo) * o[2] / pdf; /* rpv.cpp:119-122 */
        case ERTB_BSDF_RTLS: return eval_rtls(P, wi, *wo) * o[2] / pdf;
        case ERTB_BSDF_HAPKE: return eval_hapke(P, wi, *wo) * o[2] / pdf;
        case ERTB_BSDF_MQDIFFUSE: { /* mqdiffuse.cpp:109-137: a negative difference is NOT wrapped here */
            double phi_d = fmod(atan2(o[1], o[0]) - atan2(wi.y, wi.x), 2.0 * PI);
            return mq_texture(S->desc, o[2], phi_d, wi.z) * o[2] / pdf;
        }
        default: return 0.0;
    }
}
static double bsdf_sample(const scene_t *S, v3 wi, double s1, double u1, double u2, v3 *wo) {
    return bsdf_sample_tp(S, S->desc->bsdf_type, S->desc->bsdf_params, wi, s1, u1, u2, wo);
}
/* CentralPatchSurface: blendbsdf.cpp:108-165 with the 0/1 central-patch mask as weight = bsdf_1 on the patch */
static int on_patch(const scene_t *S, v3 p) {
    const ertb_scene_desc *d = S->desc;
    return d->has_patch && fabs(p.x - d->patch_rect[0]) <= d->patch_rect[2] && fabs(p.y - d->patch_rect[1]) <= d->patch_rect[3];
}

/* BSDF of the shape that was hit: the ground model or a leaf group's bilambertian */
static double surf_eval(const scene_t *S, const si_t *si, v3 wi, v3 wo) {
    if (si->shape == SHAPE_LEAF) {
        const canopy_group_t *G = &S->canopy.groups[si->group];
        double a[3] = { wi.x, wi.y, wi.z }, b[3] = { wo.x, wo.y, wo.z };
        if (si->trunk) /* diffuse.cpp:127-143: one-sided Lambertian */
            return wi.z > 0.0 && wo.z > 0.0 ? G->trunk_reflectance * INV_PI * wo.z : 0.0;
        if (si->mesh) return bilambertian_eval(si->mesh_r, si->mesh_t, a, b);
        return bilambertian_eval(G->reflectance, G->transmittance, a, b);
    }
    if (on_patch(S, si->p)) return bsdf_eval_tp(S, S->desc->patch_bsdf_type, S->desc->patch_bsdf_params, wi, wo);
    return bsdf_eval(S, wi, wo);
}
static double surf_sample(const scene_t *S, const si_t *si, v3 wi, double s1, double u1, double u2, v3 *wo) {
    if (si->shape == SHAPE_LEAF && si->trunk) { /* diffuse.cpp:100-125 */
        double o[3];
        ertbo_square_to_cosine_hemisphere(u1, u2, o);
        *wo = V(o[0], o[1], o[2]);
        return wi.z > 0.0 && o[2] > 0.0 ? S->canopy.groups[si->group].trunk_reflectance : 0.0;
    }
    if (si->shape == SHAPE_LEAF) {
        const canopy_group_t *G = &S->canopy.groups[si->group];
        double a[3] = { wi.x, wi.y, wi.z }, o[3];
        double w = si->mesh ? bilambertian_sample(si->mesh_r, si->mesh_t, a, s1, u1, u2, o)
                            : bilambertian_sample(G->reflectance, G->transmittance, a, s1, u1, u2, o);
        *wo = V(o[0], o[1], o[2]);
        return w;
    }
    if (on_patch(S, si->p))
        return bsdf_sample_tp(S, S->desc->patch_bsdf_type, S->desc->patch_bsdf_params, wi, s1, u1, u2, wo);
    return bsdf_sample(S, wi, s1, u1, u2, wo);
}

/* -------------------------------------------------------------- sensors */
static void mat_apply_vec(const double *m, v3 v, v3 *o) { /* row-major 4x4, direction */
    *o = V(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[4] * v.x + m[5] * v.y + m[6] * v.z,
           m[8] * v.x + m[9] * v.y + m[10] * v.z);
}
static void mat_apply_pt(const double *m, v3 v, v3 *o) {
    *o = V(m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3], m[4] * v.x + m[5] * v.y + m[6] * v.z + m[7],
           m[8] * v.x + m[9] * v.y + m[10] * v.z + m[11]);
}

static double sensor_ray_offset(const scene_t *S, const ertb_sensor_desc *sd) {
    /* mdistant.cpp:180-190 / hdistant.cpp:219-229 / distantflux.cpp:136-146 */
    if (sd->ray_offset >= 0.0) return sd->ray_offset;
    double rad = fmax(RAY_EPS, S->desc->bsphere_radius * (1.0 + RAY_EPS));
    return sd->target_type == ERTB_TARGET_NONE ? rad : 2.0 * rad;
}

/* sample_ray of the three distant sensors. film = adjusted film position in [0,1)^2 */
static double sensor_sample_ray(const scene_t *S, const ertb_sensor_desc *sd, double fx, double fy,
                                double ax, double ay, ray_t *ray) {
    v3 d, frame_s = V(1, 0, 0), frame_t = V(0, 1, 0);
    double weight = 1.0;
    if (sd->type == ERTB_SENSOR_PERSPECTIVE) {
        /* perspective.cpp:200-236 with sample_to_camera = inverse of sensor.h:234-269 */
        double tn = tan(0.5 * sd->x_fov_deg * PI / 180.0), aspect = (double) sd->width / (double) sd->height;
        v3 dc = vnormalize(V((1.0 - 2.0 * fx) * tn, (1.0 - 2.0 * fy) * tn / aspect, 1.0));
        mat_apply_vec(sd->to_world, dc, &d);
        double inv_z = 1.0 / dc.z, near_t = sd->near_clip * inv_z, far_t = sd->far_clip * inv_z;
        ray->o = vfma(d, near_t, V(sd->to_world[3], sd->to_world[7], sd->to_world[11]));
        ray->d = d;
        ray->maxt = far_t - near_t;
        return 1.0;
    }
    if (sd->type == ERTB_SENSOR_MRADIANCEMETER) { /* mradiancemeter.cpp:147-172 */
        int idx = (int) (fx * sd->n_directions);
        if (idx > sd->n_directions - 1) idx = sd->n_directions - 1;
        ray->o = V(sd->origins[3 * idx], sd->origins[3 * idx + 1], sd->origins[3 * idx + 2]);
        ray->d = vnormalize(V(sd->directions[3 * idx], sd->directions[3 * idx + 1], sd->directions[3 * idx + 2]));
        ray->maxt = DBL_MAX;
        return 1.0;
    }
    if (sd->type == ERTB_SENSOR_MPDISTANT) { /* mpdistant.cpp:214-262: the film sample picks the target point */
        mat_apply_vec(sd->to_world, V(0, 0, 1), &d);
        d = vnormalize(d);
        mat_apply_vec(sd->to_world, V(1, 0, 0), &frame_s);
        mat_apply_vec(sd->to_world, V(0, 1, 0), &frame_t);
        ax = fx; ay = fy;
    } else if (sd->type == ERTB_SENSOR_MDISTANT) { /* mdistant.cpp:192-242 */
        int idx = (int) (fx * sd->n_directions);
        if (idx > sd->n_directions - 1) idx = sd->n_directions - 1;
        d = vnormalize(V(sd->directions[3 * idx], sd->directions[3 * idx + 1], sd->directions[3 * idx + 2]));
        /* look_at(0, direction, up) : columns (left, new_up, dir) */
        v3 up, tmp;
        coordinate_system(d, &up, &tmp);
        v3 left = vnormalize(V(up.y * d.z - up.z * d.y, up.z * d.x - up.x * d.z, up.x * d.y - up.y * d.x));
        frame_s = left;
        frame_t = V(d.y * left.z - d.z * left.y, d.z * left.x - d.x * left.z, d.x * left.y - d.y * left.x);
    } else { /* hdistant.cpp:248-250, distantflux.cpp:163-170 */
        double h[3];
        ertbo_square_to_uniform_hemisphere(fx, fy, h);
        mat_apply_vec(sd->to_world, V(-h[0], -h[1], -h[2]), &d);
        mat_apply_vec(sd->to_world, V(1, 0, 0), &frame_s);
        mat_apply_vec(sd->to_world, V(0, 1, 0), &frame_t);
        if (sd->type == ERTB_SENSOR_DISTANTFLUX) {
            v3 nref;
            mat_apply_vec(sd->to_world, V(0, 0, 1), &nref);
            int npix = sd->width * sd->height;
            weight = vdot(vneg(d), nref) / (INV_TWO_PI * npix);
        }
    }
    double off = sensor_ray_offset(S, sd);
    ray->d = d;
    ray->maxt = DBL_MAX;
    if (sd->target_type == ERTB_TARGET_POINT) {
        ray->o = vfma(d, -off, V(sd->target[0], sd->target[1], sd->target[2]));
    } else if (sd->target_type == ERTB_TARGET_RECTANGLE) {
        /* Shape::sample_position of `rectangle`: uniform on to_world*[-1,1]^2; weight
         * 1/(pdf*area) = 1 */
        v3 p;
        mat_apply_pt(sd->target_to_world, V(2.0 * ax - 1.0, 2.0 * ay - 1.0, 0.0), &p);
        ray->o = vfma(d, -off, p);
    } else if (sd->target_type == ERTB_TARGET_DISK) {
        double x, y;
        ertbo_square_to_uniform_disk_concentric(ax, ay, &x, &y);
        v3 p;
        mat_apply_pt(sd->target_to_world, V(x, y, 0.0), &p);
        ray->o = vfma(d, -off, p);
    } else {
        double x, y;
        ertbo_square_to_uniform_disk_concentric(ax, ay, &x, &y);
        double rad = fmax(RAY_EPS, S->desc->bsphere_radius * (1.0 + RAY_EPS));
        v3 c = V(S->desc->bsphere_center[0], S->desc->bsphere_center[1], S->desc->bsphere_center[2]);
        v3 perp = vadd(vmul(frame_s, x), vmul(frame_t, y));
        ray->o = vfma(d, -off, vfma(perp, rad, c));
    }
    return weight;
}

/* ----------------------------------------------------------- integrator */
static uint64_t g_last_flights[2];
/* trips_*: iterations of the reference's loops (volpath.cpp:170-393, :454-551), stencil crossings included.
 * flights_*: free flights actually sampled (Medium::sample_interaction calls) -- the quantity a kernel that
 * has no stencil-crossing iterations counts as its loop trips.  flights_nee counts only shadow rays that can
 * contribute: a ray that ends on an opaque surface, or whose scattering value is zero, is walked by the
 * reference for nothing (a kernel may cull it without changing any estimate). */
typedef struct {
    uint64_t trips_main, trips_nee, n_scatter, n_surface;
    uint64_t flights_main, flights_nee;
    uint64_t ray_flights; int ray_opaque; /* scratch: the shadow ray being walked */
} counters_t;

/* Shading frame of a surface hit: SurfaceInteraction::initialize_sh_frame
 * (MI/include/mitsuba/render/interaction.h:278-288) with the sphere's dp_du = (-y, x, 0) 2 pi
 * (MI/src/shapes/sphere.cpp:703-720) or the rectangle's dp_du = to_world * (2, 0, 0).  Only
 * the ocean BSDF (absolute wind direction) is sensitive to it. */
static frame_t surface_frame(const scene_t *S, const si_t *si) {
    frame_t f;
    f.n = si->n;
    if (S->spherical) {
        v3 dp_du = V(-si->p.y, si->p.x, 0.0);
        if (dp_du.x == 0.0 && dp_du.y == 0.0) return make_frame(si->n);
        f.s = vnormalize(vfma(si->n, -vdot(si->n, dp_du), dp_du));
    } else {
        f.s = V(1, 0, 0);
    }
    f.t = V(f.n.y * f.s.z - f.n.z * f.s.y, f.n.z * f.s.x - f.n.x * f.s.z, f.n.x * f.s.y - f.n.y * f.s.x);
    return f;
}

static int target_medium(v3 n, v3 d) { return vdot(d, n) > 0.0 ? 0 : 1; } /* interaction.h:318-332 */

/* volpath.cpp:567-572 (power heuristic); volpathmis.cpp:657-669 combines the two full-path densities as
 * 1 / (p_a / f + p_b / f), i.e. the balance heuristic (`balance` != 0) */
static int g_mis_balance = 0;
#pragma omp threadprivate(g_mis_balance)
static double mis_weight(double pdf_a, double pdf_b) {
    if (!g_mis_balance) { pdf_a *= pdf_a; pdf_b *= pdf_b; }
    double w = pdf_a / (pdf_a + pdf_b);
    return w == w ? w : 0.0; /* detach(select(isfinite(w), w, 0)) */
}
/* Direction towards the emitter, its density and the emitter value / density: directional.cpp:171-201
 * (delta: *pdf = 0 flags it) or astroobject.cpp:141-175 (warp::square_to_uniform_cone about the axis,
 * warp.h:474-497; pdf = 1 / omega, weight = E / omega / pdf = E). */
static v3 emitter_direction_sample(const scene_t *S, double u1, double u2, double *pdf) {
    v3 axis = vneg(S->emitter_d);
    *pdf = 0.0;
    if (!S->astro) return axis;
    double x, y;
    ertbo_square_to_uniform_disk_concentric(u1, u2, &x, &y);
    double pn = x * x + y * y, omc = 1.0 - S->astro_cos;
    double z = S->astro_cos + omc * (1.0 - pn);
    double f = pn > 0.0 ? safe_sqrt((1.0 - z * z) / pn) : 0.0;
    frame_t fr = make_frame(axis);
    *pdf = 1.0 / S->astro_omega;
    return vnormalize(to_world(&fr, V(x * f, y * f, z)));
}
/* astroobject.cpp:111-124 eval for a ray leaving the scene along `d`; :206-215 pdf_direction */
static double astro_eval(const scene_t *S, v3 d) {
    return S->astro && vdot(d, vneg(S->emitter_d)) > S->astro_cos ? S->desc->irradiance / S->astro_omega : 0.0;
}

/* volpath.cpp:400-554 sample_emitter: ratio tracking toward the directional emitter.
 * `ref_n` is the zero vector for medium interactions. */
static double sample_emitter(const scene_t *S, pcg32 *rng, v3 ref_p, v3 ref_n, int medium,
                             counters_t *C, v3 *ds_d, double *ds_pdf) {
    const ertb_scene_desc *D = S->desc;
    double e1 = next_1d(rng), e2 = next_1d(rng); /* next_2d, volpath.cpp:407 */
    /* directional.cpp:171-201 / astroobject.cpp:141-175 */
    v3 d = vneg(emitter_direction_sample(S, e1, e2, ds_pdf));
    v3 c = V(D->bsphere_center[0], D->bsphere_center[1], D->bsphere_center[2]);
    double brad = fmax(RAY_EPS, D->bsphere_radius * (1.0 + RAY_EPS));
    double radius = fmax(brad, vnorm(vsub(ref_p, c)));
    double dist = 2.0 * radius;
    v3 ds_p = vfma(d, -dist, ref_p);
    *ds_d = vneg(d);
    double emitter_val = D->irradiance;

    ray_t ray = spawn_ray_to(ref_p, ref_n, ds_p);
    double max_dist = ray.maxt, total_dist = 0.0, transmittance = 1.0;
    si_t si; memset(&si, 0, sizeof si); si.t = INFINITY; si.shape = -1;
    int needs_intersection = 1, active = 1;
    C->ray_flights = 0; C->ray_opaque = 0;
    while (active) {
        double remaining = max_dist - total_dist;
        ray.maxt = remaining;
        if (!(remaining > 0.0)) break;
        C->trips_nee++;
        int escaped = 0, active_medium = medium, active_surface = !medium;
        if (active_medium) {
            C->ray_flights++;
            mei_t mei = sample_interaction(S, &ray, next_1d(rng));
            if (D->homogeneous && mei.t < INFINITY) ray.maxt = fmin(mei.t, remaining);
            if (needs_intersection) si = scene_intersect(S, &ray);
            if (si.t < mei.t) mei.t = INFINITY;
            needs_intersection = 0;
            { /* is_spectral branch (has_spectral_extinction defaults to true) */
                double t = fmin(remaining, fmin(mei.t, si.t)) - mei.mint;
                double tr = exp(-t * mei.combined);
                double pdf = (si.t < mei.t || mei.t > remaining) ? tr : tr * mei.combined;
                transmittance *= pdf > 0.0 ? tr / pdf : 0.0;
            }
            if (mei.t > remaining && mei.t < INFINITY) total_dist = dist;
            if (mei.t > remaining) mei.t = INFINITY;
            escaped = !(mei.t < INFINITY);
            active_medium = mei.t < INFINITY;
            if (active_medium) {
                total_dist += mei.t;
                ray.o = mei.p;
                si.t -= mei.t;
                transmittance *= mei.sigma_n;
            }
        }
        int intersect = active_surface && needs_intersection;
        if (intersect) { si = scene_intersect(S, &ray); needs_intersection = 0; }
        active_surface |= escaped;
        if (active_surface) total_dist += si.t;
        active_surface = active_surface && (si.t < INFINITY) && !active_medium;
        if (active_surface) {
            /* eval_null_transmission: null.cpp -> 1, every other BSDF -> 0 */
            transmittance *= si.shape == SHAPE_TOA ? 1.0 : 0.0;
            if (si.shape != SHAPE_TOA) C->ray_opaque = 1;
            ray = spawn_ray(si.p, si.n, ray.d);
            needs_intersection = 1;
        }
        ray.maxt = remaining;
        active = (active_medium || active_surface) && transmittance != 0.0;
        if (active_surface && si.shape == SHAPE_TOA) medium = target_medium(si.n, ray.d);
    }
    return transmittance * emitter_val;
}

/* piecewise_volpath.cpp:404-527 sample_emitter: one exact transmittance evaluation per medium
 * segment instead of ratio tracking. */
static double pw_sample_emitter(const scene_t *S, pcg32 *rng, v3 ref_p, v3 ref_n, int medium,
                                counters_t *C, v3 *ds_d, double *ds_pdf) {
    const ertb_scene_desc *D = S->desc;
    double e1 = next_1d(rng), e2 = next_1d(rng); /* next_2d, piecewise_volpath.cpp:413 */
    v3 d = vneg(emitter_direction_sample(S, e1, e2, ds_pdf)); /* directional.cpp:171-201 / astroobject.cpp:141-175 */
    v3 c = V(D->bsphere_center[0], D->bsphere_center[1], D->bsphere_center[2]);
    double brad = fmax(RAY_EPS, D->bsphere_radius * (1.0 + RAY_EPS));
    double radius = fmax(brad, vnorm(vsub(ref_p, c)));
    double dist = 2.0 * radius;
    v3 ds_p = vfma(d, -dist, ref_p);
    *ds_d = vneg(d);
    double emitter_val = D->irradiance;

    ray_t ray = spawn_ray_to(ref_p, ref_n, ds_p);
    double max_dist = ray.maxt, total_dist = 0.0, transmittance = 1.0;
    int active = 1;
    C->ray_flights = 0; C->ray_opaque = 0;
    while (active) {
        double remaining = max_dist - total_dist;
        ray.maxt = remaining;
        if (!(remaining > 0.0)) break;
        C->trips_nee++;
        si_t si = scene_intersect(S, &ray);
        int escaped = 0, active_medium = medium, active_surface = si.t < INFINITY;
        if (active_medium) total_dist += fmin(ray.maxt, si.t);
        if (active_surface && !active_medium) total_dist += si.t;
        if (active_medium) {
            double mint, maxt, tr, pdf;
            int hit = medium_aabb(S, &ray, &mint, &maxt);
            escaped = pw_eval_transmittance_pdf_real(S, &ray, si.t, hit, mint, maxt, &tr, &pdf);
            transmittance *= tr; /* "exact estimation" */
            C->ray_flights++;
            active_medium = !escaped;
        }
        if (active_surface) {
            transmittance *= si.shape == SHAPE_TOA ? 1.0 : 0.0; /* eval_null_transmission */
            if (si.shape != SHAPE_TOA) C->ray_opaque = 1;
            ray = spawn_ray(si.p, si.n, ray.d);
        }
        ray.maxt = remaining;
        active = (active_medium || active_surface) && transmittance != 0.0;
        if (active_surface && si.shape == SHAPE_TOA) medium = target_medium(si.n, ray.d);
    }
    return transmittance * emitter_val;
}

/* piecewise_volpath.cpp:215-246: intersect, analytic free flight, tr / pdf weight */
static double pw_medium_step(const scene_t *S, pcg32 *rng, const ray_t *ray, si_t *si,
                             int *needs_intersection, mei_t *mei) {
    if (*needs_intersection) *si = scene_intersect(S, ray);
    *needs_intersection = 0;
    double mint, maxt, tr, pdf;
    int hit = medium_aabb(S, ray, &mint, &maxt);
    *mei = pw_sample_interaction_real(S, ray, si->t, next_1d(rng), hit, mint, maxt, &tr, &pdf);
    if (si->t < mei->t) mei->t = INFINITY;
    return pdf > 0.0 ? tr / pdf : 0.0;
}

/* volpath.cpp:93-396 (mono, unpolarized).  mis != 0 selects the volpathmis.cpp
 * Russian-roulette placement (:227-231), the only difference left in mono. */
/* BSDF::pdf of the shape that was hit (only needed for MIS against a non-delta emitter: astroobject scenes) */
static double surf_pdf(const scene_t *S, const si_t *si, v3 wi, v3 wo) {
    if (si->shape == SHAPE_LEAF && !si->trunk) { /* bilambertian.cpp:161-204 */
        const canopy_group_t *G = &S->canopy.groups[si->group];
        double a[3] = { wi.x, wi.y, wi.z }, b[3] = { wo.x, wo.y, wo.z };
        return si->mesh ? bilambertian_pdf(si->mesh_r, si->mesh_t, a, b) : bilambertian_pdf(G->reflectance, G->transmittance, a, b);
    }
    if (!(wi.z > 0.0 && wo.z > 0.0)) return 0.0;
    if (si->shape == SHAPE_LEAF || on_patch(S, si->p)) return INV_PI * wo.z; /* trunk (diffuse.cpp:145-160), patch land BSDF */
    const int type = S->desc->bsdf_type;
    if (type == ERTB_BSDF_OCEAN_LEGACY) return ocean_pdf(&S->ocean, wi.x, wi.y, wi.z, wo.x, wo.y, wo.z);
    if (type == ERTB_BSDF_OCEAN_MISHCHENKO || type == ERTB_BSDF_OCEAN_GRASP) {
        double a[3] = { wi.x, wi.y, wi.z }, b[3] = { wo.x, wo.y, wo.z };
        return glint_pdf(&S->glint, a, b);
    }
    if (type == ERTB_BSDF_MEASURED_MONO) {
        double a[3] = { wi.x, wi.y, wi.z }, b[3] = { wo.x, wo.y, wo.z };
        return mm_oracle_pdf(S->desc->bsdf_table, a, b);
    }
    return INV_PI * wo.z; /* warp::square_to_cosine_hemisphere_pdf */
}
/* volpath.cpp:328-346: a ray that leaves the scene looks at the environment emitter (astroobject).  Direct for
 * camera rays and specular chains, MIS-weighted against the emitter sampling of the last event otherwise. */
static double astro_hit(const scene_t *S, v3 d, uint64_t depth, int specular_chain, double last_pdf) {
    double em = astro_eval(S, d);
    if (em == 0.0) return 0.0;
    if (S->desc->hide_emitters) { /* volpath.cpp:114, :329-330: no hit at depth 0, no specular chain from the camera */
        if (depth == 0) return 0.0;
        specular_chain = 0;
    }
    return (depth == 0 || specular_chain) ? em : em * mis_weight(last_pdf, 1.0 / S->astro_omega);
}

static double volpath_sample(const scene_t *S, pcg32 *rng, ray_t ray, int medium0, counters_t *C) {
    const ertb_scene_desc *D = S->desc;
    const int mis = D->integrator == ERTB_INTEGRATOR_VOLPATHMIS;
    const int pw = D->integrator == ERTB_INTEGRATOR_PIECEWISE_VOLPATH;
    const uint64_t max_depth = D->max_depth < 0 ? (uint64_t) 0xffffffffu : (uint64_t) D->max_depth;
    double throughput = 1.0, result = 0.0, eta = 1.0;
    int medium = medium0; /* distant sensors sit outside the atmosphere; a camera may be inside */
    uint64_t depth = 0;
    si_t si; memset(&si, 0, sizeof si); si.t = INFINITY; si.shape = -1; si.group = -1;
    int needs_intersection = 1, last_event_was_null = 0;
    int specular_chain = 1;  /* volpath.cpp:113 (hide_emitters = false) */
    double last_pdf = 0.0;   /* last_scatter_direction_pdf */
    g_mis_balance = mis;

    for (;;) {
        /* ---- termination, volpath.cpp:189-202 ---- */
        if (throughput == 0.0) break;
        double q = fmin(throughput * eta * eta, 0.95);
        int perform_rr = depth > (uint64_t) D->rr_depth && !(mis && last_event_was_null);
        double xi = next_1d(rng);
        if (perform_rr) {
            if (!(xi < q)) break;
            throughput /= q;
        }
        last_event_was_null = 0;
        if (!(depth < max_depth)) break;
        C->trips_main++;

        int active_medium = medium, active_surface = !medium;
        int escaped = 0, null_scatter = 0, medium_scatter = 0;
        mei_t mei; memset(&mei, 0, sizeof mei);

        if (active_medium && pw) {
            C->flights_main++;
            throughput *= pw_medium_step(S, rng, &ray, &si, &needs_intersection, &mei);
            escaped = !(mei.t < INFINITY);
            active_medium = mei.t < INFINITY;
            if (active_medium) { medium_scatter = 1; depth++; }
        } else if (active_medium) { /* :218-259 */
            C->flights_main++;
            mei = sample_interaction(S, &ray, next_1d(rng));
            if (D->homogeneous && mei.t < INFINITY) ray.maxt = mei.t;
            if (needs_intersection) si = scene_intersect(S, &ray);
            needs_intersection = 0;
            if (si.t < mei.t) mei.t = INFINITY;
            { /* transmittance_eval_pdf, medium.cpp:87-96 */
                double t = fmin(mei.t, si.t) - mei.mint;
                double tr = exp(-t * mei.combined);
                double pdf = si.t < mei.t ? tr : tr * mei.combined;
                throughput *= pdf > 0.0 ? tr / pdf : 0.0;
            }
            escaped = !(mei.t < INFINITY);
            active_medium = mei.t < INFINITY;
            if (active_medium) {
                double pnull = mei.sigma_n / mei.combined;
                null_scatter = next_1d(rng) < pnull;
                medium_scatter = !null_scatter;
                if (null_scatter) { throughput *= mei.sigma_n / pnull; last_event_was_null = 1; }
                else depth++;
            }
        }
        if (!(depth < max_depth)) break; /* :262-263: active false -> loop ends */

        if (null_scatter) { ray.o = mei.p; si.t -= mei.t; }

        if (medium_scatter) { /* :270-310 */
            C->n_scatter++;
            throughput *= mei.sigma_s / (mei.sigma_t / mei.combined);
            v3 ds_d;
            double ds_pdf;
            double emitted = (pw ? pw_sample_emitter : sample_emitter)(S, rng, mei.p, V(0, 0, 0), medium, C, &ds_d, &ds_pdf);
            double pv, ppdf;
            phase_eval_pdf(S, mei.layer, mei.wi, ds_d, &pv, &ppdf);
            /* mis_weight(ds.pdf, select(ds.delta, 0, phase_pdf)): 1 for a delta emitter */
            result += throughput * pv * emitted * (ds_pdf > 0.0 ? mis_weight(ds_pdf, ppdf) : 1.0);
            if (!C->ray_opaque && throughput * pv > 0.0) C->flights_nee += C->ray_flights;
            specular_chain = 0; /* :276-277 */
            v3 wo; double pw, pp;
            double s1 = next_1d(rng), u1 = next_1d(rng), u2 = next_1d(rng);
            phase_sample(S, mei.layer, mei.wi, s1, u1, u2, &wo, &pw, &pp);
            if (pp > 0.0) {
                ray = spawn_ray(mei.p, V(0, 0, 0), wo);
                needs_intersection = 1;
                last_pdf = pp;
                throughput *= pw;
            }
        }

        /* ---- surfaces, :312-389 ---- */
        active_surface |= escaped;
        if (active_surface && needs_intersection) si = scene_intersect(S, &ray);
        if (S->astro && active_surface && !(si.t < INFINITY))
            result += throughput * astro_hit(S, ray.d, depth, specular_chain, last_pdf);
        active_surface = active_surface && si.t < INFINITY;
        if (active_surface) {
            frame_t fr = si.shape == SHAPE_LEAF ? make_frame(si.mesh ? si.sh_n : si.n) : surface_frame(S, &si);
            v3 wi = to_local(&fr, vneg(ray.d));
            v3 wo_world;
            if (si.shape == SHAPE_TOA) { /* null.cpp:41-87 */
                (void) next_1d(rng); (void) next_1d(rng); (void) next_1d(rng);
                wo_world = ray.d;
            } else {
                C->n_surface++;
                if (depth + 1 < max_depth) { /* :349-363 */
                    v3 ds_d;
                    double ds_pdf;
                    double emitted = (pw ? pw_sample_emitter : sample_emitter)(S, rng, si.p, si.n, medium, C, &ds_d, &ds_pdf);
                    v3 wo = to_local(&fr, ds_d);
                    const double fv = surf_eval(S, &si, wi, wo);
                    result += throughput * fv * emitted * (ds_pdf > 0.0 ? mis_weight(ds_pdf, surf_pdf(S, &si, wi, wo)) : 1.0);
                    if (!C->ray_opaque && throughput * fv > 0.0) C->flights_nee += C->ray_flights;
                }
                double s1 = next_1d(rng), u1 = next_1d(rng), u2 = next_1d(rng);
                v3 wo;
                throughput *= surf_sample(S, &si, wi, s1, u1, u2, &wo);
                if (S->astro) last_pdf = surf_pdf(S, &si, wi, wo); /* bs.pdf, :383 */
                specular_chain = 0; /* every ground BSDF here is smooth, :387 */
                wo_world = to_world(&fr, wo);
                depth++;
            }
            ray = spawn_ray(si.p, si.n, wo_world);
            needs_intersection = 1;
            if (si.shape == SHAPE_TOA) medium = target_medium(si.n, ray.d);
        }
        if (!(active_surface || active_medium)) break;
    }
    return result;
}



/* Polarized BSDF value (times cos) in WORLD-space implicit Stokes bases.  Land BSDFs are
 * depolarizer(value) (rotation invariant).  ocean_legacy.cpp:603-634 builds the glint matrix in
 * the meridian-plane bases and rotates it to the implicit bases of -wo_hat / wi_hat (local frame);
 * volpath.cpp:357 / :369 then applies si.to_world_mueller(M, -wo, si.wi) (interaction.h:407-428). */
/* `weight` = 0: BSDF::eval; 1: the weight BSDF::sample returns for the sampled direction `wo` (eval / pdf for
 * the two-lobe ocean models, F G / G1 for ocean_mishchenko.cpp:180-181, C F for maignan.cpp:189-193).
 * `fr` = NULL: stop after the plugin's own rotation (local implicit bases: what BSDF::eval returns). */
static void bsdf_mueller(const scene_t *S, const frame_t *fr, int weight, v3 wi, v3 wo, mueller_t *M) {
    *M = mu_zero();
    const int type = S->desc->bsdf_type;
    if (type != ERTB_BSDF_OCEAN_LEGACY && !is_glint_family(type)) {
        M->m[0] = bsdf_eval(S, wi, wo);
        if (weight) M->m[0] = wo.z > 0.0 ? M->m[0] / (INV_PI * wo.z) : 0.0;
        return;
    }
    double wil[3] = { wi.x, wi.y, wi.z }, wol[3] = { wo.x, wo.y, wo.z }, dep, gl[16];
    if (type == ERTB_BSDF_OCEAN_LEGACY) {
        ocean_eval_polarized(&S->ocean, wil, wol, &dep, gl);
        if (weight) { /* ocean_legacy.cpp:553-558: eval / pdf */
            double pdf = ocean_pdf(&S->ocean, wi.x, wi.y, wi.z, wo.x, wo.y, wo.z);
            double ip = pdf > 0.0 ? 1.0 / pdf : 0.0;
            dep *= ip;
            for (int i = 0; i < 16; ++i) gl[i] *= ip;
        }
    } else {
        glint_polarized(&S->glint, weight, wil, wol, &dep, gl);
    }
    mueller_t G; memcpy(G.m, gl, sizeof gl);
    if (wi.z > 0.0 && wo.z > 0.0) {
        v3 n = V(0, 0, 1), in_fwd = vneg(wo), out_fwd = wi; /* wo_hat = wo, wi_hat = si.wi */
        v3 p_in = vnormalize(vcross(vnormalize(vcross(n, in_fwd)), in_fwd));
        v3 p_out = vnormalize(vcross(vnormalize(vcross(n, out_fwd)), out_fwd));
        if (isnan(p_in.x) || isnan(p_in.y) || isnan(p_in.z)) p_in = V(0, 1, 0);
        if (isnan(p_out.x) || isnan(p_out.y) || isnan(p_out.z)) p_out = V(0, 1, 0);
        G = rotate_mueller_basis(&G, in_fwd, p_in, stokes_basis(in_fwd), out_fwd, p_out, stokes_basis(out_fwd));
        if (fr) {
            /* to_world_mueller(M, in_forward_local = -wo, out_forward_local = si.wi) */
            v3 in_w = to_world(fr, in_fwd), out_w = to_world(fr, out_fwd);
            G = rotate_mueller_basis(&G, in_w, to_world(fr, stokes_basis(in_fwd)), stokes_basis(in_w),
                                     out_w, to_world(fr, stokes_basis(out_fwd)), stokes_basis(out_w));
        }
    }
    *M = G;
    M->m[0] += dep;
}
static void bsdf_eval_mueller(const scene_t *S, const frame_t *fr, v3 wi, v3 wo, mueller_t *M) {
    bsdf_mueller(S, fr, 0, wi, wo, M);
}

/* volpath.cpp:93-396 in a polarized variant: Spectrum = 4x4 Mueller matrix.  `throughput` is
 * right-multiplied by every weight (throughput *= w), the emitter value is depolarizer(E), so
 * only the first column of `result` is ever non-zero: it is kept as a Stokes 4-vector.  All the
 * surface BSDFs of this path return depolarizer(value), which to_world_mueller() leaves unchanged.
 * stokes.cpp:97-168 then rotates the vector from the implicit basis of -ray.d to the output basis. */
static void volpath_sample_pol(const scene_t *S, pcg32 *rng, ray_t ray, int medium0, v3 sensor_up, counters_t *C, double stokes[4]) {
    const ertb_scene_desc *D = S->desc;
    const int mis = D->integrator == ERTB_INTEGRATOR_VOLPATHMIS;
    const int pw = D->integrator == ERTB_INTEGRATOR_PIECEWISE_VOLPATH;
    const uint64_t max_depth = D->max_depth < 0 ? (uint64_t) 0xffffffffu : (uint64_t) D->max_depth;
    const v3 primary_d = ray.d;
    mueller_t T = mu_identity(1.0);
    double result[4] = { 0, 0, 0, 0 }, eta = 1.0;
    int medium = medium0;
    uint64_t depth = 0;
    si_t si; memset(&si, 0, sizeof si); si.t = INFINITY; si.shape = -1; si.group = -1;
    int needs_intersection = 1, last_event_was_null = 0;
    int specular_chain = 1;
    double last_pdf = 0.0;

    for (;;) {
        int any_nz = 0;
        for (int i = 0; i < 16; ++i) if (T.m[i] != 0.0) { any_nz = 1; break; }
        /* active &= any(unpolarized_spectrum(throughput) != 0): the (0,0) entry */
        if (T.m[0] == 0.0) break;
        (void) any_nz;
        double q = fmin(T.m[0] * eta * eta, 0.95);
        int perform_rr = depth > (uint64_t) D->rr_depth && !(mis && last_event_was_null);
        double xi = next_1d(rng);
        if (perform_rr) {
            if (!(xi < q)) break;
            T = mu_scale(&T, 1.0 / q);
        }
        last_event_was_null = 0;
        if (!(depth < max_depth)) break;
        C->trips_main++;

        int active_medium = medium, active_surface = !medium;
        int escaped = 0, null_scatter = 0, medium_scatter = 0;
        mei_t mei; memset(&mei, 0, sizeof mei);
        if (active_medium && pw) {
            C->flights_main++;
            T = mu_scale(&T, pw_medium_step(S, rng, &ray, &si, &needs_intersection, &mei));
            escaped = !(mei.t < INFINITY);
            active_medium = mei.t < INFINITY;
            if (active_medium) { medium_scatter = 1; depth++; }
        } else if (active_medium) {
C->flights_main++;
                        mei = sample_interaction(S, &ray, next_1d(rng));
            if (D->homogeneous && mei.t < INFINITY) ray.maxt = mei.t;
            if (needs_intersection) si = scene_intersect(S, &ray);
            needs_intersection = 0;
            if (si.t < mei.t) mei.t = INFINITY;
            {
                double t = fmin(mei.t, si.t) - mei.mint;
                double tr = exp(-t * mei.combined);
                double pdf = si.t < mei.t ? tr : tr * mei.combined;
                T = mu_scale(&T, pdf > 0.0 ? tr / pdf : 0.0);
            }
            escaped = !(mei.t < INFINITY);
            active_medium = mei.t < INFINITY;
            if (active_medium) {
                double pnull = mei.sigma_n / mei.combined;
                null_scatter = next_1d(rng) < pnull;
                medium_scatter = !null_scatter;
                if (null_scatter) { T = mu_scale(&T, mei.sigma_n / pnull); last_event_was_null = 1; }
                else depth++;
            }
        }
        if (!(depth < max_depth)) break;
        if (null_scatter) { ray.o = mei.p; si.t -= mei.t; }
        if (medium_scatter) {
            C->n_scatter++;
            T = mu_scale(&T, mei.sigma_s / (mei.sigma_t / mei.combined));
            v3 ds_d;
            double ds_pdf;
            double emitted = (pw ? pw_sample_emitter : sample_emitter)(S, rng, mei.p, V(0, 0, 0), medium, C, &ds_d, &ds_pdf);
            mueller_t P;
            double ppdf;
            phase_eval_mueller_pdf(S, mei.layer, mei.wi, ds_d, &P, &ppdf);
            mueller_t TP = mu_mul(&T, &P);
            const double wm = ds_pdf > 0.0 ? mis_weight(ds_pdf, ppdf) : 1.0;
            for (int i = 0; i < 4; ++i) result[i] += TP.m[4 * i] * emitted * wm; /* * depolarizer(E): column 0 */
            if (!C->ray_opaque && TP.m[0] > 0.0) C->flights_nee += C->ray_flights;
            specular_chain = 0;
            v3 wo; mueller_t W; double pp;
            double s1 = next_1d(rng), u1 = next_1d(rng), u2 = next_1d(rng);
            phase_sample_mueller(S, mei.layer, mei.wi, s1, u1, u2, &wo, &W, &pp);
            if (pp > 0.0) {
                ray = spawn_ray(mei.p, V(0, 0, 0), wo);
                needs_intersection = 1;
                last_pdf = pp;
                T = mu_mul(&T, &W);
            }
        }
        active_surface |= escaped;
        if (active_surface && needs_intersection) si = scene_intersect(S, &ray);
        if (S->astro && active_surface && !(si.t < INFINITY)) { /* depolarizer(E / omega): column 0 of T */
            const double em = astro_hit(S, ray.d, depth, specular_chain, last_pdf);
            for (int i = 0; i < 4; ++i) result[i] += T.m[4 * i] * em;
        }
        active_surface = active_surface && si.t < INFINITY;
        if (active_surface) {
            frame_t fr = surface_frame(S, &si);
            v3 wi = to_local(&fr, vneg(ray.d));
            v3 wo_world;
            if (si.shape == SHAPE_TOA) {
                (void) next_1d(rng); (void) next_1d(rng); (void) next_1d(rng);
                wo_world = ray.d;
            } else {
                C->n_surface++;
                if (depth + 1 < max_depth) {
                    v3 ds_d;
                    double ds_pdf;
                    double emitted = (pw ? pw_sample_emitter : sample_emitter)(S, rng, si.p, si.n, medium, C, &ds_d, &ds_pdf);
                    v3 wo = to_local(&fr, ds_d);
                    mueller_t B, TB;
                    bsdf_eval_mueller(S, &fr, wi, wo, &B);
                    TB = mu_mul(&T, &B);
                    const double wm = ds_pdf > 0.0 ? mis_weight(ds_pdf, surf_pdf(S, &si, wi, wo)) : 1.0;
                    for (int i = 0; i < 4; ++i) result[i] += TB.m[4 * i] * emitted * wm;
                    if (!C->ray_opaque && TB.m[0] > 0.0) C->flights_nee += C->ray_flights;
                }
                double s1 = next_1d(rng), u1 = next_1d(rng), u2 = next_1d(rng);
                v3 wo;
                double w = bsdf_sample(S, wi, s1, u1, u2, &wo);
                mueller_t Dw = mu_zero();
                if (D->bsdf_type == ERTB_BSDF_OCEAN_LEGACY || is_glint_family(D->bsdf_type)) {
                    (void) w;
                    bsdf_mueller(S, &fr, 1, wi, wo, &Dw);
                } else {
                    Dw.m[0] = w; /* depolarizer(w) */
                }
                T = mu_mul(&T, &Dw);
                if (S->astro) last_pdf = surf_pdf(S, &si, wi, wo);
                specular_chain = 0;
                wo_world = to_world(&fr, wo);
                depth++;
            }
            ray = spawn_ray(si.p, si.n, wo_world);
            needs_intersection = 1;
            if (si.shape == SHAPE_TOA) medium = target_medium(si.n, ray.d);
        }
        if (!(active_surface || active_medium)) break;
    }
    /* stokes.cpp:111-151 */
    v3 fwd = vneg(primary_d);
    v3 current = stokes_basis(fwd), target;
    if (D->meridian_align) {
        v3 tmp = vcross(V(0, 0, 1), fwd);
        if (vnorm(tmp) < RAY_EPS) target = V(1, 0, 0);
        else target = vcross(vnormalize(tmp), fwd);
    } else {
        target = vcross(primary_d, sensor_up);
    }
    mueller_t R = rotate_stokes_basis(fwd, current, target);
    for (int i = 0; i < 4; ++i) {
        stokes[i] = 0.0;
        for (int j = 0; j < 4; ++j) stokes[i] += R.m[4 * i + j] * result[j];
    }
}

/* ---------------------------------------------------------------- render */
int ertbo_render(const ertb_scene_desc *desc, int sensor, uint64_t seed, uint64_t spp,
                 uint64_t sample_offset, double *sum_wl, double *sum_l, double *sum_l2,
                 ertb_render_stats *stats, int n_threads) {
    return ertbo_render_stokes(desc, sensor, seed, spp, sample_offset, sum_wl, sum_l, sum_l2, NULL, stats, n_threads);
}

int ertbo_render_stokes(const ertb_scene_desc *desc, int sensor, uint64_t seed, uint64_t spp,
                        uint64_t sample_offset, double *sum_wl, double *sum_l, double *sum_l2,
                        double *sum_stokes, ertb_render_stats *stats, int n_threads) {
    scene_t S;
    if (scene_init(&S, desc)) return 1;
    if (sensor < 0 || sensor >= desc->n_sensors) { scene_free(&S); return fail("bad sensor index"); }
    const ertb_sensor_desc *sd = &desc->sensors[sensor];
    const int W = sd->width, H = sd->height;
    const int64_t npix = (int64_t) W * H;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#else
    (void) n_threads;
#endif
    /* work items = (pixel, chunk of samples) so a 1-pixel film still uses every core */
    const uint64_t chunk = 4096;
    const uint64_t chunks_per_pixel = (spp + chunk - 1) / chunk;
    const int64_t n_items = npix * (int64_t) chunks_per_pixel;
    double *a_wl = calloc(npix, sizeof(double)), *a_l = calloc(npix, sizeof(double)),
           *a_l2 = calloc(npix, sizeof(double)), *a_st = calloc(4 * npix, sizeof(double));
    v3 sensor_up;
    mat_apply_vec(sd->to_world, V(0, 1, 0), &sensor_up);
    if (sd->type == ERTB_SENSOR_MDISTANT || sd->type == ERTB_SENSOR_MRADIANCEMETER) sensor_up = V(0, 1, 0);
    counters_t total; memset(&total, 0, sizeof total);

#pragma omp parallel
    {
        counters_t C; memset(&C, 0, sizeof C);
#pragma omp for schedule(dynamic, 1)
        for (int64_t item = 0; item < n_items; ++item) {
            int64_t pix = item / (int64_t) chunks_per_pixel;
            uint64_t c0 = (uint64_t) (item % (int64_t) chunks_per_pixel) * chunk;
            uint64_t c1 = c0 + chunk < spp ? c0 + chunk : spp;
            int px = (int) (pix % W), py = (int) (pix / W);
            double s_wl = 0, s_l = 0, s_l2 = 0, s_st[4] = { 0, 0, 0, 0 };
            for (uint64_t s = c0; s < c1; ++s) {
                pcg32 rng;
                uint64_t gid = ((uint64_t) pix << 40) + (sample_offset + s);
                pcg_seed(&rng, mix64(seed * 0x9E3779B97F4A7C15ULL + 0x632BE59BD9B4E019ULL) ^ mix64(gid), gid);
                /* render_sample, integrator.cpp:449-520 */
                double fx = (px + next_1d(&rng)) / W, fy = (py + next_1d(&rng)) / H;
                double ax = next_1d(&rng), ay = next_1d(&rng);
                ray_t ray;
                double w = sensor_sample_ray(&S, sd, fx, fy, ax, ay, &ray);
                double L;
                const int in_medium = (sd->type == ERTB_SENSOR_PERSPECTIVE || sd->type == ERTB_SENSOR_MRADIANCEMETER) && sd->in_medium;
                if (desc->polarized) {
                    double st[4];
                    volpath_sample_pol(&S, &rng, ray, in_medium, sensor_up, &C, st);
                    L = st[0];
                    for (int k = 0; k < 4; ++k) s_st[k] += w * st[k];
                } else {
                    L = volpath_sample(&S, &rng, ray, in_medium, &C);
                }
                s_wl += w * L; s_l += L; s_l2 += L * L;
            }
#pragma omp critical
            {
                a_wl[pix] += s_wl; a_l[pix] += s_l; a_l2[pix] += s_l2;
                for (int k = 0; k < 4; ++k) a_st[k * npix + pix] += s_st[k];
            }
        }
#pragma omp critical
        {
            total.trips_main += C.trips_main; total.trips_nee += C.trips_nee;
            total.n_scatter += C.n_scatter; total.n_surface += C.n_surface;
            total.flights_main += C.flights_main; total.flights_nee += C.flights_nee;
        }
    }
    for (int64_t i = 0; i < npix; ++i) {
        if (sum_wl) sum_wl[i] = a_wl[i];
        if (sum_l) sum_l[i] = a_l[i];
        if (sum_l2) sum_l2[i] = a_l2[i];
    }
    if (sum_stokes) memcpy(sum_stokes, a_st, sizeof(double) * 4 * npix);
    free(a_wl); free(a_l); free(a_l2); free(a_st);
    if (stats) {
        memset(stats, 0, sizeof *stats);
        stats->n_paths = (uint64_t) npix * spp;
        stats->trips_main = total.trips_main; stats->trips_nee = total.trips_nee;
        stats->n_scatter = total.n_scatter; stats->n_surface = total.n_surface;
    }
    g_last_flights[0] = total.flights_main; g_last_flights[1] = total.flights_nee;
    scene_free(&S);
    return 0;
}

/* free flights sampled by the last ertbo_render* call: {main walk, contributing shadow rays} */
void ertbo_last_flights(uint64_t out[2]) { out[0] = g_last_flights[0]; out[1] = g_last_flights[1]; }

/* ------------------------------------------------------------ KAT entries */
int ertbo_bsdf_eval(const ertb_scene_desc *desc, size_t n, const double *wi, const double *wo, double *out) {
    scene_t S;
    if (scene_init(&S, desc)) return 1;
    for (size_t i = 0; i < n; ++i)
        out[i] = bsdf_eval(&S, V(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]), V(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]));
    scene_free(&S);
    return 0;
}
int ertbo_bsdf_pdf(const ertb_scene_desc *desc, size_t n, const double *wi, const double *wo, double *out) {
    scene_t S;
    if (scene_init(&S, desc)) return 1;
    si_t ground;
    memset(&ground, 0, sizeof ground);
    ground.shape = SHAPE_GROUND;
    ground.p = V(1e30, 1e30, 0.0); /* (never on a central patch) */
    for (size_t i = 0; i < n; ++i)
        out[i] = surf_pdf(&S, &ground, V(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]), V(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]));
    scene_free(&S);
    return 0;
}
int ertbo_bsdf_mueller(const ertb_scene_desc *desc, size_t n, const double *wi, const double *wo, double *mueller) {
    scene_t S;
    if (scene_init(&S, desc)) return 1;
    for (size_t i = 0; i < n; ++i) {
        mueller_t M;
        bsdf_mueller(&S, NULL, 0, V(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]), V(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]), &M);
        memcpy(mueller + 16 * i, M.m, sizeof M.m);
    }
    scene_free(&S);
    return 0;
}
int ertbo_bsdf_sample(const ertb_scene_desc *desc, size_t n, const double *wi, const double *u,
                      double *wo, double *weight) {
    scene_t S;
    if (scene_init(&S, desc)) return 1;
    for (size_t i = 0; i < n; ++i) {
        v3 o;
        weight[i] = bsdf_sample(&S, V(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]), u[3 * i], u[3 * i + 1], u[3 * i + 2], &o);
        wo[3 * i] = o.x; wo[3 * i + 1] = o.y; wo[3 * i + 2] = o.z;
    }
    scene_free(&S);
    return 0;
}
int ertbo_phase_eval(const ertb_scene_desc *desc, int leaf, size_t n, const double *c, double *out) {
    scene_t S;
    if (scene_init(&S, desc)) return 1;
    if (leaf < 0 || leaf >= desc->n_phase) { scene_free(&S); return fail("bad leaf"); }
    for (size_t i = 0; i < n; ++i) { double p; leaf_eval_pdf(&S, leaf, c[i], &out[i], &p); }
    scene_free(&S);
    return 0;
}
int ertbo_phase_sample(const ertb_scene_desc *desc, int leaf, size_t n, const double *u,
                       double *ct, double *weight, double *pdf) {
    scene_t S;
    if (scene_init(&S, desc)) return 1;
    if (leaf < 0 || leaf >= desc->n_phase) { scene_free(&S); return fail("bad leaf"); }
    for (size_t i = 0; i < n; ++i) {
        v3 wl;
        leaf_sample(&S, leaf, u[2 * i], u[2 * i + 1], &wl, &weight[i], &pdf[i]);
        ct[i] = -wl.z; /* local z = wi = -propagation direction */
    }
    scene_free(&S);
    return 0;
}
int ertbo_sensor_ray(const ertb_scene_desc *desc, int sensor, size_t n, const double *fs,
                     const double *as, double *origin, double *dir, double *weight) {
    scene_t S;
    if (scene_init(&S, desc)) return 1;
    if (sensor < 0 || sensor >= desc->n_sensors) { scene_free(&S); return fail("bad sensor index"); }
    for (size_t i = 0; i < n; ++i) {
        ray_t r;
        weight[i] = sensor_sample_ray(&S, &desc->sensors[sensor], fs[2 * i], fs[2 * i + 1], as[2 * i], as[2 * i + 1], &r);
        origin[3 * i] = r.o.x; origin[3 * i + 1] = r.o.y; origin[3 * i + 2] = r.o.z;
        dir[3 * i] = r.d.x; dir[3 * i + 1] = r.d.y; dir[3 * i + 2] = r.d.z;
    }
    scene_free(&S);
    return 0;
}
int ertbo_medium_lookup(const ertb_scene_desc *desc, size_t n, const double *p, double *st, double *al) {
    scene_t S;
    if (scene_init(&S, desc)) return 1;
    for (size_t i = 0; i < n; ++i) {
        int l = layer_index(&S, V(p[3 * i], p[3 * i + 1], p[3 * i + 2]));
        st[i] = l >= 0 ? (double) desc->sigma_t_scale * desc->sigma_t[l] : 0.0;
        al[i] = l >= 0 ? desc->albedo[l] : 0.0;
    }
    scene_free(&S);
    return 0;
}

/* Known-answer entries for ERP/tests/media/test_piecewise.py (sample_interaction_real /
 * eval_transmittance_pdf_real with ray.maxt = inf). */
int ertbo_piecewise_sample(const ertb_scene_desc *desc, double half_width, size_t n, const double *o,
                           const double *d, const double *sample, const double *si_t, double *t,
                           double *tr, double *pdf) {
    scene_t S;
    if (scene_init(&S, desc)) return 1;
    if (!S.pw_cum) { scene_free(&S); return fail("not a piecewise scene"); }
    S.pp_half_width = half_width;
    for (size_t i = 0; i < n; ++i) {
        ray_t ray; ray.o = V(o[3 * i], o[3 * i + 1], o[3 * i + 2]); ray.d = V(d[3 * i], d[3 * i + 1], d[3 * i + 2]);
        ray.maxt = INFINITY;
        double mint, maxt;
        int hit = medium_aabb(&S, &ray, &mint, &maxt);
        mei_t mei = pw_sample_interaction_real(&S, &ray, si_t[i], sample[i], hit, mint, maxt, &tr[i], &pdf[i]);
        t[i] = mei.t;
    }
    scene_free(&S);
    return 0;
}

int ertbo_piecewise_eval(const ertb_scene_desc *desc, double half_width, size_t n, const double *o,
                         const double *d, const double *si_t, double *tr, double *pdf, int *escaped) {
    scene_t S;
    if (scene_init(&S, desc)) return 1;
    if (!S.pw_cum) { scene_free(&S); return fail("not a piecewise scene"); }
    S.pp_half_width = half_width;
    for (size_t i = 0; i < n; ++i) {
        ray_t ray; ray.o = V(o[3 * i], o[3 * i + 1], o[3 * i + 2]); ray.d = V(d[3 * i], d[3 * i + 1], d[3 * i + 2]);
        ray.maxt = INFINITY;
        double mint, maxt;
        int hit = medium_aabb(&S, &ray, &mint, &maxt);
        escaped[i] = pw_eval_transmittance_pdf_real(&S, &ray, si_t[i], hit, mint, maxt, &tr[i], &pdf[i]);
    }
    scene_free(&S);
    return 0;
}

int ertbo_phase_mueller(const ertb_scene_desc *desc, int leaf, size_t n, const double *wi, const double *wo,
                        double *mueller, double *pdf) {
    scene_t S;
    if (scene_init(&S, desc)) return 1;
    if (leaf < 0 || leaf >= desc->n_phase) { scene_free(&S); return fail("bad leaf"); }
    for (size_t i = 0; i < n; ++i) {
        mueller_t M;
        leaf_eval_mueller(&S, leaf, V(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]), V(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]), &M, &pdf[i]);
        memcpy(mueller + 16 * i, M.m, sizeof M.m);
    }
    scene_free(&S);
    return 0;
}

/* ------------------------------------------------------------ canopy KATs */
int ertbo_canopy_intersect(const ertb_scene_desc *desc, size_t n, const double *o, const double *d,
                           const double *tmax, double *t, double *normal, int *group) {
    scene_t S;
    if (scene_init(&S, desc)) return 1;
    if (S.canopy.n_instances == 0) { scene_free(&S); return fail("scene has no canopy"); }
    for (size_t i = 0; i < n; ++i) {
        v3 dd = vnormalize(V(d[3 * i], d[3 * i + 1], d[3 * i + 2]));
        double dv[3] = { dd.x, dd.y, dd.z };
        canopy_hit_t h = canopy_intersect(&S.canopy, o + 3 * i, dv, tmax ? tmax[i] : INFINITY);
        t[i] = h.t;
        group[i] = h.group;
        for (int k = 0; k < 3; ++k) normal[3 * i + k] = h.t < INFINITY ? h.sh_n[k] : 0.0; /* (the shading normal) */
    }
    scene_free(&S);
    return 0;
}

/* mode 0: eval (wo given), 1: pdf, 2: sample (u = sample1, sample2) -> wo, out = weight */
int ertbo_leaf_bsdf(const ertb_scene_desc *desc, int grp, int mode, size_t n, const double *wi, double *wo,
                    const double *u, double *out) {
    if (grp < 0 || grp >= desc->n_leaf_groups) return fail("invalid leaf group");
    double r = desc->leaf_groups[grp].reflectance, t = desc->leaf_groups[grp].transmittance;
    for (size_t i = 0; i < n; ++i) {
        if (mode == 0) out[i] = bilambertian_eval(r, t, wi + 3 * i, wo + 3 * i);
        else if (mode == 1) out[i] = bilambertian_pdf(r, t, wi + 3 * i, wo + 3 * i);
        else out[i] = bilambertian_sample(r, t, wi + 3 * i, u[3 * i], u[3 * i + 1], u[3 * i + 2], wo + 3 * i);
    }
    return 0;
}
