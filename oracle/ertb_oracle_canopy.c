/*
 * ertb_oracle_canopy.c -- see ertb_oracle_canopy.h.  TEST INFRASTRUCTURE ONLY: the product
 * (eradiate_b200/) never links, loads or calls this file.
 */
#include "ertb_oracle_canopy.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define PI 3.14159265358979323846

/* extent of a disk along axis k: r * sqrt(1 - n_k^2) */
static void disk_bounds(const float *dk, double lo[3], double hi[3]) {
    for (int k = 0; k < 3; ++k) {
        double n = dk[3 + k], e = (double) dk[6] * sqrt(fmax(1.0 - n * n, 0.0));
        lo[k] = (double) dk[k] - e;
        hi[k] = (double) dk[k] + e;
    }
}

static void cyl_bounds(const float *c, double lo[3], double hi[3]) {
    double ax[3] = { c[3] - c[0], c[4] - c[1], c[5] - c[2] };
    double L = sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
    for (int k = 0; k < 3; ++k) {
        double u = ax[k] / L, e = (double) c[6] * sqrt(fmax(1.0 - u * u, 0.0));
        lo[k] = fmin(c[k], c[3 + k]) - e;
        hi[k] = fmax(c[k], c[3 + k]) + e;
    }
}

static void tri_bounds(const float *q, double lo[3], double hi[3]) {
    for (int k = 0; k < 3; ++k) {
        lo[k] = fmin(q[k], fmin(q[3 + k], q[6 + k]));
        hi[k] = fmax(q[k], fmax(q[3 + k], q[6 + k]));
    }
}

/* bounds of primitive i of a group: disks [0, n_disks), then cylinders, then triangles */
static void prim_bounds(const canopy_group_t *G, int i, double lo[3], double hi[3]) {
    if (i < G->n_disks) disk_bounds(G->disks + 7 * i, lo, hi);
    else if (i < G->n_disks + G->n_cylinders) cyl_bounds(G->cylinders + 7 * (i - G->n_disks), lo, hi);
    else tri_bounds(G->triangles + 18 * (size_t) (i - G->n_disks - G->n_cylinders), lo, hi);
}

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

static int group_init(canopy_group_t *G, const ertb_leaf_group_desc *gd) {
    memset(G, 0, sizeof *G);
    if (gd->n_disks < 0 || (gd->n_disks > 0 && !gd->disks)) return 1;
    if ((gd->n_trunk_disks > 0 && !gd->trunk_disks) || (gd->n_cylinders > 0 && !gd->cylinders)) return 1;
    if (gd->n_triangles > 0 && (!gd->triangles || !gd->triangle_bsdf || !gd->mesh_bsdfs)) return 1;
    if (gd->n_disks + gd->n_trunk_disks + gd->n_cylinders + gd->n_triangles < 1) return 1;
    G->n_triangles = gd->n_triangles;
    G->triangles = gd->triangles;
    G->triangle_bsdf = gd->triangle_bsdf;
    G->mesh_bsdfs = gd->mesh_bsdfs;
    G->n_leaf_disks = gd->n_disks;
    G->n_disks = gd->n_disks + gd->n_trunk_disks;
    G->n_cylinders = gd->n_cylinders;
    G->disks = malloc(sizeof(float) * 7 * (size_t) (G->n_disks > 0 ? G->n_disks : 1));
    if (!G->disks) return 1;
    if (gd->n_disks) memcpy(G->disks, gd->disks, sizeof(float) * 7 * (size_t) gd->n_disks);
    if (gd->n_trunk_disks)
        memcpy(G->disks + 7 * (size_t) gd->n_disks, gd->trunk_disks, sizeof(float) * 7 * (size_t) gd->n_trunk_disks);
    G->cylinders = gd->cylinders;
    G->reflectance = gd->reflectance;
    G->transmittance = gd->transmittance;
    G->trunk_reflectance = gd->trunk_reflectance;
    const int n_prims = G->n_disks + G->n_cylinders + G->n_triangles;
    for (int k = 0; k < 3; ++k) { G->lo[k] = INFINITY; G->hi[k] = -INFINITY; }
    for (int i = 0; i < n_prims; ++i) {
        double lo[3], hi[3];
        prim_bounds(G, i, lo, hi);
        for (int k = 0; k < 3; ++k) { G->lo[k] = fmin(G->lo[k], lo[k]); G->hi[k] = fmax(G->hi[k], hi[k]); }
    }
    for (int k = 0; k < 3; ++k) { /* pad: flat groups, hits exactly on the faces */
        double pad = 1e-6 * fmax(1.0, G->hi[k] - G->lo[k]);
        G->lo[k] -= pad; G->hi[k] += pad;
    }
    /* ~2 disks per cell */
    double vol = (G->hi[0] - G->lo[0]) * (G->hi[1] - G->lo[1]) * (G->hi[2] - G->lo[2]);
    double s = cbrt(vol / fmax(1.0, n_prims / 2.0));
    for (int k = 0; k < 3; ++k) {
        G->res[k] = clampi((int) ceil((G->hi[k] - G->lo[k]) / s), 1, 256);
        G->cell[k] = (G->hi[k] - G->lo[k]) / G->res[k];
    }
    const size_t ncell = (size_t) G->res[0] * G->res[1] * G->res[2];
    G->cell_start = calloc(ncell + 1, sizeof(int));
    if (!G->cell_start) return 1;
    for (int pass = 0; pass < 2; ++pass) {
        for (int i = 0; i < n_prims; ++i) {
            double lo[3], hi[3];
            prim_bounds(G, i, lo, hi);
            int a[3], b[3];
            for (int k = 0; k < 3; ++k) {
                a[k] = clampi((int) floor((lo[k] - G->lo[k]) / G->cell[k]), 0, G->res[k] - 1);
                b[k] = clampi((int) floor((hi[k] - G->lo[k]) / G->cell[k]), 0, G->res[k] - 1);
            }
            for (int z = a[2]; z <= b[2]; ++z)
                for (int y = a[1]; y <= b[1]; ++y)
                    for (int x = a[0]; x <= b[0]; ++x) {
                        size_t c = ((size_t) z * G->res[1] + y) * G->res[0] + x;
                        if (pass == 0) G->cell_start[c + 1]++;
                        else G->cell_items[G->cell_start[c]++] = i;
                    }
        }
        if (pass == 0) {
            for (size_t c = 0; c < ncell; ++c) G->cell_start[c + 1] += G->cell_start[c];
            G->cell_items = malloc(sizeof(int) * (size_t) (G->cell_start[ncell] > 0 ? G->cell_start[ncell] : 1));
            if (!G->cell_items) return 1;
        } else { /* the fill pass advanced the starts by the counts: shift them back */
            for (size_t c = ncell; c > 0; --c) G->cell_start[c] = G->cell_start[c - 1];
            G->cell_start[0] = 0;
        }
    }
    return 0;
}

int canopy_init(canopy_t *C, const ertb_scene_desc *d) {
    memset(C, 0, sizeof *C);
    if (d->n_instances <= 0) return 0;
    if (d->n_leaf_groups < 1 || !d->leaf_groups || !d->instance_group || !d->instance_offset) return 1;
    C->groups = calloc((size_t) d->n_leaf_groups, sizeof(canopy_group_t));
    if (!C->groups) return 1;
    C->n_groups = d->n_leaf_groups;
    for (int g = 0; g < d->n_leaf_groups; ++g)
        if (group_init(&C->groups[g], &d->leaf_groups[g])) return 1;
    for (int i = 0; i < d->n_instances; ++i)
        if (d->instance_group[i] < 0 || d->instance_group[i] >= d->n_leaf_groups) return 1;
    C->n_instances = d->n_instances;
    C->instance_group = d->instance_group;
    C->instance_offset = d->instance_offset;
    return 0;
}

void canopy_free(canopy_t *C) {
    for (int g = 0; g < C->n_groups; ++g) { free(C->groups[g].cell_start); free(C->groups[g].cell_items); free(C->groups[g].disks); }
    free(C->groups);
    memset(C, 0, sizeof *C);
}

/* MI/src/shapes/disk.cpp:388-407 in world space: plane hit inside the radius, 0 <= t <= maxt */
static double disk_hit(const float *dk, const double o[3], const double d[3], double maxt) {
    double nx = dk[3], ny = dk[4], nz = dk[5];
    double dn = d[0] * nx + d[1] * ny + d[2] * nz;
    double t = ((dk[0] - o[0]) * nx + (dk[1] - o[1]) * ny + (dk[2] - o[2]) * nz) / dn;
    if (!(t >= 0.0 && t <= maxt)) return INFINITY;
    double px = o[0] + t * d[0] - dk[0], py = o[1] + t * d[1] - dk[1], pz = o[2] + t * d[2] - dk[2];
    double r2 = (double) dk[6] * dk[6];
    return px * px + py * py + pz * pz <= r2 ? t : INFINITY;
}

/* MI/src/shapes/cylinder.cpp:560-615: open tube of radius r around the segment p0 p1. Quadratic in the
 * plane orthogonal to the axis (math::solve_quadratic), near root first, far root if the near one is cut
 * off by the ends; no hit when the segment [0, maxt] lies entirely inside or entirely outside. */
static double cyl_hit(const float *c, const double o[3], const double d[3], double maxt) {
    double ax[3] = { c[3] - c[0], c[4] - c[1], c[5] - c[2] };
    double L = sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
    double u[3] = { ax[0] / L, ax[1] / L, ax[2] / L };
    double w[3] = { o[0] - c[0], o[1] - c[1], o[2] - c[2] };
    double du = d[0] * u[0] + d[1] * u[1] + d[2] * u[2], wu = w[0] * u[0] + w[1] * u[1] + w[2] * u[2];
    double dp[3] = { d[0] - du * u[0], d[1] - du * u[1], d[2] - du * u[2] };
    double wp[3] = { w[0] - wu * u[0], w[1] - wu * u[1], w[2] - wu * u[2] };
    double A = dp[0] * dp[0] + dp[1] * dp[1] + dp[2] * dp[2];
    double B = 2.0 * (dp[0] * wp[0] + dp[1] * wp[1] + dp[2] * wp[2]);
    double Cc = wp[0] * wp[0] + wp[1] * wp[1] + wp[2] * wp[2] - (double) c[6] * c[6];
    if (A == 0.0) return INFINITY;
    double disc = B * B - 4.0 * A * Cc;
    if (disc < 0.0) return INFINITY;
    double temp = -0.5 * (B + copysign(sqrt(disc), B));
    double x0 = temp / A, x1 = Cc / temp;
    if (temp == 0.0) x0 = x1 = 0.0;
    double near_t = fmin(x0, x1), far_t = fmax(x0, x1);
    if (!(near_t <= maxt && far_t >= 0.0)) return INFINITY;
    if (near_t < 0.0 && far_t > maxt) return INFINITY;
    double zn = wu + du * near_t, zf = wu + du * far_t;
    if (zn >= 0.0 && zn <= L && near_t >= 0.0) return near_t;
    if (zf >= 0.0 && zf <= L && far_t <= maxt) return far_t;
    return INFINITY;
}

/* MI/include/mitsuba/render/mesh.h:481-504 (Moeller & Trumbore); `uv`: the barycentric coordinates of v1, v2 */
static double tri_hit(const float *q, const double o[3], const double d[3], double maxt, double uv[2]) {
    double e1[3], e2[3], tv[3];
    for (int k = 0; k < 3; ++k) { e1[k] = (double) q[3 + k] - q[k]; e2[k] = (double) q[6 + k] - q[k]; tv[k] = o[k] - q[k]; }
    double pv[3] = { d[1] * e2[2] - d[2] * e2[1], d[2] * e2[0] - d[0] * e2[2], d[0] * e2[1] - d[1] * e2[0] };
    double inv_det = 1.0 / (e1[0] * pv[0] + e1[1] * pv[1] + e1[2] * pv[2]);
    double u = (tv[0] * pv[0] + tv[1] * pv[1] + tv[2] * pv[2]) * inv_det;
    if (!(u >= 0.0 && u <= 1.0)) return INFINITY;
    double qv[3] = { tv[1] * e1[2] - tv[2] * e1[1], tv[2] * e1[0] - tv[0] * e1[2], tv[0] * e1[1] - tv[1] * e1[0] };
    double v = (d[0] * qv[0] + d[1] * qv[1] + d[2] * qv[2]) * inv_det;
    if (!(v >= 0.0 && u + v <= 1.0)) return INFINITY;
    double t = (e2[0] * qv[0] + e2[1] * qv[1] + e2[2] * qv[2]) * inv_det;
    if (!(t >= 0.0 && t <= maxt)) return INFINITY;
    if (uv) { uv[0] = u; uv[1] = v; }
    return t;
}

static double prim_hit(const canopy_group_t *G, int i, const double o[3], const double d[3], double maxt) {
    if (i < G->n_disks) return disk_hit(G->disks + 7 * i, o, d, maxt);
    if (i < G->n_disks + G->n_cylinders) return cyl_hit(G->cylinders + 7 * (i - G->n_disks), o, d, maxt);
    return tri_hit(G->triangles + 18 * (size_t) (i - G->n_disks - G->n_cylinders), o, d, maxt, NULL);
}

/* nearest primitive of one group (ray in the group's local coordinates); returns the primitive index */
static int group_intersect(const canopy_group_t *G, const double o[3], const double d[3], double maxt, double *t_out) {
    double t0 = 0.0, t1 = maxt;
    for (int k = 0; k < 3; ++k) {
        if (d[k] != 0.0) {
            double a = (G->lo[k] - o[k]) / d[k], b = (G->hi[k] - o[k]) / d[k];
            if (a > b) { double s = a; a = b; b = s; }
            t0 = fmax(t0, a); t1 = fmin(t1, b);
        } else if (o[k] < G->lo[k] || o[k] > G->hi[k]) return -1;
    }
    if (!(t0 <= t1)) return -1;
    int c[3], step[3];
    double tnext[3], tdelta[3];
    for (int k = 0; k < 3; ++k) {
        double p = o[k] + t0 * d[k];
        c[k] = clampi((int) floor((p - G->lo[k]) / G->cell[k]), 0, G->res[k] - 1);
        if (d[k] > 0.0) {
            step[k] = 1; tdelta[k] = G->cell[k] / d[k];
            tnext[k] = (G->lo[k] + (c[k] + 1) * G->cell[k] - o[k]) / d[k];
        } else if (d[k] < 0.0) {
            step[k] = -1; tdelta[k] = -G->cell[k] / d[k];
            tnext[k] = (G->lo[k] + c[k] * G->cell[k] - o[k]) / d[k];
        } else { step[k] = 0; tdelta[k] = INFINITY; tnext[k] = INFINITY; }
    }
    int best = -1;
    double best_t = INFINITY;
    for (;;) {
        size_t ci = ((size_t) c[2] * G->res[1] + c[1]) * G->res[0] + c[0];
        double t_exit = fmin(tnext[0], fmin(tnext[1], tnext[2]));
        for (int j = G->cell_start[ci]; j < G->cell_start[ci + 1]; ++j) {
            int i = G->cell_items[j];
            double t = prim_hit(G, i, o, d, maxt);
            if (t < best_t) { best_t = t; best = i; }
        }
        if (best_t <= t_exit || t_exit > t1) break; /* a hit inside the cells visited so far is final */
        int k = tnext[0] <= tnext[1] ? (tnext[0] <= tnext[2] ? 0 : 2) : (tnext[1] <= tnext[2] ? 1 : 2);
        c[k] += step[k];
        if (c[k] < 0 || c[k] >= G->res[k]) break;
        tnext[k] += tdelta[k];
    }
    *t_out = best_t;
    return best_t < INFINITY ? best : -1;
}

canopy_hit_t canopy_intersect(const canopy_t *C, const double o[3], const double d[3], double maxt) {
    canopy_hit_t H;
    memset(&H, 0, sizeof H);
    H.t = INFINITY; H.group = -1;
    for (int i = 0; i < C->n_instances; ++i) {
        const canopy_group_t *G = &C->groups[C->instance_group[i]];
        const double *off = C->instance_offset + 3 * i;
        double ol[3] = { o[0] - off[0], o[1] - off[1], o[2] - off[2] };
        double t;
        int k = group_intersect(G, ol, d, fmin(maxt, H.t), &t);
        if (k >= 0 && t < H.t) {
            H.t = t; H.group = C->instance_group[i];
            H.kind = k < G->n_leaf_disks ? CANOPY_LEAF : CANOPY_TRUNK;
            double p[3] = { o[0] + t * d[0], o[1] + t * d[1], o[2] + t * d[2] };
            if (k >= G->n_disks + G->n_cylinders) { /* mesh.cpp:1393-1560 */
                const int ti = k - G->n_disks - G->n_cylinders;
                const float *q = G->triangles + 18 * (size_t) ti;
                double uv[2];
                tri_hit(q, ol, d, fmin(maxt, H.t), uv);
                const double b1 = uv[0], b2 = uv[1], b0 = 1.0 - b1 - b2;
                double e1[3], e2[3], sn[3], gl = 0.0, sl = 0.0;
                for (int a = 0; a < 3; ++a) {
                    H.p[a] = q[a] * b0 + q[3 + a] * b1 + q[6 + a] * b2 + off[a]; /* re-interpolated hit point */
                    e1[a] = (double) q[3 + a] - q[a]; e2[a] = (double) q[6 + a] - q[a];
                    sn[a] = q[9 + a] * b0 + q[12 + a] * b1 + q[15 + a] * b2;
                    sl += sn[a] * sn[a];
                }
                double gn[3] = { e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0] };
                for (int a = 0; a < 3; ++a) gl += gn[a] * gn[a];
                for (int a = 0; a < 3; ++a) { H.n[a] = gn[a] / sqrt(gl); H.sh_n[a] = sn[a] / sqrt(sl); }
                H.kind = CANOPY_MESH;
                H.mesh_r = G->mesh_bsdfs[2 * G->triangle_bsdf[ti]];
                H.mesh_t = G->mesh_bsdfs[2 * G->triangle_bsdf[ti] + 1];
                continue;
            }
            if (k < G->n_disks) {
                const float *dk = G->disks + 7 * k;
                double c[3] = { dk[0] + off[0], dk[1] + off[1], dk[2] + off[2] };
                double dist = (c[0] - p[0]) * dk[3] + (c[1] - p[1]) * dk[4] + (c[2] - p[2]) * dk[5];
                for (int a = 0; a < 3; ++a) { H.n[a] = dk[3 + a]; H.p[a] = p[a] + dist * dk[3 + a]; }
            } else { /* cylinder.cpp:660-700: radial normal, hit point pulled back onto the tube */
                const float *c = G->cylinders + 7 * (k - G->n_disks);
                double ax[3] = { c[3] - c[0], c[4] - c[1], c[5] - c[2] };
                double L = sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
                double w[3] = { p[0] - c[0] - off[0], p[1] - c[1] - off[1], p[2] - c[2] - off[2] };
                double wu = (w[0] * ax[0] + w[1] * ax[1] + w[2] * ax[2]) / L;
                double r[3], rn = 0.0;
                for (int a = 0; a < 3; ++a) { r[a] = w[a] - wu * ax[a] / L; rn += r[a] * r[a]; }
                rn = sqrt(rn);
                for (int a = 0; a < 3; ++a) { H.n[a] = r[a] / rn; H.p[a] = p[a] + ((double) c[6] - rn) * H.n[a]; }
            }
        }
    }
    if (H.kind != CANOPY_MESH)
        for (int a = 0; a < 3; ++a) H.sh_n[a] = H.n[a];
    return H;
}

/* ------------------------------------------------------------------ bilambertian.cpp */
static void lobe_weights(double r, double t, double *rw, double *tw) {
    *rw = r / (r + t);
    *tw = 1.0 - *rw;
    if (isnan(*rw)) *rw = 0.0; /* r = t = 0 (:84-88) */
    if (isnan(*tw)) *tw = 0.0;
}

double bilambertian_eval(double r, double t, const double wi[3], const double wo[3]) { /* :124-159 */
    double cti = wi[2], cto = wo[2];
    int same = (cti > 0.0) == (cto > 0.0);
    return (same ? r : t) * (1.0 / PI) * fabs(cto);
}

double bilambertian_pdf(double r, double t, const double wi[3], const double wo[3]) { /* :161-204 */
    double rw, tw;
    lobe_weights(r, t, &rw, &tw);
    int same = (wi[2] > 0.0) == (wo[2] > 0.0);
    return fabs(wo[2]) * (1.0 / PI) * (same ? rw : tw);
}

void ertbo_square_to_cosine_hemisphere(double u, double v, double *o);

double bilambertian_sample(double r, double t, const double wi[3], double sample1, double u1, double u2,
                           double wo[3]) { /* :60-122 */
    double w[3], rw, tw;
    ertbo_square_to_cosine_hemisphere(u1, u2, w);
    lobe_weights(r, t, &rw, &tw);
    int sel_r = sample1 < rw;
    double value = sel_r ? r / rw : t / tw;
    double pdf = w[2] * (1.0 / PI) * (sel_r ? rw : tw);
    if (!(wi[2] > 0.0)) w[2] = -w[2]; /* incoming from "behind" */
    if (!sel_r) w[2] = -w[2];         /* transmission */
    wo[0] = w[0]; wo[1] = w[1]; wo[2] = w[2];
    return pdf > 0.0 ? value : 0.0;
}
