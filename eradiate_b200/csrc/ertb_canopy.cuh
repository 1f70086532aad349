// ertb_canopy.cuh -- plane-parallel scenes with an explicit 3D canopy (BASELINE C4):
// disk leaves (MI/src/shapes/disk.cpp:388-407) in instanced shape groups
// (src/eradiate/scenes/biosphere/_core.py:266-296) with the bilambertian leaf BSDF
// (ERP/bsdfs/bilambertian.cpp:60-215) under a 1D atmosphere, seen by the distant sensors or a
// perspective camera (MI/src/sensors/perspective.cpp:200-236).
//
// Same estimator as the 1D kernels (volpath.cpp:93-572 / piecewise_volpath.cpp:91-527: free
// flights bounded by the next surface, next-event estimation towards the sun at every real
// event, Russian roulette), different execution model, because the work is different:
//   * a path carries a world-space position (float64: kilometres of atmosphere, centimetres of
//     leaf) and direction; the atmosphere is still only a function of altitude, so the medium
//     walk is the 1D one (global-majorant delta/ratio tracking, or the analytic piecewise flight);
//   * every segment that crosses the canopy's bounding box is intersected with a two-level BVH
//     (instances -> disks) held in global memory (L2 resident: 32 B nodes, 32 B disks), in
//     float32 coordinates relative to the canopy centre;
//   * one path per lane, persistent warps, lanes refilled from the chunk queue as soon as their
//     path ends. A lane is a small state machine (segment set-up -> BVH walk -> free flight and
//     event -> shadow-ray BVH walk -> shading) and the BVH walk is RESUMABLE: every trip of the
//     main loop advances all walking lanes by at most ERTB_TRACE_STEPS nodes, then runs the other
//     stages for the lanes that are ready. Rays that need 10 node visits therefore do not wait
//     for the one that needs 300 (lock-step walks ran with 4 of 32 lanes active, ncu r01e).
//
// Tried in round 2 and withdrawn (bit-identical films, profiles/r02g_canopy_scheduler_experiment.md): a warp
// scheduler that runs ONE kind of work per trip -- BVH slices while at least half of the tracing lanes still
// trace, else steps of resumable medium walks, else the short stages.  It executed 25 % more warp
// instructions at the same 10 active lanes (box tests fell from 21 to 16 lanes): 131 against 167 Mpaths/s on
// C4.  The lanes are lost INSIDE the traversal step (leaf tests at 1.2 lanes, stack and instance
// bookkeeping at 3-5), not between the stages.
#pragma once

#include "ertb_kernel.cuh"
#include "ertb_kernel_pool.cuh" // film_flush_warp
#include "ertb_piecewise.cuh"

// 4 CTAs of 256 threads per SM = 64 registers, 32 resident warps.  The walk waits on dependent node fetches (35 % of the
// stall samples are long-scoreboard, profiles/r02t): resident warps are what hides them, and the state the register
// cap pushes into local memory belongs to the event code, not to the traversal loop.  Measured (C4 mdistant /
// C4 perspective / abstract trees, Mpaths/s): 128 x 4 (128 reg.) 150 / 192 / 696; 128 x 5 (96) 157 / 210 / 662;
// 128 x 6 (80) 161 / 217 / 722; 128 x 7 (72) 173 / 235 / 730; 128 x 8 (64) 172 / 237 / 751; 256 x 4 (64) 173 / 240 / 753;
// 128 x 10 (48) 168 / 232 / 740; 256 x 5 (48) 168 / 231 / 748; 256 x 6 (40) 160 / 204 / 711.
#ifndef ERTB_CANOPY_BLOCK
#define ERTB_CANOPY_BLOCK 256
#endif
#ifndef ERTB_CANOPY_MINB
#define ERTB_CANOPY_MINB 4
#endif
#ifndef ERTB_BVH_LEAF
#define ERTB_BVH_LEAF 1 // primitives per BVH leaf: disk tests run with ~1.5 active lanes, box tests with many
#endif

// ----------------------------------------------------------------------------
// BVH traversal
// ----------------------------------------------------------------------------
// slab test; returns the entry distance (INFINITY: missed)
__device__ __forceinline__ float aabb_entry(float4 lo, float4 hi, f3 o, f3 inv, float tmax) {
    float a = (lo.x - o.x) * inv.x, b = (hi.x - o.x) * inv.x;
    float t0 = fminf(a, b), t1 = fmaxf(a, b);
    a = (lo.y - o.y) * inv.y; b = (hi.y - o.y) * inv.y;
    t0 = fmaxf(t0, fminf(a, b)); t1 = fminf(t1, fmaxf(a, b));
    a = (lo.z - o.z) * inv.z; b = (hi.z - o.z) * inv.z;
    t0 = fmaxf(t0, fminf(a, b)); t1 = fminf(t1, fmaxf(a, b));
    t0 = fmaxf(t0, 0.f); t1 = fminf(t1, tmax);
    return t0 <= t1 ? t0 : INFINITY;
}

struct CanopyHit {
    float t;    // distance along the ray (canopy-local units = metres); INFINITY: none
    int inst;   // instance (index into C.inst)
    int disk;   // disk (index into C.disks / 2)
};

// disk.cpp:388-407: plane hit with 0 <= t <= tmax inside the radius
__device__ __forceinline__ float disk_hit(float4 c, float4 n, f3 o, f3 d, float tmax) {
    float dn = d.x * n.x + d.y * n.y + d.z * n.z;
    float t = __fdividef((c.x - o.x) * n.x + (c.y - o.y) * n.y + (c.z - o.z) * n.z, dn);
    if (!(t >= 0.f && t <= tmax)) return INFINITY;
    float px = fmaf(t, d.x, o.x) - c.x, py = fmaf(t, d.y, o.y) - c.y, pz = fmaf(t, d.z, o.z) - c.z;
    return px * px + py * py + pz * pz <= c.w * c.w ? t : INFINITY;
}

#ifndef ERTB_TRACE_STEPS
#define ERTB_TRACE_STEPS 32 // VOTE form: node visits per sub-slice of the BVH walk
#endif
#ifndef ERTB_TRACE_STEPS_FIXED
#define ERTB_TRACE_STEPS_FIXED 48 // fixed form (large groups): node visits per slice; at 64 registers 24 / 32 / 48 / 64
#endif                            // visits give C4 160 / 173 / 187 / 187 (mdistant), 224 / 240 / 253 / 246 (camera)
#ifndef ERTB_TRACE_SUBS
#define ERTB_TRACE_SUBS 2 // VOTE form: sub-slices per trip of the main loop, at most
#endif
#ifndef ERTB_TRACE_KEEP_NUM
#define ERTB_TRACE_KEEP_NUM 1 // VOTE form: the next sub-slice runs while NUM / DEN of the lanes that entered still walk
#define ERTB_TRACE_KEEP_DEN 2
#endif
#ifndef ERTB_VOTE_MAX_PRIMS
#define ERTB_VOTE_MAX_PRIMS 8192 // largest group (primitives) for which the host picks the VOTE form
#endif
#define ERTB_TRACE_STACK 40

// A BVH walk that can be suspended: everything but the stack (a per-lane local array).
// cylinder.cpp:560-615: open tube of radius c.w around the segment (c.xyz, c.xyz + ax.xyz); quadratic in the
// plane orthogonal to the axis, near root first, the far one if the near one is cut off by the ends
__device__ __forceinline__ float cylinder_hit(float4 c, float4 ax, f3 o, f3 d, float tmax) {
    const float L2 = ax.x * ax.x + ax.y * ax.y + ax.z * ax.z, iL = rsqrtf(L2), L = L2 * iL;
    const f3 u = mk3(ax.x * iL, ax.y * iL, ax.z * iL), w = mk3(o.x - c.x, o.y - c.y, o.z - c.z);
    const float du = dot3(d, u), wu = dot3(w, u);
    const f3 dp = fma3(u, -du, d), wp = fma3(u, -wu, w);
    const float A = dot3(dp, dp), B = 2.f * dot3(dp, wp), Cc = dot3(wp, wp) - c.w * c.w;
    const float disc = B * B - 4.f * A * Cc;
    if (!(A > 0.f) || disc < 0.f) return INFINITY;
    const float temp = -0.5f * (B + copysignf(sqrtf(disc), B));
    float x0 = temp / A, x1 = Cc / temp;
    if (temp == 0.f) x0 = x1 = 0.f;
    const float tn = fminf(x0, x1), tf = fmaxf(x0, x1);
    if (!(tn <= tmax && tf >= 0.f) || (tn < 0.f && tf > tmax)) return INFINITY;
    const float zn = fmaf(du, tn, wu), zf = fmaf(du, tf, wu);
    if (zn >= 0.f && zn <= L && tn >= 0.f) return tn;
    if (zf >= 0.f && zf <= L && tf <= tmax) return tf;
    return INFINITY;
}

// mesh.h:481-504 (Moeller-Trumbore): a = (v0, .), b = (e1, .), e2 from the triangle table
__device__ __forceinline__ float triangle_hit(float4 a, float4 b, float4 c, f3 o, f3 d, float tmax) {
    const f3 e1 = mk3(b.x, b.y, b.z), e2 = mk3(c.x, c.y, c.z);
    const f3 pvec = cross3(d, e2);
    const float inv_det = 1.f / dot3(e1, pvec);
    const f3 tvec = mk3(o.x - a.x, o.y - a.y, o.z - a.z);
    const float u = dot3(tvec, pvec) * inv_det;
    const f3 qvec = cross3(tvec, e1);
    const float v = dot3(d, qvec) * inv_det, t = dot3(e2, qvec) * inv_det;
    return (u >= 0.f && u <= 1.f && v >= 0.f && u + v <= 1.f && t >= 0.f && t <= tmax) ? t : INFINITY;
}

// `MESH` = false: the instances the host launches for canopies without triangles over a ground BSDF of SURVEY 8a
// (disc / cylinder scenes such as C4 keep the instruction stream they had before meshes and the later BSDF plugins
// existed: C4 lost 17 % with the triangle code and another 5 % with the measured BSDFs compiled in)
template <bool MESH>
__device__ __forceinline__ float prim_hit(const ErtbCanopy &C, float4 a, float4 b, f3 o, f3 d, float tmax) {
    const int kind = __float_as_int(b.w);
    if (MESH && kind == 3) return triangle_hit(a, b, __ldg(C.tris + 4 * __float_as_int(a.w)), o, d, tmax);
    return kind == 2 ? cylinder_hit(a, b, o, d, tmax) : disk_hit(a, b, o, d, tmax);
}

struct TraceState {
    f3 o;        // ray origin, canopy-local
    float tmax;
    int node, ii, sp; // current node, its instance (< 0: top level), stack height
    int skip_inst, skip_disk; // the leaf the ray starts on
    CanopyHit H;
};

__device__ __forceinline__ void trace_begin(TraceState &T, f3 o, float tmax, int skip_inst, int skip_disk) {
    T.o = o; T.tmax = tmax; T.node = 0; T.ii = -1; T.sp = 0;
    T.skip_inst = skip_inst; T.skip_disk = skip_disk;
    T.H.t = INFINITY; T.H.inst = -1; T.H.disk = -1;
}

// Advance the walk by at most `steps` nodes; returns true when it is over (T.H holds the nearest leaf,
// or, with `any`, the first leaf found). ONE loop walks both levels: a stack entry is (node, instance),
// instance < 0 meaning a node of the top-level tree, so the lanes of a warp always execute the same
// few instructions (fetch a node, test its two child boxes, push) whichever level each of them is in.
// VOTE: every lane of the warp calls (`active`: the lanes that walk); sub-slices of `steps` visits follow one another
// while enough of the lanes that entered still walk (see the kernel's BVH stage).
template <bool MESH, bool VOTE = false>
__device__ __forceinline__ bool trace_run(const ErtbCanopy &C, TraceState &T, int2 *stack, f3 d, bool any, int steps,
                                          bool active = true) {
    const float big = 1e30f;
    const f3 inv = mk3(fabsf(d.x) > 1e-30f ? 1.f / d.x : copysignf(big, d.x), fabsf(d.y) > 1e-30f ? 1.f / d.y : copysignf(big, d.y),
                       fabsf(d.z) > 1e-30f ? 1.f / d.z : copysignf(big, d.z));
    int node = T.node, ii = T.ii, sp = T.sp;
    f3 ol = T.o;
    if (ii >= 0) { const float4 in = __ldg(C.inst + ii); ol = mk3(T.o.x - in.x, T.o.y - in.y, T.o.z - in.z); }
    bool done = !active;
    const int n_enter = VOTE ? __popc(__ballot_sync(0xffffffffu, active)) : 0;
    for (int step = 0; step < (VOTE ? steps * ERTB_TRACE_SUBS : steps); ++step) {
        if (VOTE && step > 0 && step % steps == 0 &&
            __popc(__ballot_sync(0xffffffffu, !done)) * ERTB_TRACE_KEEP_DEN < n_enter * ERTB_TRACE_KEEP_NUM) break;
        if (VOTE && done) continue;
        const float4 *q = reinterpret_cast<const float4 *>((ii < 0 ? C.tlas : C.blas) + node);
        const float4 l0 = __ldg(q), h0 = __ldg(q + 1), l1 = __ldg(q + 2), h1 = __ldg(q + 3);
        const float lim = fminf(T.tmax, T.H.t);
        float e[2] = { aabb_entry(l0, h0, ol, inv, lim), aabb_entry(l1, h1, ol, inv, lim) };
        const int c[2] = { __float_as_int(l0.w), __float_as_int(l1.w) }, n[2] = { __float_as_int(h0.w), __float_as_int(h1.w) };
        int next = -1, next_ii = ii;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            if (!(e[s] < INFINITY) || n[s] == 0) continue;
            if (ii < 0) { // instances: their group's root goes on the stack
                for (int k = c[s]; k < c[s] + n[s]; ++k)
                    if (sp < ERTB_TRACE_STACK) stack[sp++] = make_int2(__ldg(C.blas_root + __float_as_int(__ldg(C.inst + k).w)), k);
            } else { // disks: intersected on the spot
                for (int k = c[s]; k < c[s] + n[s]; ++k) {
                    if (k == T.skip_disk && ii == T.skip_inst) continue;
                    float t = prim_hit<MESH>(C, __ldg(C.disks + 2 * k), __ldg(C.disks + 2 * k + 1), ol, d, fminf(T.tmax, T.H.t));
                    if (t < T.H.t) { T.H.t = t; T.H.inst = ii; T.H.disk = k; }
                }
            }
            e[s] = INFINITY;
        }
        if (any && T.H.inst >= 0) { done = true; if (VOTE) continue; else break; }
        // inner children: the nearer one next, the other on the stack
        if (e[0] < INFINITY && e[1] < INFINITY) {
            const bool first0 = e[0] <= e[1];
            if (sp < ERTB_TRACE_STACK) stack[sp++] = make_int2(first0 ? c[1] : c[0], ii);
            next = first0 ? c[0] : c[1];
        } else if (e[0] < INFINITY) next = c[0];
        else if (e[1] < INFINITY) next = c[1];
        else {
            if (sp == 0) { done = true; if (VOTE) continue; else break; }
            const int2 top = stack[--sp];
            next = top.x; next_ii = top.y;
        }
        if (next_ii != ii) {
            ii = next_ii;
            if (ii < 0) ol = T.o;
            else { const float4 in = __ldg(C.inst + ii); ol = mk3(T.o.x - in.x, T.o.y - in.y, T.o.z - in.z); }
        }
        node = next;
    }
    T.node = node; T.ii = ii; T.sp = sp;
    return done && active;
}

// whole walk in one call (KAT entry point)
__device__ __noinline__ CanopyHit canopy_trace(const ErtbCanopy &C, f3 o, f3 d, float tmax, bool any, int skip_inst, int skip_disk) {
    TraceState T;
    int2 stack[ERTB_TRACE_STACK];
    trace_begin(T, o, tmax, skip_inst, skip_disk);
    while (!trace_run<true>(C, T, stack, d, any, ERTB_TRACE_STEPS)) { }
    return T.H;
}

// Clip the segment p + t d, 0 <= t <= tmax, to the canopy's bounding box (float64, world space).
__device__ __forceinline__ bool canopy_clip(const ErtbCanopy &C, const double p[3], f3 d, double tmax, double &t0, double &t1) {
    t0 = 0.0; t1 = tmax;
    const float dd[3] = { d.x, d.y, d.z };
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (dd[k] != 0.f) {
            double inv = 1.0 / (double) dd[k];
            double a = (C.lo[k] - p[k]) * inv, b = (C.hi[k] - p[k]) * inv;
            t0 = fmax(t0, fmin(a, b)); t1 = fmin(t1, fmax(a, b));
        } else if (p[k] < C.lo[k] || p[k] > C.hi[k]) return false;
    }
    return t0 <= t1;
}

// canopy-local float origin of the point p + t0 d
__device__ __forceinline__ f3 canopy_local(const ErtbCanopy &C, const double p[3], f3 d, double t0) {
    return mk3((float) (p[0] + t0 * (double) d.x - C.origin[0]), (float) (p[1] + t0 * (double) d.y - C.origin[1]),
               (float) (p[2] + t0 * (double) d.z - C.origin[2]));
}

// surface (shading) normal and kind (0 leaf, 1 trunk cap, 2 trunk tube, 3 mesh triangle) of the primitive hit at
// world point q; `mat`: row of a triangle's BSDF in the mesh table
template <bool MESH = true>
__device__ __forceinline__ f3 canopy_normal(const ErtbCanopy &C, const double q[3], const CanopyHit &H, int &kind, int &mat) {
    const float4 a = __ldg(C.disks + 2 * H.disk), b = __ldg(C.disks + 2 * H.disk + 1);
    kind = __float_as_int(b.w);
    mat = 0;
    if (kind < 2) return mk3(b.x, b.y, b.z);
    const float4 in = __ldg(C.inst + H.inst);
    const f3 w = mk3((float) (q[0] - C.origin[0]) - in.x - a.x, (float) (q[1] - C.origin[1]) - in.y - a.y,
                     (float) (q[2] - C.origin[2]) - in.z - a.z);
    if (MESH && kind == 3) { // mesh.cpp:1500-1535: barycentric interpolation of the vertex normals, renormalised
        const float4 *t = C.tris + 4 * __float_as_int(a.w);
        const float4 c = __ldg(t), n0 = __ldg(t + 1), n1 = __ldg(t + 2), n2 = __ldg(t + 3);
        mat = __float_as_int(c.w);
        const f3 e1 = mk3(b.x, b.y, b.z), e2 = mk3(c.x, c.y, c.z);
        const float d11 = dot3(e1, e1), d12 = dot3(e1, e2), d22 = dot3(e2, e2), w1 = dot3(w, e1), w2 = dot3(w, e2);
        const float inv = 1.f / (d11 * d22 - d12 * d12);
        const float b1 = fminf(fmaxf((d22 * w1 - d12 * w2) * inv, 0.f), 1.f), b2 = fminf(fmaxf((d11 * w2 - d12 * w1) * inv, 0.f), 1.f);
        const float b0 = 1.f - b1 - b2;
        return normalize3(mk3(fmaf(n2.x, b2, fmaf(n1.x, b1, n0.x * b0)), fmaf(n2.y, b2, fmaf(n1.y, b1, n0.y * b0)),
                              fmaf(n2.z, b2, fmaf(n1.z, b1, n0.z * b0))));
    }
    const float iL = rsqrtf(b.x * b.x + b.y * b.y + b.z * b.z);
    const f3 u = mk3(b.x * iL, b.y * iL, b.z * iL);
    return normalize3(fma3(u, -dot3(w, u), w)); // radial, pointing outwards
}

// nearest leaf along a world-space segment; returns the distance from p (INFINITY: none)
__device__ __forceinline__ double canopy_nearest(const ErtbCanopy &C, const double p[3], f3 d, double tmax,
                                                 int skip_inst, int skip_disk, CanopyHit &H) {
    H.t = INFINITY; H.inst = -1; H.disk = -1;
    double t0, t1;
    if (C.n_instances == 0 || !canopy_clip(C, p, d, tmax, t0, t1)) return INFINITY;
    H = canopy_trace(C, canopy_local(C, p, d, t0), d, (float) (t1 - t0), false, skip_inst, skip_disk);
    return H.inst >= 0 ? t0 + (double) H.t : INFINITY;
}

// ----------------------------------------------------------------------------
// bilambertian.cpp (local frame: z = leaf normal; ci / co = cosines of wi / wo with it)
// ----------------------------------------------------------------------------
// value * |cos(theta_o)| (:124-159)
__device__ __forceinline__ float bilambertian_eval(float r, float t, float ci, float co) {
    return ((ci > 0.f) == (co > 0.f) ? r : t) * ERTB_INV_PI * fabsf(co);
}
// (:60-122) returns weight = value / pdf and the sampled local direction
__device__ __forceinline__ float bilambertian_sample(float r, float t, float ci, float s1, float u1, float u2, f3 &wo) {
    wo = cosine_hemisphere(u1, u2);
    float rw = r + t > 0.f ? __fdividef(r, r + t) : 0.f;
    float tw = r + t > 0.f ? 1.f - rw : 0.f;
    const bool sel_r = s1 < rw;
    float value = sel_r ? __fdividef(r, rw) : (tw > 0.f ? __fdividef(t, tw) : 0.f);
    float pdf = wo.z * ERTB_INV_PI * (sel_r ? rw : tw);
    if (!(ci > 0.f)) wo.z = -wo.z;
    if (!sel_r) wo.z = -wo.z;
    return pdf > 0.f ? value : 0.f;
}

// land BSDF value (without the cosine) for an explicit (type, parameter array): the patch of a
// CentralPatchSurface; same device functions as bsdf_f()
__device__ __forceinline__ float land_bsdf(int type, const float *prm, float ci, float co, float cdphi) {
    switch (type) {
        case ERTB_BSDF_DIFFUSE: return prm[0] * ERTB_INV_PI;
        case ERTB_BSDF_RPV: return rpv_eval(prm, ci, co, cdphi);
        case ERTB_BSDF_RTLS: return rtls_eval(prm, ci, co, cdphi);
        case ERTB_BSDF_HAPKE: return hapke_eval(prm, ci, co, cdphi);
        default: return 0.f;
    }
}

// ----------------------------------------------------------------------------
// primary rays: distant sensors aim at a target point T (mdistant.cpp:192-242,
// hdistant.cpp:232-275, distantflux.cpp:148-195), the camera starts at its pinhole
// ----------------------------------------------------------------------------
// returns false when the sample contributes L = 0. Distant sensors: `p` is a point ON the ray (the
// target), the caller moves it back to where the ray enters the scene.
__device__ __forceinline__ bool canopy_primary(const ErtbParams &P, unsigned pix, Pcg32 &rng, double p[3], f3 &d,
                                               float &wray, float &maxt) {
    const ErtbSensor &S = P.sensor;
    unsigned px = pix % (unsigned) S.width, py = pix / (unsigned) S.width;
    float fx = __fdividef((float) px + pcg_float(rng), (float) S.width);
    float fy = __fdividef((float) py + pcg_float(rng), (float) S.height);
    float ax = pcg_float(rng), ay = pcg_float(rng);
    wray = 1.f;
    maxt = INFINITY;
    if (S.type == ERTB_SENSOR_PERSPECTIVE) {
        f3 dc = normalize3(mk3((1.f - 2.f * fx) * S.tan_half_fov, (1.f - 2.f * fy) * S.tan_half_fov / S.aspect, 1.f));
        const float *M = S.to_world;
        d = normalize3(mk3(M[0] * dc.x + M[1] * dc.y + M[2] * dc.z, M[3] * dc.x + M[4] * dc.y + M[5] * dc.z,
                           M[6] * dc.x + M[7] * dc.y + M[8] * dc.z));
        float inv_z = 1.f / dc.z;
        float near_t = S.near_clip * inv_z;
        p[0] = S.cam_origin[0] + (double) (near_t * d.x);
        p[1] = S.cam_origin[1] + (double) (near_t * d.y);
        p[2] = S.cam_origin[2] + (double) (near_t * d.z);
        maxt = (S.far_clip - S.near_clip) * inv_z;
        return true;
    }
    if (S.type == ERTB_SENSOR_MRADIANCEMETER) { // mradiancemeter.cpp:147-172: explicit rays
        const float4 *t4 = reinterpret_cast<const float4 *>(S.table) + 2u * pix;
        float4 a = __ldg(t4), c4 = __ldg(t4 + 1);
        d = mk3(a.w, c4.x, c4.y);
        p[0] = S.origins[3u * pix]; p[1] = S.origins[3u * pix + 1]; p[2] = S.origins[3u * pix + 2];
        return c4.z != 0.f;
    }
    f3 fs = mk3(1.f, 0.f, 0.f), ft = mk3(0.f, 1.f, 0.f);
    if (S.type == ERTB_SENSOR_MPDISTANT) { // mpdistant.cpp:214-262: the film sample picks the target point
        const float *M = S.to_world;
        d = normalize3(mk3(M[2], M[5], M[8]));
        fs = mk3(M[0], M[3], M[6]);
        ft = mk3(M[1], M[4], M[7]);
        ax = fx; ay = fy;
    } else if (S.type == ERTB_SENSOR_MDISTANT) {
        const float4 *t4 = reinterpret_cast<const float4 *>(S.table) + 2u * pix;
        float4 a = __ldg(t4), c4 = __ldg(t4 + 1);
        d = mk3(a.w, c4.x, c4.y);
        onb(d, fs, ft);
    } else {
        f3 hv = uniform_hemisphere(fx, fy);
        const float *M = S.to_world;
        d = mk3(-(M[0] * hv.x + M[1] * hv.y + M[2] * hv.z), -(M[3] * hv.x + M[4] * hv.y + M[5] * hv.z),
                -(M[6] * hv.x + M[7] * hv.y + M[8] * hv.z));
        fs = mk3(M[0], M[3], M[6]);
        ft = mk3(M[1], M[4], M[7]);
        if (S.type == ERTB_SENSOR_DISTANTFLUX) wray = hv.z * S.flux_norm;
    }
    if (S.target_type == ERTB_TARGET_POINT) {
        p[0] = S.target[0]; p[1] = S.target[1]; p[2] = S.target[2];
    } else if (S.target_type == ERTB_TARGET_NONE) {
        float ox, oy;
        disk_concentric(ax, ay, ox, oy);
        p[0] = S.bs_center[0] + ((double) fs.x * ox + (double) ft.x * oy) * S.bs_radius;
        p[1] = S.bs_center[1] + ((double) fs.y * ox + (double) ft.y * oy) * S.bs_radius;
        p[2] = S.bs_center[2] + ((double) fs.z * ox + (double) ft.z * oy) * S.bs_radius;
    } else {
        float lx, ly;
        if (S.target_type == ERTB_TARGET_RECTANGLE) { lx = fmaf(2.f, ax, -1.f); ly = fmaf(2.f, ay, -1.f); }
        else disk_concentric(ax, ay, lx, ly);
        const double *T = S.target_to_world;
        p[0] = T[0] * lx + T[1] * ly + T[3];
        p[1] = T[4] * lx + T[5] * ly + T[7];
        p[2] = T[8] * lx + T[9] * ly + T[11];
    }
    return d.z < 0.f; // distant sensors look down on a plane-parallel scene
}

// Atmospheric transmittance from altitude h towards the sun, sun.z > 0 (volpath.cpp:400-554 ratio
// tracking / piecewise_volpath.cpp:392-527 exact); occlusion by leaves is the shadow-ray BVH walk.
template <bool PW, bool STATS>
__device__ __forceinline__ float sun_medium_transmittance(const ErtbParams &P, const float *tb, float h, f3 sun,
                                                          Pcg32 &rng, unsigned &st_nee) {
    if (PW) {
        if (STATS) st_nee++;
        return pw_transmittance_up(P, tb, h, sun.z);
    }
    float tr = 1.f, t = 0.f;
    const float t_top = __fdividef(fmaxf(P.H - h, 0.f), sun.z);
    for (;;) {
        if (STATS) st_nee++;
        t += -__logf(1.f - pcg_float(rng)) * P.inv_majorant;
        if (!(t < t_top)) break;
        tr *= 1.f - tb[P.off_preal + layer_of(P, fmaf(t, sun.z, h))];
        if (tr == 0.f) break;
    }
    return tr;
}

// astroobject.cpp:141-175: a direction uniform in the cone of the solar disc about `sun` (weight: the irradiance)
__device__ __forceinline__ f3 astro_sample(const ErtbParams &P, f3 sun, Pcg32 &rng) {
    float ox, oy;
    disk_concentric(pcg_float(rng), pcg_float(rng), ox, oy);
    const float t = P.astro_omc * fmaf(ox, ox, oy * oy); // 1 - cos(theta) of the sample
    const float sc = sqrtf(P.astro_omc * (2.f - t));     // sin(theta) / sqrt(ox^2 + oy^2)
    f3 fs, ft;
    onb(sun, fs, ft);
    return normalize3(fma3(fs, ox * sc, fma3(ft, oy * sc, scale3(sun, 1.f - t))));
}
// astroobject.cpp:111-124: radiance of the disc seen along `d` (0 for a directional emitter or with hide_emitters)
__device__ __forceinline__ float astro_direct(const ErtbParams &P, f3 d, f3 sun) {
    const f3 cx = cross3(d, sun);
    return (P.astro_radiance > 0.f && dot3(d, sun) > 0.f && dot3(cx, cx) < P.astro_sin2) ? P.astro_radiance : 0.f;
}

enum : int { CEV_NONE = 0, CEV_COLLISION = 1, CEV_GROUND = 2, CEV_LEAF = 3, CEV_END = 4, CEV_CLIP = 5 };
// lane state machine
enum : int { LP_NEW = 0, LP_SEGMENT = 1, LP_TRACE = 2, LP_FLIGHT = 3, LP_SHADE = 4 };

// VOTE: the form of the BVH stage for canopies of small groups (see there).
template <bool STATS, bool PW, bool MESH = false, bool VOTE = false>
__global__ void __launch_bounds__(ERTB_CANOPY_BLOCK, ERTB_CANOPY_MINB) ertb_canopy_kernel(const ErtbParams P) {
    extern __shared__ __align__(16) float tb[]; // table blob
    __shared__ __align__(8) unsigned long long mbar;
    if (P.blob_bytes > 0) tma_stage(tb, P.blob, (unsigned) P.blob_bytes, &mbar);

    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const f3 sun = mk3(P.sun[0], P.sun[1], P.sun[2]);
    const ErtbCanopy &C = P.canopy;
    const double z_ground = P.Rd;
    const bool medium = P.has_medium != 0;

    // per-lane path state
    int phase = LP_NEW;
    bool in_medium = false, first_segment = false, shadow = false;
    double p[3] = { 0.0, 0.0, 0.0 };
    f3 d = mk3(0.f, 0.f, -1.f);
    float thr = 0.f, res = 0.f, wray = 1.f, maxt = INFINITY, nee = 0.f;
    // next-event direction of the last event: the sun, or (general instances, astroobject.cpp:141-175) a point of the
    // solar disc drawn at the event -- it has to outlive the shadow-ray walk
    f3 sun_e = sun;
#define SUN_E (MESH ? sun_e : sun)
    unsigned depth = 0, pix = 0;
    int on_inst = -1, on_disk = -1; // the leaf the current ray starts on
    // the segment being resolved: distance / kind of the next analytic surface, start of the BVH walk
    double t_geo = 0.0, t_clip0 = 0.0;
    int ev_geo = CEV_END;
    TraceState T;
    int2 stack[ERTB_TRACE_STACK];
    trace_begin(T, mk3(0.f, 0.f, 0.f), 0.f, -1, -1);
    Pcg32 rng;
    rng.state = 0; rng.inc = 1;

    double acc_wl = 0.0, acc_l = 0.0, acc_l2 = 0.0;
    unsigned acc_pix = 0xffffffffu;
    unsigned long long cur_next = 0, cur_end = 0;
    unsigned cur_pix = 0;
    bool exhausted = false;
    unsigned st_main = 0, st_nee = 0, st_scatter = 0, st_surface = 0, st_paths = 0;

#define CANOPY_FINISH()                                 \
    do {                                                \
        acc_wl += (double) (wray * res);                \
        acc_l += (double) res;                          \
        acc_l2 += (double) res * (double) res;          \
        phase = LP_NEW;                                 \
    } while (0)

    for (;;) {
        // ================= regeneration: finished lanes pop the next samples (warp-aggregated) =================
        {
            unsigned pending = exhausted ? 0u : __ballot_sync(0xffffffffu, phase == LP_NEW);
            bool got = false;
            unsigned long long my_sample = 0;
            unsigned my_pix = 0;
            while (pending && !exhausted) {
                if (cur_next >= cur_end) {
                    unsigned long long c = 0;
                    if (lane == 0) c = atomicAdd(P.work_counter, 1ULL);
                    c = __shfl_sync(0xffffffffu, c, 0);
                    if (c >= P.n_chunks) { exhausted = true; break; }
                    cur_pix = (unsigned) (c % P.n_pixels);
                    unsigned long long k = c / P.n_pixels;
                    cur_next = k * P.chunk;
                    cur_end = min(cur_next + (unsigned long long) P.chunk, P.spp);
                }
                unsigned long long avail = cur_end - cur_next;
                unsigned take = (unsigned) min((unsigned long long) __popc(pending), avail);
                bool mine = (pending >> lane) & 1u;
                unsigned rank = __popc(pending & lt_mask);
                if (mine && rank < take && !got) { got = true; my_sample = cur_next + rank; my_pix = cur_pix; }
                cur_next += take;
                pending &= ~__ballot_sync(0xffffffffu, mine && rank < take);
            }
            if (got) {
                pix = my_pix;
                if (pix != acc_pix) {
                    if (acc_pix != 0xffffffffu) film_flush_lane<false>(P, acc_pix, acc_wl, acc_l, acc_l2, 0.0, 0.0, 0.0);
                    acc_wl = acc_l = acc_l2 = 0.0;
                    acc_pix = pix;
                }
                pcg_seed(rng, P.seed, ((unsigned long long) pix << 40) + (P.sample_offset + my_sample));
                if (STATS) st_paths++;
                thr = 1.f; res = 0.f; depth = 0; on_inst = on_disk = -1;
                const bool valid = canopy_primary(P, pix, rng, p, d, wray, maxt);
                first_segment = true;
                if (P.sensor.type == ERTB_SENSOR_PERSPECTIVE || P.sensor.type == ERTB_SENSOR_MRADIANCEMETER) {
                    in_medium = P.sensor.in_medium != 0;
                } else if (valid) {
                    // distant sensors sit outside the scene: bring the origin down to where the ray
                    // enters it (top of the atmosphere, or just above the canopy without one)
                    in_medium = false;
                    double z_in = medium ? z_ground + (double) P.H : fmax(C.n_instances ? C.hi[2] : z_ground, p[2]) + 1.0;
                    double t_in = (z_in - p[2]) / (double) d.z; // <= 0: move backwards along the ray
                    p[0] += t_in * (double) d.x; p[1] += t_in * (double) d.y; p[2] = z_in;
                }
                // (an invalid primary ray is a sample with L = 0: nothing to accumulate)
                if (valid) phase = LP_SEGMENT;
            }
        }
        if (exhausted && !__any_sync(0xffffffffu, phase != LP_NEW)) break;

        // ================= segment set-up: termination, analytic surfaces, start of the BVH walk =================
        if (phase == LP_SEGMENT) {
            // ---- termination (volpath.cpp:189-202) ----
            bool dead = thr == 0.f || depth >= P.max_depth;
            if (!dead && depth > P.rr_depth) {
                float q = fminf(thr, 0.95f);
                if (pcg_float(rng) >= q) dead = true;
                else thr = __fdividef(thr, q);
            }
            // a ray above the atmosphere travels through vacuum down to its top
            if (!dead && !in_medium && medium && (float) (p[2] - z_ground) >= P.H) {
                if (!(d.z < 0.f)) { // leaves the scene
                    dead = true;
                    if (MESH && depth == 0u) res += astro_direct(P, d, sun);
                }
                else {
                    double t_in = ((double) P.H - (p[2] - z_ground)) / (double) d.z;
                    p[0] += t_in * (double) d.x; p[1] += t_in * (double) d.y; p[2] = z_ground + (double) P.H;
                    in_medium = true;
                    if (first_segment) maxt -= (float) t_in;
                }
            }
            if (dead) CANOPY_FINISH();
            else {
                // ---- next analytic surface: ground plane / top of the slab ----
                const bool down = d.z < 0.f;
                t_geo = down ? (p[2] - z_ground) / (double) -d.z
                             : (in_medium && d.z > 0.f ? fmax((double) P.H - (p[2] - z_ground), 0.0) / (double) d.z : 1e30);
                ev_geo = down ? CEV_GROUND : CEV_END; // upward: leaves through the top (or to infinity)
                if (t_geo > 1e9) { t_geo = 1e9; ev_geo = CEV_END; } // grazing rays: the reference's slab is 1e9 m wide
                // far clip. (General instances: a ray clipped INSIDE the medium carries no throughput in the reference --
                // transmittance_eval_pdf over an infinite distance, volpath.cpp:229-233 -- so it does not see the solar
                // disc of an astroobject either; a ray clipped in vacuum does)
                if (first_segment && (double) maxt < t_geo) {
                    t_geo = (double) maxt;
                    ev_geo = (MESH && in_medium && medium) ? CEV_CLIP : CEV_END;
                }
                // ---- nearest leaf: walk the BVH over the part of the segment inside the canopy's box ----
                double t1;
                shadow = false;
                T.H.t = INFINITY; T.H.inst = -1; T.H.disk = -1;
                phase = LP_FLIGHT;
                if (C.n_instances && canopy_clip(C, p, d, t_geo, t_clip0, t1)) {
                    trace_begin(T, canopy_local(C, p, d, t_clip0), (float) (t1 - t_clip0), on_inst, on_disk);
                    phase = LP_TRACE;
                }
            }
        }

        // ================= BVH walk, one slice (nearest leaf of a path segment, or any leaf on a shadow ray) =================
        // Two forms.  Fixed slices (VOTE = false): the lanes that walk advance by ERTB_TRACE_STEPS_FIXED node visits.  Every
        // trip of the main loop costs the other stages' votes and divergent code -- and the set-up of the walk itself --
        // whatever the number of lanes that need them (C4: 112 / 158 / 168 / 160 Mpaths/s at slices of 8 / 32 / 64 /
        // 128 visits): short slices pay it too often, long ones keep finished lanes waiting.  VOTE = true: EVERY lane
        // enters the walk's loop (idle ones skip the body), whose trip count is then warp-uniform, and a second
        // sub-slice follows while at least half of the lanes that entered still walk.  Canopies of small groups
        // (abstract / mesh trees, a few hundred primitives: walks of a dozen visits) gain 13-20 % from it, the 60 k-leaf
        // groups of C4 lose 0-4 % (profiles/r02t_canopy_slices.md): the host picks the form from the group sizes.
        if (__any_sync(0xffffffffu, phase == LP_TRACE)) {
            if (VOTE) {
                if (trace_run<MESH, true>(C, T, stack, shadow ? SUN_E : d, shadow, ERTB_TRACE_STEPS, phase == LP_TRACE))
                    phase = shadow ? LP_SHADE : LP_FLIGHT;
            } else if (phase == LP_TRACE && trace_run<MESH>(C, T, stack, shadow ? SUN_E : d, shadow, ERTB_TRACE_STEPS_FIXED))
                phase = shadow ? LP_SHADE : LP_FLIGHT;
        }

        // ================= free flight to the next event, the event, the sun sample it generates =================
        if (phase == LP_FLIGHT) {
            float h = (float) (p[2] - z_ground); // altitude above the ground
            double t_seg = t_geo;
            int ev = ev_geo;
            if (T.H.inst >= 0) { t_seg = t_clip0 + (double) T.H.t; ev = CEV_LEAF; }
            double t_ev = t_seg;
            nee = 0.f;
            // ---- free flight through the medium, bounded by t_seg (medium.cpp:42-82 / piecewise.cpp:183-332) ----
            if (STATS && !(in_medium && medium)) st_main++; // a vacuum segment is one loop trip
            if (in_medium && medium) {
                if (PW) {
                    if (STATS) st_main++;
                    float s, hn;
                    int r = pw_flight(P, tb, h, d.z, -__logf(1.f - pcg_float(rng)), s, hn);
                    if (r == PW_COLLISION && (double) s < t_seg) { t_ev = (double) s; ev = CEV_COLLISION; }
                } else {
                    float t = 0.f;
                    const float ts = (float) t_seg;
                    for (;;) {
                        if (STATS) st_main++;
                        t += -__logf(1.f - pcg_float(rng)) * P.inv_majorant;
                        if (!(t < ts)) break;
                        float preal = tb[P.off_preal + layer_of(P, fmaf(t, d.z, h))];
                        if (pcg_float(rng) >= 1.f - preal) { t_ev = (double) t; ev = CEV_COLLISION; break; }
                    }
                }
            }
            first_segment = false;
            const bool escaped = ev == CEV_END; // no event on this segment: the ray leaves the scene
            if (MESH && ev == CEV_CLIP) ev = CEV_END;
            if (MESH && P.astro_omc > 0.f && !escaped) sun_e = astro_sample(P, sun, rng);
            if (ev != CEV_END) {
                p[0] += t_ev * (double) d.x; p[1] += t_ev * (double) d.y; p[2] += t_ev * (double) d.z;
                if (ev == CEV_GROUND) p[2] = z_ground;
                h = fmaxf((float) (p[2] - z_ground), 0.f);
            }
            if (ev == CEV_COLLISION) {
                // ---- real collision (volpath.cpp:261-310) ----
                int l = layer_of(P, h);
                thr *= tb[P.off_albedo + l];
                depth++;
                on_inst = on_disk = -1;
                if (STATS) st_scatter++;
                if (depth >= P.max_depth || thr == 0.f) ev = CEV_END;
                else {
                    float ct_sun = dot3(d, SUN_E);
                    float pv = 0.f;
                    int leaf = 0;
                    if (P.n_phase == 1) pv = leaf_eval(tb, P.leaf[0], ct_sun);
                    else {
                        float u0 = pcg_float(rng), prev = 0.f;
                        leaf = P.n_phase - 1;
                        bool picked = false;
                        for (int i = 0; i < P.n_phase; ++i) {
                            float cum = (i < P.n_phase - 1) ? tb[P.off_cumw + i * P.n_layers + l] : 1.f;
                            float w = cum - prev;
                            prev = cum;
                            if (w > 0.f) pv = fmaf(w, leaf_eval(tb, P.leaf[i], ct_sun), pv);
                            if (!picked && u0 < cum) { leaf = i; picked = true; }
                        }
                    }
                    nee = thr * pv * P.irradiance;
                    float u1 = pcg_float(rng), u2 = pcg_float(rng);
                    float pw, ppdf;
                    float ct = leaf_sample(tb, P.leaf[leaf], u1, pw, ppdf);
                    if (ppdf > 0.f) {
                        float st = safe_sqrtf(1.f - ct * ct), sp, cp;
                        __sincosf(2.f * ERTB_PI * u2, &sp, &cp);
                        f3 fs, ft;
                        onb(d, fs, ft);
                        d = normalize3(fma3(fs, st * cp, fma3(ft, st * sp, scale3(d, ct))));
                        if (MESH && P.phase_mis) { // multiphase.cpp:176-200: mixture value / mixture pdf
                            float num = 0.f, den = 0.f, prev = 0.f;
                            for (int i = 0; i < P.n_phase; ++i) {
                                float cum = (i < P.n_phase - 1) ? tb[P.off_cumw + i * P.n_layers + l] : 1.f;
                                float w = cum - prev;
                                prev = cum;
                                if (w > 0.f) {
                                    num = fmaf(w, leaf_eval(tb, P.leaf[i], ct), num);
                                    den = fmaf(w, leaf_pdf(tb, P.leaf[i], ct), den);
                                }
                            }
                            pw = den > 1e-8f ? __fdividef(num, den) : 0.f;
                        }
                        thr *= pw;
                    }
                }
            } else if (ev == CEV_GROUND) {
                // ---- ground BSDF (volpath.cpp:344-389), local frame = world axes ----
                if (STATS) st_surface++;
                on_inst = on_disk = -1;
                float ci = -d.z;
                if (!(ci > 0.f) || P.bsdf_type == ERTB_BSDF_BLACK) ev = CEV_END;
                else if (C.patch_type >= 0 && fabs(p[0] - C.patch_rect[0]) <= C.patch_rect[2] &&
                         fabs(p[1] - C.patch_rect[1]) <= C.patch_rect[3]) {
                    // CentralPatchSurface (blendbsdf.cpp:108-165 with a 0/1 mask): the patch's own land BSDF
                    const float *pp = tb + C.off_patch_bsdf;
                    if (depth + 1u < P.max_depth && SUN_E.z > 0.f)
                        nee = thr * land_bsdf(C.patch_type, pp, ci, SUN_E.z, cos_dphi(ci, SUN_E.z, -dot3(d, SUN_E))) * SUN_E.z * P.irradiance;
                    float u1 = pcg_float(rng), u2 = pcg_float(rng);
                    f3 wl = cosine_hemisphere(u1, u2);
                    float weight = 0.f;
                    if (wl.z > 0.f) // value * cos / pdf = value * pi (rpv.cpp:119-122)
                        weight = land_bsdf(C.patch_type, pp, ci, wl.z, cos_dphi(ci, wl.z, -dot3(d, wl))) * ERTB_PI;
                    d = normalize3(wl); // local frame of the ground = world axes
                    thr *= weight;
                    depth++;
                } else {
                    float f_sun, weight;
                    const f3 up = mk3(0.f, 0.f, 1.f);
                    surface_interact<false, false, MESH>(P, up, SUN_E, ci, depth + 1u < P.max_depth, rng, d, f_sun, weight);
                    nee = thr * f_sun * P.irradiance;
                    thr *= weight;
                    depth++;
                }
            } else if (ev == CEV_LEAF) {
                // ---- leaf: bilambertian reflection / transmission; the medium does not change ----
                if (STATS) st_surface++;
                float4 in = __ldg(C.inst + T.H.inst);
                const float *lb = tb + C.off_leaf_bsdf + 4 * __float_as_int(in.w);
                int kind, mat;
                f3 n = canopy_normal<MESH>(C, p, T.H, kind, mat);
                float ci = -dot3(n, d);
                on_inst = T.H.inst; on_disk = T.H.disk;
                // leaves: bilambertian (r, t); trunk parts: one-sided Lambertian = (rho, 0) seen from the front only;
                // mesh triangles: the bilambertian of their element, about the interpolated shading normal
                float r_ = lb[0], t_ = lb[1];
                if (MESH && kind == 3) { r_ = tb[C.off_mesh_bsdf + 2 * mat]; t_ = tb[C.off_mesh_bsdf + 2 * mat + 1]; }
                else if (kind != 0) { r_ = ci > 0.f ? lb[2] : 0.f; t_ = 0.f; }
                if (depth + 1u < P.max_depth) nee = thr * bilambertian_eval(r_, t_, ci, dot3(n, SUN_E)) * P.irradiance;
                float s1 = pcg_float(rng), u1 = pcg_float(rng), u2 = pcg_float(rng);
                f3 wl;
                thr *= bilambertian_sample(r_, t_, ci, s1, u1, u2, wl);
                f3 fs, ft;
                onb(n, fs, ft);
                d = normalize3(fma3(fs, wl.x, fma3(ft, wl.y, scale3(n, wl.z))));
                depth++;
            }
            if (ev == CEV_END) {
                // astroobject: a primary ray that met nothing looks at the sky and sees the disc if it points into it
                // (volpath.cpp:328-346, throughput 1; scattered rays reach the disc by next-event estimation only)
                if (MESH && escaped && depth == 0u) res += astro_direct(P, d, sun);
                CANOPY_FINISH();
            } else {
                // ---- next-event estimation (volpath.cpp:400-554): leaves and the ground are opaque to the
                //      shadow ray, the medium attenuates it. The occlusion test is the second BVH walk. ----
                phase = LP_SEGMENT;
                if (STATS && nee != 0.f && !(in_medium && medium)) st_nee++; // a vacuum shadow ray is one loop trip
                if (!(SUN_E.z > 0.f)) nee = 0.f; // the ground plane is in the way
                if (nee != 0.f) { // (either sign: a BSDF such as RTLS may be negative at grazing angles)
                    double t1;
                    shadow = true;
                    T.H.inst = -1;
                    phase = LP_SHADE;
                    if (C.n_instances && canopy_clip(C, p, SUN_E, 1e30, t_clip0, t1)) {
                        trace_begin(T, canopy_local(C, p, SUN_E, t_clip0), (float) (t1 - t_clip0), on_inst, on_disk);
                        phase = LP_TRACE;
                    }
                }
            }
        }

        // ================= shading: the sun sample of the last event, if nothing was in the way =================
        if (phase == LP_SHADE) {
            if (T.H.inst < 0) {
                float tr = 1.f;
                if (in_medium && medium)
                    tr = sun_medium_transmittance<PW, STATS>(P, tb, fmaxf((float) (p[2] - z_ground), 0.f), SUN_E, rng, st_nee);
                res = fmaf(nee, tr, res);
            }
            phase = LP_SEGMENT;
        }
    }
#undef CANOPY_FINISH
#undef SUN_E

    film_flush_warp<false>(P, lane, true, acc_pix, acc_wl, acc_l, acc_l2, 0.0, 0.0, 0.0);
    if (STATS && P.stats) {
        unsigned v[5] = { st_paths, st_main, st_nee, st_scatter, st_surface };
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            unsigned long long x = v[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            if (lane == 0) atomicAdd(&P.stats[i], x);
        }
    }
}
