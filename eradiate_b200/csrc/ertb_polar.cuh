// ertb_polar.cuh -- polarized (Stokes / Mueller) building blocks of the sm_100a path tracer.
//
// Reference: MI/include/mitsuba/render/mueller.h (rotator :164-173, stokes_basis :286,
// rotate_stokes_basis :316-324, rotate_mueller_basis :362-372),
// ERP/phase/rayleigh_polarized.cpp:55-127 (Hansen & Travis 1974 eq. 2.15 + basis rotation),
// ERP/phase/tabphase_polarized.cpp:178-206, :318-368 (m11 .. m44 interpolants),
// MI/src/integrators/stokes.cpp:111-151 (output basis).  The scalar plugins return
// Spectrum(value) = value * Identity in a polarized variant, the land BSDFs depolarizer(value).
//
// Conventions: matrices are row-major float[16]; a Stokes vector lives in the implicit basis
// stokes_basis(forward) = first vector of the Duff et al. frame of its propagation direction
// (the same construction as onb()); rotations between bases never need trigonometric calls:
// with unit vectors, cos(theta) = b_cur . b_tgt and sin(theta) = forward . (b_cur x b_tgt).
#pragma once

#include "ertb_device.cuh"

__device__ __forceinline__ f3 cross3(f3 a, f3 b) {
    return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ f3 neg3(f3 a) { return mk3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ f3 stokes_basis(f3 forward) {
    f3 s, t;
    onb(forward, s, t);
    return s;
}

// (cos 2theta, sin 2theta) of the rotation taking basis_current to basis_target around forward
// (rotate_stokes_basis, mueller.h:316-324).  Returns false when a basis is degenerate.
__device__ __forceinline__ bool basis_rotation(f3 forward, f3 bc, f3 bt, float &c2, float &s2) {
    float nc = dot3(bc, bc), nt = dot3(bt, bt);
    if (!(nc > 1e-20f) || !(nt > 1e-20f)) { c2 = 1.f; s2 = 0.f; return false; }
    float inv = rsqrtf(nc * nt);
    float c = dot3(bc, bt) * inv, s = dot3(forward, cross3(bc, bt)) * inv;
    float n = c * c + s * s; // = 1 for bases perpendicular to a unit forward; renormalise
    if (!(n > 1e-20f)) { c2 = 1.f; s2 = 0.f; return false; }
    float invn = 1.f / n;
    c2 = (c * c - s * s) * invn;
    s2 = 2.f * c * s * invn;
    return true;
}

// M <- R(out) * M * R(in)^T with R = rotator (mueller.h:164-173): only rows/columns 1,2 mix.
__device__ __forceinline__ void rotate_mueller(float *M, float ci, float si, float co, float so) {
    // columns: (M R_in^T)[:,1] = c M[:,1] + s M[:,2];  [:,2] = -s M[:,1] + c M[:,2]
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        float a = M[4 * r + 1], b = M[4 * r + 2];
        M[4 * r + 1] = ci * a + si * b;
        M[4 * r + 2] = -si * a + ci * b;
    }
    // rows: (R_out M)[1,:] = c M[1,:] + s M[2,:];  [2,:] = -s M[1,:] + c M[2,:]
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        float a = M[4 + c], b = M[8 + c];
        M[4 + c] = co * a + so * b;
        M[8 + c] = -so * a + co * b;
    }
}

__device__ __forceinline__ void mueller_mul(const float *A, const float *B, float *C) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
            C[4 * i + j] = fmaf(A[4 * i], B[j], fmaf(A[4 * i + 1], B[4 + j], fmaf(A[4 * i + 2], B[8 + j], A[4 * i + 3] * B[12 + j])));
}

// Mueller-valued eval_pdf of one phase-function leaf in world space, Radiance mode: light arrives
// along -wo and leaves along +wi (wi = -propagation direction of the camera path).
__device__ __forceinline__ void leaf_mueller(const float *tb, const ErtbPhaseLeaf &L, f3 wi, f3 wo, float *M, float &pdf) {
#pragma unroll
    for (int i = 0; i < 16; ++i) M[i] = 0.f;
    float ct = -dot3(wo, wi); // physics convention
    ct = clampf(ct, -1.f, 1.f);
    bool polarized_leaf = false;
    if (L.type == ERTB_PHASE_RAYLEIGH_POLARIZED) {
        float rho = L.p0;
        float r1 = __fdividef(1.f - rho, 1.f + 0.5f * rho), r2 = __fdividef(1.f + rho, 1.f - rho),
              r3 = __fdividef(1.f - 2.f * rho, 1.f - rho);
        float k = (3.f / 16.f) * ERTB_INV_PI * r1;
        float ct2 = ct * ct;
        M[0] = k * (r2 + ct2); M[1] = k * (ct2 - 1.f); M[4] = M[1]; M[5] = k * (ct2 + 1.f);
        M[10] = k * 2.f * ct; M[15] = M[10] * r3;
        pdf = rayleigh_pdf(ct);
        polarized_leaf = true;
    } else if (L.type == ERTB_PHASE_TABULATED_POLARIZED) {
        const float *nodes = tb + L.off_nodes, *m11v = tb + L.off_pdf;
        float norm = L.normalization * ERTB_INV_TWO_PI;
        float ms[6] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f };
        if (ct >= nodes[0] && ct <= nodes[L.n_nodes - 1]) {
            int lo = 0, hi = L.n_nodes - 1;
            while (hi - lo > 1) {
                int mid = (lo + hi) >> 1;
                if (nodes[mid] < ct) lo = mid; else hi = mid;
            }
            float t = __fdividef(ct - nodes[lo], nodes[lo + 1] - nodes[lo]);
            ms[0] = fmaf(t, m11v[lo + 1] - m11v[lo], m11v[lo]);
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const float *a = tb + L.off_mueller + k * L.mueller_stride;
                ms[k + 1] = fmaf(t, a[lo + 1] - a[lo], a[lo]);
            }
        }
        M[0] = ms[0] * norm; M[1] = ms[1] * norm; M[4] = M[1]; M[5] = ms[2] * norm;
        M[10] = ms[3] * norm; M[11] = ms[4] * norm; M[14] = -M[11]; M[15] = ms[5] * norm;
        pdf = M[0];
        polarized_leaf = true;
    } else {
        float v = leaf_eval(tb, L, ct);
        M[0] = M[5] = M[10] = M[15] = v;
        pdf = (L.type == ERTB_PHASE_RAYLEIGH) ? rayleigh_pdf(ct) : v;
    }
    if (polarized_leaf) {
        // scattering-plane frame -> implicit bases of -wo (incident) and wi (outgoing)
        f3 fin = neg3(wo), fout = wi;
        f3 x_hat = cross3(fin, fout);
        f3 p_in = cross3(x_hat, fin), p_out = cross3(x_hat, fout);
        float ci, si, co, so;
        bool ok = basis_rotation(fin, p_in, stokes_basis(fin), ci, si);
        ok = basis_rotation(fout, p_out, stokes_basis(fout), co, so) && ok;
        if (!ok) { // collinear directions: the reference zeroes the NaN matrix
#pragma unroll
            for (int i = 0; i < 16; ++i) M[i] = 0.f;
        } else {
            rotate_mueller(M, ci, si, co, so);
        }
    }
}

// Rotator taking the implicit basis of -d (primary ray) to the output basis of the stokes
// integrator (stokes.cpp:111-151), as a full Mueller matrix.
__device__ __forceinline__ void stokes_output_rotation(const ErtbParams &P, f3 d, float *R) {
    f3 fwd = neg3(d);
    f3 current = stokes_basis(fwd), target;
    if (P.meridian_align) {
        f3 tmp = cross3(mk3(0.f, 0.f, 1.f), fwd);
        float n2 = dot3(tmp, tmp);
        if (n2 < 1e-12f) target = mk3(1.f, 0.f, 0.f);
        else target = cross3(scale3(tmp, rsqrtf(n2)), fwd);
    } else {
        target = cross3(d, mk3(P.sensor_up[0], P.sensor_up[1], P.sensor_up[2]));
    }
    float c2, s2;
    basis_rotation(fwd, current, target, c2, s2);
#pragma unroll
    for (int i = 0; i < 16; ++i) R[i] = 0.f;
    R[0] = 1.f; R[5] = c2; R[6] = s2; R[9] = -s2; R[10] = c2; R[15] = 1.f;
}

#include "ertb_ocean.cuh"

// Polarized value of a local-frame BSDF (6SV ocean, ocean_mishchenko, ocean_grasp, maignan) in WORLD implicit
// Stokes bases: BSDF::eval (ocean_legacy.cpp:561-661, ocean_mishchenko.cpp:228-296, ocean_grasp.cpp:354-455,
// maignan.cpp:105-166) followed by SurfaceInteraction::to_world_mueller.  The reference rotates twice about the
// same propagation directions (meridian plane -> local implicit basis -> world implicit basis); the two rotations
// compose, so the meridian-plane axes are carried to world space and rotated once.
// `weight`: return the weight BSDF::sample gives for the sampled `wo` instead (eval / pdf for the two-lobe
// oceans, F G / G1 for ocean_mishchenko, C F for maignan).
// `wi`, `wo` local (wi = si.wi, wo = towards the light / sampled), (fs, ft, n) = shading frame.
template <bool GENERAL = true>
__device__ __forceinline__ void lf_eval_mueller(const ErtbParams &P, bool weight, f3 wi, f3 wo, f3 fs, f3 ft, f3 n, float *M) {
#pragma unroll
    for (int i = 0; i < 16; ++i) M[i] = 0.f;
    if (!(wi.z > 0.f && wo.z > 0.f)) return;
    float g, dep;
    if (!GENERAL || P.bsdf_type == ERTB_BSDF_OCEAN_LEGACY) {
        const float *tdn = P.ocean_tables, *tup = P.ocean_tables + ERTB_OC_RES * ERTB_OC_RES;
        float wc = P.bsdf[OC_WHITECAP], ul = 0.f;
        if (P.bsdf[OC_UNDERLIGHT_ON] != 0.f)
            ul = P.bsdf[OC_UL_NORM] * oc_transmittance(P, tup, wi.z, wo.x, wo.y) * oc_transmittance(P, tdn, wo.z, wo.x, wo.y);
        float scale = wo.z * ERTB_INV_PI;
        if (weight) { // ocean_legacy.cpp:553-558
            float pdf = oc_pdf(P, wi, wo);
            scale = pdf > 0.f ? __fdividef(scale, pdf) : 0.f;
        }
        // glint geometry factor without Fresnel (eval_glint(wi := wo, wo := si.wi), :405-420)
        f3 m = normalize3(mk3(wi.x + wo.x, wi.y + wo.y, wi.z + wo.z));
        g = oc_beckmann_D(P, m) * oc_gram_charlier(P, m) / (4.f * wi.z * wo.z);
        if (P.bsdf[OC_SHADOWING] != 0.f) {
            float G = 1.f / (1.f + oc_lambda(P, wi) + oc_lambda(P, wo));
            if (dot3(wi, m) * wi.z <= 0.f || dot3(wo, m) * wo.z <= 0.f) G = 0.f;
            g *= G;
        }
        g *= ERTB_PI * (1.f - P.bsdf[OC_COVERAGE]) * scale;
        dep = (wc + (1.f - wc) * ul) * scale;
    } else {
        float dscale = 1.f;
        g = weight ? gl_weight_geometry(P, wi, wo, dscale) : gl_geometry(P, wi, wo);
        dep = gl_dep(P, wo) * dscale;
    }
    f3 in_fwd = neg3(wo), out_fwd = wi; // light arrives along -wo and leaves along si.wi
    oc_fresnel_mueller(P.bsdf[OC_N_REAL], P.bsdf[OC_N_IMAG], in_fwd, out_fwd, M);
#pragma unroll
    for (int i = 0; i < 16; ++i) M[i] *= g;
    // meridian-plane axes (local, normal = z): p = normalize((z x f) x f); fallback (0, 1, 0)
    f3 z = mk3(0.f, 0.f, 1.f);
    f3 a_in = cross3(z, in_fwd), a_out = cross3(z, out_fwd);
    f3 p_in = dot3(a_in, a_in) > 1e-20f ? cross3(a_in, in_fwd) : mk3(0.f, 1.f, 0.f);
    f3 p_out = dot3(a_out, a_out) > 1e-20f ? cross3(a_out, out_fwd) : mk3(0.f, 1.f, 0.f);
    // to world
    f3 in_w = fma3(fs, in_fwd.x, fma3(ft, in_fwd.y, scale3(n, in_fwd.z)));
    f3 out_w = fma3(fs, out_fwd.x, fma3(ft, out_fwd.y, scale3(n, out_fwd.z)));
    f3 p_in_w = fma3(fs, p_in.x, fma3(ft, p_in.y, scale3(n, p_in.z)));
    f3 p_out_w = fma3(fs, p_out.x, fma3(ft, p_out.y, scale3(n, p_out.z)));
    float ci, si, co, so;
    basis_rotation(in_w, p_in_w, stokes_basis(in_w), ci, si);
    basis_rotation(out_w, p_out_w, stokes_basis(out_w), co, so);
    rotate_mueller(M, ci, si, co, so);
    M[0] += dep; // depolarizer part is rotation invariant
}
