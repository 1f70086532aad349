// ertb_device.cuh -- device-side building blocks of the sm_100a path tracer:
// RNG, warps, tabulated distributions, phase functions and BSDFs (fp32).
//
// Each function cites the reference plugin it implements ("MI" =
// ext/mitsuba, "ERP" = MI/src/eradiate_plugins).  Nothing here is shared with the
// CPU oracle (oracle/): the oracle is an independent double-precision
// restatement used only to check this code.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/eradiate_b200.h"

#define ERTB_PI 3.14159265358979323846f
#define ERTB_INV_PI 0.31830988618379067154f
#define ERTB_INV_TWO_PI 0.15915494309189533577f
#define ERTB_INV_FOUR_PI 0.07957747154594766788f

// ----------------------------------------------------------------------------
// Kernel parameters (passed by value -> constant bank)
// ----------------------------------------------------------------------------

struct ErtbPhaseLeaf {
    int type;       // enum ertb_phase_type
    int n_nodes;    // tabulated
    float p0;       // hg: g ; rayleigh: depolarization
    float integral; // tabulated: integral of the unnormalised pdf (cdf[n-2])
    float normalization;
    float inv_interval; // regular grid: (n-1)/2
    int off_pdf;    // float offsets into the shared-memory table blob
    int off_cdf;
    int off_nodes;  // irregular only (else -1)
    int valid0, valid1; // first / last interval with non-zero mass (distr_1d.h:548-600)
    int off_mueller;    // tabulated_polarized: m12 | m22 | m33 | m34 | m44, `mueller_stride` floats apart
    int mueller_stride;
};

struct ErtbSensor {
    int type;          // enum ertb_sensor_type
    int width, height;
    int target_type;
    int use_table;     // mdistant + point target: per-pixel primary rays precomputed on the host
    const float *table; // [n_pixels][8]: entry n (3), direction d (3), valid, unused
    float to_world[9]; // rotation part, row-major (hdistant / distantflux)
    double target[3];
    double target_to_world[12]; // 3x4 row-major
    double bs_center[3];
    double bs_radius;
    double ray_offset; // origin = target - d * ray_offset (mdistant.cpp:180-190)
    float flux_norm;   // distantflux: 2*pi / n_pixels
    // perspective (perspective.cpp:200-236): pinhole at cam_origin, to_world = camera rotation
    double cam_origin[3];
    float tan_half_fov, aspect, near_clip, far_clip;
    int in_medium;     // the camera / the radiancemeters sit inside the atmosphere
    const double *origins; // mradiancemeter: [n][3] ray origins (3D kernel; the 1D kernels use `table`)
};

// Explicit canopy (ertb_canopy.cuh): two-level BVH over translated instances of disk-leaf groups.
// Coordinates are float32 relative to `origin` (the centre of the canopy's bounding box), so leaf
// geometry keeps sub-millimetre resolution whatever the extent of the atmosphere around it.
struct ErtbBvhNode { // 64 B: the boxes of BOTH children, so one fetch decides where to go next
    float lo0[3]; int c0; // child 0: n0 == 0 -> inner node index c0; n0 > 0 -> leaf of n0 primitives from c0
    float hi0[3]; int n0;
    float lo1[3]; int c1; // child 1 (an unreachable point box when the tree is a single leaf)
    float hi1[3]; int n1;
};
struct ErtbCanopy {
    int n_instances;          // 0: the scene has no canopy
    const ErtbBvhNode *tlas;  // over the instances (leaf primitives index `inst`)
    const float4 *inst;       // (offset xyz relative to `origin`, group index as int bits)
    const ErtbBvhNode *blas;  // all groups' trees, concatenated
    const int *blas_root;     // per group: root node index in `blas`
    const float4 *disks;      // 2 per primitive, in BVH leaf order: disk (centre, radius) (normal, kind 0 leaf / 1 trunk
                              // cap); cylinder (p0, radius) (axis vector p1 - p0, kind 2)
                              // triangle (v0, index into `tris` as int bits) (e1 = v1 - v0, kind 3)
    const float4 *tris;       // 4 per triangle: (e2 = v2 - v0, row of its BSDF in the mesh table as int bits), n0, n1, n2
    int off_leaf_bsdf;        // table blob: per group (leaf reflectance, leaf transmittance, trunk reflectance, -)
    int off_mesh_bsdf;        // table blob: per mesh BSDF (reflectance, transmittance)
    double origin[3];
    double lo[3], hi[3];      // world-space bounding box of all instances
    // CentralPatchSurface: a second ground BSDF (type, 16 params in the table blob) inside a rectangle
    int patch_type;           // < 0: none
    int off_patch_bsdf;
    double patch_rect[4];     // cx, cy, hx, hy
};

struct ErtbParams {
    // geometry (altitudes are relative to the ground surface)
    int spherical;
    float R;          // ground sphere radius (spherical shell)
    double Rd;        // same, double (primary-ray setup)
    float H;          // top-of-atmosphere altitude above the ground (0 without medium)
    float h_off;      // ground altitude - grid bottom  (layer lookup offset)
    float inv_dz;     // n_layers / (grid top - grid bottom)
    // medium
    int has_medium;
    int n_layers;
    float majorant, inv_majorant;
    int n_phase;
    ErtbPhaseLeaf leaf[ERTB_MAX_PHASE];
    // table blob (global), staged to shared memory with one TMA bulk copy
    const float *blob;
    int blob_bytes;   // multiple of 16
    int off_preal;    // sigma_t/majorant per layer
    int off_albedo;
    int off_cumw;     // (n_phase-1) x n_layers cumulative leaf probabilities
    // banded majorant (pool kernel, BANDS instances): the layer stack is cut into n_bands altitude bands,
    // each with its own majorant. Blob: band_lo[n_bands + 1] (altitudes above the ground of the band
    // boundaries), band_ratio[n_bands] (global majorant / band majorant), band_iratio[n_bands] (its inverse)
    int n_bands;
    int off_band_lo, off_band_ratio, off_band_iratio;
    // piecewise medium (ertb_piecewise.cuh): sigma_t per layer, vertical optical depth above each of
    // the n_layers + 1 layer boundaries, layer thickness, optical depth above the ground level
    int piecewise;
    int off_sigma, off_tau;
    float dz, tau_ground;
    // surface
    int bsdf_type;
    float bsdf[ERTB_MAX_BSDF_PARAMS];
    const float *ocean_tables; // ocean_legacy: [2][64*64] down/up-welling transmittance (global memory)
    // emitter
    float sun[3];     // unit vector pointing towards the sun (= -emitter direction)
    float irradiance;
    // astroobject (astroobject.cpp:54-242): a uniform disc of angular radius a around `sun`; 0 = directional.
    // 1 - cos a and sin^2 a are derived in double on the host (cos a = 1 - 1.1e-5 for the Sun)
    float astro_omc, astro_sin2, astro_radiance; // 1 - cos a, sin^2 a, irradiance / solid angle
    // integrator
    int polarized;    // Mueller/Stokes transport (scalar_mono_polarized variant)
    int phase_mis;    // multiphase.cpp:176-200: mixture weight of a sampled direction (GEN instances of the pool kernel)
    int meridian_align; // stokes.cpp: output basis in the meridian plane
    float sensor_up[3]; // else: sensor world_transform * (0, 1, 0)
    int mis;          // volpathmis Russian-roulette placement
    unsigned rr_depth;
    unsigned max_depth; // 0xffffffff = unbounded
    // sensor + work decomposition
    ErtbSensor sensor;
    unsigned long long seed;
    unsigned long long spp;           // samples per pixel rendered by this launch
    unsigned long long sample_offset; // first sample index (multi-GPU sharding)
    unsigned n_pixels;
    int tw;                           // scheduler: min. walking lanes for a free-flight trip
    int twi;                          // pool kernel: keep stepping while >= twi loaded records walk
    unsigned chunk;                   // samples per work chunk
    unsigned chunks_per_pixel;
    unsigned long long n_chunks;
    unsigned long long *work_counter; // device, zeroed before launch
    double *accum;                    // [3][n_pixels]: sum w*L | sum L | sum L^2
    unsigned long long *stats;        // [8] or nullptr
    ErtbCanopy canopy;
};

// ----------------------------------------------------------------------------
// small vector helpers
// ----------------------------------------------------------------------------
struct f3 { float x, y, z; };
__device__ __forceinline__ f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ float dot3(f3 a, f3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
__device__ __forceinline__ f3 fma3(f3 a, float s, f3 b) { return mk3(fmaf(a.x, s, b.x), fmaf(a.y, s, b.y), fmaf(a.z, s, b.z)); }
__device__ __forceinline__ f3 scale3(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ f3 normalize3(f3 a) { return scale3(a, rsqrtf(dot3(a, a))); }
__device__ __forceinline__ float safe_sqrtf(float x) { return sqrtf(fmaxf(x, 0.f)); }
__device__ __forceinline__ float clampf(float x, float a, float b) { return fminf(fmaxf(x, a), b); }

// Branchless orthonormal basis (Duff et al. 2017), same construction as
// MI/include/mitsuba/core/vector.h:118-140 coordinate_system().
__device__ __forceinline__ void onb(f3 n, f3 &s, f3 &t) {
    float sign = copysignf(1.f, n.z);
    float a = -1.f / (sign + n.z);
    float b = n.x * n.y * a;
    s = mk3(fmaf(sign * n.x * n.x, a, 1.f), sign * b, -sign * n.x);
    t = mk3(b, fmaf(n.y * n.y, a, sign), -n.y);
}

// ----------------------------------------------------------------------------
// RNG: one stream per path keyed by (seed, pixel, sample index), so results do not depend on how
// paths are scheduled over lanes, CTAs or GPUs. The reference's sampler is PCG32
// (MI/ext/drjit/include/drjit/random.h:108-195, kept under -DERTB_RNG_PCG32 and used by the oracle);
// its 64-bit multiply-add costs ~13 % of the C2 step in 32-bit integer instructions
// (profiles/r01d_pool_regions.md). The default generator is xoroshiro64** (Blackman & Vigna 2018): same
// 64 bits of state (the `inc` word is unused), 32-bit operations only, +3.7 % on C2. Estimates are
// statistically equivalent, not bit-identical, to any other generator's -- as between two seeds of the
// reference (SURVEY 8b "Seeds").
// ----------------------------------------------------------------------------
struct Pcg32 {
    unsigned long long state, inc;
};
#define ERTB_PCG_MULT 6364136223846793005ULL

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
__device__ __forceinline__ unsigned pcg_next(Pcg32 &r) {
#ifndef ERTB_RNG_PCG32
    // xoroshiro64** (Blackman & Vigna 2018) on the two halves of `state`: 32-bit operations only
    unsigned s0 = (unsigned) r.state, s1 = (unsigned) (r.state >> 32);
    const unsigned x = s0 * 0x9E3779BBu;
    const unsigned result = __funnelshift_l(x, x, 5) * 5u;
    s1 ^= s0;
    s0 = __funnelshift_l(s0, s0, 26) ^ s1 ^ (s1 << 9);
    s1 = __funnelshift_l(s1, s1, 13);
    r.state = (unsigned long long) s0 | ((unsigned long long) s1 << 32);
    return result;
#else
    unsigned long long old = r.state;
    r.state = old * ERTB_PCG_MULT + r.inc;
    unsigned xs = (unsigned) (((old >> 18u) ^ old) >> 27u);
    unsigned rot = (unsigned) (old >> 59u);
    return __funnelshift_r(xs, xs, rot);
#endif
}
__device__ __forceinline__ void pcg_seed(Pcg32 &r, unsigned long long seed, unsigned long long gid) {
    // distinct hashing constants from the oracle's: the two realisations are independent
    r.inc = (gid << 1u) | 1u;
    r.state = mix64(seed + 0xD1B54A32D192ED03ULL * (gid + 1ULL));
#ifndef ERTB_RNG_PCG32
    if (r.state == 0ULL) r.state = 0x9E3779B97F4A7C15ULL;
#endif
    pcg_next(r);
}
// uniform in [0, 1): 24 random bits (MI/src/samplers/independent.cpp:77-86, float variant)
#ifdef ERTB_FLOAT_I2F
// (the reference's conversion: 24 bits through an int -> float conversion and a multiply)
__device__ __forceinline__ float pcg_float(Pcg32 &r) { return (float) (pcg_next(r) >> 8) * 5.9604644775390625e-8f; }
#else
// 23 random mantissa bits under the exponent of 1.0, minus 1: no conversion instruction (quarter-rate pipe); C2 +1.2 %
__device__ __forceinline__ float pcg_float(Pcg32 &r) { return __uint_as_float(0x3f800000u | (pcg_next(r) >> 9)) - 1.f; }
#endif

// ----------------------------------------------------------------------------
// warps (MI/include/mitsuba/core/warp.h)
// ----------------------------------------------------------------------------
// warp.h:54-90 square_to_uniform_disk_concentric
__device__ __forceinline__ void disk_concentric(float u, float v, float &ox, float &oy) {
    float x = fmaf(2.f, u, -1.f), y = fmaf(2.f, v, -1.f);
    bool is_zero = (x == 0.f) && (y == 0.f);
    bool q13 = fabsf(x) < fabsf(y);
    float r = q13 ? y : x, rp = q13 ? x : y;
    float phi = 0.25f * ERTB_PI * __fdividef(rp, r);
    if (q13) phi = 0.5f * ERTB_PI - phi;
    if (is_zero) phi = 0.f;
    float s, c;
    __sincosf(phi, &s, &c);
    ox = r * c;
    oy = r * s;
}
// warp.h:412-433
__device__ __forceinline__ f3 cosine_hemisphere(float u, float v) {
    float x, y;
    disk_concentric(u, v, x, y);
    return mk3(x, y, safe_sqrtf(1.f - x * x - y * y));
}
// warp.h:374-388
__device__ __forceinline__ f3 uniform_hemisphere(float u, float v) {
    float x, y;
    disk_concentric(u, v, x, y);
    float z = 1.f - (x * x + y * y);
    float s = sqrtf(z + 1.f);
    return mk3(x * s, y * s, z);
}

// ----------------------------------------------------------------------------
// tabulated 1D distributions in shared memory (MI/include/mitsuba/core/distr_1d.h)
// ----------------------------------------------------------------------------
// eval_pdf, regular grid on [-1,1] (:370-392), normalised
__device__ __forceinline__ float tab_eval(const float *tb, const ErtbPhaseLeaf &L, float x) {
    const float *pdf = tb + L.off_pdf;
    if (L.off_nodes < 0) {
        if (!(x >= -1.f && x <= 1.f)) return 0.f;
        float xs = (x + 1.f) * L.inv_interval;
        int i = min(max((int) xs, 0), L.n_nodes - 2);
        float w1 = xs - (float) i;
        return fmaf(w1, pdf[i + 1] - pdf[i], pdf[i]) * L.normalization;
    }
    // irregular (:712-735): binary search over the nodes
    const float *nodes = tb + L.off_nodes;
    if (!(x >= nodes[0] && x <= nodes[L.n_nodes - 1])) return 0.f;
    int lo = 0, hi = L.n_nodes - 1; // find last i with nodes[i] < x  (clamped)
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (nodes[mid] < x) lo = mid; else hi = mid;
    }
    float x0 = nodes[lo], x1 = nodes[lo + 1];
    float t = __fdividef(x - x0, x1 - x0);
    return fmaf(t, pdf[lo + 1] - pdf[lo], pdf[lo]) * L.normalization;
}
// sample (:429-459 regular, :790-822 irregular): CDF inversion + linear-segment solve
__device__ __forceinline__ float tab_sample(const float *tb, const ErtbPhaseLeaf &L, float u) {
    const float *pdf = tb + L.off_pdf, *cdf = tb + L.off_cdf;
    float sample = u * L.integral;
    int lo = L.valid0, hi = L.valid1; // first index in [valid0, valid1] with cdf[i] >= sample
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (cdf[mid] < sample) lo = mid + 1; else hi = mid;
    }
    int i = lo;
    float y0 = pdf[i], y1 = pdf[i + 1];
    float c0 = i > 0 ? cdf[i - 1] : 0.f;
    float x0, w;
    if (L.off_nodes < 0) {
        w = __frcp_rn(L.inv_interval);
        x0 = fmaf((float) i, w, -1.f);
    } else {
        const float *nodes = tb + L.off_nodes;
        x0 = nodes[i];
        w = nodes[i + 1] - x0;
    }
    sample = (sample - c0) / w;
    float t_lin = (y0 - safe_sqrtf(fmaf(y0, y0, 2.f * sample * (y1 - y0)))) / (y0 - y1);
    float t_const = sample / y0;
    float t = (y0 == y1) ? t_const : t_lin;
    t = clampf(t, 0.f, 1.f);
    return fmaf(t, w, x0);
}

// ----------------------------------------------------------------------------
// phase functions.  Convention used throughout the kernel: `ct` is the cosine of
// the scattering angle between the propagation direction before and after the
// event (physics convention, +1 = forward).  The reference's "graphics" cosine is
// dot(wo, wi) = -ct.
// ----------------------------------------------------------------------------
// rayleigh.cpp:61-71
__device__ __forceinline__ float rayleigh_eval(float ct, float rho) {
    float r1 = __fdividef(1.f - rho, 1.f + 0.5f * rho), r2 = __fdividef(1.f + rho, 1.f - rho);
    return (3.f / 16.f) * ERTB_INV_PI * r1 * (r2 + ct * ct);
}
__device__ __forceinline__ float rayleigh_pdf(float ct) { return (3.f / 16.f) * ERTB_INV_PI * (1.f + ct * ct); }
// hg.cpp:64-68 with graphics cosine c = -ct
__device__ __forceinline__ float hg_eval(float g, float ct) {
    float temp = fmaf(-2.f * g, ct, 1.f + g * g);
    return ERTB_INV_FOUR_PI * (1.f - g * g) * __frcp_rn(temp * sqrtf(temp));
}

// eval of one leaf (value == pdf except for depolarised Rayleigh)
__device__ __forceinline__ float leaf_eval(const float *tb, const ErtbPhaseLeaf &L, float ct) {
    switch (L.type) {
        case ERTB_PHASE_ISOTROPIC: return ERTB_INV_FOUR_PI;          // isotropic.cpp:52-58
        case ERTB_PHASE_RAYLEIGH_POLARIZED:
        case ERTB_PHASE_RAYLEIGH: return rayleigh_eval(ct, L.p0);    // rayleigh.cpp:97-107
        case ERTB_PHASE_HG: return hg_eval(L.p0, ct);                // hg.cpp:92-99
        default: return tab_eval(tb, L, ct) * ERTB_INV_TWO_PI;       // tabphase.cpp:107-118
    }
}

// pdf of one leaf at the physics-convention cosine: the value, except for (depolarized) Rayleigh (rayleigh.cpp:97-107)
__device__ __forceinline__ float leaf_pdf(const float *tb, const ErtbPhaseLeaf &L, float ct) {
    if (L.type == ERTB_PHASE_RAYLEIGH || L.type == ERTB_PHASE_RAYLEIGH_POLARIZED) return rayleigh_pdf(ct);
    return leaf_eval(tb, L, ct);
}

// sample one leaf: returns ct (physics convention), weight = value/pdf, pdf
__device__ __forceinline__ float leaf_sample(const float *tb, const ErtbPhaseLeaf &L, float u,
                                             float &weight, float &pdf) {
    weight = 1.f;
    switch (L.type) {
        case ERTB_PHASE_ISOTROPIC: { // warp::square_to_uniform_sphere
            pdf = ERTB_INV_FOUR_PI;
            return fmaf(-2.f, u, 1.f);
        }
        case ERTB_PHASE_RAYLEIGH_POLARIZED: // rayleigh_polarized.cpp:133-160: same inversion
        case ERTB_PHASE_RAYLEIGH: { // rayleigh.cpp:75-95: Cardano inversion of the CDF
            float z = 2.f * fmaf(2.f, u, -1.f);
            float tmp = sqrtf(fmaf(z, z, 1.f));
            float A = cbrtf(z + tmp), B = -cbrtf(tmp - z);
            float c = A + B; // cos w.r.t. wi; the lobe is symmetric
            c = clampf(c, -1.f, 1.f);
            pdf = rayleigh_pdf(c);
            weight = __fdividef(rayleigh_eval(c, L.p0), pdf);
            return -c;
        }
        case ERTB_PHASE_HG: { // hg.cpp:70-90
            float g = L.p0;
            float ct;
            if (fabsf(g) < 1.1920929e-7f) {
                ct = fmaf(-2.f, u, 1.f);
            } else {
                float sq = (1.f - g * g) / fmaf(2.f * g, u, 1.f - g);
                ct = (1.f + g * g - sq * sq) / (2.f * g);
            }
            ct = clampf(ct, -1.f, 1.f);
            pdf = hg_eval(g, ct);
            return ct;
        }
        default: { // tabphase.cpp:77-105 / tabphase_irregular.cpp:111-135
            float ct = tab_sample(tb, L, u);
            pdf = tab_eval(tb, L, ct) * ERTB_INV_TWO_PI;
            return ct;
        }
    }
}

// ----------------------------------------------------------------------------
// BSDFs: all land-surface models are azimuthally symmetric, so they are written
// in terms of (cos_i, cos_o, cos_dphi) -- no local frame needed to evaluate them.
// Returned value excludes the foreshortening factor unless stated.
// ----------------------------------------------------------------------------
// ERP/bsdfs/rpv.cpp:128-167
__device__ __forceinline__ float rpv_eval(const float *P, float ci, float co, float cdphi) {
    float rho_0 = P[0], k = P[1], g = P[2], rho_c = P[3];
    float si = safe_sqrtf(1.f - ci * ci), so = safe_sqrtf(1.f - co * co);
    float ti = __fdividef(si, ci), to = __fdividef(so, co);
    float cT = fmaf(si * so, cdphi, ci * co);
    float base = fmaf(2.f * g, cT, 1.f + g * g);
    float F = __fdividef(1.f - g * g, base * sqrtf(base));
    float G = safe_sqrtf(fmaf(-2.f * ti * to, cdphi, ti * ti + to * to));
    float Hs = 1.f + __fdividef(1.f - rho_c, 1.f + G);
    float M = __powf(ci * co * (ci + co), k - 1.f);
    return rho_0 * M * F * Hs * ERTB_INV_PI;
}

// ERP/bsdfs/rtls.cpp:116-243.  sin(dphi) enters squared only.
__device__ __forceinline__ float rtls_eval(const float *P, float cti, float cto, float cdphi) {
    float f_iso = P[0], f_vol = P[1], f_geo = P[2], h = P[3], r = P[4], b = P[5];
    float sti = safe_sqrtf(1.f - cti * cti), sto = safe_sqrtf(1.f - cto * cto);
    float tti = sti / cti, tto = sto / cto;
    float sdphi2 = fmaxf(1.f - cdphi * cdphi, 0.f);
    float cpsi = clampf(fmaf(sti * sto, cdphi, cti * cto), -1.f, 1.f);
    float spsi = sqrtf(fmaxf(1.f - cpsi * cpsi, 0.f)), psi = acosf(cpsi);
    float K_vol = ((0.5f * ERTB_PI - psi) * cpsi + spsi) / (cti + cto) - 0.25f * ERTB_PI;
    float ci = cti, co = cto, ti = tti, to = tto, cp = cpsi;
    if (fabsf(r - b) > 1.1920929e-7f) { // rtls.cpp:197-214
        ti = b / r * tti; to = b / r * tto;
        ci = rsqrtf(fmaf(ti, ti, 1.f)); co = rsqrtf(fmaf(to, to, 1.f)); // cos(atan(x))
        float sip = ti * ci, sop = to * co;
        cp = fmaf(sip * sop, cdphi, ci * co);
    }
    float sec_i = 1.f / ci, sec_o = 1.f / co, sec_sum = sec_i + sec_o;
    float D2 = fmaxf(fmaf(-2.f * ti * to, cdphi, ti * ti + to * to), 0.f);
    float tsp2 = ti * ti * to * to * sdphi2;
    float cos_t = clampf((h / b) * sqrtf(D2 + tsp2) / sec_sum, -1.f, 1.f);
    float t = acosf(cos_t), sin_t = sqrtf(fmaxf(1.f - cos_t * cos_t, 0.f));
    float O = ERTB_INV_PI * (t - sin_t * cos_t) * sec_sum;
    float K_geo = O - sec_sum + 0.5f * (1.f + cp) * sec_i * sec_o;
    return (f_iso + f_vol * K_vol + f_geo * K_geo) * ERTB_INV_PI;
}

// ERP/bsdfs/hapke.cpp:120-332
__device__ __forceinline__ float hapke_H(float w, float x) {
    float gamma = sqrtf(1.f - w), ro = (1.f - gamma) / (1.f + gamma);
    return 1.f / (1.f - w * x * (ro + (1.f - 2.f * ro * x) * 0.5f * logf((1.f + x) / x)));
}
__device__ __forceinline__ float hapke_mu(float tt, float tan_e, float tan_i, float cos_x, float sin_x,
                                          float phi, float opt_cos_phi, float sign) {
    // E1/E2 take tan(angle) directly (tan(atan(x)) = x)
    float chi = rsqrtf(fmaf(ERTB_PI * tt, tt, 1.f));
    float E1e = expf(-2.f * ERTB_INV_PI / tt / tan_e), E1i = expf(-2.f * ERTB_INV_PI / tt / tan_i);
    float E2e = expf(-ERTB_INV_PI / (tt * tt) / (tan_e * tan_e)), E2i = expf(-ERTB_INV_PI / (tt * tt) / (tan_i * tan_i));
    float s2 = sinf(0.5f * phi);
    return chi * (cos_x + sin_x * tt * (opt_cos_phi * E2e + sign * s2 * s2 * E2i) /
                              (2.f - E1e - phi * ERTB_INV_PI * E1i));
}
__device__ __forceinline__ float hapke_eval(const float *P, float mu_0, float mu, float cos_phi) {
    float w = P[0], b = P[1], c = P[2], theta = P[3] * (ERTB_PI / 180.f), B0 = P[4], h = P[5];
    float tt = tanf(theta);
    float sin_i = safe_sqrtf(1.f - mu_0 * mu_0), sin_e = safe_sqrtf(1.f - mu * mu);
    float tan_i = sin_i / mu_0, tan_e = sin_e / mu;
    cos_phi = clampf(cos_phi, -1.f, 1.f);
    float phi = fabsf(acosf(cos_phi));
    bool e_le_i = tan_e <= tan_i; // e <= i  (atan is monotonic)
    float ta = e_le_i ? tan_i : tan_e, tb = e_le_i ? tan_e : tan_i;
    float mu_0eG = hapke_mu(tt, ta, tb, mu_0, sin_i, phi, e_le_i ? 1.f : cos_phi, e_le_i ? -1.f : 1.f);
    float mu_eG = hapke_mu(tt, ta, tb, mu, sin_e, phi, e_le_i ? cos_phi : 1.f, e_le_i ? 1.f : -1.f);
    float mu_ratio = mu_0eG / (mu_0eG + mu_eG) / mu_0;
    float cos_g = clampf(fmaf(sin_i * sin_e, cos_phi, mu_0 * mu), -1.f, 1.f);
    float g = acosf(cos_g);
    float num = 1.f - b * b;
    float d1 = 1.f + 2.f * b * cos_g + b * b, d2 = 1.f - 2.f * b * cos_g + b * b;
    float Pf = (1.f - c) * num / (d1 * sqrtf(d1)) + c * num / (d2 * sqrtf(d2));
    float B = B0 / (1.f + tanf(0.5f * g) / h);
    float M = hapke_H(w, mu_0eG) * hapke_H(w, mu_eG) - 1.f;
    float half = clampf(0.5f * phi, 0.f, 0.5f * ERTB_PI - 1.1920929e-7f);
    float f = expf(-2.f * tanf(half));
    float chi = rsqrtf(fmaf(ERTB_PI * tt, tt, 1.f));
    float E1e = expf(-2.f * ERTB_INV_PI / tt / tan_e), E1i = expf(-2.f * ERTB_INV_PI / tt / tan_i);
    float E2e = expf(-ERTB_INV_PI / (tt * tt) / (tan_e * tan_e)), E2i = expf(-ERTB_INV_PI / (tt * tt) / (tan_i * tan_i));
    float eta_0e = chi * (mu_0 + sin_i * tt * E2i / (2.f - E1i));
    float eta_e = chi * (mu + sin_e * tt * E2e / (2.f - E1e));
    bool e_lt_i = tan_e < tan_i;
    float opt_mu = e_lt_i ? mu : mu_0, opt_eta = e_lt_i ? eta_e : eta_0e;
    float S = (mu_eG * mu_0 * chi) / (eta_e * eta_0e * (1.f - f + f * chi * opt_mu / opt_eta));
    return w * 0.25f * ERTB_INV_PI * mu_ratio * (Pf * (1.f + B) + M) * S;
}

// BSDF value WITHOUT the cosine factor for the azimuthally symmetric models.
__device__ __forceinline__ float bsdf_f(const ErtbParams &P, float ci, float co, float cdphi) {
    switch (P.bsdf_type) {
        case ERTB_BSDF_DIFFUSE: return P.bsdf[0] * ERTB_INV_PI;          // diffuse.cpp:127-143
        case ERTB_BSDF_RPV: return rpv_eval(P.bsdf, ci, co, cdphi);       // rpv.cpp:169-181
        case ERTB_BSDF_RTLS: return rtls_eval(P.bsdf, ci, co, cdphi);     // rtls.cpp:245-257
        case ERTB_BSDF_HAPKE: return hapke_eval(P.bsdf, ci, co, cdphi);   // hapke.cpp:334-347
        default: return 0.f;
    }
}

// cos(dphi) between two unit vectors given their cosines with the normal and their dot product
__device__ __forceinline__ float cos_dphi(float ci, float co, float wi_dot_wo) {
    float si2 = fmaxf(1.f - ci * ci, 0.f), so2 = fmaxf(1.f - co * co, 0.f);
    float den = si2 * so2;
    // Frame3f::sincos_phi returns (0, 1) for a vector along the normal -> dphi = phi_other
    if (den < 1e-14f) return 1.f;
    return clampf((wi_dot_wo - ci * co) * rsqrtf(den), -1.f, 1.f);
}
