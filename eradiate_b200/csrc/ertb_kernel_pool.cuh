// ertb_kernel_pool.cuh -- wavefront megakernel with warp-private shared-memory path pools.
//
// Same estimator as ertb_kernel.cuh (null-collision volpath, MI/src/integrators/
// volpath.cpp:93-572); different execution model.  The register-resident kernel keeps
// one path per lane, so a warp can only batch 32 paths per phase and runs with ~12-19 of
// 32 lanes active (profiles/r01b).  Here every warp owns a pool of ERTB_POOL_NS path
// records laid out SoA in shared memory (field-major: conflict-free for consecutive
// slots).  Each trip the warp
//   1. reads the mode word of every slot and ballots the per-phase membership masks,
//   2. picks ONE phase -- the free-flight walk while >= tw records can walk, otherwise
//      the event phase (surface / scatter / regenerate) holding the most records,
//   3. compacts up to 32 member slots into a list (ballot + popc ranks) so that lane i
//      works on the i-th member: lanes stay dense whatever the divergence of path depths,
//   4. loads only the fields that phase needs, runs it, stores the fields it changed.
// The walk phase keeps stepping the records it loaded while enough of them are still
// walking, which amortises the load/store over several free flights.  Path state moves
// through shared memory only; HBM sees the film atomics and the work-queue counter.
//
// PW = true is the piecewise integrator (ERP/integrators/piecewise_volpath.cpp:91-527): there
// are no null collisions, so the walk phase disappears.  Every event (and the regeneration)
// ends with the analytic free flight to the NEXT event (ertb_piecewise.cuh), and the shadow ray
// of an event is one exact transmittance evaluation instead of a ratio-tracking walk; a record
// is therefore always in one of the three event modes and each trip of a lane is one real event.
#pragma once

#include <cuda_fp16.h>

#include "ertb_kernel.cuh"
#include "ertb_polar.cuh"
#include "ertb_piecewise.cuh"

#ifndef ERTB_POOL_NS
#define ERTB_POOL_NS 64 // records per warp (multiple of 32)
#endif
// Launch shape of the scalar instances: 2 CTAs of 448 threads = 28 resident warps at 72 registers.  Round 1 ran 6 CTAs
// of 128 threads (24 warps, 80 registers; registers and shared memory both allowed exactly 6).  A third of the stall
// samples of the C2 launch are fixed-latency waits (profiles/r02u), which more warps hide; warps cost pool space, and
// big CTAs pay for it by staging two copies of the table blob instead of six.  Measured (STATS instances, Mpaths/s;
// C2 / C3 banded / plane-parallel AFGL + RPV): 128 x 6: 9214 / 3535 / 7488; 128 x 7: 9622 / - / 7433; 256 x 4: 9476 /
// 3599 / 6822; 448 x 2: 9613 / 4042 / 7448; 512 x 2: 9470 / 3973 / 6864; 896 x 1: 9677 / - / 7394; 1024 x 1: 9616 / - /
// 6772; banded 256 x 3: 3855, 384 x 2: 3872, 320 x 3: 3898, 192 x 4: 3862.
#ifndef ERTB_POOL_BLOCK
#define ERTB_POOL_BLOCK 448
#endif
#ifndef ERTB_POOL_MINB
#define ERTB_POOL_MINB 2
#endif
// Polarized instances: the table blob is replicated per CTA, so a few big CTAs leave more shared memory to the
// pools than many small ones (warps never synchronise with each other after the tables are staged)
#ifndef ERTB_POOL_BLOCK_POL
#define ERTB_POOL_BLOCK_POL 640 // one CTA of 20 warps per SM (<= 102 registers): C5 +7 % over 4 CTAs of 4 warps
#endif
#ifndef ERTB_POOL_MINB_POL
#define ERTB_POOL_MINB_POL 1 // polarized instances: CTAs per SM at ERTB_POOL_BLOCK_POL threads
#endif
// Banded instances (the record is two fields longer and the table blob holds the aerosol tables as well): fewer, bigger
// CTAs stage fewer copies of the blob and leave room for more warps
#ifndef ERTB_POOL_BLOCK_BANDS
#define ERTB_POOL_BLOCK_BANDS ERTB_POOL_BLOCK
#endif
#ifndef ERTB_POOL_MINB_BANDS
#define ERTB_POOL_MINB_BANDS ERTB_POOL_MINB
#endif
// launch shape of an instance
__host__ __device__ constexpr int ertb_pool_block(bool pol, bool bands) {
    return pol ? ERTB_POOL_BLOCK_POL : (bands ? ERTB_POOL_BLOCK_BANDS : ERTB_POOL_BLOCK);
}
__host__ __device__ constexpr int ertb_pool_minb(bool pol, bool bands) {
    return pol ? ERTB_POOL_MINB_POL : (bands ? ERTB_POOL_MINB_BANDS : ERTB_POOL_MINB);
}
#ifndef ERTB_WALK_UNROLL
#define ERTB_WALK_UNROLL 2 // free flights per loop-control vote in the walk phase
#endif
#ifndef ERTB_POOL_NS_POL
#define ERTB_POOL_NS_POL 64 // records per warp of the polarized instances
#endif
// Polarized records keep the normalised Mueller throughput and the Q/I, U/I, V/I ratios of the pending
// next-event term as fp16 pairs (entries are O(1) by construction; storage rounding is unbiased and 5e-4
// relative, far below the Monte Carlo noise of Q, U, V): 13 extra words per record instead of 22, which is
// what lets a fourth CTA fit into an SM's shared memory.  The intensity path (thr, wnee, res) stays fp32.
#ifndef ERTB_POL_PACK16
#define ERTB_POL_PACK16 1
#endif

enum : int {
    PF_FLAGS = 0, PF_H0, PF_B, PF_S, PF_SMAX, PF_THR, PF_WNEE, PF_RES, PF_RNG0, PF_RNG1,
#ifdef ERTB_RNG_PCG32
    PF_INC0, PF_INC1, // (PCG32 carries a per-stream increment; xoroshiro64** has 64 bits of state in all)
#endif
    PF_B2, PF_SMAX2, PF_N0X, PF_N0Y, PF_N0Z, PF_DX, PF_DY, PF_DZ, PF_PIX, PF_WRAY, PF_COUNT,
    // polarized records only: throughput Mueller matrix normalised by its (0,0) entry (PF_THR holds
    // that entry, so the walk phase is identical in both modes), Q/I U/I V/I of the pending NEE
    // Stokes vector, and the Q, U, V components of the result (PF_RES holds I)
#if ERTB_POL_PACK16
    PF_T0 = PF_COUNT, PF_QN0 = PF_T0 + 8, PF_RQ = PF_QN0 + 2, PF_COUNT_POL = PF_RQ + 3
#else
    PF_T0 = PF_COUNT, PF_QN0 = PF_T0 + 16, PF_RQ = PF_QN0 + 3, PF_COUNT_POL = PF_RQ + 3
#endif
};
enum : unsigned {
    PM_DEAD = 0, PM_IDLE = 1, PM_WALK_MAIN = 2, PM_WALK_NEE = 3, PM_SURF = 4, PM_SCAT = 5,
    PFL_MODE_MASK = 7u, PFL_KIND = 8u, PFL_KIND2 = 16u, PFL_LAST_NULL = 32u, PFL_VACUUM = 64u,
    PFL_NEE_QUV = 128u, // polarized: PF_WNEE holds a finished NEE term whose Q, U, V are still to be added
    PFL_DEPTH_SHIFT = 8
};

// BANDS instances append two fields to the record: the distance of the next band boundary along the
// segment, and (band index | descending << 8)
// (`threads`: CTA size of the launch, 0 = the instance's own; the host launches smaller CTAs when the tables are large)
__host__ __device__ inline size_t ertb_pool_smem_bytes(size_t blob_bytes, bool pol = false, bool bands = false, int threads = 0) {
    size_t blob = (blob_bytes + 15) & ~size_t(15);
    size_t warps = (size_t) (threads > 0 ? threads : ertb_pool_block(pol, bands)) / 32;
    return blob + warps * (size_t) ((pol ? PF_COUNT_POL : PF_COUNT) + (bands ? 2 : 0)) * (pol ? ERTB_POOL_NS_POL : ERTB_POOL_NS) * 4 + warps * 32 * 4;
}

// free flight that follows a regeneration or an event of the piecewise integrator
// (piecewise_volpath.cpp:215-246): returns the mode of the next event of the record
__device__ __forceinline__ unsigned pw_advance(const ErtbParams &P, const float *tb, Pcg32 &rng, float &h0, float mu) {
    float E = -__logf(1.f - pcg_float(rng));
    float s, h;
    int r = pw_flight(P, tb, h0, mu, E, s, h);
    if (r == PW_COLLISION) { h0 = h; return PM_SCAT; }
    return r == PW_GROUND ? PM_SURF : PM_IDLE;
}

// ----------------------------------------------------------------------------
// Banded majorant (BANDS instances). The reference samples free flights against ONE majorant,
// scale * max(sigma_t) over the whole grid (heterogeneous.cpp:163): an aerosol layer of optical
// depth 0.5 between 1 and 2 km makes every kilometre of the 120 km column as expensive as a
// kilometre of aerosol -- 476 loop trips per path on BASELINE C3, > 95 % of them null collisions.
// Delta and ratio tracking stay unbiased with ANY majorant that bounds sigma_t locally, and free
// flights are memoryless, so the host cuts the layer stack into a few altitude bands (dynamic
// programme over the profile, ertb_cuda.cu) and the walk samples against the majorant of the band
// it is in; a flight that would leave the band stops at the boundary and continues from there
// with the next band's majorant. Same estimator in expectation (lower variance for the ratio
// tracker), an order of magnitude fewer trips when the profile has a thin dense layer; with one
// band (clear-sky profiles such as C2) this is the reference's walk, trip for trip, and the
// BANDS = false instances are used.
// ----------------------------------------------------------------------------
// distance, along the segment (h0, b), at which a record in band `bi` reaches the band's boundary
// (`down`: it is descending; cleared when a spherical ray passes its perigee inside the band)
// band holding altitude h (<= 8 bands: a handful of compares, cheaper than a per-layer table in shared memory)
__device__ __forceinline__ int band_of(const ErtbParams &P, const float *tb, float h) {
    const float *lo = tb + P.off_band_lo;
    int bi = 0;
    for (int k = 1; k < P.n_bands; ++k) bi += h >= lo[k] ? 1 : 0;
    return bi;
}

// Distance along the segment (h0, b) at which a record in band `bi` leaves it. `down`: the record is
// descending; cleared when a spherical ray passes its perigee inside the band (it then leaves through the
// band's top). Spherical shell: the boundary at altitude hk is met where s (s + 2 b) = c,
// c = (hk - h0)(hk + h0 + 2 R) -- the cancellation-free roots of segment_setup.
template <bool SPH>
__device__ __forceinline__ float band_exit(const ErtbParams &P, const float *tb, float h0, float b, int bi, bool &down,
                                           float smax) {
    const float *lo = tb + P.off_band_lo;
    if (down) {
        const float hk = lo[bi]; // (lo[0] = 0: band 0 ends at the ground)
        if (SPH) {
            const float c = (hk - h0) * (hk + h0 + 2.f * P.R); // <= 0
            const float disc = fmaf(b, b, c);
            if (disc >= 0.f) return bi > 0 ? fminf(__fdividef(-c, fast_sqrt(disc) - b), smax) : smax;
            down = false; // perigee inside this band
        } else {
            return bi > 0 ? fminf(__fdividef(hk - h0, b), smax) : smax;
        }
    }
    if (bi >= P.n_bands - 1) return smax;
    const float hk = lo[bi + 1];
    if (SPH) {
        const float c = fmaxf((hk - h0) * (hk + h0 + 2.f * P.R), 0.f);
        const float sq = fast_sqrt(fmaf(b, b, c));
        return fminf(b > 0.f ? __fdividef(c, b + sq) : sq - b, smax);
    }
    return fminf(__fdividef(hk - h0, b), smax);
}

// polarized record fields (field-major pool `wp` of NS slots): normalised throughput and NEE ratios
template <int NS>
__device__ __forceinline__ void pol_load_T(const float *wp, int slot, float scale, float *T) {
#if ERTB_POL_PACK16
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float2 v = __half22float2(*reinterpret_cast<const __half2 *>(&wp[(PF_T0 + k) * NS + slot]));
        T[2 * k] = scale * v.x; T[2 * k + 1] = scale * v.y;
    }
#else
#pragma unroll
    for (int k = 0; k < 16; ++k) T[k] = scale * wp[(PF_T0 + k) * NS + slot];
#endif
}
template <int NS>
__device__ __forceinline__ void pol_store_T(float *wp, int slot, const float *T, float scale) {
#if ERTB_POL_PACK16
#pragma unroll
    for (int k = 0; k < 8; ++k)
        *reinterpret_cast<__half2 *>(&wp[(PF_T0 + k) * NS + slot]) = __floats2half2_rn(T[2 * k] * scale, T[2 * k + 1] * scale);
#else
#pragma unroll
    for (int k = 0; k < 16; ++k) wp[(PF_T0 + k) * NS + slot] = T[k] * scale;
#endif
}
template <int NS>
__device__ __forceinline__ void pol_load_qn(const float *wp, int slot, float *qn) {
#if ERTB_POL_PACK16
    const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&wp[PF_QN0 * NS + slot]));
    const float2 b = __half22float2(*reinterpret_cast<const __half2 *>(&wp[(PF_QN0 + 1) * NS + slot]));
    qn[0] = a.x; qn[1] = a.y; qn[2] = b.x;
#else
#pragma unroll
    for (int k = 0; k < 3; ++k) qn[k] = wp[(PF_QN0 + k) * NS + slot];
#endif
}
template <int NS>
__device__ __forceinline__ void pol_store_qn(float *wp, int slot, const float *qn) {
#if ERTB_POL_PACK16
    *reinterpret_cast<__half2 *>(&wp[PF_QN0 * NS + slot]) = __floats2half2_rn(qn[0], qn[1]);
    *reinterpret_cast<__half2 *>(&wp[(PF_QN0 + 1) * NS + slot]) = __floats2half2_rn(qn[2], 0.f);
#else
#pragma unroll
    for (int k = 0; k < 3; ++k) wp[(PF_QN0 + k) * NS + slot] = qn[k];
#endif
}

#ifndef ERTB_FLUSH_COLLECTIVE_MIN
#define ERTB_FLUSH_COLLECTIVE_MIN 6
#endif

__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

// Film flush. Per-lane atomics made small renders (a few hundred paths per warp) spend most of
// their time serialised in the L2 atomic unit: every lane of every warp ends up adding to the
// same few film addresses. When many lanes flush at once (a whole chunk finishing, the end of
// the kernel) the flush is therefore a warp collective (all 32 lanes call it): the participating
// lanes' float64 sums are grouped by pixel, reduced with shuffles, and one lane per pixel issues
// the atomics. A lone straggler switching pixel still flushes on its own -- the collective costs
// ~100 warp instructions whatever the number of participants.
template <bool POL>
__device__ __forceinline__ void film_flush_lane(const ErtbParams &P, unsigned pix, double wl, double l, double l2,
                                                double q, double u, double v) {
    atomicAdd(&P.accum[pix], wl);
    atomicAdd(&P.accum[P.n_pixels + pix], l);
    atomicAdd(&P.accum[2u * P.n_pixels + pix], l2);
    if (POL) {
        atomicAdd(&P.accum[3u * P.n_pixels + pix], wl);
        atomicAdd(&P.accum[4u * P.n_pixels + pix], q);
        atomicAdd(&P.accum[5u * P.n_pixels + pix], u);
        atomicAdd(&P.accum[6u * P.n_pixels + pix], v);
    }
}

template <bool POL>
__device__ __noinline__ void film_flush_warp(const ErtbParams &P, unsigned lane, bool mine, unsigned acc_pix,
                                             double acc_wl, double acc_l, double acc_l2, double acc_q, double acc_u,
                                             double acc_v) {
    if (!mine) acc_pix = 0xffffffffu;
    unsigned todo = __ballot_sync(0xffffffffu, acc_pix != 0xffffffffu);
    while (todo) {
        const int leader = __ffs(todo) - 1;
        const unsigned pix = __shfl_sync(0xffffffffu, acc_pix, leader);
        const bool in = acc_pix == pix;
        double wl = warp_sum(in ? acc_wl : 0.0), l = warp_sum(in ? acc_l : 0.0), l2 = warp_sum(in ? acc_l2 : 0.0);
        double q = 0.0, u = 0.0, v = 0.0;
        if (POL) { q = warp_sum(in ? acc_q : 0.0); u = warp_sum(in ? acc_u : 0.0); v = warp_sum(in ? acc_v : 0.0); }
        if ((int) lane == leader) film_flush_lane<POL>(P, pix, wl, l, l2, q, u, v);
        todo &= ~__ballot_sync(0xffffffffu, in);
    }
}

// COLL: collective mid-kernel film flushes, for renders whose chunks are small (many pixel
// switches per lane). It is a template parameter because any call inside the scheduler loop
// perturbs the register allocation of the walk phase (-3..7 % on the C2 headline, measured).
// GEN: general primary rays -- `mradiancemeter` (explicit rays whose origin may lie inside the atmosphere:
// class 3 of the per-pixel table carries the start altitude) and `mpdistant` (the film sample picks the
// target point). A template parameter for the same reason as COLL: the C2 instance must not change.
template <bool SPH, bool STATS, bool POL, bool PW = false, bool COLL = false, bool BANDS = false, bool GEN = false>
__global__ void __launch_bounds__(ertb_pool_block(POL, BANDS), ertb_pool_minb(POL, BANDS)) ertb_render_pool_kernel(const ErtbParams P) {
    static_assert(!(PW && BANDS), "the piecewise integrator has no null collisions");
    static_assert(!(PW && SPH), "the piecewise medium is a plane-parallel layer stack");
    constexpr int NS = POL ? ERTB_POOL_NS_POL : ERTB_POOL_NS; // records per warp
    constexpr int NK = NS / 32;
#ifdef ERTB_WALK_UNROLL_ALL
    constexpr int WALK_UNROLL = ERTB_WALK_UNROLL;
#else
    constexpr int WALK_UNROLL = BANDS ? 1 : ERTB_WALK_UNROLL; // (see the walk phase)
#endif
    constexpr int NF = (POL ? PF_COUNT_POL : PF_COUNT) + (BANDS ? 2 : 0);
    constexpr int PF_SB = NF - 2, PF_BAND = NF - 1; // (BANDS only)
    extern __shared__ __align__(16) float smem[];
    __shared__ __align__(8) unsigned long long mbar;
    float *tb = smem; // table blob first
    const unsigned blob_words = ((unsigned) P.blob_bytes + 15u) / 16u * 4u;
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    float *wp = smem + blob_words + warp * (NF * NS);
    unsigned *wpu = reinterpret_cast<unsigned *>(wp);
    const int NW = (int) (blockDim.x >> 5); // warps per CTA (<= the instance's launch bound: large tables get smaller CTAs)
    int *list = reinterpret_cast<int *>(smem + blob_words + NW * (NF * NS)) + warp * 32;

    if (P.blob_bytes > 0) tma_stage(tb, P.blob, (unsigned) P.blob_bytes, &mbar);

    const f3 sun = mk3(P.sun[0], P.sun[1], P.sun[2]);

    // every record starts "finished with nothing to accumulate"
#pragma unroll
    for (int j = 0; j < NK; ++j) {
        wpu[PF_FLAGS * NS + j * 32 + lane] = PM_IDLE;
        wpu[PF_PIX * NS + j * 32 + lane] = 0xffffffffu;
        wp[PF_RES * NS + j * 32 + lane] = 0.f;
        wp[PF_WRAY * NS + j * 32 + lane] = 0.f;
    }
    __syncwarp();

    // lane-local film accumulators, tagged with the pixel they belong to
    double acc_wl = 0.0, acc_l = 0.0, acc_l2 = 0.0;
    double acc_q = 0.0, acc_u = 0.0, acc_v = 0.0; // polarized: sums of w * (S1, S2, S3)
    unsigned acc_pix = 0xffffffffu;
    // warp-uniform work-queue cursor
    unsigned long long cur_next = 0, cur_end = 0;
    unsigned cur_pix = 0;
    bool exhausted = false;
    unsigned st_main = 0, st_nee = 0, st_scatter = 0, st_surface = 0, st_paths = 0;

#define FLD(f, slot) wp[(f) * NS + (slot)]
#define FLDU(f, slot) wpu[(f) * NS + (slot)]

    for (;;) {
        // ---- 1. membership masks -------------------------------------------------
        unsigned md[NK];
        unsigned sel[NK];
        int n_walk = 0;
#pragma unroll
        for (int j = 0; j < NK; ++j) {
            md[j] = FLDU(PF_FLAGS, j * 32 + lane) & PFL_MODE_MASK;
            sel[j] = __ballot_sync(0xffffffffu, md[j] == PM_WALK_MAIN || md[j] == PM_WALK_NEE);
            n_walk += __popc(sel[j]);
        }
        // ---- 2. phase selection ----------------------------------------------------
        unsigned phase = PM_WALK_MAIN; // walk
        int n_sel = n_walk;
        if (PW || n_walk < P.tw) {
            unsigned ms[NK], mc[NK], mi[NK];
            int ns = 0, nc = 0, ni = 0;
#pragma unroll
            for (int j = 0; j < NK; ++j) {
                ms[j] = __ballot_sync(0xffffffffu, md[j] == PM_SURF);
                mc[j] = __ballot_sync(0xffffffffu, md[j] == PM_SCAT);
                mi[j] = __ballot_sync(0xffffffffu, md[j] == PM_IDLE);
                ns += __popc(ms[j]); nc += __popc(mc[j]); ni += __popc(mi[j]);
            }
            int best = n_walk; // an event phase must hold more records than can walk
            if (ni > best) { best = ni; phase = PM_IDLE; }
            if (ns > best) { best = ns; phase = PM_SURF; }
            if (nc > best) { best = nc; phase = PM_SCAT; }
            if (best == 0) break; // every record is dead: done
            if (phase != PM_WALK_MAIN) {
                n_sel = best;
#pragma unroll
                for (int j = 0; j < NK; ++j)
                    sel[j] = phase == PM_IDLE ? mi[j] : (phase == PM_SURF ? ms[j] : mc[j]);
            }
        }
        // ---- 3. compaction: lane i <- i-th member slot -------------------------------
        {
            int base = 0;
#pragma unroll
            for (int j = 0; j < NK; ++j) {
                int pos = base + __popc(sel[j] & lt_mask);
                if (((sel[j] >> lane) & 1u) && pos < 32) list[pos] = j * 32 + (int) lane;
                base += __popc(sel[j]);
            }
        }
        __syncwarp();
        if (n_sel > 32) n_sel = 32;
        const bool have = (int) lane < n_sel;
        const int slot = have ? list[lane] : 0;
        __syncwarp();

        if (!PW && phase == PM_WALK_MAIN) {
            // =====================================================================
            // free-flight walk (delta tracking / ratio tracking), medium.cpp:42-82
            // =====================================================================
            unsigned flags = 0;
            float h0 = 0.f, b = 0.f, s = 0.f, smax = 0.f, thr = 0.f, wnee = 0.f, res = 0.f, b2 = 0.f, smax2 = 0.f;
            Pcg32 rng; rng.state = 0; rng.inc = 1;
            if (have) {
                flags = FLDU(PF_FLAGS, slot);
                h0 = FLD(PF_H0, slot); b = FLD(PF_B, slot); s = FLD(PF_S, slot); smax = FLD(PF_SMAX, slot);
                thr = FLD(PF_THR, slot); wnee = FLD(PF_WNEE, slot); res = FLD(PF_RES, slot);
                b2 = FLD(PF_B2, slot); smax2 = FLD(PF_SMAX2, slot);
                rng.state = (unsigned long long) FLDU(PF_RNG0, slot) | ((unsigned long long) FLDU(PF_RNG1, slot) << 32);
#ifdef ERTB_RNG_PCG32
                rng.inc = (unsigned long long) FLDU(PF_INC0, slot) | ((unsigned long long) FLDU(PF_INC1, slot) << 32);
#endif
            }
            float sb = 0.f;     // BANDS: where the current band ends along the segment
            unsigned band = 0u; //        band index | descending << 8
            if (BANDS && have) { sb = FLD(PF_SB, slot); band = FLDU(PF_BAND, slot); }
            unsigned mode = have ? (flags & PFL_MODE_MASK) : PM_DEAD;
            const int keep = max(1, min(P.twi, n_sel));
            // depth is constant during a walk: hoist the Russian-roulette / max-depth predicates
            const bool rr_on = !P.mis && (flags >> PFL_DEPTH_SHIFT) > P.rr_depth;
            const bool depth_done = (flags >> PFL_DEPTH_SHIFT) >= P.max_depth;
            // shared-memory address of the p_real table, formed once: indexed with a per-lane layer the generic
            // pointer made the compiler rebuild the window base (S2UR + 3 uniform ops) on every trip
            const unsigned preal_s = smem_u32(tb + P.off_preal);
            for (;;) {
                if (__popc(__ballot_sync(0xffffffffu, mode == PM_WALK_MAIN || mode == PM_WALK_NEE)) < keep) break;
                // ERTB_WALK_UNROLL free flights per vote: the vote + count + compare + branch of the loop control
                // are a fifth of a trip's instructions; a lane that stops walking idles for at most one extra trip
                // (C2 +4.8 % at 2, nothing more at 3, -5 % at 4; the banded walk, whose trip holds the band loop, loses
                // 5-18 % when unrolled: it keeps one flight per vote)
#pragma unroll
                for (int rep = 0; rep < WALK_UNROLL; ++rep)
                if (mode == PM_WALK_MAIN || mode == PM_WALK_NEE) {
                    const bool is_main = mode == PM_WALK_MAIN;
                    bool alive = true;
                    if (is_main && rr_on) { // volpath.cpp:194-198: every main loop trip
                        float q = fminf(thr, 0.95f);
                        if (pcg_float(rng) >= q) { thr = 0.f; mode = PM_IDLE; alive = false; }
                        else thr = __fdividef(thr, q);
                    }
                    if (alive) {
                        if (STATS) { if (is_main) st_main++; else st_nee++; }
                        float bratio = 1.f;
                        const float E = -__logf(1.f - pcg_float(rng)); // majorant optical depth of this flight
                        bool nee_over = false;
                        if (!BANDS) {
                            s = fmaf(E, P.inv_majorant, s);
                        } else {
                            if (s == 0.f) { // a new segment (every event starts one with s = 0) in the event's band
                                const int bi = (int) ((band >> 16) & 255u);
                                bool down = b < 0.f;
                                sb = band_exit<SPH>(P, tb, h0, b, bi, down, smax);
                                band = (band & 0xff0000u) | (unsigned) bi | (down ? 256u : 0u);
                            }
                            // Spend E band by band: a flight that outlives the band it is in pays that band's
                            // share (majorant x remaining length) and carries on in the next band with its
                            // majorant -- the tentative collision falls where the piecewise-constant majorant
                            // optical depth along the segment reaches E (no event and no random number at a
                            // boundary; a crossing costs one boundary distance).
                            float Er = E * P.inv_majorant; // remaining depth, in metres at the global majorant
#pragma unroll 1
                            for (int it = 2 * P.n_bands + 2; it > 0; --it) {
                                const unsigned bi = band & 255u;
                                bratio = tb[P.off_band_ratio + bi];
                                const float sn = fmaf(Er, bratio, s);
                                if (sn < sb || !(sb < smax)) { s = sn; break; }
                                Er = fmaxf(fmaf(s - sb, tb[P.off_band_iratio + bi], Er), 0.f);
                                s = sb;
                                bool down = (band & 256u) != 0u;
                                const int nb = (int) bi + (down ? -1 : 1);
                                sb = band_exit<SPH>(P, tb, h0, b, nb, down, smax);
                                band = (band & 0xff0000u) | (unsigned) nb | (down ? 256u : 0u);
                            }
                        }
                        if (!(s < smax)) {
                            if (is_main) {
                                mode = (flags & PFL_KIND) ? PM_SURF : PM_IDLE; // ground hit / left through the TOA
                            } else {
                                res += wnee; // shadow ray reached the TOA (I component)
                                nee_over = true;
                                if (POL) flags |= PFL_NEE_QUV; // Q, U, V = wnee * (stored ratios), added by the next event
                                else wnee = 0.f;
                            }
                        } else {
                            float h = altitude_at<SPH>(P, h0, b, s);
                            float preal;
                            asm("ld.shared.f32 %0, [%1];" : "=f"(preal) : "r"(preal_s + 4u * (unsigned) layer_of(P, h)));
                            if (BANDS) preal = fminf(preal * bratio, 1.f); // relative to the band's majorant
                            if (is_main) {
                                if (pcg_float(rng) >= 1.f - preal) mode = PM_SCAT; // real collision
                                else flags |= PFL_LAST_NULL;
                            } else {
                                wnee *= 1.f - preal; // ratio tracking
                                nee_over = wnee == 0.f;
                            }
                        }
                        if (nee_over) {
                            // NEE walk over: continue with the main segment prepared by the event
                            if (thr == 0.f || depth_done) mode = PM_IDLE;
                            else {
                                b = b2; smax = smax2; s = 0.f;
                                flags = (flags & ~PFL_KIND) | ((flags & PFL_KIND2) ? PFL_KIND : 0u);
                                mode = PM_WALK_MAIN;
                            }
                        }
                    }
                }
            }
            if (have) {
                FLDU(PF_FLAGS, slot) = (flags & ~PFL_MODE_MASK) | mode;
                FLD(PF_B, slot) = b; FLD(PF_S, slot) = s; FLD(PF_SMAX, slot) = smax;
                FLD(PF_THR, slot) = thr; FLD(PF_WNEE, slot) = wnee; FLD(PF_RES, slot) = res;
                FLDU(PF_RNG0, slot) = (unsigned) rng.state; FLDU(PF_RNG1, slot) = (unsigned) (rng.state >> 32);
                if (BANDS) { FLD(PF_SB, slot) = sb; FLDU(PF_BAND, slot) = band; }
            }
            __syncwarp();
            continue;
        }

        if (phase == PM_IDLE) {
            // =====================================================================
            // finish the previous path of the record + regenerate (integrator.cpp:449-520)
            // =====================================================================
            {
                // lanes about to accumulate a different pixel flush their sums first
                unsigned pix_chk = have ? FLDU(PF_PIX, slot) : 0xffffffffu;
                const bool sw = pix_chk != 0xffffffffu && acc_pix != 0xffffffffu && pix_chk != acc_pix;
                if (COLL) {
                    if (__popc(__ballot_sync(0xffffffffu, sw)) >= ERTB_FLUSH_COLLECTIVE_MIN)
                        film_flush_warp<POL>(P, lane, sw, acc_pix, acc_wl, acc_l, acc_l2, acc_q, acc_u, acc_v);
                    else if (sw)
                        film_flush_lane<POL>(P, acc_pix, acc_wl, acc_l, acc_l2, acc_q, acc_u, acc_v);
                } else if (sw) {
                    film_flush_lane<POL>(P, acc_pix, acc_wl, acc_l, acc_l2, acc_q, acc_u, acc_v);
                }
                if (sw) {
                    acc_wl = acc_l = acc_l2 = 0.0;
                    acc_q = acc_u = acc_v = 0.0;
                    acc_pix = 0xffffffffu;
                }
            }
            if (have) {
                unsigned pix_old = FLDU(PF_PIX, slot);
                if (pix_old != 0xffffffffu) {
                    acc_pix = pix_old; // (a lane holding another pixel's sums was flushed just above)
                    float r = FLD(PF_RES, slot);
                    float wr = FLD(PF_WRAY, slot);
                    if (GEN && P.astro_radiance > 0.f && (FLDU(PF_FLAGS, slot) >> PFL_DEPTH_SHIFT) == 0u) {
                        // astroobject: a primary ray that left the scene without any event looks at the sky and
                        // sees the disc if it points into it (volpath.cpp:328-346, count_direct; throughput 1)
                        f3 dd = mk3(FLD(PF_DX, slot), FLD(PF_DY, slot), FLD(PF_DZ, slot));
                        f3 cx = cross3(dd, sun);
                        if (dot3(dd, sun) > 0.f && dot3(cx, cx) < P.astro_sin2) r += P.astro_radiance;
                    }
                    acc_wl += (double) (wr * r);
                    acc_l += (double) r;
                    acc_l2 += (double) r * (double) r;
                    if (POL) {
                        float rq = FLD(PF_RQ, slot), ru = FLD(PF_RQ + 1, slot), rv = FLD(PF_RQ + 2, slot);
                        if (FLDU(PF_FLAGS, slot) & PFL_NEE_QUV) {
                            float wn = FLD(PF_WNEE, slot), q3[3];
                            pol_load_qn<NS>(wp, slot, q3);
                            rq = fmaf(wn, q3[0], rq);
                            ru = fmaf(wn, q3[1], ru);
                            rv = fmaf(wn, q3[2], rv);
                        }
                        acc_q += (double) (wr * rq); acc_u += (double) (wr * ru); acc_v += (double) (wr * rv);
                    }
                }
            }
            // warp-aggregated pops from the chunk queue
            bool got = false;
            unsigned long long my_sample = 0;
            unsigned my_pix = 0;
            unsigned pending = __ballot_sync(0xffffffffu, have);
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
                if (mine && rank < take) { got = true; my_sample = cur_next + rank; my_pix = cur_pix; }
                cur_next += take;
                pending &= ~__ballot_sync(0xffffffffu, mine && rank < take);
            }
            if (have && !got) {
                FLDU(PF_FLAGS, slot) = PM_DEAD;
                FLDU(PF_PIX, slot) = 0xffffffffu;
            }
            if (got) {
                const unsigned pix = my_pix;
                Pcg32 rng;
                unsigned long long gid = ((unsigned long long) pix << 40) + (P.sample_offset + my_sample);
                pcg_seed(rng, P.seed, gid);
                if (STATS) st_paths++;
                float thr = 1.f, wray = 1.f;
                f3 n0 = mk3(0.f, 0.f, 1.f), d = mk3(0.f, 0.f, -1.f);
                const ErtbSensor &S = P.sensor;
                int valid = 1;
                float h_start = P.H;
                if (S.use_table) {
                    const float4 *t4 = reinterpret_cast<const float4 *>(S.table) + 2u * pix;
                    float4 a = __ldg(t4), c4 = __ldg(t4 + 1);
                    n0 = mk3(a.x, a.y, a.z);
                    d = mk3(a.w, c4.x, c4.y);
                    valid = (int) c4.z;
                    if (GEN) h_start = c4.w;
                } else {
                    unsigned px = pix % (unsigned) S.width, py = pix / (unsigned) S.width;
                    float fx = __fdividef((float) px + pcg_float(rng), (float) S.width);
                    float fy = __fdividef((float) py + pcg_float(rng), (float) S.height);
                    float ax = pcg_float(rng), ay = pcg_float(rng);
                    f3 fs = mk3(1.f, 0.f, 0.f), ft = mk3(0.f, 1.f, 0.f);
                    if (GEN && S.type == ERTB_SENSOR_MPDISTANT) { // mpdistant.cpp:214-262
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
                        d = mk3(-(M[0] * hv.x + M[1] * hv.y + M[2] * hv.z),
                                -(M[3] * hv.x + M[4] * hv.y + M[5] * hv.z),
                                -(M[6] * hv.x + M[7] * hv.y + M[8] * hv.z));
                        fs = mk3(M[0], M[3], M[6]);
                        ft = mk3(M[1], M[4], M[7]);
                        if (S.type == ERTB_SENSOR_DISTANTFLUX) wray = hv.z * S.flux_norm;
                    }
                    if (SPH) {
                        double tx, ty, tz;
                        if (S.target_type == ERTB_TARGET_POINT) {
                            tx = S.target[0]; ty = S.target[1]; tz = S.target[2];
                        } else if (S.target_type == ERTB_TARGET_NONE) {
                            float ox, oy;
                            disk_concentric(ax, ay, ox, oy);
                            tx = S.bs_center[0] + ((double) fs.x * ox + (double) ft.x * oy) * S.bs_radius;
                            ty = S.bs_center[1] + ((double) fs.y * ox + (double) ft.y * oy) * S.bs_radius;
                            tz = S.bs_center[2] + ((double) fs.z * ox + (double) ft.z * oy) * S.bs_radius;
                        } else {
                            float lx, ly;
                            if (S.target_type == ERTB_TARGET_RECTANGLE) { lx = fmaf(2.f, ax, -1.f); ly = fmaf(2.f, ay, -1.f); }
                            else disk_concentric(ax, ay, lx, ly);
                            const double *T = S.target_to_world;
                            tx = T[0] * lx + T[1] * ly + T[3];
                            ty = T[4] * lx + T[5] * ly + T[7];
                            tz = T[8] * lx + T[9] * ly + T[11];
                        }
                        valid = primary_entry_sph(P, tx, ty, tz, d, S.ray_offset, n0);
                    } else {
                        double tz = S.target_type == ERTB_TARGET_POINT ? S.target[2]
                                  : S.target_type == ERTB_TARGET_NONE ? S.bs_center[2] : S.target_to_world[11];
                        double oz = tz - (double) d.z * S.ray_offset - P.Rd;
                        valid = !(d.z < 0.f) || oz < 0.0 ? 0 : (oz >= (double) P.H ? 1 : 2);
                    }
                }
                if (GEN && valid == 3) valid = 1; // starts inside the atmosphere at h_start, looking anywhere
                else {
                    if (!SPH && !(d.z < 0.f)) valid = 0;
                    h_start = P.H;
                }
                unsigned flags;
                float h0 = GEN ? h_start : P.H, b = 0.f, smax = 0.f;
                if (valid == 1 && PW) {
                    if (STATS) st_main++;
                    b = d.z;
                    flags = pw_advance(P, tb, rng, h0, d.z);
                } else if (valid == 1) {
                    int kind;
                    segment_setup<SPH>(P, n0, h0, d, b, smax, kind);
                    flags = PM_WALK_MAIN | (kind == KIND_GROUND ? PFL_KIND : 0u);
                } else if (valid == 2) {
                    h0 = 0.f;
                    flags = PM_SURF | PFL_VACUUM;
                } else {
                    thr = 0.f;
                    flags = PM_IDLE; // L = 0, still counted as a sample
                }
                FLDU(PF_FLAGS, slot) = flags;
                FLD(PF_H0, slot) = h0; FLD(PF_B, slot) = b; FLD(PF_S, slot) = 0.f; FLD(PF_SMAX, slot) = smax;
                FLD(PF_THR, slot) = thr; FLD(PF_WNEE, slot) = 0.f; FLD(PF_RES, slot) = 0.f;
                FLDU(PF_RNG0, slot) = (unsigned) rng.state; FLDU(PF_RNG1, slot) = (unsigned) (rng.state >> 32);
#ifdef ERTB_RNG_PCG32
                FLDU(PF_INC0, slot) = (unsigned) rng.inc; FLDU(PF_INC1, slot) = (unsigned) (rng.inc >> 32);
#endif
                FLD(PF_B2, slot) = b; FLD(PF_SMAX2, slot) = smax;
                FLD(PF_N0X, slot) = n0.x; FLD(PF_N0Y, slot) = n0.y; FLD(PF_N0Z, slot) = n0.z;
                FLD(PF_DX, slot) = d.x; FLD(PF_DY, slot) = d.y; FLD(PF_DZ, slot) = d.z;
                FLDU(PF_PIX, slot) = pix; FLD(PF_WRAY, slot) = wray;
                // band of the segment's origin (bits 16-23): the top band for rays entering at the TOA
                if (BANDS) FLDU(PF_BAND, slot) = (unsigned) (GEN ? band_of(P, tb, h0) : P.n_bands - 1) << 16;
                if (POL) {
                    // throughput starts as the rotation to the output Stokes basis (stokes.cpp:111-151):
                    // the accumulated vector is then directly expressed in that basis
                    float R[16];
                    stokes_output_rotation(P, d, R);
                    pol_store_T<NS>(wp, slot, R, 1.f);
                    FLD(PF_RQ, slot) = 0.f; FLD(PF_RQ + 1, slot) = 0.f; FLD(PF_RQ + 2, slot) = 0.f;
                }
            }
            __syncwarp();
            continue;
        }

        // =========================================================================
        // heavy events: surface interaction (volpath.cpp:344-389) / real collision (:261-296)
        // =========================================================================
        if (have) {
            unsigned flags = FLDU(PF_FLAGS, slot);
            f3 n0 = mk3(FLD(PF_N0X, slot), FLD(PF_N0Y, slot), FLD(PF_N0Z, slot));
            f3 d = mk3(FLD(PF_DX, slot), FLD(PF_DY, slot), FLD(PF_DZ, slot));
            float h0 = FLD(PF_H0, slot), thr = FLD(PF_THR, slot), res = FLD(PF_RES, slot);
            Pcg32 rng;
            rng.state = (unsigned long long) FLDU(PF_RNG0, slot) | ((unsigned long long) FLDU(PF_RNG1, slot) << 32);
#ifdef ERTB_RNG_PCG32
            rng.inc = (unsigned long long) FLDU(PF_INC0, slot) | ((unsigned long long) FLDU(PF_INC1, slot) << 32);
#else
            rng.inc = 1ULL;
#endif
            unsigned depth = flags >> PFL_DEPTH_SHIFT;
            float wnee = 0.f;
            bool dead = false;
            // Next-event direction: the sun, or a point of the solar disc (astroobject.cpp:141-175: uniform in
            // the cone, weight E).  The disc is sampled by next-event estimation ONLY: no MIS weight here and no
            // emitter hit for scattered rays -- the same expectation as the reference's MIS pair (volpath.cpp:
            // 286-290, :336-343), whose second technique carries (pdf_scatter x solid angle)^2 of the weight.
            f3 sun_e = sun;
            if (GEN && P.astro_omc > 0.f) { // (astroobject scenes run in the GEN instances)
                float ox, oy;
                disk_concentric(pcg_float(rng), pcg_float(rng), ox, oy);
                float pn = fmaf(ox, ox, oy * oy);
                float t = P.astro_omc * pn;           // 1 - cos(theta) of the sample
                float sc = sqrtf(P.astro_omc * (2.f - t)); // sin(theta) / sqrt(pn)
                f3 fs, ft;
                onb(sun, fs, ft);
                sun_e = normalize3(fma3(fs, ox * sc, fma3(ft, oy * sc, scale3(sun, 1.f - t))));
            }
            // polarized state: T = thr * That (actual Mueller throughput), result (Q, U, V), NEE ratios
            float T[POL ? 16 : 1], rquv[3] = { 0.f, 0.f, 0.f }, qn[3] = { 0.f, 0.f, 0.f };
            if (POL) {
                pol_load_T<NS>(wp, slot, thr, T);
#pragma unroll
                for (int k = 0; k < 3; ++k) rquv[k] = FLD(PF_RQ + k, slot);
                if (flags & PFL_NEE_QUV) { // finish the previous NEE term
                    float wn = FLD(PF_WNEE, slot), q3[3];
                    pol_load_qn<NS>(wp, slot, q3);
#pragma unroll
                    for (int k = 0; k < 3; ++k) rquv[k] = fmaf(wn, q3[k], rquv[k]);
                }
            }

            if (phase == PM_SURF) {
                float sm = (flags & PFL_VACUUM) ? 0.f : FLD(PF_SMAX, slot);
                if (SPH) n0 = normalize3(fma3(d, sm, scale3(n0, P.R + h0)));
                h0 = 0.f;
                if (STATS) st_surface++;
                float ci = -dot3(n0, d);
                if (!(ci > 0.f) || P.bsdf_type == ERTB_BSDF_BLACK) {
                    thr = 0.f; dead = true;
                    depth++; // absorbed by the ground: not an escape (the astroobject direct view tests depth == 0)
                } else {
                    float f_sun, weight;
                    if (POL && bsdf_is_mueller_t<GEN>(P.bsdf_type)) {
                        // Mueller-valued BSDF (polarized Fresnel matrix): ocean_legacy.cpp:561-661,
                        // ocean_mishchenko.cpp:228-296, ocean_grasp.cpp:354-455, maignan.cpp:105-166
                        f3 fs, ft;
                        surface_frame<SPH>(n0, fs, ft);
                        f3 wi = mk3(-dot3(d, fs), -dot3(d, ft), ci);
                        float Mb[16], v[4] = { 0.f, 0.f, 0.f, 0.f };
                        if (depth + 1u < P.max_depth) {
                            f3 ws = mk3(dot3(sun_e, fs), dot3(sun_e, ft), dot3(sun_e, n0));
                            if (ws.z > 0.f) {
                                lf_eval_mueller<GEN>(P, false, wi, ws, fs, ft, n0, Mb);
#pragma unroll
                                for (int r = 0; r < 4; ++r)
                                    v[r] = fmaf(T[4 * r], Mb[0], fmaf(T[4 * r + 1], Mb[4], fmaf(T[4 * r + 2], Mb[8], T[4 * r + 3] * Mb[12])));
                            }
                        }
                        wnee = v[0] * P.irradiance;
                        float inv = v[0] != 0.f ? __fdividef(1.f, v[0]) : 0.f;
#pragma unroll
                        for (int k = 0; k < 3; ++k) qn[k] = v[k + 1] * inv;
                        float s1 = pcg_float(rng), u1 = pcg_float(rng), u2 = pcg_float(rng);
                        f3 wo;
                        if (!GEN || P.bsdf_type == ERTB_BSDF_OCEAN_LEGACY) oc_sample(P, wi, s1, u1, u2, wo);
                        else wo = gl_sample_dir(P, wi, s1, u1, u2);
                        float Tn[16];
                        if (wo.z > 0.f) { // the weight BSDF::sample returns, as a matrix
                            lf_eval_mueller<GEN>(P, true, wi, wo, fs, ft, n0, Mb);
                            mueller_mul(T, Mb, Tn);
                        } else {
#pragma unroll
                            for (int k = 0; k < 16; ++k) Tn[k] = 0.f;
                        }
#pragma unroll
                        for (int k = 0; k < 16; ++k) T[k] = Tn[k];
                        d = normalize3(fma3(fs, wo.x, fma3(ft, wo.y, scale3(n0, wo.z))));
                        weight = thr != 0.f ? __fdividef(T[0], thr) : 0.f; // so that thr * weight = T00 below
                        if (!(weight > 0.f)) weight = 0.f;
                    } else {
                    surface_interact<SPH, POL, GEN>(P, n0, sun_e, ci, depth + 1u < P.max_depth, rng, d, f_sun, weight);
                    wnee = thr * f_sun * P.irradiance;
                    if (POL) {
                        // depolarizer(f): the NEE Stokes vector is T[:,0] * f * E; then T <- T * depolarizer(w)
                        float inv = thr != 0.f ? __fdividef(1.f, thr) : 0.f;
#pragma unroll
                        for (int k = 0; k < 3; ++k) qn[k] = T[4 * (k + 1)] * inv;
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            T[4 * r] *= weight; T[4 * r + 1] = 0.f; T[4 * r + 2] = 0.f; T[4 * r + 3] = 0.f;
                        }
                    }
                    }
                    if (flags & PFL_VACUUM) {
                        res += wnee;
                        if (POL) {
#pragma unroll
                            for (int k = 0; k < 3; ++k) rquv[k] = fmaf(wnee, qn[k], rquv[k]);
                        }
                        wnee = 0.f; thr = 0.f; dead = true;
                    }
                    thr *= weight;
                    depth++;
                }
            } else { // PM_SCAT
                float bb = FLD(PF_B, slot), s = FLD(PF_S, slot);
                float h = altitude_at<SPH>(P, h0, bb, s);
                if (SPH) n0 = normalize3(fma3(d, s, scale3(n0, P.R + h0)));
                h0 = h;
                int l = layer_of(P, h);
                {
                    float al = tb[P.off_albedo + l];
                    thr *= al;
                    if (POL) {
#pragma unroll
                        for (int k = 0; k < 16; ++k) T[k] *= al;
                    }
                }
                depth++;
                if (depth >= P.max_depth) {
                    thr = 0.f; dead = true;
                } else {
                    if (STATS) st_scatter++;
                    if (thr == 0.f) {
                        dead = true;
                    } else {
                        float ct_sun = dot3(d, sun_e);
                        float pv = 0.f;
                        int leaf = 0;
                        const f3 wi_w = mk3(-d.x, -d.y, -d.z);
                        float Pm[POL ? 16 : 1]; // mixture phase matrix towards the sun
                        if (POL) {
#pragma unroll
                            for (int k = 0; k < 16; ++k) Pm[k] = 0.f;
                        }
                        if (P.n_phase == 1) {
                            if (POL) { float pp; leaf_mueller(tb, P.leaf[0], wi_w, sun_e, Pm, pp); }
                            else pv = leaf_eval(tb, P.leaf[0], ct_sun);
                        } else {
                            float u0 = pcg_float(rng);
                            float prev = 0.f;
                            leaf = P.n_phase - 1;
                            bool picked = false;
                            for (int i = 0; i < P.n_phase; ++i) {
                                float cum = (i < P.n_phase - 1) ? tb[P.off_cumw + i * P.n_layers + l] : 1.f;
                                float w = cum - prev;
                                prev = cum;
                                if (w > 0.f) {
                                    if (POL) {
                                        float Mi[16], pp;
                                        leaf_mueller(tb, P.leaf[i], wi_w, sun_e, Mi, pp);
#pragma unroll
                                        for (int k = 0; k < 16; ++k) Pm[k] = fmaf(w, Mi[k], Pm[k]);
                                    } else {
                                        pv = fmaf(w, leaf_eval(tb, P.leaf[i], ct_sun), pv);
                                    }
                                }
                                if (!picked && u0 < cum) { leaf = i; picked = true; }
                            }
                        }
                        if (POL) {
                            // NEE Stokes vector: first column of T * P, times E
                            float v[4];
#pragma unroll
                            for (int r = 0; r < 4; ++r)
                                v[r] = fmaf(T[4 * r], Pm[0], fmaf(T[4 * r + 1], Pm[4], fmaf(T[4 * r + 2], Pm[8], T[4 * r + 3] * Pm[12])));
                            wnee = v[0] * P.irradiance;
                            float inv = v[0] != 0.f ? __fdividef(1.f, v[0]) : 0.f;
#pragma unroll
                            for (int k = 0; k < 3; ++k) qn[k] = v[k + 1] * inv;
                            if (!(wnee == wnee)) wnee = 0.f; // (NaN guard; the sign is the estimator's business)
                        } else {
                            wnee = thr * pv * P.irradiance;
                        }
                        float u1 = pcg_float(rng), u2 = pcg_float(rng);
                        float pw, ppdf;
                        float ct = leaf_sample(tb, P.leaf[leaf], u1, pw, ppdf);
                        if (ppdf > 0.f) {
                            float st = safe_sqrtf(1.f - ct * ct);
                            float sp, cp;
                            __sincosf(2.f * ERTB_PI * u2, &sp, &cp);
                            f3 fs, ft;
                            onb(d, fs, ft);
                            d = normalize3(fma3(fs, st * cp, fma3(ft, st * sp, scale3(d, ct))));
                            if (POL) {
                                // weight = Mueller value / pdf (rayleigh_polarized.cpp:155-157), T <- T * W
                                float W[16], pp, Tn[16];
                                if (GEN && P.phase_mis) {
                                    // multiphase.cpp:176-200: mixture value / mixture pdf over all leaves
#pragma unroll
                                    for (int k = 0; k < 16; ++k) W[k] = 0.f;
                                    pp = 0.f;
                                    float prev = 0.f;
                                    for (int i = 0; i < P.n_phase; ++i) {
                                        float cum = (i < P.n_phase - 1) ? tb[P.off_cumw + i * P.n_layers + l] : 1.f;
                                        float w = cum - prev;
                                        prev = cum;
                                        if (w > 0.f) {
                                            float Mi[16], pi_;
                                            leaf_mueller(tb, P.leaf[i], wi_w, d, Mi, pi_);
#pragma unroll
                                            for (int k = 0; k < 16; ++k) W[k] = fmaf(w, Mi[k], W[k]);
                                            pp = fmaf(w, pi_, pp);
                                        }
                                    }
                                    if (!(pp > 1e-8f)) pp = 0.f;
                                } else {
                                    leaf_mueller(tb, P.leaf[leaf], wi_w, d, W, pp);
                                }
                                float ip = pp > 0.f ? __fdividef(1.f, pp) : 0.f;
#pragma unroll
                                for (int k = 0; k < 16; ++k) W[k] *= ip;
                                mueller_mul(T, W, Tn);
#pragma unroll
                                for (int k = 0; k < 16; ++k) T[k] = Tn[k];
                                thr = T[0];
                            } else {
                                if (GEN && P.phase_mis) { // multiphase.cpp:176-200: mixture value / mixture pdf
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
                    }
                }
            }
            // volpathmis.cpp:227-231: Russian roulette after a real event only
            // (piecewise_volpath.cpp:187-193: every loop trip, i.e. once per real event as well)
            if (!dead && (PW || P.mis) && depth > P.rr_depth) {
                float q = fminf(thr, 0.95f);
                if (pcg_float(rng) >= q) thr = 0.f; else thr = __fdividef(thr, q);
            }
            // A BSDF may be negative (RTLS at grazing angles, rtls.cpp:231-243 does not clamp): throughput and
            // next-event terms keep their sign as in the reference, where only unpolarized(throughput) == 0 ends
            // the path (volpath.cpp:191) and Russian roulette removes a negative throughput once depth > rr_depth
            if (POL && !(thr == thr)) thr = 0.f;
            if (thr == 0.f || depth >= P.max_depth) dead = true; // the NEE walk (if any) still runs
            if (PW) {
                // ---- exact shadow-ray transmittance (piecewise_volpath.cpp:404-527), then the free
                //      flight to the next event ----
                if (wnee != 0.f) {
                    if (STATS) st_nee++;
                    float c = sun_e.z > 0.f ? wnee * pw_transmittance_up(P, tb, h0, sun_e.z) : 0.f;
                    res += c;
                    if (POL) {
#pragma unroll
                        for (int k = 0; k < 3; ++k) rquv[k] = fmaf(c, qn[k], rquv[k]);
                    }
                }
                unsigned mode = PM_IDLE;
                if (dead) thr = 0.f;
                else {
                    if (STATS) st_main++;
                    mode = pw_advance(P, tb, rng, h0, d.z);
                }
                FLDU(PF_FLAGS, slot) = mode | (depth << PFL_DEPTH_SHIFT);
                FLD(PF_H0, slot) = h0; FLD(PF_B, slot) = d.z; FLD(PF_S, slot) = 0.f; FLD(PF_SMAX, slot) = 0.f;
                FLD(PF_THR, slot) = thr; FLD(PF_WNEE, slot) = 0.f; FLD(PF_RES, slot) = res;
                FLDU(PF_RNG0, slot) = (unsigned) rng.state; FLDU(PF_RNG1, slot) = (unsigned) (rng.state >> 32);
                FLD(PF_DX, slot) = d.x; FLD(PF_DY, slot) = d.y; FLD(PF_DZ, slot) = d.z;
                if (POL) {
                    float inv = thr != 0.f ? __fdividef(1.f, T[0]) : 0.f;
                    pol_store_T<NS>(wp, slot, T, inv);
#pragma unroll
                    for (int k = 0; k < 3; ++k) FLD(PF_RQ + k, slot) = rquv[k];
                }
            } else {
            // ---- segment set-up for the NEE walk and for the main walk that follows it ----
            float b = 0.f, smax = 0.f, b2 = 0.f, smax2 = 0.f;
            int kind = KIND_TOA, kind2 = KIND_TOA;
            if (!dead) segment_setup<SPH>(P, n0, h0, d, b2, smax2, kind2);
            if (wnee != 0.f) {
                segment_setup<SPH>(P, n0, h0, sun_e, b, smax, kind);
                if (kind == KIND_GROUND) wnee = 0.f; // sun below the local horizon
            }
            unsigned mode;
            if (wnee != 0.f) {
                mode = PM_WALK_NEE;
                kind = KIND_TOA;
            } else if (dead) {
                mode = PM_IDLE;
            } else {
                mode = PM_WALK_MAIN;
                b = b2; smax = smax2; kind = kind2;
            }
            if (dead) thr = 0.f;
            flags = mode | (kind == KIND_GROUND ? PFL_KIND : 0u) | (kind2 == KIND_GROUND ? PFL_KIND2 : 0u) |
                    (depth << PFL_DEPTH_SHIFT);
            FLDU(PF_FLAGS, slot) = flags;
            FLD(PF_H0, slot) = h0; FLD(PF_B, slot) = b; FLD(PF_S, slot) = 0.f; FLD(PF_SMAX, slot) = smax;
            FLD(PF_THR, slot) = thr; FLD(PF_WNEE, slot) = wnee; FLD(PF_RES, slot) = res;
            FLDU(PF_RNG0, slot) = (unsigned) rng.state; FLDU(PF_RNG1, slot) = (unsigned) (rng.state >> 32);
            FLD(PF_B2, slot) = b2; FLD(PF_SMAX2, slot) = smax2;
            FLD(PF_N0X, slot) = n0.x; FLD(PF_N0Y, slot) = n0.y; FLD(PF_N0Z, slot) = n0.z;
            FLD(PF_DX, slot) = d.x; FLD(PF_DY, slot) = d.y; FLD(PF_DZ, slot) = d.z;
            // both segments of this event start in its band: the ground's, or the one the collision fell in
            if (BANDS) FLDU(PF_BAND, slot) = phase == PM_SURF ? 0u : (FLDU(PF_BAND, slot) & 255u) << 16;
            if (POL) {
                float inv = thr != 0.f ? __fdividef(1.f, T[0]) : 0.f;
                pol_store_T<NS>(wp, slot, T, inv);
                pol_store_qn<NS>(wp, slot, qn);
#pragma unroll
                for (int k = 0; k < 3; ++k) FLD(PF_RQ + k, slot) = rquv[k];
            }
            } // !PW
        }
        __syncwarp();
    }
#undef FLD
#undef FLDU

    film_flush_warp<POL>(P, lane, true, acc_pix, acc_wl, acc_l, acc_l2, acc_q, acc_u, acc_v);
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
