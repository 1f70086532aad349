// ertb_kernel.cuh -- the persistent wavefront megakernel.
//
// One launch renders samples [sample_offset, sample_offset + spp) of every pixel of
// one sensor.  The integrator is the reference's null-collision volumetric path
// tracer (MI/src/integrators/volpath.cpp:93-572; delta tracking for the walk, ratio
// tracking for next-event estimation, Russian roulette after rr_depth), re-designed
// for the GPU rather than translated:
//
//   * persistent CTAs (grid = SMs x resident CTAs); every lane owns one path whose
//     whole state lives in registers -- nothing is streamed through HBM per bounce;
//   * the 1D atmosphere (sigma_t/majorant, albedo, blend weights) and the tabulated
//     phase pdf/cdf are staged once per CTA into shared memory with a single TMA
//     bulk copy (cp.async.bulk ... mbarrier::complete_tx);
//   * the geometry is analytic and reduced: a path segment is (unit "up" vector n0
//     at the origin, altitude h0, direction) and the altitude along the segment is
//     h(s) = h0 + q/(r + r0), q = s (s + 2 r0 mu0) -- exact in fp32 at planetary
//     scale, no world-space xyz, no quadric per collision;
//   * the main walk and the NEE shadow walk share ONE free-flight step, so lanes in
//     either mode stay converged; the rare heavy events (surface BSDF, phase
//     sampling, primary-ray generation) are compacted with warp ballots;
//   * terminated lanes are refilled immediately (path regeneration) from a chunked
//     global work queue, so divergent path depths never leave lanes idle.
#pragma once

#include "ertb_device.cuh"
#include "ertb_ocean.cuh"

enum : int {
    MODE_IDLE = 0,
    MODE_WALK_MAIN = 1,
    MODE_WALK_NEE = 2,
    MODE_EV_SCATTER = 3,
    MODE_EV_SURFACE = 4,
    MODE_SETUP_MAIN = 5,
    MODE_SETUP_NEE = 6,
};
enum : int { KIND_TOA = 0, KIND_GROUND = 1 };
enum : int { PHASE_STEP = 0, PHASE_SETUP = 1, PHASE_REGEN = 2, PHASE_SURFACE = 3, PHASE_SCATTER = 4 };

#ifndef ERTB_TW
#define ERTB_TW 4 // minimum number of walking lanes for a free-flight trip (tuned on B200)
#endif

#define ERTB_BLOCK 256
#ifndef ERTB_MINB
#define ERTB_MINB 4
#endif

__device__ __forceinline__ float fast_sqrt(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned) __cvta_generic_to_shared(p); }

// Stage `bytes` (multiple of 16) from global to shared memory with one TMA bulk copy.
__device__ __forceinline__ void tma_stage(void *dst, const void *src, unsigned bytes,
                                          unsigned long long *mbar) {
    unsigned mb = smem_u32(mbar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mb), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
            ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(mb) : "memory");
    }
    // all threads wait for phase 0 to complete
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "ERTB_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra ERTB_WAIT_%=;\n"
        "}\n" ::"r"(mb), "r"(0) : "memory");
}

// Segment set-up: distance to the next boundary along `v` from (n0, h0).
// Spherical shell: roots of s^2 + 2 b s - c = 0 with b = r0 mu0 and
// c = (h_s - h0)(2R + h_s + h0), written in their cancellation-free forms.
template <bool SPH>
__device__ __forceinline__ void segment_setup(const ErtbParams &P, f3 n0, float h0, f3 v, float &b,
                                              float &smax, int &kind) {
    if (SPH) {
        float mu = dot3(n0, v);
        float r0 = P.R + h0;
        b = r0 * mu;
        kind = KIND_TOA;
        float cg = h0 * (2.f * P.R + h0);
        if (mu < 0.f) {
            float disc = fmaf(b, b, -cg);
            if (disc >= 0.f) {
                smax = __fdividef(cg, fast_sqrt(disc) - b);
                kind = KIND_GROUND;
            }
        }
        if (kind == KIND_TOA) {
            float ct = fmaxf((P.H - h0) * (2.f * P.R + P.H + h0), 0.f);
            float sq = fast_sqrt(fmaf(b, b, ct));
            smax = b > 0.f ? __fdividef(ct, b + sq) : sq - b;
        }
    } else {
        b = v.z;
        if (v.z < 0.f) {
            kind = KIND_GROUND;
            smax = __fdividef(h0, -v.z);
        } else {
            kind = KIND_TOA;
            smax = v.z > 0.f ? __fdividef(fmaxf(P.H - h0, 0.f), v.z) : 0.f;
        }
    }
}

// altitude after travelling s along the segment
template <bool SPH>
__device__ __forceinline__ float altitude_at(const ErtbParams &P, float h0, float b, float s) {
    if (SPH) {
        float r0 = P.R + h0;
        float q = s * fmaf(2.f, b, s);
        float r = fast_sqrt(fmaf(r0, r0, q));
        return h0 + __fdividef(q, r + r0);
    }
    return fmaf(s, b, h0);
}

__device__ __forceinline__ int layer_of(const ErtbParams &P, float h) {
    int i = __float2int_rd((h + P.h_off) * P.inv_dz);
    return min(max(i, 0), P.n_layers - 1);
}

// Primary ray through target point T (double, world space) with direction d.  The
// reference places the origin at o = T - d * ray_offset (mdistant.cpp:192-242,
// hdistant.cpp:232-275, distantflux.cpp:148-195) and the sensor sits in vacuum:
//   return 1: o is outside the outer stencil and the ray enters it at n0 (h0 = H);
//   return 2: o lies INSIDE the outer stencil (possible with target-less sensors or a
//             short ray_offset): the ray travels through vacuum (sensor medium = null)
//             and hits the ground at n0 without crossing the atmosphere;
//   return 0: nothing is hit, L = 0.
__device__ __forceinline__ int primary_entry_sph(const ErtbParams &P, double tx, double ty, double tz,
                                                 f3 d, double ray_offset, f3 &n0) {
    double dx = d.x, dy = d.y, dz = d.z;
    double ox = tx - dx * ray_offset, oy = ty - dy * ray_offset, oz = tz - dz * ray_offset;
    double Rt = P.Rd + (double) P.H;
    double b = ox * dx + oy * dy + oz * dz;
    double oo = ox * ox + oy * oy + oz * oz;
    if (oo >= Rt * Rt) {
        double disc = b * b - (oo - Rt * Rt);
        if (disc < 0.0 || b > 0.0) return 0; // missed, or the sphere is behind the origin
        double t0 = -b - sqrt(disc);
        double inv = 1.0 / Rt;
        n0 = mk3((float) ((ox + t0 * dx) * inv), (float) ((oy + t0 * dy) * inv), (float) ((oz + t0 * dz) * inv));
        return 1;
    }
    if (oo < P.Rd * P.Rd) return 0; // origin below the surface: back-facing hit, BSDF = 0
    double disc = b * b - (oo - P.Rd * P.Rd);
    if (disc < 0.0 || b > 0.0) return 0; // leaves through the outer stencil
    double t0 = -b - sqrt(disc);
    double inv = 1.0 / P.Rd;
    n0 = mk3((float) ((ox + t0 * dx) * inv), (float) ((oy + t0 * dy) * inv), (float) ((oz + t0 * dz) * inv));
    return 2;
}

// Local shading frame of the ground: ERP/shapes/arectangle.cpp (world x, y) for the slab,
// MI/src/shapes/sphere.cpp:703-720 + interaction.h:278-288 (s = normalised dp_du = east) for the
// sphere.  Only the ocean BSDF (wind direction) depends on it.
template <bool SPH>
__device__ __forceinline__ void surface_frame(f3 n, f3 &s, f3 &t) {
    if (SPH) {
        float rd2 = n.x * n.x + n.y * n.y;
        if (rd2 > 0.f) {
            float inv = rsqrtf(rd2);
            s = mk3(-n.y * inv, n.x * inv, 0.f);
            t = mk3(n.y * s.z - n.z * s.y, n.z * s.x - n.x * s.z, n.x * s.y - n.y * s.x);
        } else {
            onb(n, s, t);
        }
    } else {
        s = mk3(1.f, 0.f, 0.f);
        t = mk3(0.f, 1.f, 0.f);
    }
}

// Surface interaction (volpath.cpp:344-375): BSDF value towards the sun for next-event
// estimation (`f_sun` = f * cos, 0 when not wanted) and BSDF sampling (new direction `d`,
// weight = f * cos / pdf).  `ci` = cos of the incident direction with the normal (> 0).
// `SCALAR_ONLY`: the caller handles the Mueller-valued BSDFs itself (polarized instances of the pool kernel), so
// their scalar code is not compiled into it; the local-frame BSDFs left here are the table-driven ones.
// `GENERAL` = false: only the plugin set of SURVEY 8a (see bsdf_is_local_t).
template <bool SPH, bool SCALAR_ONLY = false, bool GENERAL = true>
__device__ __forceinline__ void surface_interact(const ErtbParams &P, f3 n0, f3 sun, float ci, bool want_nee,
                                                 Pcg32 &rng, f3 &d, float &f_sun, float &weight) {
    f_sun = 0.f;
    weight = 0.f;
    if (SCALAR_ONLY ? (GENERAL && bsdf_is_table(P.bsdf_type)) : bsdf_is_local_t<GENERAL>(P.bsdf_type)) { // 6SV ocean, glint family, mqdiffuse, measured_mono
        f3 fs, ft;
        surface_frame<SPH>(n0, fs, ft);
        f3 wi = mk3(-dot3(d, fs), -dot3(d, ft), ci);
        if (want_nee) {
            f3 ws = mk3(dot3(sun, fs), dot3(sun, ft), dot3(sun, n0));
            if (ws.z > 0.f) f_sun = SCALAR_ONLY ? tb_eval(P, wi, ws) : lf_eval<GENERAL>(P, wi, ws);
        }
        float s1 = pcg_float(rng), u1 = pcg_float(rng), u2 = pcg_float(rng);
        f3 wo;
        weight = SCALAR_ONLY ? tb_sample(P, wi, u1, u2, wo) : lf_sample<GENERAL>(P, wi, s1, u1, u2, wo);
        if (!(wo.z > 0.f)) weight = 0.f;
        d = normalize3(fma3(fs, wo.x, fma3(ft, wo.y, scale3(n0, wo.z))));
        return;
    }
    if (want_nee) {
        float co = dot3(n0, sun);
        if (co > 0.f) f_sun = bsdf_f(P, ci, co, cos_dphi(ci, co, -dot3(d, sun))) * co;
    }
    float u1 = pcg_float(rng), u2 = pcg_float(rng);
    f3 wl = cosine_hemisphere(u1, u2);
    f3 fs, ft;
    onb(n0, fs, ft);
    f3 nd = fma3(fs, wl.x, fma3(ft, wl.y, scale3(n0, wl.z)));
    if (wl.z > 0.f) // value * cos / pdf = value * pi   (rpv.cpp:119-122)
        weight = bsdf_f(P, ci, wl.z, cos_dphi(ci, wl.z, -dot3(d, nd))) * ERTB_PI;
    d = normalize3(nd);
}

template <bool SPH, bool STATS>
__global__ void __launch_bounds__(ERTB_BLOCK, ERTB_MINB) ertb_render_kernel(const ErtbParams P) {
    extern __shared__ __align__(16) float tb[]; // table blob
    __shared__ __align__(8) unsigned long long mbar;

    if (P.blob_bytes > 0) tma_stage(tb, P.blob, (unsigned) P.blob_bytes, &mbar);

    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const f3 sun = mk3(P.sun[0], P.sun[1], P.sun[2]);

    // ---- per-lane path state (registers) ----
    int mode = MODE_IDLE;
    int kind = KIND_TOA;
    f3 n0 = mk3(0.f, 0.f, 1.f);
    f3 d = mk3(0.f, 0.f, -1.f);
    float h0 = 0.f, b = 0.f, s = 0.f, smax = 0.f;
    float thr = 0.f, res = 0.f, wnee = 0.f, wray = 1.f;
    unsigned depth = 0;
    bool last_null = false;
    bool vacuum = false; // primary ray reached the ground without crossing the medium
    Pcg32 rng;
    rng.state = 0; rng.inc = 1;

    // lane-local film accumulators, tagged with the pixel they belong to
    double acc_wl = 0.0, acc_l = 0.0, acc_l2 = 0.0;
    unsigned acc_pix = 0xffffffffu, pix = 0;

    // warp-uniform work-queue cursor
    unsigned long long cur_next = 0, cur_end = 0; // sample indices within the chunk's pixel
    unsigned cur_pix = 0;
    bool exhausted = false;

    unsigned st_main = 0, st_nee = 0, st_scatter = 0, st_surface = 0, st_paths = 0;

    unsigned m_idle = 0u;
    for (;;) {
        // ------------------------------------------------------------------
        // Warp-level phase scheduler.  Executing every phase every trip leaves ~6 of 32
        // lanes active per instruction (ncu, profiles/r01): the heavy event code runs for
        // one or two lanes at a time.  Instead each trip runs ONE phase: the free-flight
        // step while at least ERTB_TW lanes can walk, otherwise the non-walking phase that
        // currently holds the most lanes (events are thereby batched across trips).
        // ------------------------------------------------------------------
        const unsigned m_walk = __ballot_sync(0xffffffffu, mode == MODE_WALK_MAIN || mode == MODE_WALK_NEE);
        int phase = PHASE_STEP;
        if (__popc(m_walk) < P.tw) {
            const unsigned m_setup = __ballot_sync(0xffffffffu, mode == MODE_SETUP_MAIN || mode == MODE_SETUP_NEE);
            const unsigned m_surf = __ballot_sync(0xffffffffu, mode == MODE_EV_SURFACE);
            const unsigned m_scat = __ballot_sync(0xffffffffu, mode == MODE_EV_SCATTER);
            m_idle = exhausted ? 0u : __ballot_sync(0xffffffffu, mode == MODE_IDLE);
            int best = __popc(m_setup);
            phase = PHASE_SETUP;
            if (__popc(m_idle) > best) { best = __popc(m_idle); phase = PHASE_REGEN; }
            if (__popc(m_surf) > best) { best = __popc(m_surf); phase = PHASE_SURFACE; }
            if (__popc(m_scat) > best) { best = __popc(m_scat); phase = PHASE_SCATTER; }
            if (best == 0) {
                if (m_walk == 0u) break; // nothing left to do for this warp
                phase = PHASE_STEP;
            }
        }

        if (phase == PHASE_STEP) {
            // ------------------------------------------------------------------
            // D. one free-flight step: delta tracking (main) / ratio tracking (NEE)
            //    medium.cpp:42-82 with the global majorant of heterogeneous.cpp:163
            // ------------------------------------------------------------------
            if (mode == MODE_WALK_MAIN || mode == MODE_WALK_NEE) {
                const bool is_main = mode == MODE_WALK_MAIN;
                bool alive = true;
                if (is_main && !P.mis && depth > P.rr_depth) { // volpath.cpp:194-198, every loop trip
                    float q = fminf(thr, 0.95f);
                    if (pcg_float(rng) >= q) { thr = 0.f; mode = MODE_SETUP_MAIN; alive = false; }
                    else thr = __fdividef(thr, q);
                }
                if (alive) {
                    if (STATS) { if (is_main) st_main++; else st_nee++; }
                    float u = pcg_float(rng);
                    float t = -__logf(1.f - u) * P.inv_majorant; // inv_majorant = +inf without medium
                    s += t;
                    if (!(s < smax)) {
                        // boundary reached
                        if (is_main) {
                            if (kind == KIND_GROUND) mode = MODE_EV_SURFACE;
                            else { // left through the TOA: the path ends, film accumulation
                                acc_wl += (double) (wray * res);
                                acc_l += (double) res;
                                acc_l2 += (double) res * (double) res;
                                mode = MODE_IDLE;
                            }
                        } else {
                            res += wnee; // shadow ray reached the TOA (ground hits were culled at set-up)
                            mode = MODE_SETUP_MAIN;
                        }
                    } else {
                        float h = altitude_at<SPH>(P, h0, b, s);
                        float preal = tb[P.off_preal + layer_of(P, h)];
                        if (is_main) {
                            float u2 = pcg_float(rng);
                            if (u2 >= 1.f - preal) mode = MODE_EV_SCATTER; // real collision
                            else last_null = true;
                        } else {
                            wnee *= 1.f - preal; // ratio tracking: T *= sigma_n / majorant
                            if (wnee == 0.f) mode = MODE_SETUP_MAIN;
                        }
                    }
                }
            }
            continue;
        }

        // ------------------------------------------------------------------
        // A. path regeneration (warp-aggregated pops from the chunk queue)
        // ------------------------------------------------------------------
        if (phase == PHASE_REGEN) {
            const unsigned need = m_idle;
            bool got = false;
            unsigned long long my_sample = 0;
            unsigned my_pix = 0;
            unsigned pending = need;
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
                unsigned npend = __popc(pending);
                unsigned take = (unsigned) min((unsigned long long) npend, avail);
                bool mine = (pending >> lane) & 1u;
                unsigned rank = __popc(pending & lt_mask);
                if (mine && rank < take && !got) {
                    got = true;
                    my_sample = cur_next + rank;
                    my_pix = cur_pix;
                }
                cur_next += take;
                // clear the `take` lowest set bits of pending
                unsigned assigned = __ballot_sync(0xffffffffu, mine && rank < take);
                pending &= ~assigned;
            }
            if (got) {
                pix = my_pix;
                if (pix != acc_pix) {
                    if (acc_pix != 0xffffffffu) {
                        atomicAdd(&P.accum[acc_pix], acc_wl);
                        atomicAdd(&P.accum[P.n_pixels + acc_pix], acc_l);
                        atomicAdd(&P.accum[2u * P.n_pixels + acc_pix], acc_l2);
                    }
                    acc_wl = acc_l = acc_l2 = 0.0;
                    acc_pix = pix;
                }
                unsigned long long gid = ((unsigned long long) pix << 40) + (P.sample_offset + my_sample);
                pcg_seed(rng, P.seed, gid);
                thr = 1.f; res = 0.f; wnee = 0.f; wray = 1.f; depth = 0; last_null = false;
                vacuum = false;
                if (STATS) st_paths++;
                // ---- primary ray (render_sample, integrator.cpp:449-520) ----
                const ErtbSensor &S = P.sensor;
                int valid = 1;
                float h_start = P.H;
                if (S.use_table) {
                    const float4 *t4 = reinterpret_cast<const float4 *>(S.table) + 2u * pix;
                    float4 a = __ldg(t4), c4 = __ldg(t4 + 1);
                    n0 = mk3(a.x, a.y, a.z);
                    d = mk3(a.w, c4.x, c4.y);
                    valid = (int) c4.z;
                    h_start = c4.w; // mradiancemeter: altitude of an origin inside the atmosphere (class 3)
                } else {
                    unsigned px = pix % (unsigned) S.width, py = pix / (unsigned) S.width;
                    float fx = __fdividef((float) px + pcg_float(rng), (float) S.width);
                    float fy = __fdividef((float) py + pcg_float(rng), (float) S.height);
                    float ax = pcg_float(rng), ay = pcg_float(rng);
                    f3 fs = mk3(1.f, 0.f, 0.f), ft = mk3(0.f, 1.f, 0.f);
                    if (S.type == ERTB_SENSOR_MPDISTANT) { // mpdistant.cpp:214-262
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
                        if (S.type == ERTB_SENSOR_DISTANTFLUX) // distantflux.cpp:168-170
                            wray = hv.z * S.flux_norm;
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
                            if (S.target_type == ERTB_TARGET_RECTANGLE) {
                                lx = fmaf(2.f, ax, -1.f); ly = fmaf(2.f, ay, -1.f);
                            } else {
                                disk_concentric(ax, ay, lx, ly);
                            }
                            const double *T = S.target_to_world;
                            tx = T[0] * lx + T[1] * ly + T[3];
                            ty = T[4] * lx + T[5] * ly + T[7];
                            tz = T[8] * lx + T[9] * ly + T[11];
                        }
                        valid = primary_entry_sph(P, tx, ty, tz, d, S.ray_offset, n0);
                    } else {
                        // plane-parallel: only the origin height matters
                        double tz = S.target_type == ERTB_TARGET_POINT ? S.target[2]
                                  : S.target_type == ERTB_TARGET_NONE ? S.bs_center[2] : S.target_to_world[11];
                        double oz = tz - (double) d.z * S.ray_offset - P.Rd; // height above the ground
                        valid = !(d.z < 0.f) || oz < 0.0 ? 0 : (oz >= (double) P.H ? 1 : 2);
                    }
                }
                if (valid == 3) valid = 1; // starts inside the atmosphere at h_start, looking anywhere
                else {
                    if (!SPH && !(d.z < 0.f)) valid = 0; // upward-looking rays never see the ground
                    h_start = P.H;
                }
                h0 = h_start;
                mode = MODE_SETUP_MAIN;
                if (valid == 0) thr = 0.f; // L = 0, still counted as a sample
                if (valid == 2) { vacuum = true; h0 = 0.f; s = 0.f; smax = 0.f; mode = MODE_EV_SURFACE; }
            }
        }

        // ------------------------------------------------------------------
        // B. heavy events: real collision in the medium / surface interaction
        // ------------------------------------------------------------------
        if (phase == PHASE_SCATTER || phase == PHASE_SURFACE) {
            if (phase == PHASE_SCATTER && mode == MODE_EV_SCATTER) {
                // ---- volpath.cpp:261-296 ----
                float h = altitude_at<SPH>(P, h0, b, s);
                if (SPH) n0 = normalize3(fma3(d, s, scale3(n0, P.R + h0)));
                h0 = h;
                int l = layer_of(P, h);
                thr *= tb[P.off_albedo + l];
                depth++;
                last_null = false;
                if (depth >= P.max_depth) {
                    mode = MODE_SETUP_MAIN; thr = 0.f;
                } else if (thr == 0.f) {
                    if (STATS) st_scatter++;
                    mode = MODE_SETUP_MAIN;
                } else {
                    if (STATS) st_scatter++;
                    // emitter sampling: phase value towards the sun (delta emitter -> MIS weight 1)
                    float ct_sun = dot3(d, sun);
                    float pv = 0.f;
                    int leaf = 0;
                    if (P.n_phase == 1) {
                        pv = leaf_eval(tb, P.leaf[0], ct_sun);
                    } else {
                        float u0 = pcg_float(rng);
                        float prev = 0.f;
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
                    wnee = thr * pv * P.irradiance;
                    // phase sampling (blendphase.cpp:100-141: the component's own weight)
                    float u1 = pcg_float(rng), u2 = pcg_float(rng);
                    float pw, ppdf;
                    float ct = leaf_sample(tb, P.leaf[leaf], u1, pw, ppdf);
                    if (ppdf > 0.f) {
                        float st = safe_sqrtf(1.f - ct * ct);
                        float sp, cp;
                        __sincosf(2.f * ERTB_PI * u2, &sp, &cp);
                        f3 fs, ft;
                        onb(d, fs, ft);
                        f3 nd = fma3(fs, st * cp, fma3(ft, st * sp, scale3(d, ct)));
                        d = normalize3(nd);
                        thr *= pw;
                    }
                    mode = wnee != 0.f ? MODE_SETUP_NEE : MODE_SETUP_MAIN; // (a BSDF may be negative: rtls.cpp)
                }
            } else if (phase == PHASE_SURFACE && mode == MODE_EV_SURFACE) {
                // ---- volpath.cpp:344-389 ----
                if (SPH) n0 = normalize3(fma3(d, smax, scale3(n0, P.R + h0)));
                h0 = 0.f;
                if (STATS) st_surface++;
                float ci = -dot3(n0, d);
                if (!(ci > 0.f) || P.bsdf_type == ERTB_BSDF_BLACK) {
                    thr = 0.f;
                    mode = MODE_SETUP_MAIN;
                } else {
                    float f_sun, weight;
                    surface_interact<SPH>(P, n0, sun, ci, depth + 1u < P.max_depth, rng, d, f_sun, weight);
                    wnee = thr * f_sun * P.irradiance;
                    if (vacuum) { // no medium on either leg: the sample is the direct surface term
                        res += wnee; wnee = 0.f; thr = 0.f;
                    }
                    thr *= weight;
                    depth++;
                    last_null = false;
                    mode = wnee != 0.f ? MODE_SETUP_NEE : MODE_SETUP_MAIN; // (a BSDF may be negative: rtls.cpp)
                }
            }
        }

        // ------------------------------------------------------------------
        // C. segment set-up (shared by the main walk and the NEE walk)
        // ------------------------------------------------------------------
        if (phase == PHASE_SETUP && mode == MODE_SETUP_NEE) {
            segment_setup<SPH>(P, n0, h0, sun, b, smax, kind);
            s = 0.f;
            if (kind == KIND_GROUND) { wnee = 0.f; mode = MODE_SETUP_MAIN; } // sun below the local horizon
            else mode = MODE_WALK_NEE;
        }
        if (phase == PHASE_SETUP && mode == MODE_SETUP_MAIN) {
            if (thr == 0.f || depth >= P.max_depth) {
                // ---- path finished: film accumulation (imageblock.cpp:174, moment.cpp:101-104) ----
                acc_wl += (double) (wray * res);
                acc_l += (double) res;
                acc_l2 += (double) res * (double) res;
                mode = MODE_IDLE;
            } else {
                segment_setup<SPH>(P, n0, h0, d, b, smax, kind);
                s = 0.f;
                mode = MODE_WALK_MAIN;
                if (P.mis && depth > P.rr_depth) { // volpathmis.cpp:227-231 (after a real event only)
                    float q = fminf(thr, 0.95f);
                    if (pcg_float(rng) >= q) { thr = 0.f; mode = MODE_SETUP_MAIN; }
                    else thr = __fdividef(thr, q);
                }
            }
        }

    }

    // ---- flush lane accumulators and statistics ----
    if (acc_pix != 0xffffffffu) {
        atomicAdd(&P.accum[acc_pix], acc_wl);
        atomicAdd(&P.accum[P.n_pixels + acc_pix], acc_l);
        atomicAdd(&P.accum[2u * P.n_pixels + acc_pix], acc_l2);
    }
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
