// ertb_piecewise.cuh -- analytic free flight and exact transmittance through a stack of
// homogeneous plane-parallel layers (ERP/media/piecewise.cpp:183-429).
//
// The reference walks its two running-sum tables (bottom-up / top-down, :445-507) with a
// binary search per call and rebuilds distances layer by layer.  Here ONE table is staged in
// shared memory with the rest of the scene blob: tau[i] = optical depth, measured vertically,
// from layer boundary i UP TO THE TOP of the grid (n_layers + 1 floats, accumulated in double on
// the host, tau[n] = 0).  Counting from the top keeps float32 resolution where an atmosphere
// needs it: the optically thin upper layers sit next to tau = 0, the thick lower ones are wide
// in tau.  A flight of optical length E along a direction with vertical cosine mu ends where
// tau(h) = tau(h0) - mu * E, so both directions use the same table and the same search:
//   mu > 0: target <= 0            -> the path leaves through the top of the atmosphere
//   mu < 0: target >= tau(ground)  -> the path reaches the surface
//   else  : the layer holding `target` is found by bisection over at most 8 shared-memory loads
// and the collision altitude follows from the linear tau(h) inside that layer.  Collisions
// that stay in the starting layer take s = E / sigma directly (no cancellation for grazing rays).
#pragma once

#include "ertb_device.cuh"

enum : int { PW_COLLISION = 0, PW_GROUND = 1, PW_TOA = 2 };

// vertical optical depth between altitude h (relative to the ground) and the top of the grid
__device__ __forceinline__ float pw_tau_at(const ErtbParams &P, const float *tb, float h, int &l, float &sig) {
    float x = fmaxf((h + P.h_off) * P.inv_dz, 0.f);
    l = min((int) x, P.n_layers - 1);
    sig = tb[P.off_sigma + l];
    float frac = fminf(x - (float) l, 1.f);
    return fmaf(-sig * frac, P.dz, tb[P.off_tau + l]);
}

// Free flight of optical length E from altitude h0 along vertical cosine mu.
// PW_COLLISION: s = distance, h = collision altitude.  PW_GROUND: s = distance to the surface.
__device__ __forceinline__ int pw_flight(const ErtbParams &P, const float *tb, float h0, float mu, float E,
                                         float &s, float &h) {
    const float *tau = tb + P.off_tau;
    const int n = P.n_layers;
    int l0;
    float sig0;
    const float tau0 = pw_tau_at(P, tb, h0, l0, sig0);
    const float target = fmaf(-mu, E, tau0);
    s = 0.f; h = h0;
    if (mu >= 0.f) {
        if (!(target > 0.f)) return PW_TOA;
        if (target > tau[l0 + 1] || mu == 0.f) {
            s = __fdividef(E, sig0);
            h = fmaf(mu, s, h0);
            return isfinite(s) ? PW_COLLISION : PW_TOA; // horizontal flight through a vacuum layer
        }
    } else {
        if (!(target < P.tau_ground)) {
            s = __fdividef(h0, -mu);
            h = 0.f;
            return PW_GROUND;
        }
        if (!(target > tau[l0])) {
            s = __fdividef(E, sig0);
            h = fmaf(mu, s, h0);
            return PW_COLLISION;
        }
    }
    // largest l with tau[l] >= target (tau decreases with l), on the side of l0 the flight goes to
    int lo = mu > 0.f ? l0 + 1 : 0, hi = mu > 0.f ? n - 1 : l0 - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (tau[mid] >= target) lo = mid; else hi = mid - 1;
    }
    const float sig = tb[P.off_sigma + lo];
    const float hb = fmaf((float) lo, P.dz, -P.h_off); // bottom boundary of layer lo
    float dh = sig > 0.f ? __fdividef(tau[lo] - target, sig) : 0.f;
    dh = fminf(fmaxf(dh, 0.f), P.dz);
    h = fmaxf(hb + dh, 0.f);
    s = __fdividef(h - h0, mu);
    return PW_COLLISION;
}

// exp(-optical depth) from altitude h to the top of the atmosphere along vertical cosine mu > 0
// (eval_transmittance_pdf_real, piecewise.cpp:335-429, for a shadow ray that leaves the slab)
__device__ __forceinline__ float pw_transmittance_up(const ErtbParams &P, const float *tb, float h, float mu) {
    int l; float sig;
    float t = pw_tau_at(P, tb, h, l, sig);
    return __expf(-__fdividef(fmaxf(t, 0.f), mu));
}
