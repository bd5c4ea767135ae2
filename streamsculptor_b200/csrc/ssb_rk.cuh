// Adaptive Dormand-Prince stepping for second-order systems  x' = p, p' = F(x, t)  held in registers.
//
// Algorithm = diffrax.diffeqsolve(Dopri5|Dopri8, PIDController(rtol, atol, dtmin, dtmax, force_dtmin=True),
// dt0=None, SaveAt(ts)) as configured at /root/reference/streamsculptor/main.py:139-162 and fields.py:85-98:
// identical Butcher tableaus (ssb_tableau.h), I-controller (safety 0.9, factor in [0.2 | 1 after an accepted
// step, 10], exponent 1/order), RMS error norm with scale atol + rtol*max(|y0|,|y1|), HNW initial step,
// clip-to-end at 1e-10, steps at dt <= dtmin always accepted.
//
// B200 mapping: every orbit on the path is a second-order system whose force depends on (x, t) only, so the
// tableau is applied in its algebraically identical Nystrom form (ssb_tableau.h: aa = A.A, ea = e^T A):
// only the S force stages (3 doubles each) live in registers instead of S full 6-vectors, which halves the
// register footprint of Dopri8 (42 instead of 84 doubles) and is what lets 16 warps/SM stay resident.
#ifndef SSB_RK_CUH
#define SSB_RK_CUH
#include <cuda_runtime.h>
#include <math.h>

#define SSB_TABLEAU_QUAL static __device__ constexpr
#include "ssb_tableau.h"                // ssb_tab::  compile-time values: zero tests while unrolling, rolled loops
// ... and the same numbers once more in the constant bank (ssb_ctab::).  A 64-bit immediate cannot be encoded in DFMA: with the constexpr
// copy every tableau coefficient of the unrolled stepper was materialised by two UMOVs into a uniform register pair before each use
// (500 UMOV + ~700 moves beside 2224 FP64 instructions in orbit_kernel<8>); a __constant__ entry with a compile-time index is an operand.
#undef SSB_TABLEAU_QUAL
#define SSB_TABLEAU_QUAL static __constant__
#define SSB_TABLEAU_NS ssb_ctab
#define SSB_TABLEAU_AGAIN
#include "ssb_tableau.h"
#undef SSB_TABLEAU_AGAIN
#undef SSB_TABLEAU_NS
#undef SSB_TABLEAU_QUAL
#define SSB_TABLEAU_QUAL static __device__ constexpr
#ifndef SSB_TABLEAU_CONSTBANK
#define SSB_TABLEAU_CONSTBANK 1      // 0: operands from the constexpr copy (immediates materialised by UMOV pairs), for A/B builds
#endif
#if SSB_TABLEAU_CONSTBANK
#define SSB_CTAB ssb_ctab
#else
#define SSB_CTAB ssb_tab
#endif
#include "ssb_fastmath.cuh"

namespace ssb {

template <int SOLVER> struct Tab;
template <> struct Tab<5> {
    static constexpr int S = 7, ORDER = 5;
    static __device__ __forceinline__ constexpr double a(int i, int j) { return ssb_tab::d5_a[i][j]; }
    static __device__ __forceinline__ constexpr double aa(int i, int j) { return ssb_tab::d5_aa[i][j]; }
    static __device__ __forceinline__ constexpr double c(int i) { return ssb_tab::d5_c[i]; }
    static __device__ __forceinline__ constexpr double rs(int i) { return ssb_tab::d5_rs[i]; }
    static __device__ __forceinline__ constexpr double e(int i) { return ssb_tab::d5_e[i]; }
    static __device__ __forceinline__ constexpr double ea(int i) { return ssb_tab::d5_ea[i]; }
    static __device__ __forceinline__ constexpr double esum() { return ssb_tab::d5_esum; }
    // operands from the constant bank (same values)
    static __device__ __forceinline__ double av(int i, int j) { return SSB_CTAB::d5_a[i][j]; }
    static __device__ __forceinline__ double aav(int i, int j) { return SSB_CTAB::d5_aa[i][j]; }
    static __device__ __forceinline__ double rsv(int i) { return SSB_CTAB::d5_rs[i]; }
    static __device__ __forceinline__ double ev(int i) { return SSB_CTAB::d5_e[i]; }
    static __device__ __forceinline__ double eav(int i) { return SSB_CTAB::d5_ea[i]; }
    static __device__ __forceinline__ double esumv() { return SSB_CTAB::d5_esum; }
};
template <> struct Tab<8> {
    static constexpr int S = 14, ORDER = 8;
    static __device__ __forceinline__ constexpr double a(int i, int j) { return ssb_tab::d8_a[i][j]; }
    static __device__ __forceinline__ constexpr double aa(int i, int j) { return ssb_tab::d8_aa[i][j]; }
    static __device__ __forceinline__ constexpr double c(int i) { return ssb_tab::d8_c[i]; }
    static __device__ __forceinline__ constexpr double rs(int i) { return ssb_tab::d8_rs[i]; }
    static __device__ __forceinline__ constexpr double e(int i) { return ssb_tab::d8_e[i]; }
    static __device__ __forceinline__ constexpr double ea(int i) { return ssb_tab::d8_ea[i]; }
    static __device__ __forceinline__ constexpr double esum() { return ssb_tab::d8_esum; }
    static __device__ __forceinline__ double av(int i, int j) { return SSB_CTAB::d8_a[i][j]; }
    static __device__ __forceinline__ double aav(int i, int j) { return SSB_CTAB::d8_aa[i][j]; }
    static __device__ __forceinline__ double rsv(int i) { return SSB_CTAB::d8_rs[i]; }
    static __device__ __forceinline__ double ev(int i) { return SSB_CTAB::d8_e[i]; }
    static __device__ __forceinline__ double eav(int i) { return SSB_CTAB::d8_ea[i]; }
    static __device__ __forceinline__ double esumv() { return SSB_CTAB::d8_esum; }
};

struct CtrlDev { double rtol, atol, dtmin, dtmax; int max_steps; };

// The helpers are written for systems of D second-order components (D = 3 for an orbit or a response block, D = 6 for the
// coupled first+second-order mass block); D is deduced from the array arguments.

// ---- stages 2..S-1 of one step attempt.  F[0] must hold the force at (x, t) (FSAL). ----
template <int SOLVER, class Force, int D>
__device__ __forceinline__ void rk_stages(Force& force, const double (&x)[D], const double (&p)[D], double t, double h,
                                          double (&F)[Tab<SOLVER>::S][D]) {
    typedef Tab<SOLVER> T;
    const double h2 = h * h;
    double hp[D];
#pragma unroll
    for (int k = 0; k < D; ++k) hp[k] = h * p[k];
#pragma unroll
    for (int i = 1; i < T::S - 1; ++i) {
        double X[D];
#pragma unroll
        for (int k = 0; k < D; ++k) {
            double acc = 0.0;
#pragma unroll
            for (int l = 0; l < i; ++l)
                if (T::aa(i, l) != 0.0) acc = fma(T::aav(i, l), F[l][k], acc);
            X[k] = fma(h2, acc, fma(T::rsv(i), hp[k], x[k]));            // x + c_i h p + h^2 sum_l (A.A)_il F_l
        }
        force(X, t + T::c(i) * h, F[i]);
    }
}

// candidate state = last stage value (the last tableau row equals b); caller then sets F[S-1] = force(x1, t + h)
template <int SOLVER, int D>
__device__ __forceinline__ void rk_candidate(const double (&x)[D], const double (&p)[D], double h, const double (&F)[Tab<SOLVER>::S][D],
                                             double (&x1)[D], double (&p1)[D]) {
    typedef Tab<SOLVER> T;
    constexpr int L = T::S - 1;
    const double h2 = h * h;
#pragma unroll
    for (int k = 0; k < D; ++k) {
        double ax = 0.0, ap = 0.0;
#pragma unroll
        for (int l = 0; l < L; ++l) {
            if (T::aa(L, l) != 0.0) ax = fma(T::aav(L, l), F[l][k], ax);
            if (T::a(L, l) != 0.0) ap = fma(T::av(L, l), F[l][k], ap);
        }
        x1[k] = fma(h2, ax, fma(T::rs(L) * h, p[k], x[k]));
        p1[k] = fma(h, ap, p[k]);
    }
}

// embedded error estimate  y_err = h sum e_i f_i  (needs all S force stages)
template <int SOLVER, int D>
__device__ __forceinline__ void rk_error(const double (&p)[D], double h, const double (&F)[Tab<SOLVER>::S][D], double (&ex)[D], double (&ep)[D]) {
    typedef Tab<SOLVER> T;
#pragma unroll
    for (int k = 0; k < D; ++k) {
        double bx = 0.0, bp = 0.0;
#pragma unroll
        for (int l = 0; l < T::S; ++l) {
            if (T::ea(l) != 0.0) bx = fma(T::eav(l), F[l][k], bx);
            if (T::e(l) != 0.0) bp = fma(T::ev(l), F[l][k], bp);
        }
        ex[k] = h * fma(h, bx, T::esumv() * p[k]);
        ep[k] = h * bp;
    }
}

// stage weights of the Dopri8 dense output at theta: wa[l] (positions, Nystrom form), wb[l] (momenta), wsum (coefficient of h p)
static __device__ __noinline__ void d8_dense_weights(double theta, double* __restrict__ wa, double* __restrict__ wb, double* __restrict__ wsum) {
#pragma unroll 1
    for (int l = 0; l < 14; ++l) {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int q = 6; q >= 0; --q) { a = fma(a, theta, ssb_tab::d8_dense_a[l][q]); b = fma(b, theta, ssb_tab::d8_dense[l][q]); }
        wa[l] = a * theta; wb[l] = b * theta;
    }
    double s = 0.0;
#pragma unroll
    for (int q = 6; q >= 0; --q) s = fma(s, theta, ssb_tab::d8_dense_sum[q]);
    *wsum = s * theta;
}

// dense output at theta in (0,1): the same interpolants the oracle uses (orc_solver.h), in Nystrom form
template <int SOLVER>
__device__ __forceinline__ void rk_dense(const double x[3], const double p[3], const double x1[3], const double p1[3], double h,
                                         const double (&F)[Tab<SOLVER>::S][3], double theta, double xo[3], double po[3]) {
    if constexpr (SOLVER == 5) {
        // diffrax Dopri5: quartic through y0, y1, k1 = h f0, k7 = h f1 and y_mid = y0 + h sum cmid_i f_i
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            double mx = 0.0, mp = 0.0;
#pragma unroll
            for (int l = 0; l < 7; ++l) {
                if (ssb_tab::d5_cmida[l] != 0.0) mx = fma(ssb_tab::d5_cmida[l], F[l][k], mx);
                if (ssb_tab::d5_cmid[l] != 0.0) mp = fma(ssb_tab::d5_cmid[l], F[l][k], mp);
            }
            const double xm = fma(h, fma(h, mx, ssb_tab::d5_cmidsum * p[k]), x[k]);
            const double pm = fma(h, mp, p[k]);
            {   // position component: f0 = h p, f1 = h p1
                const double f0 = h * p[k], f1 = h * p1[k], y0 = x[k], y1 = x1[k];
                const double a = 2 * (f1 - f0) - 8 * (y1 + y0) + 16 * xm;
                const double b = 5 * f0 - 3 * f1 + 18 * y0 + 14 * y1 - 32 * xm;
                const double c = f1 - 4 * f0 - 11 * y0 - 5 * y1 + 16 * xm;
                xo[k] = (((a * theta + b) * theta + c) * theta + f0) * theta + y0;
            }
            {   // momentum component: f0 = h F_1, f1 = h F_7
                const double f0 = h * F[0][k], f1 = h * F[6][k], y0 = p[k], y1 = p1[k];
                const double a = 2 * (f1 - f0) - 8 * (y1 + y0) + 16 * pm;
                const double b = 5 * f0 - 3 * f1 + 18 * y0 + 14 * y1 - 32 * pm;
                const double c = f1 - 4 * f0 - 11 * y0 - 5 * y1 + 16 * pm;
                po[k] = (((a * theta + b) * theta + c) * theta + f0) * theta + y0;
            }
        }
    } else {
        // our C1 5th-order continuous extension of RK8(7)13M (tools/derive_dopri8_dense.py).  The 28 stage weights come from a small
        // out-of-line routine with rolled loops over the coefficient tables: inlining ~200 immediate-operand FMAs here pushed the
        // saving kernels' step loop out of the instruction cache (ncu: 2.5 "no instruction" stall cycles per issue).
        double wa[14], wb[14], wsum;
        d8_dense_weights(theta, wa, wb, &wsum);
        double accx[3] = {0, 0, 0}, accp[3] = {0, 0, 0};
#pragma unroll
        for (int l = 0; l < 14; ++l) {
            bool anya = false, anyb = false;
#pragma unroll
            for (int q = 6; q >= 0; --q) { anya |= ssb_tab::d8_dense_a[l][q] != 0.0; anyb |= ssb_tab::d8_dense[l][q] != 0.0; }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (anya) accx[k] = fma(wa[l], F[l][k], accx[k]);
                if (anyb) accp[k] = fma(wb[l], F[l][k], accp[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            xo[k] = fma(h, fma(h, accx[k], wsum * p[k]), x[k]);
            po[k] = fma(h, accp[k], p[k]);
        }
    }
}

// sum of squared scaled errors of one 6-vector (x,p): sc = atol + rtol*max(|y0|,|y1|)  (NaN candidate -> y0)
template <int D>
__device__ __forceinline__ double err_sq(const double (&x)[D], const double (&p)[D], const double (&x1)[D], const double (&p1)[D],
                                         const double (&ex)[D], const double (&ep)[D], double rtol, double atol, bool nan_cand) {
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < D; ++k) {
        const double xc = nan_cand ? x[k] : x1[k], pc = nan_cand ? p[k] : p1[k];
        const double sx = fma(rtol, fmax(fabs(x[k]), fabs(xc)), atol);
        const double sp = fma(rtol, fmax(fabs(p[k]), fabs(pc)), atol);
        const double qx = ex[k] * frcp(sx), qp = ep[k] * frcp(sp);
        acc = fma(qx, qx, acc); acc = fma(qp, qp, acc);
    }
    return acc;
}
__device__ __forceinline__ double err_sq6(const double (&x)[3], const double (&p)[3], const double (&x1)[3], const double (&p1)[3],
                                          const double (&ex)[3], const double (&ep)[3], double rtol, double atol, bool nan_cand) {
    return err_sq<3>(x, p, x1, p1, ex, ep, rtol, atol, nan_cand);
}

// PIDController.adapt_step_size with pcoeff = dcoeff = 0: returns keep, updates h_next / at_dtmin
template <int ORDER>
__device__ __forceinline__ bool pid_update(double err, double dt, const CtrlDev& c, bool& at_dtmin, double& h_next, bool& bad) {
    const bool keep = (err < 1.0) || at_dtmin;
    double factor = 0.9 * inv_root<ORDER>(err);
    bad = isnan(factor);
    factor = fmin(fmax(factor, keep ? 1.0 : 0.2), 10.0);
    double hn = fmin(dt * factor, c.dtmax);
    at_dtmin = hn <= c.dtmin;
    h_next = fmax(hn, c.dtmin);
    return keep;
}

// Hairer-Norsett-Wanner initial step from d0 = rms(y0/sc), d1 = rms(f0/sc); second RHS evaluation by caller
__device__ __forceinline__ double hnw_h0(double d0, double d1) {
    return (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
}
template <int ORDER>
__device__ __forceinline__ double hnw_h1(double h0, double d1, double d2) {
    const double md = fmax(d1, d2);
    const double h1 = (md <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : pow(0.01 / md, 1.0 / ORDER);
    return fmin(100.0 * h0, h1);
}

}  // namespace ssb
#endif
