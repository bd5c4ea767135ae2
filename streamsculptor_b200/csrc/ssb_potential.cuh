// Closed-form Phi, grad Phi, Hess Phi for the component set of the hot path (sm_100a device code).
//
// The reference obtains every derivative by autodiff of a scalar potential
// (/root/reference/streamsculptor/main.py:37-65, fields.py:193); a fixed kernel needs them in closed form.
// Formulas follow potential.py (line numbers at each case) and SURVEY.md Appendix C; tests compare them with
// the oracle's AD of the same scalar formulas.
#ifndef SSB_POTENTIAL_CUH
#define SSB_POTENTIAL_CUH
#include <cuda_runtime.h>
#include <math.h>

#include "../../include/ssb200.h"
#include "ssb_fastmath.cuh"
#include "ssb_jet.cuh"

namespace ssb {

enum { WANT_PHI = 1, WANT_GRAD = 2, WANT_HESS = 4 };

// Hessian storage: symmetric 6 = {xx, yy, zz, xy, xz, yz}
struct Sym3 { double xx, yy, zz, xy, xz, yz; };

// ---------------------------------------------------------------------------------------------
// tracks (potential.py:581-600 linear; interpax 'cubic' streamhelpers.py:520)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int lower_bound_d(const double* __restrict__ a, int n, double v) {   // first i with a[i] >= v
    int lo = 0, hi = n;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (__ldg(a + mid) < v) lo = mid + 1; else hi = mid; }
    return lo;
}
__device__ __forceinline__ int upper_bound_d(const double* __restrict__ a, int n, double v) {   // first i with a[i] > v
    int lo = 0, hi = n;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (__ldg(a + mid) <= v) lo = mid + 1; else hi = mid; }
    return lo;
}

// Largest i with a[i] < v (= lower_bound - 1) / a[i] <= v (= upper_bound - 1), found from a proportional guess:
// the tables on the path are (near-)uniform in time (potential.py:581-600: 1000 knots on [-14000, 0]; Chen25 tracks:
// linspace + one extra knot), so the guess is exact up to +-1 and two or three cached loads replace a 10-deep
// dependent binary search per force evaluation.  Non-uniform tables fall back to the binary search.
// Segment lookup.  The tables on the path are (near-)uniform in time (potential.py:581-600: 1000 knots on [-14000, 0]; Chen25
// tracks: linspace + one extra knot), so a proportional guess is exact up to +-1; the two knot times that bracket the guess are
// the ones the interpolation needs anyway, so the common case costs ONE round of loads instead of a 10-deep dependent binary
// search per force evaluation.  T.t_first / T.inv_dt are prepared once per CTA by stage_potential (0 = not prepared).
//   LINEAR: i = clip(searchsorted(t, v, 'left') - 1, 0, n-2):  t[i] <  v <= t[i+1] inside the table
//   CUBIC : i = clip(searchsorted(t, v, 'right'), 1, n-1) - 1: t[i] <= v <  t[i+1] inside the table
template <bool CUBIC>
__device__ __forceinline__ int segment_guess(const ssb_track& T, double v) {
    double a0 = T.t_first, sc = T.inv_dt;
    if (sc == 0.0) { a0 = __ldg(T.t); sc = (double)(T.n - 1) / (__ldg(T.t + T.n - 1) - a0); }
    const int i = (int)((v - a0) * sc);
    return min(max(i, 0), T.n - 2);
}
// correct a guessed segment given its end times (ta, tb); returns the true segment, reloading ta/tb only if it moved
template <bool CUBIC>
__device__ __forceinline__ int segment_fix(const ssb_track& T, double v, int i, double& ta, double& tb) {
    const double* __restrict__ a = T.t;
    const int n = T.n;
    int walk = 0;
    if (CUBIC) {
        while (i > 0 && ta > v && walk < 4) { --i; tb = ta; ta = __ldg(a + i); ++walk; }
        while (i < n - 2 && tb <= v && walk < 4) { ++i; ta = tb; tb = __ldg(a + i + 1); ++walk; }
        if (walk >= 4) { i = min(max(upper_bound_d(a, n, v), 1), n - 1) - 1; ta = __ldg(a + i); tb = __ldg(a + i + 1); }
    } else {
        while (i > 0 && ta >= v && walk < 4) { --i; tb = ta; ta = __ldg(a + i); ++walk; }
        while (i < n - 2 && tb < v && walk < 4) { ++i; ta = tb; tb = __ldg(a + i + 1); ++walk; }
        if (walk >= 4) { i = min(max(lower_bound_d(a, n, v) - 1, 0), n - 2); ta = __ldg(a + i); tb = __ldg(a + i + 1); }
    }
    return i;
}

template <bool DERIV>
__device__ __forceinline__ void track_eval(const ssb_track& T, double tq, double c[3], double dc[3]) {
    const int n = T.n;
    if (T.kind == SSB_TRACK_LINEAR) {
        // knot times AND knot values of the guessed segment are fetched in one round of loads; a wrong guess (rare) reloads
        const int ig = segment_guess<false>(T, tq);
        double ta = __ldg(T.t + ig), tb = __ldg(T.t + ig + 1);
        double y0[3], y1[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) { y0[k] = __ldg(T.y + 3 * ig + k); y1[k] = __ldg(T.y + 3 * ig + 3 + k); }
        const int i = segment_fix<false>(T, tq, ig, ta, tb);
        if (i != ig) {
#pragma unroll
            for (int k = 0; k < 3; ++k) { y0[k] = __ldg(T.y + 3 * i + k); y1[k] = __ldg(T.y + 3 * i + 3 + k); }
        }
        const double ih = frcp(tb - ta), w = (tq - ta) * ih;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            c[k] = (1.0 - w) * y0[k] + w * y1[k];
            if (DERIV) dc[k] = (y1[k] - y0[k]) * ih;
        }
    } else {
        if (!(tq >= __ldg(T.t) && tq <= __ldg(T.t + n - 1))) {     // interpax extrap=False -> NaN
            const double qn = __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll
            for (int k = 0; k < 3; ++k) { c[k] = qn; if (DERIV) dc[k] = qn; }
            return;
        }
        const int ig = segment_guess<true>(T, tq);
        double ta = __ldg(T.t + ig), tb = __ldg(T.t + ig + 1);
        const int i = segment_fix<true>(T, tq, ig, ta, tb) + 1;     // searchsorted(side='right'), clipped to [1, n-1]
        const double dx = tb - ta, dxi = dx == 0.0 ? 0.0 : frcp(dx);
        const double u = (tq - ta) * dxi;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double f0 = __ldg(T.y + 3 * (i - 1) + k), f1 = __ldg(T.y + 3 * i + k);
            const double m0 = __ldg(T.s + 3 * (i - 1) + k) * dx, m1 = __ldg(T.s + 3 * i + k) * dx;
            const double c2 = 3.0 * (f1 - f0) - 2.0 * m0 - m1;
            const double c3 = 2.0 * (f0 - f1) + m0 + m1;
            c[k] = f0 + u * (m0 + u * (c2 + u * c3));
            if (DERIV) dc[k] = (m0 + u * (2.0 * c2 + 3.0 * u * c3)) * dxi;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// spherical building block: given x, the scalars q = Phi'/r and w = (Phi'' - Phi'/r)/r^2 give
//   grad = q x,   H_ij = q delta_ij + w x_i x_j
// ---------------------------------------------------------------------------------------------
template <int MODE>
__device__ __forceinline__ void add_spherical(const double x[3], double phi, double q, double w,
                                              double& P, double g[3], Sym3& H) {
    if (MODE & WANT_PHI) P += phi;
    if (MODE & WANT_GRAD) { g[0] = fma(q, x[0], g[0]); g[1] = fma(q, x[1], g[1]); g[2] = fma(q, x[2], g[2]); }
    if (MODE & WANT_HESS) {
        H.xx += fma(w * x[0], x[0], q); H.yy += fma(w * x[1], x[1], q); H.zz += fma(w * x[2], x[2], q);
        H.xy = fma(w * x[0], x[1], H.xy); H.xz = fma(w * x[0], x[2], H.xz); H.yz = fma(w * x[1], x[2], H.yz);
    }
}

// NFW (potential.py:80-84): Phi = -(GM/rs) ln(1+m)/m, m = r/rs  =>  Phi = -GM u / r, u = log1p(r/rs)
//   q  = Phi'/r = GM (u/r - 1/(r+rs)) / r^2
//   q' = GM [3/(r^3 (r+rs)) - 3u/r^4 + 1/(r^2 (r+rs)^2)],   w = q'/r
template <int MODE>
__device__ __forceinline__ void nfw_terms(double GM, double rs, double r2, double& phi, double& q, double& w) {
    const double ir = frsqrt(r2), r = r2 * ir;
    const double u = flog1p_pos(r * frcp(rs));
    const double irs = frcp(r + rs);
    const double ir2 = ir * ir;
    if (MODE & WANT_PHI) phi = -GM * u * ir;
    const double a = u * ir;                       // u/r
    q = GM * (a - irs) * ir2;
    if (MODE & WANT_HESS) w = GM * (3.0 * ir * (irs - a) + irs * irs) * ir2 * ir;     // q'/r
}

// Hernquist (potential.py:136-138): r = sqrt(|x|^2 + soft), Phi = -GM/(r+a)
//   q = GM / ((r+a)^2 r),  dq/dr = -GM [2/((r+a)^3 r) + 1/((r+a)^2 r^2)],  w = (dq/dr)/r
template <int MODE>
__device__ __forceinline__ void hernquist_terms(double GM, double a, double r2soft, double& phi, double& q, double& w) {
    const double ir = frsqrt(r2soft), r = r2soft * ir;
    const double ira = frcp(r + a);
    if (MODE & WANT_PHI) phi = -GM * ira;
    q = GM * ira * ira * ir;
    if (MODE & WANT_HESS) w = -q * ir * (2.0 * ira + ir);
}

// Plummer (potential.py:128-130): Phi = -GM (r^2+a^2)^(-1/2);  q = GM s^3, w = -3 GM s^5, s = (r^2+a^2)^(-1/2)
template <int MODE>
__device__ __forceinline__ void plummer_terms(double GM, double a, double r2, double& phi, double& q, double& w) {
    const double s = frsqrt(fma(a, a, r2));
    const double s2 = s * s;
    if (MODE & WANT_PHI) phi = -GM * s;
    q = GM * s * s2;
    if (MODE & WANT_HESS) w = -3.0 * q * s2;
}

// Isochrone (potential.py:120-122): Phi = -GM/(a + ww), ww = sqrt(r^2 + a^2): Hernquist form in ww
template <int MODE>
__device__ __forceinline__ void isochrone_terms(double GM, double a, double r2, double& phi, double& q, double& w) {
    hernquist_terms<MODE>(GM, a, fma(a, a, r2), phi, q, w);
}

// d/d r_s of the subhalo profiles and the x-gradient of that (potential.py:862-864, 1226-1228):
//   returns psi = dPhi/dr_s and qd with grad(psi) = qd * x
__device__ __forceinline__ void profile_dradius(int profile, double GM, double a, double r2, double& psi, double& qd) {
    if (profile == SSB_PROFILE_PLUMMER) {            // psi = GM a s^3, grad = -3 GM a s^5 x
        const double s = frsqrt(fma(a, a, r2)), s2 = s * s;
        psi = GM * a * s * s2; qd = -3.0 * psi * s2;
    } else if (profile == SSB_PROFILE_HERNQUIST) {   // psi = GM/(r+a)^2, grad = -2 GM/((r+a)^3 r) x
        const double ir = frsqrt(r2), r = r2 * ir, ira = frcp(r + a);
        psi = GM * ira * ira; qd = -2.0 * psi * ira * ir;
    } else {                                         // NFW: psi = GM/(a (a+r)), grad = -GM/(a (a+r)^2 r) x
        const double ir = frsqrt(r2), r = r2 * ir, ira = frcp(r + a);
        psi = GM * ira * frcp(a); qd = -psi * ira * ir;
    }
}
template <int MODE>
__device__ __forceinline__ void profile_terms(int profile, double GM, double a, double r2, double& phi, double& q, double& w) {
    if (profile == SSB_PROFILE_PLUMMER) plummer_terms<MODE>(GM, a, r2, phi, q, w);
    else if (profile == SSB_PROFILE_HERNQUIST) hernquist_terms<MODE>(GM, a, r2, phi, q, w);
    else nfw_terms<MODE>(GM, a, r2, phi, q, w);
}

// Miyamoto-Nagai (potential.py:70-72): zeta = sqrt(z^2+b^2), D = R^2 + (a+zeta)^2, Phi = -GM D^(-1/2)
template <int MODE>
__device__ __forceinline__ void add_miyamoto(double GM, double a, double b, const double x[3], double& P, double g[3], Sym3& H) {
    const double zb2 = fma(x[2], x[2], b * b);
    const double iz = frsqrt(zb2), zeta = zb2 * iz;
    const double az = a + zeta;
    const double D = fma(x[0], x[0], fma(x[1], x[1], az * az));
    const double id = frsqrt(D), id2 = id * id, id3 = id * id2;
    const double s = az * iz;                                       // (a+zeta)/zeta
    const double q = GM * id3;
    if (MODE & WANT_PHI) P -= GM * id;
    if (MODE & WANT_GRAD) { g[0] = fma(q, x[0], g[0]); g[1] = fma(q, x[1], g[1]); g[2] = fma(q * s, x[2], g[2]); }
    if (MODE & WANT_HESS) {
        const double q5 = 3.0 * q * id2;                            // 3 GM D^(-5/2)
        const double zs = x[2] * s;
        H.xx += q - q5 * x[0] * x[0]; H.yy += q - q5 * x[1] * x[1];
        H.xy -= q5 * x[0] * x[1]; H.xz -= q5 * x[0] * zs; H.yz -= q5 * x[1] * zs;
        H.zz += q * (s - a * x[2] * x[2] * iz * iz * iz) - q5 * zs * zs;
    }
}

// one subhalo of a SubhaloLine* set at relative position (potential.py:813-830); window gate strict '<'
template <int MODE>
__device__ __forceinline__ void add_subhalos(const ssb_subhalos& S, const double x[3], double t, double& P, double g[3], Sym3& H) {
    for (int j = 0; j < S.n; ++j) {
        const double dt = t - __ldg(S.t0 + j);
        if (!(fabs(dt) < __ldg(S.tw + j))) continue;
        double rel[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) rel[k] = x[k] - fma(__ldg(S.v + 3 * j + k), dt, __ldg(S.x0 + 3 * j + k));
        const double r2 = fma(rel[0], rel[0], fma(rel[1], rel[1], rel[2] * rel[2]));
        double phi = 0, q = 0, w = 0;
        profile_terms<MODE>(S.profile, S.G * __ldg(S.m + j), __ldg(S.rs + j), r2, phi, q, w);
        add_spherical<MODE>(rel, phi, q, w, P, g, H);
    }
}

// a set of moving spheres on one shared time grid (ssb_perturbers).  `pc` (optional): the n centres ALREADY interpolated at this t
// ([n][3]; kernels whose particles share the stage times prepare them once per stage and CTA).
#define SSB_PSET_FROZEN_MAX 104
__device__ __forceinline__ void pset_segment(const ssb_perturbers& S, double t, int& i, double& w) {
    const double* __restrict__ a = S.t;
    const int n = S.n_knots;
    const double a0 = __ldg(a), a1 = __ldg(a + n - 1);
    int g = (int)((t - a0) * ((double)(n - 1) / (a1 - a0)));
    g = min(max(g, 0), n - 2);
    double ta = __ldg(a + g), tb = __ldg(a + g + 1);
    int walk = 0;                                       // searchsorted(t, v, 'left') - 1 clipped to [0, n-2], as the linear tracks
    while (g > 0 && ta >= t && walk < 4) { --g; tb = ta; ta = __ldg(a + g); ++walk; }
    while (g < n - 2 && tb < t && walk < 4) { ++g; ta = tb; tb = __ldg(a + g + 1); ++walk; }
    if (walk >= 4) { g = min(max(lower_bound_d(a, n, t) - 1, 0), n - 2); ta = __ldg(a + g); tb = __ldg(a + g + 1); }
    i = g;
    w = (t - ta) * frcp(tb - ta);
}
__device__ __forceinline__ void pset_centres(const ssb_perturbers& S, double t, int j0, int j1, double* out /*[(j1-j0)][3]*/) {
    int i; double w;
    pset_segment(S, t, i, w);
    const double* __restrict__ y0 = S.y + (size_t)i * S.n * 3;
    const double* __restrict__ y1 = y0 + (size_t)S.n * 3;
    for (int j = j0; j < j1; ++j)
#pragma unroll
        for (int k = 0; k < 3; ++k) out[3 * (j - j0) + k] = (1.0 - w) * __ldg(y0 + 3 * j + k) + w * __ldg(y1 + 3 * j + k);
}
template <int MODE>
__device__ __forceinline__ void add_perturbers(const ssb_perturbers& S, const double x[3], double t, double& P, double g[3], Sym3& H, const double* pc) {
    int i = 0; double w = 0.0;
    const double* __restrict__ y0 = nullptr; const double* __restrict__ y1 = nullptr;
    if (!pc) { pset_segment(S, t, i, w); y0 = S.y + (size_t)i * S.n * 3; y1 = y0 + (size_t)S.n * 3; }
    const double w1 = 1.0 - w;
    for (int j = 0; j < S.n; ++j) {
        double rel[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) rel[k] = x[k] - (pc ? pc[3 * j + k] : (w1 * __ldg(y0 + 3 * j + k) + w * __ldg(y1 + 3 * j + k)));
        const double r2 = fma(rel[0], rel[0], fma(rel[1], rel[1], rel[2] * rel[2]));
        double phi = 0, q = 0, ww = 0;
        profile_terms<MODE>(S.profile, __ldg(S.GM + j), __ldg(S.rs + j), r2, phi, q, ww);
        add_spherical<MODE>(rel, phi, q, ww, P, g, H);
    }
}

// ---------------------------------------------------------------------------------------------
// total field: sum over the program (Potential_Combine.gradient_func, potential.py:1291-1296)
// ---------------------------------------------------------------------------------------------
// `frozen` (optional): [SSB_MAX_TRACK][6] = value c[3] and derivative dc[3] of every track ALREADY evaluated at this t - used where many
// evaluation points share one time (the shared-step kernels), so that the segment search and interpolation run once per stage, not per tracer.
// BARS = false: the instantiation contains no code for the rotating bars (jets, out-of-line calls).  The interpreter TAILS of the fused-signature
// kernels use it - their hot force functions stay as they were (a call inside them costs prologue spills on every evaluation: measured
// +23 % on the production response batch) - and ssb_canonicalize gives a program with a bar no fused signature, so it runs on the generic
// kernels, which carry the bar code.
template <int MODE, bool BARS = true>
__device__ __forceinline__ void pot_eval(const ssb_potential& Pt, const double x[3], double t, double& P, double g[3], Sym3& H, int first = 0,
                                         const double* frozen = nullptr, const double* frozen_pc = nullptr) {
    if (MODE & WANT_PHI) P = 0.0;
    if (MODE & WANT_GRAD) { g[0] = g[1] = g[2] = 0.0; }
    if (MODE & WANT_HESS) { H.xx = H.yy = H.zz = H.xy = H.xz = H.yz = 0.0; }
    const int nc = Pt.n_comp;
    for (int ic = first; ic < nc; ++ic) {
        const ssb_component& c = Pt.comp[ic];
        const int type = c.type;
        if (type == SSB_PERTURBERS) {
            add_perturbers<MODE>(Pt.pset[c.sh], x, t, P, g, H, frozen_pc);
            continue;
        }
        if (type == SSB_UNIFORM_ACC) {                              // potential.py:497-499
            if (MODE & WANT_GRAD) {
                double cv[3], dv[3];
                if (frozen) { dv[0] = frozen[6 * c.track + 3]; dv[1] = frozen[6 * c.track + 4]; dv[2] = frozen[6 * c.track + 5]; }
                else track_eval<true>(Pt.track[c.track], t, cv, dv);
                g[0] += dv[0]; g[1] += dv[1]; g[2] += dv[2];
            }
            continue;
        }
        double xs[3] = {x[0], x[1], x[2]};
        if (c.track >= 0) {                                         // potential.py:460-462
            double ctr[3];
            if (frozen) { ctr[0] = frozen[6 * c.track]; ctr[1] = frozen[6 * c.track + 1]; ctr[2] = frozen[6 * c.track + 2]; }
            else track_eval<false>(Pt.track[c.track], t, ctr, ctr);
            xs[0] -= ctr[0]; xs[1] -= ctr[1]; xs[2] -= ctr[2];
        }
        double gm = c.p[0];                                         // G m, times the tabulated growth factor (potential.py:474-477)
        if (c.growth > 0) {
            double gf[3];
            if (frozen) gf[0] = frozen[6 * (c.growth - 1)];
            else track_eval<false>(Pt.track[c.growth - 1], t, gf, gf);
            gm *= gf[0];
        }
        double phi = 0, q = 0, w = 0;
        switch (type) {
            case SSB_NFW: {
                const double r2 = fma(xs[0], xs[0], fma(xs[1], xs[1], xs[2] * xs[2]));
                nfw_terms<MODE>(gm, c.p[1], r2, phi, q, w);
                add_spherical<MODE>(xs, phi, q, w, P, g, H);
            } break;
            case SSB_HERNQUIST: {
                const double r2 = fma(xs[0], xs[0], fma(xs[1], xs[1], fma(xs[2], xs[2], c.p[2])));
                hernquist_terms<MODE>(gm, c.p[1], r2, phi, q, w);
                add_spherical<MODE>(xs, phi, q, w, P, g, H);
            } break;
            case SSB_PLUMMER: {
                const double r2 = fma(xs[0], xs[0], fma(xs[1], xs[1], xs[2] * xs[2]));
                plummer_terms<MODE>(gm, c.p[1], r2, phi, q, w);
                add_spherical<MODE>(xs, phi, q, w, P, g, H);
            } break;
            case SSB_ISOCHRONE: {
                const double r2 = fma(xs[0], xs[0], fma(xs[1], xs[1], xs[2] * xs[2]));
                isochrone_terms<MODE>(gm, c.p[1], r2, phi, q, w);
                add_spherical<MODE>(xs, phi, q, w, P, g, H);
            } break;
            case SSB_MIYAMOTO:
                add_miyamoto<MODE>(gm, c.p[1], c.p[2], xs, P, g, H);
                break;
            case SSB_TRIAXNFW: {                                    // potential.py:94: x_i / q_i
                const double i1 = frcp(c.p[2]), i2 = frcp(c.p[3]), i3 = frcp(c.p[4]);
                const double xq[3] = {xs[0] * i1, xs[1] * i2, xs[2] * i3};
                const double r2 = fma(xq[0], xq[0], fma(xq[1], xq[1], xq[2] * xq[2]));
                nfw_terms<MODE>(gm, c.p[1], r2, phi, q, w);
                if (MODE & WANT_PHI) P += phi;
                if (MODE & WANT_GRAD) { g[0] = fma(q * i1, xq[0], g[0]); g[1] = fma(q * i2, xq[1], g[1]); g[2] = fma(q * i3, xq[2], g[2]); }
                if (MODE & WANT_HESS) {
                    H.xx += (q + w * xq[0] * xq[0]) * i1 * i1; H.yy += (q + w * xq[1] * xq[1]) * i2 * i2; H.zz += (q + w * xq[2] * xq[2]) * i3 * i3;
                    H.xy += w * xq[0] * xq[1] * i1 * i2; H.xz += w * xq[0] * xq[2] * i1 * i3; H.yz += w * xq[1] * xq[2] * i2 * i3;
                }
            } break;
            case SSB_SUBHALOS:
                add_subhalos<MODE>(Pt.sh[c.sh], xs, t, P, g, H);
                break;
            case SSB_BAR:
            case SSB_DEHNEN_BAR: if constexpr (BARS) {              // rotating bars: derivatives by jets (ssb_jet.cuh), out of line
                constexpr int ORD = (MODE & WANT_HESS) ? 2 : 1;
                double J[jet_n(ORD)];
                if (type == SSB_BAR) bar_jet<ORD>(c.p, gm, xs[0], xs[1], xs[2], t, J);
                else dehnen_bar_jet<ORD>(c.p, gm, xs[0], xs[1], xs[2], t, J);
                if (MODE & WANT_PHI) P += J[0];
                if (MODE & WANT_GRAD) { g[0] += J[1]; g[1] += J[2]; g[2] += J[3]; }
                if (MODE & WANT_HESS) {                             // Taylor coefficients -> derivatives: xx, yy, zz carry 1/2!
                    H.xx += 2.0 * J[4]; H.xy += J[5]; H.xz += J[6]; H.yy += 2.0 * J[7]; H.yz += J[8]; H.zz += 2.0 * J[9];
                }
            } break;
            default: break;
        }
    }
}

// "Fast extras": the moving-perturber terms of the reference next to a fused galaxy signature - a spherical component translating
// on a LINEAR table (MW_LMC_Potential's LMC, potential.py:581-650; BASELINE config C3) or a uniform frame acceleration
// (potential.py:480-502, 646-650).  Their constants are gathered once per CTA into a compact shared-memory record; the evaluation
// is one straight-line block (segment guess -> ONE round of table loads -> interpolation -> force) that the scheduler overlaps with
// the galaxy arithmetic; a wrong segment guess (non-uniform table, or t within rounding of a knot) takes a rare out-of-line path.
struct FastX {
    double t_first, inv_dt, GM, a, soft;
    const double* t; const double* y;
    int n, type;
    const double* s; double t_last;       // cubic tracks (fastx_grad_cubic): knot slopes, last knot
};
// (inlined: a call site in the step loop, even a rarely taken one, costs spills around it in every stage)
static __device__ __forceinline__ void fastx_slow(const double* __restrict__ t, const double* __restrict__ y, int n, double v, double* out /*ta, tb, y0[3], y1[3]*/) {
    const int i = min(max(lower_bound_d(t, n, v) - 1, 0), n - 2);
    out[0] = __ldg(t + i); out[1] = __ldg(t + i + 1);
    for (int k = 0; k < 3; ++k) { out[2 + k] = __ldg(y + 3 * i + k); out[5 + k] = __ldg(y + 3 * i + 3 + k); }
}
__device__ __forceinline__ void fastx_fill(FastX* fx, const ssb_potential& Pt, int ic) {
    const ssb_component& c = Pt.comp[ic];
    const ssb_track& T = Pt.track[c.track];
    fx->t_first = T.t_first; fx->inv_dt = T.inv_dt; fx->t = T.t; fx->y = T.y; fx->n = T.n; fx->type = c.type;
    fx->GM = c.p[0]; fx->a = c.p[1]; fx->soft = c.type == SSB_HERNQUIST ? c.p[2] : 0.0;
    fx->s = T.s; fx->t_last = T.t[T.n - 1];
}
__device__ __forceinline__ void fastx_grad(const FastX& fx, const double x[3], double v, double g[3]) {
    const int n = fx.n;
    int i = (int)((v - fx.t_first) * fx.inv_dt);
    i = min(max(i, 0), n - 2);
    const double* __restrict__ tp = fx.t + i;
    const double* __restrict__ yp = fx.y + 3 * i;
    double ta = __ldg(tp), tb = __ldg(tp + 1);
    double y0[3], y1[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { y0[k] = __ldg(yp + k); y1[k] = __ldg(yp + 3 + k); }
    // searchsorted(t, v, 'left') - 1 clipped to [0, n-2]: t[i] < v <= t[i+1] inside the table
    const bool ok = (ta < v || i == 0) && (v <= tb || i == n - 2);
    if (!ok) {
        double o[8];
        fastx_slow(fx.t, fx.y, n, v, o);
        ta = o[0]; tb = o[1];
#pragma unroll
        for (int k = 0; k < 3; ++k) { y0[k] = o[2 + k]; y1[k] = o[5 + k]; }
    }
    const double ih = frcp(tb - ta), w = (v - ta) * ih, w1 = 1.0 - w;
    if (fx.type == SSB_UNIFORM_ACC) {                              // gradient = d v_frame / dt = slope of the segment
#pragma unroll
        for (int k = 0; k < 3; ++k) g[k] = fma(y1[k] - y0[k], ih, g[k]);
        return;
    }
    double xs[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) xs[k] = x[k] - (w1 * y0[k] + w * y1[k]);
    const double r2 = fma(xs[0], xs[0], fma(xs[1], xs[1], fma(xs[2], xs[2], fx.soft)));
    double phi, q, wq;
    if (fx.type == SSB_PLUMMER) plummer_terms<WANT_GRAD>(fx.GM, fx.a, r2, phi, q, wq);
    else if (fx.type == SSB_HERNQUIST) hernquist_terms<WANT_GRAD>(fx.GM, fx.a, r2, phi, q, wq);
    else if (fx.type == SSB_NFW) nfw_terms<WANT_GRAD>(fx.GM, fx.a, r2, phi, q, wq);
    else isochrone_terms<WANT_GRAD>(fx.GM, fx.a, r2, phi, q, wq);
    g[0] = fma(q, xs[0], g[0]); g[1] = fma(q, xs[1], g[1]); g[2] = fma(q, xs[2], g[2]);
}
// The same for ONE spherical component on a CUBIC track (interpax 'cubic': the progenitor's own potential in the Chen25 stream models,
// streamhelpers.py:520; NotAKnotTrack): segment guess, one round of 14 loads, the Hermite form of track_eval, rare out-of-line fix-up.
static __device__ __forceinline__ void fastx_slow_cubic(const double* __restrict__ t, const double* __restrict__ y, const double* __restrict__ s, int n, double v,
                                                     double* out /*ta, tb, y0[3], y1[3], s0[3], s1[3]*/) {
    const int i = min(max(upper_bound_d(t, n, v), 1), n - 1) - 1;
    out[0] = __ldg(t + i); out[1] = __ldg(t + i + 1);
    for (int k = 0; k < 3; ++k) {
        out[2 + k] = __ldg(y + 3 * i + k); out[5 + k] = __ldg(y + 3 * i + 3 + k);
        out[8 + k] = __ldg(s + 3 * i + k); out[11 + k] = __ldg(s + 3 * i + 3 + k);
    }
}
__device__ __forceinline__ void fastx_grad_cubic(const FastX& fx, const double x[3], double v, double g[3]) {
    const int n = fx.n;
    if (!(v >= fx.t_first && v <= fx.t_last)) {                    // interpax extrap=False: NaN outside the knots
        const double qn = __longlong_as_double(0x7ff8000000000000LL);
        g[0] = g[1] = g[2] = qn;
        return;
    }
    int i = (int)((v - fx.t_first) * fx.inv_dt);
    i = min(max(i, 0), n - 2);
    const double* __restrict__ tp = fx.t + i;
    double ta = __ldg(tp), tb = __ldg(tp + 1);
    double y0[3], y1[3], s0[3], s1[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        y0[k] = __ldg(fx.y + 3 * i + k); y1[k] = __ldg(fx.y + 3 * i + 3 + k);
        s0[k] = __ldg(fx.s + 3 * i + k); s1[k] = __ldg(fx.s + 3 * i + 3 + k);
    }
    // searchsorted(t, v, 'right') clipped to [1, n-1], minus 1: t[i] <= v < t[i+1] inside the table
    const bool ok = (ta <= v || i == 0) && (v < tb || i == n - 2);
    if (!ok) {
        double o[14];
        fastx_slow_cubic(fx.t, fx.y, fx.s, n, v, o);
        ta = o[0]; tb = o[1];
#pragma unroll
        for (int k = 0; k < 3; ++k) { y0[k] = o[2 + k]; y1[k] = o[5 + k]; s0[k] = o[8 + k]; s1[k] = o[11 + k]; }
    }
    const double dx = tb - ta, dxi = dx == 0.0 ? 0.0 : frcp(dx);
    const double u = (v - ta) * dxi;
    double xs[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {                                  // the Hermite form of track_eval
        const double f0 = y0[k], f1 = y1[k];
        const double m0 = s0[k] * dx, m1 = s1[k] * dx;
        const double c2 = 3.0 * (f1 - f0) - 2.0 * m0 - m1;
        const double c3 = 2.0 * (f0 - f1) + m0 + m1;
        xs[k] = x[k] - (f0 + u * (m0 + u * (c2 + u * c3)));
    }
    const double r2 = fma(xs[0], xs[0], fma(xs[1], xs[1], fma(xs[2], xs[2], fx.soft)));
    double phi, q, wq;
    if (fx.type == SSB_PLUMMER) plummer_terms<WANT_GRAD>(fx.GM, fx.a, r2, phi, q, wq);
    else if (fx.type == SSB_HERNQUIST) hernquist_terms<WANT_GRAD>(fx.GM, fx.a, r2, phi, q, wq);
    else if (fx.type == SSB_NFW) nfw_terms<WANT_GRAD>(fx.GM, fx.a, r2, phi, q, wq);
    else isochrone_terms<WANT_GRAD>(fx.GM, fx.a, r2, phi, q, wq);
    g[0] = fma(q, xs[0], g[0]); g[1] = fma(q, xs[1], g[1]); g[2] = fma(q, xs[2], g[2]);
}
// host side: exactly one extra, spherical, on a cubic track -> 1
static inline int ssb_fast_extra_cubic(const ssb_potential* p, int nf) {
    if (p->n_comp - nf != 1) return 0;
    const ssb_component& c = p->comp[nf];
    const bool kind_ok = c.type == SSB_NFW || c.type == SSB_HERNQUIST || c.type == SSB_PLUMMER || c.type == SSB_ISOCHRONE;
    return (kind_ok && c.growth == 0 && c.track >= 0 && p->track[c.track].kind == SSB_TRACK_CUBIC) ? 1 : 0;
}
// host side: do the components beyond the first `nf` qualify?  returns their number (1 or 2) or 0
static inline int ssb_fast_extras(const ssb_potential* p, int nf) {
    const int nx = p->n_comp - nf;
    if (nx < 1 || nx > 2) return 0;
    for (int i = nf; i < p->n_comp; ++i) {
        const ssb_component& c = p->comp[i];
        const bool kind_ok = c.type == SSB_NFW || c.type == SSB_HERNQUIST || c.type == SSB_PLUMMER || c.type == SSB_ISOCHRONE || c.type == SSB_UNIFORM_ACC;
        if (!kind_ok || c.growth != 0 || c.track < 0 || p->track[c.track].kind != SSB_TRACK_LINEAR) return 0;
    }
    return nx;
}

// ---------------------------------------------------------------------------------------------
// third derivatives of Phi (needed by jacfwd(release_model), perturbative.py:281-296, through main.py:76-104, and by the
// second-order response).  Not on the per-step hot path: plain IEEE arithmetic, runtime loops.
//   spherical:  Phi_ijk = u3 x_i x_j x_k + w (d_ij x_k + d_ik x_j + d_jk x_i),   w = q'/r,  u3 = w'/r,  q = Phi'/r
//   storage: 10 unique components {xxx, yyy, zzz, xxy, xxz, xyy, yyz, xzz, yzz, xyz}
// ---------------------------------------------------------------------------------------------
struct Sym3x3 { double xxx, yyy, zzz, xxy, xxz, xyy, yyz, xzz, yzz, xyz; };

__device__ inline void add_spherical_third(const double x[3], double w, double u3, double sc0, double sc1, double sc2, Sym3x3& T) {
    // sc*: per-axis chain-rule factors (1 except for the triaxial NFW)
    T.xxx += (u3 * x[0] * x[0] * x[0] + 3.0 * w * x[0]) * sc0 * sc0 * sc0;
    T.yyy += (u3 * x[1] * x[1] * x[1] + 3.0 * w * x[1]) * sc1 * sc1 * sc1;
    T.zzz += (u3 * x[2] * x[2] * x[2] + 3.0 * w * x[2]) * sc2 * sc2 * sc2;
    T.xxy += (u3 * x[0] * x[0] * x[1] + w * x[1]) * sc0 * sc0 * sc1;
    T.xxz += (u3 * x[0] * x[0] * x[2] + w * x[2]) * sc0 * sc0 * sc2;
    T.xyy += (u3 * x[0] * x[1] * x[1] + w * x[0]) * sc0 * sc1 * sc1;
    T.yyz += (u3 * x[1] * x[1] * x[2] + w * x[2]) * sc1 * sc1 * sc2;
    T.xzz += (u3 * x[0] * x[2] * x[2] + w * x[0]) * sc0 * sc2 * sc2;
    T.yzz += (u3 * x[1] * x[2] * x[2] + w * x[1]) * sc1 * sc2 * sc2;
    T.xyz += (u3 * x[0] * x[1] * x[2]) * sc0 * sc1 * sc2;
}
// radial scalars w = q'/r and u3 = w'/r
__device__ inline void nfw_wu(double GM, double rs, double r2, double& w, double& u3) {
    const double r = sqrt(r2), s1 = r + rs, u = log1p(r / rs);
    const double r3 = r2 * r, r4 = r2 * r2;
    w = GM * (3.0 / (r4 * s1) - 3.0 * u / (r4 * r) + 1.0 / (r3 * s1 * s1));
    u3 = GM * (-15.0 / (r4 * r * s1) - 6.0 / (r4 * s1 * s1) - 2.0 / (r3 * s1 * s1 * s1) + 15.0 * u / (r4 * r2)) / r;
}
__device__ inline void hernquist_wu(double GM, double a, double r2soft, double& w, double& u3) {
    const double r = sqrt(r2soft), ir = 1.0 / r, ira = 1.0 / (r + a);
    const double q = GM * ira * ira * ir;
    w = -q * ir * (2.0 * ira + ir);
    u3 = 3.0 * q * ir * ir * (2.0 * ira * ira + 2.0 * ira * ir + ir * ir);
}
__device__ inline void plummer_wu(double GM, double a, double r2, double& w, double& u3) {
    const double s = 1.0 / sqrt(r2 + a * a), s2 = s * s, s5 = s2 * s2 * s;
    w = -3.0 * GM * s5;
    u3 = 15.0 * GM * s5 * s2;
}
__device__ inline void profile_wu(int profile, double GM, double a, double r2, double& w, double& u3) {
    if (profile == SSB_PROFILE_PLUMMER) plummer_wu(GM, a, r2, w, u3);
    else if (profile == SSB_PROFILE_HERNQUIST) hernquist_wu(GM, a, r2, w, u3);
    else nfw_wu(GM, a, r2, w, u3);
}

__device__ inline void pot_third(const ssb_potential& Pt, const double x[3], double t, Sym3x3& T) {
    T.xxx = T.yyy = T.zzz = T.xxy = T.xxz = T.xyy = T.yyz = T.xzz = T.yzz = T.xyz = 0.0;
    for (int ic = 0; ic < Pt.n_comp; ++ic) {
        const ssb_component& c = Pt.comp[ic];
        if (c.type == SSB_UNIFORM_ACC) continue;
        if (c.type == SSB_PERTURBERS) {
            const ssb_perturbers& S = Pt.pset[c.sh];
            for (int j = 0; j < S.n; ++j) {
                double ctr[3], rel[3], w, u3;
                pset_centres(S, t, j, j + 1, ctr);
                for (int k = 0; k < 3; ++k) rel[k] = x[k] - ctr[k];
                profile_wu(S.profile, S.GM[j], S.rs[j], rel[0] * rel[0] + rel[1] * rel[1] + rel[2] * rel[2], w, u3);
                add_spherical_third(rel, w, u3, 1, 1, 1, T);
            }
            continue;
        }
        double xs[3] = {x[0], x[1], x[2]};
        if (c.track >= 0) {
            double ctr[3];
            track_eval<false>(Pt.track[c.track], t, ctr, ctr);
            xs[0] -= ctr[0]; xs[1] -= ctr[1]; xs[2] -= ctr[2];
        }
        const double r2 = xs[0] * xs[0] + xs[1] * xs[1] + xs[2] * xs[2];
        double w, u3;
        double gm = c.p[0];
        if (c.growth > 0) { double gf[3]; track_eval<false>(Pt.track[c.growth - 1], t, gf, gf); gm *= gf[0]; }
        switch (c.type) {
            case SSB_NFW: nfw_wu(gm, c.p[1], r2, w, u3); add_spherical_third(xs, w, u3, 1, 1, 1, T); break;
            case SSB_HERNQUIST: hernquist_wu(gm, c.p[1], r2 + c.p[2], w, u3); add_spherical_third(xs, w, u3, 1, 1, 1, T); break;
            case SSB_PLUMMER: plummer_wu(gm, c.p[1], r2, w, u3); add_spherical_third(xs, w, u3, 1, 1, 1, T); break;
            case SSB_ISOCHRONE: hernquist_wu(gm, c.p[1], r2 + c.p[1] * c.p[1], w, u3); add_spherical_third(xs, w, u3, 1, 1, 1, T); break;
            case SSB_TRIAXNFW: {
                const double i1 = 1.0 / c.p[2], i2 = 1.0 / c.p[3], i3 = 1.0 / c.p[4];
                const double xq[3] = {xs[0] * i1, xs[1] * i2, xs[2] * i3};
                nfw_wu(gm, c.p[1], xq[0] * xq[0] + xq[1] * xq[1] + xq[2] * xq[2], w, u3);
                add_spherical_third(xq, w, u3, i1, i2, i3, T);
            } break;
            case SSB_MIYAMOTO: {
                // Phi = f(D), f = -GM D^(-1/2), D = x^2 + y^2 + (a + zeta)^2, zeta = sqrt(z^2 + b^2):
                // Phi_ijk = f''' D_i D_j D_k + f'' (D_ij D_k + D_ik D_j + D_jk D_i) + f' D_ijk
                const double GM = gm, a = c.p[1], b = c.p[2];
                const double zeta = sqrt(xs[2] * xs[2] + b * b), az = a + zeta;
                const double D = xs[0] * xs[0] + xs[1] * xs[1] + az * az;
                const double sD = sqrt(D);
                const double f1 = 0.5 * GM / (D * sD), f2 = -0.75 * GM / (D * D * sD), f3 = 1.875 * GM / (D * D * D * sD);
                const double Dx = 2.0 * xs[0], Dy = 2.0 * xs[1], Dz = 2.0 * xs[2] * az / zeta;
                const double Dzz = 2.0 + 2.0 * a * b * b / (zeta * zeta * zeta);
                const double Dzzz = -6.0 * a * b * b * xs[2] / (zeta * zeta * zeta * zeta * zeta);
                T.xxx += f3 * Dx * Dx * Dx + f2 * 3.0 * 2.0 * Dx;
                T.yyy += f3 * Dy * Dy * Dy + f2 * 3.0 * 2.0 * Dy;
                T.zzz += f3 * Dz * Dz * Dz + f2 * 3.0 * Dzz * Dz + f1 * Dzzz;
                T.xxy += f3 * Dx * Dx * Dy + f2 * 2.0 * Dy;
                T.xxz += f3 * Dx * Dx * Dz + f2 * 2.0 * Dz;
                T.xyy += f3 * Dx * Dy * Dy + f2 * 2.0 * Dx;
                T.yyz += f3 * Dy * Dy * Dz + f2 * 2.0 * Dz;
                T.xzz += f3 * Dx * Dz * Dz + f2 * Dzz * Dx;
                T.yzz += f3 * Dy * Dz * Dz + f2 * Dzz * Dy;
                T.xyz += f3 * Dx * Dy * Dz;
            } break;
            case SSB_BAR:
            case SSB_DEHNEN_BAR: {
                double J[20];
                if (c.type == SSB_BAR) bar_jet<3>(c.p, gm, xs[0], xs[1], xs[2], t, J);
                else dehnen_bar_jet<3>(c.p, gm, xs[0], xs[1], xs[2], t, J);
                // xxx xxy xxz xyy xyz xzz yyy yyz yzz zzz at 10..19; derivative = alpha! x coefficient
                T.xxx += 6.0 * J[10]; T.xxy += 2.0 * J[11]; T.xxz += 2.0 * J[12]; T.xyy += 2.0 * J[13]; T.xyz += J[14];
                T.xzz += 2.0 * J[15]; T.yyy += 6.0 * J[16]; T.yyz += 2.0 * J[17]; T.yzz += 2.0 * J[18]; T.zzz += 6.0 * J[19];
            } break;
            case SSB_SUBHALOS: {
                const ssb_subhalos& S = Pt.sh[c.sh];
                for (int j = 0; j < S.n; ++j) {
                    const double dt = t - S.t0[j];
                    if (!(fabs(dt) < S.tw[j])) continue;
                    double rel[3];
                    for (int k = 0; k < 3; ++k) rel[k] = xs[k] - (S.x0[3 * j + k] + S.v[3 * j + k] * dt);
                    profile_wu(S.profile, S.G * S.m[j], S.rs[j], rel[0] * rel[0] + rel[1] * rel[1] + rel[2] * rel[2], w, u3);
                    add_spherical_third(rel, w, u3, 1, 1, 1, T);
                }
            } break;
            default: break;
        }
    }
}
__device__ inline double third_at(const Sym3x3& T, int i, int j, int k) {
    int c[3] = {0, 0, 0}; c[i]++; c[j]++; c[k]++;
    const int key = c[0] * 16 + c[1] * 4 + c[2];
    switch (key) {
        case 48: return T.xxx; case 12: return T.yyy; case 3: return T.zzz; case 36: return T.xxy; case 33: return T.xxz;
        case 24: return T.xyy; case 9: return T.yyz; case 18: return T.xzz; case 6: return T.yzz; default: return T.xyz;
    }
}

// ---------------------------------------------------------------------------------------------
// fused static signatures: the reference's canonical galaxy models evaluated without the interpreter.
// The host (ssb_canonicalize) moves the matching static components to the front of the program in the order below
// and stores two derived constants (NFW p[2] = 1/r_s, Miyamoto p[3] = b^2); parameters are read straight from the
// kernel-parameter constant bank (compile-time offsets), the radius is shared by the spherical terms.
//   SIG_N    : [NFW]                                    (tests.ipynb's pot_NFW)
//   SIG_NHM  : [NFW, Hernquist, Miyamoto]               (MW3: MW_LMC_Potential's Milky Way, BASELINE configs C1-C4)
//   SIG_NHHM : [NFW, Hernquist, Hernquist, Miyamoto]    (GalaMilkyWayPotential, potential.py:390-418)
// ---------------------------------------------------------------------------------------------
enum { SIG_GENERIC = 0, SIG_N = 1, SIG_NHM = 2, SIG_NHHM = 3 };
template <int SIG> struct SigInfo { static constexpr int NF = SIG == SIG_N ? 1 : SIG == SIG_NHM ? 3 : SIG == SIG_NHHM ? 4 : 0; };

template <int SIG, int MODE>
__device__ __forceinline__ void fused_eval(const ssb_potential& P, const double x[3], double g[3], Sym3& H) {
    const double R2 = fma(x[0], x[0], x[1] * x[1]);           // shared by the spherical radius and the Miyamoto-Nagai denominator
    const double r2 = fma(x[2], x[2], R2);
    const double ir = frsqrt(r2), r = r2 * ir, ir2 = ir * ir;
    // spherical components share x: accumulate q = sum Phi'/r and w = sum (Phi'' - Phi'/r)/r^2
    // NFW (comp 0): 1/(1 + r/r_s) = r_s/(r + r_s) feeds the log1p rounding correction
    const double irs = frcp(r + P.comp[0].p[1]);
    const double u = flog1p_tab(r * P.comp[0].p[2], P.comp[0].p[1] * irs);
    const double a = u * ir;
    double q = P.comp[0].p[0] * (a - irs) * ir2, w = 0.0;
    if (MODE & WANT_HESS) w = P.comp[0].p[0] * (3.0 * ir * (irs - a) + irs * irs) * ir2 * ir;
    if (SIG >= SIG_NHM) {                         // Hernquist (comp 1), soft == 0 guaranteed by the host
        const double ira = frcp(r + P.comp[1].p[1]);
        const double qh = P.comp[1].p[0] * ira * ira * ir;
        q += qh;
        if (MODE & WANT_HESS) w -= qh * ir * (2.0 * ira + ir);
    }
    if (SIG >= SIG_NHHM) {                        // second Hernquist (comp 2)
        const double ira = frcp(r + P.comp[2].p[1]);
        const double qh = P.comp[2].p[0] * ira * ira * ir;
        q += qh;
        if (MODE & WANT_HESS) w -= qh * ir * (2.0 * ira + ir);
    }
    if (MODE & WANT_HESS) {
        H.xx = fma(w * x[0], x[0], q); H.yy = fma(w * x[1], x[1], q); H.zz = fma(w * x[2], x[2], q);
        H.xy = w * x[0] * x[1]; H.xz = w * x[0] * x[2]; H.yz = w * x[1] * x[2];
    }
    if (SIG >= SIG_NHM) {                         // Miyamoto-Nagai (last fused comp)
        constexpr int M = SigInfo<SIG>::NF - 1;
        const double zb2 = fma(x[2], x[2], P.comp[M].p[3]);
        const double iz = frsqrt(zb2);
        const double az = fma(zb2, iz, P.comp[M].p[1]);
        const double D = fma(az, az, R2);
        const double id = frsqrt(D), id2 = id * id;
        const double qm = P.comp[M].p[0] * id * id2;
        const double s = az * iz;
        if (MODE & WANT_GRAD) { const double qx = q + qm; g[0] = qx * x[0]; g[1] = qx * x[1]; g[2] = fma(qm, s, q) * x[2]; }
        if (MODE & WANT_HESS) {
            const double q5 = 3.0 * qm * id2, zs = x[2] * s;
            H.xx += qm - q5 * x[0] * x[0]; H.yy += qm - q5 * x[1] * x[1];
            H.xy -= q5 * x[0] * x[1]; H.xz -= q5 * x[0] * zs; H.yz -= q5 * x[1] * zs;
            H.zz += qm * (s - P.comp[M].p[1] * x[2] * x[2] * iz * iz * iz) - q5 * zs * zs;
        }
    } else if (MODE & WANT_GRAD) { g[0] = q * x[0]; g[1] = q * x[1]; g[2] = q * x[2]; }
}
template <int SIG>
__device__ __forceinline__ void fused_grad(const ssb_potential& P, const double x[3], double g[3]) {
    Sym3 H;
    fused_eval<SIG, WANT_GRAD>(P, x, g, H);
}

// Move a fused static signature to the front of the program (summation order is free) and fill its derived constants
// (host side).  Returns the signature id; `out` is the program the kernels receive.
static inline int ssb_canonicalize(const ssb_potential* in, ssb_potential* out) {
    *out = *in;
    int idx_n = -1, idx_m = -1, idx_h[2] = {-1, -1}, nh = 0, nn = 0, nm = 0;
    for (int i = 0; i < in->n_comp; ++i) {
        const ssb_component& c = in->comp[i];
        if (c.track >= 0 || c.growth != 0) continue;
        if (c.type == SSB_NFW) { if (nn++ == 0) idx_n = i; }
        else if (c.type == SSB_HERNQUIST && c.p[2] == 0.0) { if (nh < 2) idx_h[nh] = i; nh++; }
        else if (c.type == SSB_MIYAMOTO) { if (nm++ == 0) idx_m = i; }
    }
    int sig = SIG_GENERIC, order[4], nf = 0;
    for (int i = 0; i < in->n_comp; ++i)
        if (in->comp[i].type == SSB_BAR || in->comp[i].type == SSB_DEHNEN_BAR) return sig;   // bars: generic kernels only (pot_eval<.., BARS>)
    if (nn >= 1 && nh >= 2 && nm >= 1) { sig = SIG_NHHM; order[0] = idx_n; order[1] = idx_h[0]; order[2] = idx_h[1]; order[3] = idx_m; nf = 4; }
    else if (nn >= 1 && nh >= 1 && nm >= 1) { sig = SIG_NHM; order[0] = idx_n; order[1] = idx_h[0]; order[2] = idx_m; nf = 3; }
    else if (nn >= 1) { sig = SIG_N; order[0] = idx_n; nf = 1; }
    if (sig == SIG_GENERIC) return sig;
    bool used[SSB_MAX_COMP] = {false};
    int k = 0;
    for (int j = 0; j < nf; ++j) { out->comp[k++] = in->comp[order[j]]; used[order[j]] = true; }
    for (int i = 0; i < in->n_comp; ++i) if (!used[i]) out->comp[k++] = in->comp[i];
    out->comp[0].p[2] = 1.0 / out->comp[0].p[1];                                             // NFW: 1 / r_s
    if (nf >= 3) out->comp[nf - 1].p[3] = out->comp[nf - 1].p[2] * out->comp[nf - 1].p[2];   // Miyamoto-Nagai: b^2
    return sig;
}

}  // namespace ssb
#endif
