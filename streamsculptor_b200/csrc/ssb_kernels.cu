// libssb200: sm_100a kernels + C ABI (include/ssb200.h) for the streamsculptor hot path.
//
//   K1 orbit_kernel        one thread per orbit, stepper state in registers        (main.py:125-202, A3/A4)
//   K0 dense_step_kernel + dense_eval_kernel   one orbit saved at M times            (main.py:289, A7)
//   K2 release_kernel      particle-spray ICs incl. jax threefry normals            (main.py:209-306, A6)
//   potential_eval_kernel, subhalo_eval_kernel, track kernels                       (main.py:37-65, A1/A5/A10/A11)
// The linear-response kernel lives in ssb_response.cu.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <limits>
#include <vector>

// tableau coefficients as immediates in this translation unit (the one-thread-per-orbit kernels); measured on B200, constant bank vs
// immediates: orbit_kernel<8> final state 11.56 vs 11.30 ms, Dopri5 8.41 vs 8.00 ms (immediates win), saving kernel 21.7 vs 22.5 ms,
// response kernel 31.4 vs 33.7 ms (constant bank wins: ssb_response.cu keeps the default)
#ifndef SSB_TABLEAU_CONSTBANK
#define SSB_TABLEAU_CONSTBANK 0
#endif
#include "ssb_common.cuh"

using namespace ssb;

// =============================================================================================
// force functors
// =============================================================================================
// The force is a real function call (not inlined 13x into the unrolled stepper): the stepper body stays inside
// the instruction cache and ptxas allocates the stage registers once.
template <bool BARS>
__device__ __noinline__ double3 accel_call(const ssb_potential* P, int first, double x, double y, double z, double t) {
    const double X[3] = {x, y, z};
    double phi, g[3];
    Sym3 H;
    pot_eval<WANT_GRAD, BARS>(*P, X, t, phi, g, H, first);
    return make_double3(-g[0], -g[1], -g[2]);
}
template <int SIG>
__device__ __noinline__ double3 fused_call(const ssb_potential* P, double x, double y, double z) {
    const double X[3] = {x, y, z};
    double g[3];
    fused_grad<SIG>(*P, X, g);
    return make_double3(-g[0], -g[1], -g[2]);
}
// Potential.velocity_acceleration (main.py:116-120) in mirrored time.  SIG != 0: the leading NF components are a fused
// static signature evaluated inline from the constant bank (Pc); any remaining components go through the interpreter (P).
template <int SIG, int XS = 0, int INL = SSB_FUSED_INLINE>
struct OrbitForce {
    const ssb_potential* P;      // shared-memory copy (dynamic indexing)
    const ssb_potential* Pc;     // kernel parameter (constant bank)
    double dir;
    bool extra;                  // components beyond the fused ones exist
    const FastX* fx;             // XS > 0: the XS extras are "fast extras" (ssb_potential.cuh), evaluated inline
    __device__ __forceinline__ void operator()(const double X[3], double tau, double A[3]) const {
        if (SIG == SIG_GENERIC) {
            const double3 a = accel_call<true>(P, 0, X[0], X[1], X[2], tau * dir);
            A[0] = a.x; A[1] = a.y; A[2] = a.z;
        } else {
            if (INL) {         // 13 inlined copies of the fused force: fastest while the unrolled step loop still fits the instruction cache (final-state kernel)
                double g[3];
                fused_grad<SIG>(*Pc, X, g);
                A[0] = -g[0]; A[1] = -g[1]; A[2] = -g[2];
            } else {           // one out-of-line copy: 16 KB less code in the step loop (the saving kernel's loop + dense output would overflow the cache)
                const double3 f = fused_call<SIG>(Pc, X[0], X[1], X[2]);
                A[0] = f.x; A[1] = f.y; A[2] = f.z;
            }
            if (XS != 4 && extra) {       // XS == 4: the program IS the fused signature (the headline stream): no extras path in the step loop at all
                if (XS > 0) {
                    double g2[3] = {0.0, 0.0, 0.0};
                    if (XS == 3) {                                                   // one extra on a cubic track
                        fastx_grad_cubic(fx[0], X, tau * dir, g2);
                    } else {
#pragma unroll
                        for (int e = 0; e < (XS == 3 ? 0 : XS); ++e) fastx_grad(fx[e], X, tau * dir, g2);
                    }
                    A[0] -= g2[0]; A[1] -= g2[1]; A[2] -= g2[2];
                } else {
                    const double3 a = accel_call<false>(P, SigInfo<SIG>::NF, X[0], X[1], X[2], tau * dir);
                    A[0] += a.x; A[1] += a.y; A[2] += a.z;
                }
            }
        }
    }
};

// =============================================================================================
// warp-cooperative dense output of the saving orbit kernel (K1 MODE 0: integrate_orbit_batch_vmapped with ts[N,M], main.py:186-202)
// =============================================================================================
// srec: the CTA's step records, field f of thread t at srec[f * SSB_ORBIT_THREADS + t]:
//   0 tprev, 1 tnext (mirrored time), 2..4 x, 5..7 p (step start), 8..10 x1, 11..13 p1 (step end), 14 + 3 l + k: force stage l.
// nsave: save times of THIS lane inside its published step (0: none); they are ts[save_idx .. save_idx + nsave) of its row tsp and go
// to rows save_idx.. of its output block ys.  Task j of the warp (j < sum of nsave) belongs to the lane s with excl_s <= j < incl_s
// (inclusive scan); tasks are dealt out j = lane, lane + 32, ...: every pass keeps all lanes busy and the lanes that serve one orbit
// write adjacent 48-byte rows.
// Coefficient tables of the dense output staged in shared memory (s_coef: Dopri8 d8_dense_a[14][7], d8_dense[14][7], d8_dense_sum[7];
// Dopri5 d5_cmida[7], d5_cmid[7]): the rolled stage loop indexes them dynamically, and indexed loads from the constant bank miss the
// small constant cache that the step loop's kernel parameters occupy (measured: ~10 k cycles per pass before, as long as a whole step).
template <int SOLVER>
__device__ __forceinline__ void dense_coef_init(double* s_coef) {
    if constexpr (SOLVER == 8) {
        for (int i = threadIdx.x; i < 98; i += blockDim.x) { s_coef[i] = ssb_tab::d8_dense_a[i / 7][i % 7]; s_coef[98 + i] = ssb_tab::d8_dense[i / 7][i % 7]; }
        if (threadIdx.x < 7) s_coef[196 + threadIdx.x] = ssb_tab::d8_dense_sum[threadIdx.x];
    } else {
        if (threadIdx.x < 7) { s_coef[threadIdx.x] = ssb_tab::d5_cmida[threadIdx.x]; s_coef[7 + threadIdx.x] = ssb_tab::d5_cmid[threadIdx.x]; }
    }
    __syncthreads();
}
#ifndef SSB_COOP_INLINE
#define SSB_COOP_INLINE 0          // 1: the cooperative dense output inlined into the step loop (A/B: no call, but the loop body grows)
#endif
#if SSB_COOP_INLINE
#define SSB_COOP_QUAL __device__ __forceinline__
#else
#define SSB_COOP_QUAL __device__ __noinline__
#endif
template <int SOLVER>
SSB_COOP_QUAL void coop_dense(const double* __restrict__ srec, const double* __restrict__ s_coef, int nsave, int save_idx, const double* tsp, double* ys,
                                        double dir) {
    constexpr int NT = SSB_ORBIT_THREADS;
    constexpr int S = Tab<SOLVER>::S;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, wbase = threadIdx.x & ~31;
    int incl = nsave;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(full, incl, o); if (lane >= o) incl += v; }
    const int total = __shfl_sync(full, incl, 31);
    const int excl = incl - nsave;
    for (int j0 = 0; j0 < total; j0 += 32) {
        const int j = j0 + lane;
        int lo = 0, hi = 31;                                  // smallest s with incl_s > j
#pragma unroll
        for (int it = 0; it < 5; ++it) { const int mid = (lo + hi) >> 1; const int v = __shfl_sync(full, incl, mid); if (v > j) hi = mid; else lo = mid + 1; }
        const int s = lo;
        const int r = j - __shfl_sync(full, excl, s);
        const int m = __shfl_sync(full, save_idx, s) + r;
        const double* ts_s = reinterpret_cast<const double*>(__shfl_sync(full, reinterpret_cast<unsigned long long>(tsp), s));
        double* ys_s = reinterpret_cast<double*>(__shfl_sync(full, reinterpret_cast<unsigned long long>(ys), s));
        const double dir_s = __shfl_sync(full, dir, s);
        if (j >= total) continue;
        const double* R = srec + wbase + s;
        const double tq = ts_s[m] * dir_s;
        const double ta = R[0], tb = R[NT], h = tb - ta;
        double xo[3], po[3];
        if (tq == tb) {
#pragma unroll
            for (int k = 0; k < 3; ++k) { xo[k] = R[(8 + k) * NT]; po[k] = R[(11 + k) * NT]; }
        } else {
            const double theta = (tq - ta) / h;
            double x0[3], p0[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) { x0[k] = R[(2 + k) * NT]; p0[k] = R[(5 + k) * NT]; }
            if constexpr (SOLVER == 5) {
                // diffrax Dopri5: quartic through y0, y1, k1 = h f0, k7 = h f1 and y_mid = y0 + h sum cmid_i f_i (as rk_dense<5>)
                double mx[3] = {0, 0, 0}, mp[3] = {0, 0, 0};
#pragma unroll 1
                for (int l = 0; l < S; ++l) {
                    const double ca = s_coef[l], cb = s_coef[7 + l];
#pragma unroll
                    for (int k = 0; k < 3; ++k) { const double f = R[(14 + 3 * l + k) * NT]; mx[k] = fma(ca, f, mx[k]); mp[k] = fma(cb, f, mp[k]); }
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const double x1k = R[(8 + k) * NT], p1k = R[(11 + k) * NT];
                    const double xm = fma(h, fma(h, mx[k], ssb_tab::d5_cmidsum * p0[k]), x0[k]);
                    const double pm = fma(h, mp[k], p0[k]);
                    {
                        const double f0 = h * p0[k], f1 = h * p1k, y0 = x0[k], y1 = x1k;
                        const double a = 2 * (f1 - f0) - 8 * (y1 + y0) + 16 * xm;
                        const double b = 5 * f0 - 3 * f1 + 18 * y0 + 14 * y1 - 32 * xm;
                        const double cc = f1 - 4 * f0 - 11 * y0 - 5 * y1 + 16 * xm;
                        xo[k] = (((a * theta + b) * theta + cc) * theta + f0) * theta + y0;
                    }
                    {
                        const double f0 = h * R[(14 + k) * NT], f1 = h * R[(14 + 3 * (S - 1) + k) * NT], y0 = p0[k], y1 = p1k;
                        const double a = 2 * (f1 - f0) - 8 * (y1 + y0) + 16 * pm;
                        const double b = 5 * f0 - 3 * f1 + 18 * y0 + 14 * y1 - 32 * pm;
                        const double cc = f1 - 4 * f0 - 11 * y0 - 5 * y1 + 16 * pm;
                        po[k] = (((a * theta + b) * theta + cc) * theta + f0) * theta + y0;
                    }
                }
            } else {
                // our C1 5th-order continuous extension of RK8(7)13M (tools/derive_dopri8_dense.py), stage weights by Horner in theta:
                // rolled loop over the stages, coefficients from the constant tables, stage forces from the shared record
                double accx[3] = {0, 0, 0}, accp[3] = {0, 0, 0};
#pragma unroll 1
                for (int l = 0; l < S; ++l) {
                    double wa = 0.0, wb = 0.0;
#pragma unroll
                    for (int q = 6; q >= 0; --q) { wa = fma(wa, theta, s_coef[7 * l + q]); wb = fma(wb, theta, s_coef[98 + 7 * l + q]); }
                    wa *= theta; wb *= theta;
#pragma unroll
                    for (int k = 0; k < 3; ++k) { const double f = R[(14 + 3 * l + k) * NT]; accx[k] = fma(wa, f, accx[k]); accp[k] = fma(wb, f, accp[k]); }
                }
                double wsum = 0.0;
#pragma unroll
                for (int q = 6; q >= 0; --q) wsum = fma(wsum, theta, s_coef[196 + q]);
                wsum *= theta;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    xo[k] = fma(h, fma(h, accx[k], wsum * p0[k]), x0[k]);
                    po[k] = fma(h, accp[k], p0[k]);
                }
            }
        }
        double* o = ys_s + (size_t)m * 6;
#pragma unroll
        for (int k = 0; k < 3; ++k) { o[k] = xo[k]; o[3 + k] = dir_s * po[k]; }
    }
}

// =============================================================================================
// K1: batch of independent adaptive solves
// =============================================================================================
struct OrbitArgs {
    int64_t N;
    const double *w0, *t0, *t1, *ts;
    int M, ts_per_orbit;
    double* ys;
    int32_t *status, *nsteps;
    CtrlDev c;
};

// per-thread integration of one orbit; REC != nullptr records accepted steps (K0) instead of saving
// MODE: 0 = SaveAt(ts) with dense output, 1 = RECORD accepted steps (K0), 2 = final state only (ts == t1; no dense-output
// code in the instruction stream - the stream-generation hot path)
template <int SOLVER, int MODE, int SIG, int XS = 0>
__device__ __forceinline__ void integrate_one(const ssb_potential* P, const ssb_potential* Pc, const double* w0, double t0_in, double t1_in,
                                              const double* tsp, int M, double* ys, const CtrlDev& c, bool valid,
                                              int& status, int& n_steps, int& n_acc, int& n_rej, double* rec, int rec_cap,
                                              const FastX* fxp = nullptr, double t_stop = HUGE_VAL, double* trace = nullptr, int trace_cap = 0,
                                              double* srec = nullptr) {
    typedef Tab<SOLVER> T;
    constexpr int S = T::S;
    const double dir = (t0_in < t1_in) ? 1.0 : -1.0;              // diffrax: direction = where(t0 < t1, 1, -1)
    const double T0 = t0_in * dir, T1 = t1_in * dir;
    OrbitForce<SIG, XS, (MODE == 0 ? SSB_SNAP_FUSED_INLINE : SSB_FUSED_INLINE)> force{P, Pc, dir, Pc->n_comp > SigInfo<SIG>::NF, fxp};
    double x[3], p[3], F[S][3];
    status = 0; n_steps = 0; n_acc = 0; n_rej = 0;
    double tprev = T0, tnext = T0, h = 0.0;
    bool at_dtmin = false;
    int save_idx = 0;
    if (valid) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { x[k] = w0[k]; p[k] = dir * w0[3 + k]; }
        // ---- PIDController.init: Hairer-Norsett-Wanner initial step ----
        force(x, T0, F[0]);
        double sx[3], sp[3], d0 = 0.0, d1 = 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            sx[k] = fma(c.rtol, fabs(x[k]), c.atol); sp[k] = fma(c.rtol, fabs(p[k]), c.atol);
            double q;
            q = x[k] / sx[k]; d0 = fma(q, q, d0); q = p[k] / sp[k]; d0 = fma(q, q, d0);
            q = p[k] / sx[k]; d1 = fma(q, q, d1); q = F[0][k] / sp[k]; d1 = fma(q, q, d1);
        }
        d0 = sqrt(d0 / 6.0); d1 = sqrt(d1 / 6.0);
        const double h0 = hnw_h0(d0, d1);
        double X1[3], F1[3], d2 = 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) X1[k] = fma(h0, p[k], x[k]);
        force(X1, T0 + h0, F1);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            double q;
            q = (fma(h0, F[0][k], p[k]) - p[k]) / sx[k]; d2 = fma(q, q, d2);     // f1 - f0, position rows: p1 - p
            q = (F1[k] - F[0][k]) / sp[k]; d2 = fma(q, q, d2);
        }
        d2 = sqrt(d2 / 6.0) / h0;
        h = fmin(hnw_h1<T::ORDER>(h0, d1, d2), c.dtmax);
        at_dtmin = h <= c.dtmin;
        h = fmax(h, c.dtmin);
        tnext = fmin(T0 + h, T1);
    }
    double tq_a = __longlong_as_double(0x7ff0000000000000LL), tq_b = tq_a;      // MODE 0: the next two save times (mirrored time)
    if (MODE == 0 && valid) {
        if (M > 0) tq_a = tsp[0] * dir;
        if (M > 1) tq_b = tsp[1] * dir;
    }
    if constexpr (MODE == 0) {
        // SaveAt(ts) with WARP-COOPERATIVE dense output.  A lane whose accepted step covers save times publishes the step (times,
        // end points, force stages) in its shared-memory record; the (orbit, save time) pairs of the whole warp are then dealt out
        // evenly over the 32 lanes (coop_dense), so the interpolation runs converged whatever the lanes' step sequences are and
        // adjacent rows of one orbit are written by adjacent lanes.
        constexpr int NT = SSB_ORBIT_THREADS;
        double* myrec = srec + threadIdx.x;
        const double* s_coef = srec + (14 + 3 * S) * NT;          // filled by the kernel (dense_coef_init)
        for (;;) {
            bool active = valid && status == 0 && tprev < T1;
            if (active && n_steps >= c.max_steps) { status = 1; active = false; }
            if (SSB_ORBIT_CTA_ALIGN & 1) { if (!__syncthreads_or(active)) break; }          // the CTA's warps stay on the same iteration: shared instruction fetch
            else if (!__any_sync(0xffffffffu, active)) break;
            int nsave = 0;
            if (active) {
                const double dt = tnext - tprev;
                double x1[3], p1[3], ex[3], ep[3];
                rk_stages<SOLVER>(force, x, p, tprev, dt, F);
                rk_candidate<SOLVER>(x, p, dt, F, x1, p1);
                force(x1, tprev + T::c(S - 1) * dt, F[S - 1]);
                // an attempt that would cover a save time (known before it is computed) is published BEFORE the controller runs: the force
                // stages are still live for the error estimate here, whereas a store after the accept / reject logic would stretch 39
                // doubles over the controller code and spill; the record of a rejected attempt is simply never read
                if (tq_a <= tnext) {
                    myrec[0] = tprev; myrec[NT] = tnext;
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        myrec[(2 + k) * NT] = x[k]; myrec[(5 + k) * NT] = p[k]; myrec[(8 + k) * NT] = x1[k]; myrec[(11 + k) * NT] = p1[k];
                    }
#pragma unroll
                    for (int l = 0; l < S; ++l)
#pragma unroll
                        for (int k = 0; k < 3; ++k) myrec[(14 + 3 * l + k) * NT] = F[l][k];
                }
                rk_error<SOLVER>(p, dt, F, ex, ep);
                bool nan_cand = false, finite = true;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    nan_cand |= isnan(x1[k]) | isnan(p1[k]);
                    finite &= isfinite(x1[k]) & isfinite(p1[k]);
                }
                const double err = sqrt(err_sq6(x, p, x1, p1, ex, ep, c.rtol, c.atol, nan_cand) / 6.0);
                double hn; bool bad;
                const bool keep = pid_update<T::ORDER>(err, dt, c, at_dtmin, hn, bad);
                n_steps++;
                if (bad) { status = 2; n_rej++; }
                else {
                    if (keep) {
                        n_acc++;
                        if (!finite) status = 2;
                        else {
                            if (tq_a <= tnext) {           // every ts[save_idx + r] <= tnext is interpolated inside this accepted step
                                // number of save times inside the step (ts is monotone).  The common cases - one save time, the next one
                                // beyond the step - are decided from the two prefetched values; otherwise gallop, then bisect
                                nsave = 1;
                                if (tq_b <= tnext) {
                                    const int left = M - save_idx;
                                    int hi = 2;
                                    while (hi < left && tsp[save_idx + hi] * dir <= tnext) hi <<= 1;
                                    int lo = hi >> 1;                   // row lo is inside the step ...
                                    if (hi > left) hi = left;           // ... row hi is outside (or the end of the list)
                                    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (tsp[save_idx + mid] * dir <= tnext) lo = mid; else hi = mid; }
                                    nsave = hi;
                                }
                            }
#pragma unroll
                            for (int k = 0; k < 3; ++k) { x[k] = x1[k]; p[k] = p1[k]; F[0][k] = F[S - 1][k]; }    // FSAL
                            tprev = tnext;
                        }
                    } else {
                        n_rej++;
                    }
                    if (status == 0) {
                        tprev = fmin(tprev, T1);
                        double tn = tprev + hn;
                        if (tn > T1 - 1e-10) tn = keep ? T1 : tprev + 0.5 * (T1 - tprev);     // diffrax _clip_to_end (f64)
                        tnext = tn;
                    }
                }
            }
            if (__any_sync(0xffffffffu, nsave > 0)) {
                __syncwarp();
                coop_dense<SOLVER>(srec, s_coef, nsave, save_idx, tsp, ys, dir);
                __syncwarp();
                if (nsave > 0) {           // reload the two look-ahead save times: not needed before the end of the next step
                    save_idx += nsave;
                    tq_a = (save_idx < M) ? tsp[save_idx] * dir : __longlong_as_double(0x7ff0000000000000LL);
                    tq_b = (save_idx + 1 < M) ? tsp[save_idx + 1] * dir : __longlong_as_double(0x7ff0000000000000LL);
                }
            }
        }
        if (valid) {               // rows never reached stay +inf (diffrax SaveAt semantics); every saved row was written exactly once
            const double inf = __longlong_as_double(0x7ff0000000000000LL);
            for (int m = save_idx; m < M; ++m)
#pragma unroll
                for (int k = 0; k < 6; ++k) ys[(size_t)m * 6 + k] = inf;
        }
        return;
    }
    for (;;) {
        bool active = valid && status == 0 && tprev < T1;
        if (MODE == 1) active = active && tprev < t_stop;            // K0 part runs (t_stop in mirrored time): same steps as the full solve, cut short
        if (active && n_steps >= c.max_steps) { status = 1; active = false; }
        if (MODE == 2 && (SSB_ORBIT_CTA_ALIGN & 2)) { if (!__syncthreads_or(active)) break; }
        else if (!__any_sync(0xffffffffu, active)) break;
        if (!active) continue;
        const double dt = tnext - tprev;
        double x1[3], p1[3], ex[3], ep[3];
        rk_stages<SOLVER>(force, x, p, tprev, dt, F);
        rk_candidate<SOLVER>(x, p, dt, F, x1, p1);
        force(x1, tprev + T::c(S - 1) * dt, F[S - 1]);
        rk_error<SOLVER>(p, dt, F, ex, ep);
        bool nan_cand = false, finite = true;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            nan_cand |= isnan(x1[k]) | isnan(p1[k]);
            finite &= isfinite(x1[k]) & isfinite(p1[k]);
        }
        const double err = sqrt(err_sq6(x, p, x1, p1, ex, ep, c.rtol, c.atol, nan_cand) / 6.0);
        double hn; bool bad;
        const bool keep = pid_update<T::ORDER>(err, dt, c, at_dtmin, hn, bad);
        if (MODE == 1 && trace && n_steps < trace_cap) {        // lock-step parity tests: every ATTEMPT {tprev, dt, err, keep} (mirrored time)
            double* q = trace + 4 * (size_t)n_steps;
            q[0] = tprev; q[1] = dt; q[2] = err; q[3] = keep ? 1.0 : 0.0;
        }
        n_steps++;
        if (bad) { status = 2; n_rej++; continue; }
        if (keep) {
            n_acc++;
            if (!finite) { status = 2; continue; }
            if (MODE == 1) {
                if (n_acc <= rec_cap) {
                    double* r = rec + (size_t)(n_acc - 1) * SSB_REC_STRIDE;
                    r[0] = tprev; r[1] = tnext;
#pragma unroll
                    for (int k = 0; k < 3; ++k) { r[2 + k] = x[k]; r[5 + k] = p[k]; r[8 + k] = x1[k]; r[11 + k] = p1[k]; }
#pragma unroll
                    for (int l = 0; l < S; ++l)
#pragma unroll
                        for (int k = 0; k < 3; ++k) r[14 + 3 * l + k] = F[l][k];
                }
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) { x[k] = x1[k]; p[k] = p1[k]; F[0][k] = F[S - 1][k]; }    // FSAL
            tprev = tnext;
        } else {
            n_rej++;
        }
        tprev = fmin(tprev, T1);
        double tn = tprev + hn;
        if (tn > T1 - 1e-10) tn = keep ? T1 : tprev + 0.5 * (T1 - tprev);     // diffrax _clip_to_end (f64)
        tnext = tn;
    }
    if ((MODE == 2 || (MODE == 1 && ys != nullptr)) && valid) {              // final state (or +inf if the end was not reached: diffrax leaves unsaved rows at inf)
        const bool done = status == 0 && tprev >= T1 && T0 < T1;
        const double inf = __longlong_as_double(0x7ff0000000000000LL);
#pragma unroll
        for (int k = 0; k < 3; ++k) { ys[k] = done ? x[k] : inf; ys[3 + k] = done ? dir * p[k] : inf; }
    }
}

template <int SOLVER, int MODE, int SIG, int XS>
__global__ void __launch_bounds__(SSB_ORBIT_THREADS, MODE == 0 ? SSB_SNAP_MIN_BLOCKS : SSB_ORBIT_MIN_BLOCKS) orbit_kernel(const __grid_constant__ ssb_potential Pin, const OrbitArgs a) {
    __shared__ ssb_potential sP;
    stage_potential(&sP, &Pin);
    logtab_init();
    constexpr int NX = XS == 3 ? 1 : (XS == 4 ? 0 : XS);   // XS: 1, 2 = that many fast extras on linear tracks; 3 = one on a cubic track; 4 = no extras at all
    __shared__ FastX sfx[NX > 0 ? NX : 1];
    if (NX > 0) {
        if ((int)threadIdx.x < NX) fastx_fill(&sfx[threadIdx.x], sP, SigInfo<SIG>::NF + threadIdx.x);
        __syncthreads();
    }
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < a.N;
    const int64_t ii = valid ? i : 0;
    int status, n_steps, n_acc, n_rej;
    const double* tsp = a.ts + (a.ts_per_orbit ? ii * a.M : 0);
    if (MODE == 2) {
        // Final-state mode (M == 1): the warp's 32 x 6 doubles and 32 x 3 step counters are contiguous in the outputs.  They are staged
        // through shared memory and written as whole 256-byte / 128-byte runs, so that the same kernel can write straight into pinned
        // HOST memory (ssb_gen_stream_host's zero-copy outputs: full PCIe write transactions, overlapped with the running orbits).
        __shared__ double s_fin[SSB_ORBIT_THREADS * 6];
        __shared__ int32_t s_cnt[SSB_ORBIT_THREADS * 3];
        double fin[6] = {0, 0, 0, 0, 0, 0};
        integrate_one<SOLVER, MODE, SIG, XS>(&sP, &Pin, a.w0 + ii * 6, a.t0[ii], a.t1[ii], tsp, a.M, fin, a.c, valid,
                                             status, n_steps, n_acc, n_rej, nullptr, 0, sfx);
        const int lane = threadIdx.x & 31, wbase = threadIdx.x & ~31;
        double* sf = s_fin + wbase * 6;
        int32_t* sc = s_cnt + wbase * 3;
#pragma unroll
        for (int k = 0; k < 6; ++k) sf[lane * 6 + k] = fin[k];
        sc[lane * 3] = n_steps; sc[lane * 3 + 1] = n_acc; sc[lane * 3 + 2] = n_rej;
        __syncwarp();
        const int64_t w_first = (int64_t)blockIdx.x * blockDim.x + wbase;          // first orbit of this warp
        const int64_t n_here = a.N - w_first < 32 ? a.N - w_first : 32;            // valid orbits of this warp (<= 0: none)
        double* yo = a.ys + (size_t)w_first * 6;
        int32_t* no = a.nsteps + (size_t)w_first * 3;
#pragma unroll
        for (int j = 0; j < 6; ++j) { const int q = j * 32 + lane; if (q < n_here * 6) yo[q] = sf[q]; }
#pragma unroll
        for (int j = 0; j < 3; ++j) { const int q = j * 32 + lane; if (q < n_here * 3) no[q] = sc[q]; }
        if (valid) a.status[i] = status;
        return;
    }
    extern __shared__ double s_steprec[];          // MODE 0: one published step record per thread (coop_dense), (14 + 3 S) x SSB_ORBIT_THREADS doubles, + coefficient tables
    if (MODE == 0) dense_coef_init<SOLVER>(s_steprec + (14 + 3 * Tab<SOLVER>::S) * SSB_ORBIT_THREADS);
    integrate_one<SOLVER, MODE, SIG, XS>(&sP, &Pin, a.w0 + ii * 6, a.t0[ii], a.t1[ii], tsp, a.M, a.ys + (size_t)ii * a.M * 6, a.c, valid,
                                     status, n_steps, n_acc, n_rej, nullptr, 0, sfx, HUGE_VAL, nullptr, 0, s_steprec);
    if (valid) {
        a.status[i] = status;
        a.nsteps[3 * i] = n_steps; a.nsteps[3 * i + 1] = n_acc; a.nsteps[3 * i + 2] = n_rej;
    }
}

// lock-step parity diagnostics: every step ATTEMPT of every orbit {tprev, dt, err, keep} (mirrored time) + the final state
template <int SOLVER, int SIG>
__global__ void __launch_bounds__(SSB_ORBIT_THREADS, SSB_ORBIT_MIN_BLOCKS) trace_kernel(const __grid_constant__ ssb_potential Pin, int64_t N, const double* w0,
                                                                                        const double* t0, const double* t1, CtrlDev c, double* trace,
                                                                                        int trace_cap, double* yfin, int32_t* status_out, int32_t* nsteps_out) {
    __shared__ ssb_potential sP;
    stage_potential(&sP, &Pin);
    logtab_init();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < N;
    const int64_t ii = valid ? i : 0;
    int status, n_steps, n_acc, n_rej;
    double fin[6] = {0, 0, 0, 0, 0, 0};
    integrate_one<SOLVER, 1, SIG>(&sP, &Pin, w0 + 6 * ii, t0[ii], t1[ii], nullptr, 0, fin, c, valid, status, n_steps, n_acc, n_rej, nullptr, 0, nullptr, HUGE_VAL,
                                  trace + (size_t)ii * 4 * trace_cap, trace_cap);
    if (valid) {
        for (int k = 0; k < 6; ++k) yfin[6 * i + k] = fin[k];
        status_out[i] = status;
        nsteps_out[3 * i] = n_steps; nsteps_out[3 * i + 1] = n_acc; nsteps_out[3 * i + 2] = n_rej;
    }
}

// =============================================================================================
// K0: one orbit with dense output at M save times (progenitor at all stripping times, main.py:289)
//   scratch layout: hdr[8] doubles {n_acc, status, n_steps, n_rej, dir}, then rec[max_steps][SSB_REC_STRIDE]
// =============================================================================================
template <int SOLVER, int SIG, int XS = 0>
__global__ void __launch_bounds__(32) dense_step_kernel(const __grid_constant__ ssb_potential Pin, const double* w0, double t0, double t1,
                                                        const double* t0p, const double* t1p, CtrlDev c, double* scratch, int rec_cap,
                                                        int32_t* status_out, int32_t* nsteps_out, const double* tstop_p = nullptr) {
    __shared__ ssb_potential sP;
    stage_potential(&sP, &Pin);
    logtab_init();
    if (t0p) { t0 = *t0p; t1 = *t1p; }              // interval ends read on the device (no host round trip in gen_stream)
    const bool valid = threadIdx.x == 0;
    int status, n_steps, n_acc, n_rej;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    const double t_stop = tstop_p ? *tstop_p * ((t0 < t1) ? 1.0 : -1.0) : inf;      // mirrored time, as integrate_one runs
    integrate_one<SOLVER, 1, SIG, XS>(&sP, &Pin, w0, t0, t1, nullptr, 0, nullptr, c, valid, status, n_steps, n_acc, n_rej, scratch + 8, rec_cap, nullptr, t_stop);
    if (valid) {
        scratch[0] = (double)min(n_acc, rec_cap); scratch[1] = (double)status; scratch[4] = (t0 < t1) ? 1.0 : -1.0;
        if (status_out) *status_out = status;
        if (nsteps_out) { nsteps_out[0] = n_steps; nsteps_out[1] = n_acc; nsteps_out[2] = n_rej; }
    }
}

// ts_stride > 1: row m of ys is the state at ts[ts_begin + m * ts_stride] (a rank's own stripping times of a sharded stream)
template <int SOLVER>
__global__ void dense_eval_kernel(const double* scratch, const double* ts, int64_t M, double* ys, int64_t ts_begin = 0, int64_t ts_stride = 1) {
    constexpr int S = Tab<SOLVER>::S;
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    ts += ts_begin + m * (ts_stride - 1);               // ts[m] below reads ts[ts_begin + m * ts_stride]
    const int n = (int)scratch[0];
    const double dir = scratch[4];
    const double* rec = scratch + 8;
    const double tq = ts[m] * dir;
    double* o = ys + m * 6;
    // first accepted step whose end time >= tq (the step inside which diffrax saves this ts)
    int lo = 0, hi = n;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (rec[(size_t)mid * SSB_REC_STRIDE + 1] < tq) lo = mid + 1; else hi = mid; }
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    if (lo >= n || n == 0 || tq < rec[0]) {
#pragma unroll
        for (int k = 0; k < 6; ++k) o[k] = inf;
        return;
    }
    const double* r = rec + (size_t)lo * SSB_REC_STRIDE;
    const double ta = r[0], tb = r[1];
    double x[3], p[3], x1[3], p1[3], F[S][3], xo[3], po[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { x[k] = r[2 + k]; p[k] = r[5 + k]; x1[k] = r[8 + k]; p1[k] = r[11 + k]; }
#pragma unroll
    for (int l = 0; l < S; ++l)
#pragma unroll
        for (int k = 0; k < 3; ++k) F[l][k] = r[14 + 3 * l + k];
    if (tq == tb) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { xo[k] = x1[k]; po[k] = p1[k]; }
    } else {
        rk_dense<SOLVER>(x, p, x1, p1, tb - ta, F, (tq - ta) / (tb - ta), xo, po);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) { o[k] = xo[k]; o[3 + k] = dir * po[k]; }
}

// ---- batched dense solutions (gen_stream_*_dense, main.py:376-430; integrate_orbit(dense=True) under vmap) ----
// record layout per orbit: hdr[8] {n_recorded, status, -, -, dir} + rec[cap][SSB_REC_STRIDE], i.e. the K0 layout repeated N times
template <int SOLVER, int SIG>
__global__ void __launch_bounds__(SSB_ORBIT_THREADS, SSB_ORBIT_MIN_BLOCKS) record_kernel(const __grid_constant__ ssb_potential Pin, int64_t N, const double* w0,
                                                                                         const double* t0, const double* t1, CtrlDev c, double* recs,
                                                                                         int rec_cap, int32_t* status_out, int32_t* nsteps_out) {
    __shared__ ssb_potential sP;
    stage_potential(&sP, &Pin);
    logtab_init();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < N;
    const int64_t ii = valid ? i : 0;
    double* scratch = recs + (size_t)ii * (8 + (size_t)rec_cap * SSB_REC_STRIDE);
    int status, n_steps, n_acc, n_rej;
    integrate_one<SOLVER, 1, SIG>(&sP, &Pin, w0 + 6 * ii, t0[ii], t1[ii], nullptr, 0, nullptr, c, valid, status, n_steps, n_acc, n_rej, scratch + 8, rec_cap);
    if (valid) {
        if (status == 0 && n_acc > rec_cap) status = 1;             // more accepted steps than record slots: treated like max_steps
        scratch[0] = (double)min(n_acc, rec_cap); scratch[1] = (double)status; scratch[4] = (t0[ii] < t1[ii]) ? 1.0 : -1.0;
        status_out[i] = status;
        nsteps_out[3 * i] = n_steps; nsteps_out[3 * i + 1] = n_acc; nsteps_out[3 * i + 2] = n_rej;
    }
}
// ys[i] = orbit i evaluated at tq[i] (per_orbit) or tq[0]; +inf outside the recorded interval
template <int SOLVER>
__global__ void record_eval_kernel(const double* recs, int rec_cap, int64_t N, const double* tq, int per_orbit, double* ys) {
    constexpr int S = Tab<SOLVER>::S;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double* scratch = recs + (size_t)i * (8 + (size_t)rec_cap * SSB_REC_STRIDE);
    const int n = (int)scratch[0];
    const double dir = scratch[4];
    const double* rec = scratch + 8;
    const double t = tq[per_orbit ? i : 0] * dir;
    double* o = ys + 6 * i;
    int lo = 0, hi = n;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (rec[(size_t)mid * SSB_REC_STRIDE + 1] < t) lo = mid + 1; else hi = mid; }
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    if (lo >= n || n == 0 || t < rec[0]) {
        for (int k = 0; k < 6; ++k) o[k] = inf;
        return;
    }
    const double* r = rec + (size_t)lo * SSB_REC_STRIDE;
    const double ta = r[0], tb = r[1];
    double x[3], p[3], x1[3], p1[3], F[S][3], xo[3], po[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { x[k] = r[2 + k]; p[k] = r[5 + k]; x1[k] = r[8 + k]; p1[k] = r[11 + k]; }
#pragma unroll
    for (int l = 0; l < S; ++l)
#pragma unroll
        for (int k = 0; k < 3; ++k) F[l][k] = r[14 + 3 * l + k];
    if (t == tb) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { xo[k] = x1[k]; po[k] = p1[k]; }
    } else {
        rk_dense<SOLVER>(x, p, x1, p1, tb - ta, F, (t - ta) / (tb - ta), xo, po);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) { o[k] = xo[k]; o[3 + k] = dir * po[k]; }
}

// =============================================================================================
// field evaluation kernels
// =============================================================================================
__global__ void potential_eval_kernel(const __grid_constant__ ssb_potential Pin, int64_t n, const double* xyz, const double* t,
                                      double* phi, double* grad, double* hess) {
    __shared__ ssb_potential sP;
    stage_potential(&sP, &Pin);
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double X[3] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    double P, g[3];
    Sym3 H;
    pot_eval<WANT_PHI | WANT_GRAD | WANT_HESS>(sP, X, t[i], P, g, H);
    if (phi) phi[i] = P;
    if (grad) { grad[3 * i] = g[0]; grad[3 * i + 1] = g[1]; grad[3 * i + 2] = g[2]; }
    if (hess) {
        double* h = hess + 9 * i;
        h[0] = H.xx; h[1] = H.xy; h[2] = H.xz; h[3] = H.xy; h[4] = H.yy; h[5] = H.yz; h[6] = H.xz; h[7] = H.yz; h[8] = H.zz;
    }
}

__global__ void subhalo_eval_kernel(const ssb_subhalos S, int dradius, double x0, double x1, double x2, double t, double* phi, double* grad) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= S.n) return;
    double ph = 0.0, g[3] = {0, 0, 0};
    const double dt = t - S.t0[j];
    if (fabs(dt) < S.tw[j]) {
        const double X[3] = {x0, x1, x2};
        double rel[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) rel[k] = X[k] - fma(S.v[3 * j + k], dt, S.x0[3 * j + k]);
        const double r2 = fma(rel[0], rel[0], fma(rel[1], rel[1], rel[2] * rel[2]));
        double q, w = 0;
        if (dradius) profile_dradius(S.profile, S.G * S.m[j], S.rs[j], r2, ph, q);
        else profile_terms<WANT_PHI | WANT_GRAD>(S.profile, S.G * S.m[j], S.rs[j], r2, ph, q, w);
#pragma unroll
        for (int k = 0; k < 3; ++k) g[k] = q * rel[k];
    }
    if (phi) phi[j] = ph;
    if (grad) { grad[3 * j] = g[0]; grad[3 * j + 1] = g[1]; grad[3 * j + 2] = g[2]; }
}

__global__ void track_slopes_kernel(int64_t n, const double* t, const double* y, double* s) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    auto sec = [&](int64_t j, int k) {           // secant slope of segment j (interpax approx_df 'cubic')
        const double dx = t[j + 1] - t[j];
        const double dxi = dx == 0.0 ? 0.0 : 1.0 / dx;
        return dxi * (y[3 * (j + 1) + k] - y[3 * j + k]);
    };
    for (int k = 0; k < 3; ++k) {
        double v;
        if (i == 0) v = sec(0, k);
        else if (i == n - 1) v = sec(n - 2, k);
        else v = 0.5 * (sec(i - 1, k) + sec(i, k));
        s[3 * i + k] = v;
    }
}

__global__ void track_eval_kernel(const ssb_track T, int64_t nq, const double* tq, double* out, double* dout) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    double c[3], dc[3];
    track_eval<true>(T, tq[i], c, dc);
    for (int k = 0; k < 3; ++k) { out[3 * i + k] = c[k]; if (dout) dout[3 * i + k] = dc[k]; }
}

// =============================================================================================
// K2: release model (main.py:209-280) with jax.random threefry2x32 normals (main.py:220-228, 263-266)
// =============================================================================================
__device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
__device__ void threefry2x32(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t& o0, uint32_t& o1) {
    const uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
    const int R[2][4] = {{13, 15, 26, 6}, {17, 29, 16, 24}};
    uint32_t x0 = c0 + ks[0], x1 = c1 + ks[1];
#pragma unroll
    for (int g = 0; g < 5; ++g) {
#pragma unroll
        for (int r = 0; r < 4; ++r) { x0 += x1; x1 = rotl32(x1, R[g & 1][r]); x1 ^= x0; }
        x0 += ks[(g + 1) % 3];
        x1 += ks[(g + 2) % 3] + (uint32_t)(g + 1);
    }
    o0 = x0; o1 = x1;
}
// jax.random.normal(PRNGKey(seed), (1,), float64): bits64 = threefry(key, (0,1)); mantissa trick; sqrt(2) erfinv(u)
__device__ double jax_normal1(int64_t seed) {
    const uint64_t u64 = (uint64_t)seed;
    uint32_t o0, o1;
    threefry2x32((uint32_t)(u64 >> 32), (uint32_t)(u64 & 0xFFFFFFFFu), 0u, 1u, o0, o1);
    const uint64_t bits = ((uint64_t)o0 << 32) | (uint64_t)o1;
    const double f = __longlong_as_double((long long)((bits >> 12) | 0x3FF0000000000000ull)) - 1.0;
    const double lo = -0.99999999999999989;          // nextafter(-1, 0)
    const double u = fmax(lo, fma(f, 1.0 - lo, lo));
    return 1.4142135623730951 * erfinv(u);
}

struct ReleaseArgs {
    int64_t N;
    const double *prog, *Msat, *t, *normals;
    const int64_t* idx;
    int64_t r[4];                 // jax.random.randint(PRNGKey(seed), (5,), 0, 1000)[:4], computed on the host
    double kv[8];
    double G;
    double *pos_lead, *pos_trail, *vel_lead, *vel_trail;
    double *w0_packed, *t0_packed;   // optional [2,N,6] / [2,N] in the orbit-kernel layout (lead block, trail block)
    // sharded streams: thread k handles stripping time g = sel_begin + k * sel_stride; t / Msat / normals / the draw index are read at g,
    // prog and every output at k (sel_stride = 0: g = k).  t1_packed[2,N] (optional) is filled with *t_end.
    int64_t sel_begin, sel_stride;
    double* t1_packed; const double* t_end;
    int64_t i0, cnt;              // cnt > 0: this launch handles the compact indices [i0, i0 + cnt) of the N (gen_stream's two-part pipeline)
    // compact inputs (ssb_gen_stream_host on a shard): t / Msat / normals hold ONLY this shard's stripping times, row k <-> stripping
    // index draw_begin + k * draw_stride (the index that seeds jax.random, main.py:223-228).  draw_stride = 0: the array index itself.
    int64_t draw_begin, draw_stride;
};

// ---- forward-mode dual numbers for jacfwd(release_model) (perturbative.py:281-296): value + 6 partials d/d(x, v) ----
struct D6 {
    double v, d[6];
    __device__ D6() : v(0.0) { for (int i = 0; i < 6; ++i) d[i] = 0.0; }
    __device__ D6(double c) : v(c) { for (int i = 0; i < 6; ++i) d[i] = 0.0; }
};
__device__ inline D6 operator+(const D6& a, const D6& b) { D6 r; r.v = a.v + b.v; for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
__device__ inline D6 operator-(const D6& a, const D6& b) { D6 r; r.v = a.v - b.v; for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
__device__ inline D6 operator-(const D6& a) { D6 r; r.v = -a.v; for (int i = 0; i < 6; ++i) r.d[i] = -a.d[i]; return r; }
__device__ inline D6 operator*(const D6& a, const D6& b) { D6 r; r.v = a.v * b.v; for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
__device__ inline D6 operator/(const D6& a, const D6& b) {
    D6 r; const double inv = 1.0 / b.v; r.v = a.v * inv;
    for (int i = 0; i < 6; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
    return r;
}
__device__ inline D6 dsqrt(const D6& a) { D6 r; r.v = sqrt(a.v); const double f = 0.5 / r.v; for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] * f; return r; }
__device__ inline double dsqrt(double a) { return sqrt(a); }
__device__ inline D6 dcbrt(const D6& a) { D6 r; r.v = pow(a.v, 1.0 / 3.0); const double f = r.v / (3.0 * a.v); for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] * f; return r; }
__device__ inline double dcbrt(double a) { return pow(a, 1.0 / 3.0); }

// release_model (main.py:230-278) on scalar type T; Hrr = rhat^T Hess(Phi) rhat supplied by the caller.
// out[12] = pos_lead(3), pos_trail(3), vel_lead(3), vel_trail(3)
template <class T>
__device__ inline void release_math(const T x[3], const T v[3], const T H[3][3], double GMsat, const double kv[8], const double nr[4], T out[12]) {
    const T rad2 = x[0] * x[0] + x[1] * x[1] + x[2] * x[2];
    const T L[3] = {x[1] * v[2] - x[2] * v[1], x[2] * v[0] - x[0] * v[2], x[0] * v[1] - x[1] * v[0]};
    const T Lmag = dsqrt(L[0] * L[0] + L[1] * L[1] + L[2] * L[2]);
    const T omega = Lmag / rad2;                                                 // main.py:88-96
    const T r = dsqrt(rad2);
    const T rhat[3] = {x[0] / r, x[1] / r, x[2] / r};
    T d2 = T(0.0);                                                               // main.py:76-85 (rhat held fixed inside the derivative)
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) d2 = d2 + rhat[i] * H[i][j] * rhat[j];
    const T rt = dcbrt(T(GMsat) / (omega * omega - d2));                         // main.py:104
    const T vcirc = omega * rt;                                                  // main.py:238,242
    const T zhat[3] = {L[0] / Lmag, L[1] / Lmag, L[2] / Lmag};
    const T vr = v[0] * rhat[0] + v[1] * rhat[1] + v[2] * rhat[2];
    const T pv[3] = {v[0] - vr * rhat[0], v[1] - vr * rhat[1], v[2] - vr * rhat[2]};
    const T pn = dsqrt(pv[0] * pv[0] + pv[1] * pv[1] + pv[2] * pv[2]);
    const T phat[3] = {pv[0] / pn, pv[1] / pn, pv[2] / pn};
    const double kr = kv[0] + nr[0] * kv[4];                                     // main.py:263-266
    const double kvphi = kr * (kv[1] + nr[1] * kv[5]);
    const double kz = kv[2] + nr[2] * kv[6];
    const double kvz = kv[3] + nr[3] * kv[7];
    for (int k = 0; k < 3; ++k) {
        T pt = x[k] + T(kr) * rhat[k] * rt;                                      // main.py:269-272 (trailing)
        pt = pt + zhat[k] * T(kz) * rt;
        T vt = v[k] + T(kvphi) * vcirc * phat[k];
        vt = vt + T(kvz) * vcirc * zhat[k];
        T pl = x[k] - T(kr) * rhat[k] * rt;                                      // main.py:275-278 (leading)
        pl = pl - zhat[k] * T(kz) * rt;
        T vl = v[k] - T(kvphi) * vcirc * phat[k];
        vl = vl - T(kvz) * vcirc * zhat[k];
        out[k] = pl; out[3 + k] = pt; out[6 + k] = vl; out[9 + k] = vt;
    }
}

__device__ inline void release_draws(const ReleaseArgs& a, int64_t i, double nr[4]) {
    if (a.normals) { for (int q = 0; q < 4; ++q) nr[q] = a.normals[4 * i + q]; }
    else { const int64_t id = a.idx ? a.idx[i] : (a.draw_stride ? a.draw_begin + i * a.draw_stride : i); for (int q = 0; q < 4; ++q) nr[q] = jax_normal1(id * a.r[q]); }
}

__global__ void release_kernel(const __grid_constant__ ssb_potential Pin, const ReleaseArgs a) {
    __shared__ ssb_potential sP;
    stage_potential(&sP, &Pin);
    const int64_t k0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a.cnt > 0 && k0 >= a.cnt) return;
    const int64_t i = a.i0 + k0;                                             // compact index: prog row, output row
    if (i >= a.N) return;
    const int64_t gi = a.sel_stride ? a.sel_begin + i * a.sel_stride : i;   // stripping-time index
    const double* w = a.prog + 6 * i;
    const double x[3] = {w[0], w[1], w[2]}, v[3] = {w[3], w[4], w[5]};
    const double t = a.t[gi];
    double nr[4];
    release_draws(a, gi, nr);
    double P, g[3];
    Sym3 Hs;
    pot_eval<WANT_HESS>(sP, x, t, P, g, Hs);
    const double H[3][3] = {{Hs.xx, Hs.xy, Hs.xz}, {Hs.xy, Hs.yy, Hs.yz}, {Hs.xz, Hs.yz, Hs.zz}};
    double out[12];
    release_math<double>(x, v, H, a.G * a.Msat[gi], a.kv, nr, out);
    for (int k = 0; k < 3; ++k) {
        if (a.pos_lead) { a.pos_lead[3 * i + k] = out[k]; a.pos_trail[3 * i + k] = out[3 + k]; a.vel_lead[3 * i + k] = out[6 + k]; a.vel_trail[3 * i + k] = out[9 + k]; }
        if (a.w0_packed) {
            a.w0_packed[6 * i + k] = out[k]; a.w0_packed[6 * i + 3 + k] = out[6 + k];
            a.w0_packed[6 * (a.N + i) + k] = out[3 + k]; a.w0_packed[6 * (a.N + i) + 3 + k] = out[9 + k];
        }
    }
    if (a.t0_packed) { a.t0_packed[i] = t; a.t0_packed[a.N + i] = t; }
    if (a.t1_packed) { const double te = *a.t_end; a.t1_packed[i] = te; a.t1_packed[a.N + i] = te; }
}

// ---- Chen+25 release (streamhelpers.py:352-432): 6 correlated normals per stripping time -> (Dr, phi, theta, Dv, alpha, beta) ----
struct Chen25Args {
    int64_t N;
    const double *prog, *Msat, *t, *normals;     // normals: optional [N,6] standard normals
    uint32_t key[2];                             // jax PRNG key; per-release keys = jax.random.split(key, N) (streamhelpers.py:455)
    double mean[6], factor[36];                  // multivariate_normal(mean, cov, method='svd'): sample = mean + factor @ z
    double G;
    double *pos_lead, *pos_trail, *vel_lead, *vel_trail;
};
__global__ void release_chen25_kernel(const __grid_constant__ ssb_potential Pin, const Chen25Args a) {
    __shared__ ssb_potential sP;
    stage_potential(&sP, &Pin);
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.N) return;
    const double* w = a.prog + 6 * i;
    const double x[3] = {w[0], w[1], w[2]}, v[3] = {w[3], w[4], w[5]};
    double z[6];
    if (a.normals) { for (int q = 0; q < 6; ++q) z[q] = a.normals[6 * i + q]; }
    else {
        // key_i = split(key, N)[i]: threefry(key, iota(2N)) reshaped (N, 2)
        uint32_t ki[2];
        for (int h = 0; h < 2; ++h) {
            const int64_t f = 2 * i + h;                       // flat index into concat(o0[0..N-1], o1[0..N-1])
            const int64_t j = f < a.N ? f : f - a.N;
            uint32_t o0, o1;
            threefry2x32(a.key[0], a.key[1], (uint32_t)j, (uint32_t)(a.N + j), o0, o1);
            ki[h] = f < a.N ? o0 : o1;
        }
        for (int q = 0; q < 6; ++q) {                          // jax.random.normal(key_i, (6,)) float64
            uint32_t o0, o1;
            threefry2x32(ki[0], ki[1], (uint32_t)q, (uint32_t)(6 + q), o0, o1);
            const uint64_t bits = ((uint64_t)o0 << 32) | (uint64_t)o1;
            const double f = __longlong_as_double((long long)((bits >> 12) | 0x3FF0000000000000ull)) - 1.0;
            const double lo = -0.99999999999999989;
            z[q] = 1.4142135623730951 * erfinv(fmax(lo, fma(f, 1.0 - lo, lo)));
        }
    }
    double pv6[6];
    for (int r = 0; r < 6; ++r) { double acc = a.mean[r]; for (int c = 0; c < 6; ++c) acc += a.factor[6 * r + c] * z[c]; pv6[r] = acc; }
    // tidal radius (main.py:98-104)
    const double rad2 = x[0] * x[0] + x[1] * x[1] + x[2] * x[2];
    const double L[3] = {x[1] * v[2] - x[2] * v[1], x[2] * v[0] - x[0] * v[2], x[0] * v[1] - x[1] * v[0]};
    const double Lmag = sqrt(L[0] * L[0] + L[1] * L[1] + L[2] * L[2]);
    const double omega = Lmag / rad2, r = sqrt(rad2);
    const double xh[3] = {x[0] / r, x[1] / r, x[2] / r};
    double P, g[3];
    Sym3 H;
    pot_eval<WANT_HESS>(sP, x, a.t[i], P, g, H);
    const double d2 = xh[0] * (H.xx * xh[0] + H.xy * xh[1] + H.xz * xh[2]) + xh[1] * (H.xy * xh[0] + H.yy * xh[1] + H.yz * xh[2]) +
                      xh[2] * (H.xz * xh[0] + H.yz * xh[1] + H.zz * xh[2]);
    const double GM = a.G * a.Msat[i];
    const double rt = pow(GM / (omega * omega - d2), 1.0 / 3.0);
    const double zh[3] = {L[0] / Lmag, L[1] / Lmag, L[2] / Lmag};
    const double vr = v[0] * xh[0] + v[1] * xh[1] + v[2] * xh[2];
    const double pvv[3] = {v[0] - vr * xh[0], v[1] - vr * xh[1], v[2] - vr * xh[2]};
    const double pn = sqrt(pvv[0] * pvv[0] + pvv[1] * pvv[1] + pvv[2] * pvv[2]);
    const double yh[3] = {pvv[0] / pn, pvv[1] / pn, pvv[2] / pn};
    const double Dr = pv6[0] * rt;                                   // streamhelpers.py:389-398
    const double Dv = pv6[3] * sqrt(2.0 * GM / Dr);
    const double d2r = 0.017453292519943295;
    double sphi, cphi, sth, cth, sal, cal, sbe, cbe;
    sincos(pv6[1] * d2r, &sphi, &cphi); sincos(pv6[2] * d2r, &sth, &cth); sincos(pv6[4] * d2r, &sal, &cal); sincos(pv6[5] * d2r, &sbe, &cbe);
    for (int k = 0; k < 3; ++k) {
        a.pos_trail[3 * i + k] = x[k] + (Dr * cth * cphi) * xh[k] + (Dr * cth * sphi) * yh[k] + (Dr * sth) * zh[k];     // streamhelpers.py:405-416
        a.vel_trail[3 * i + k] = v[k] + (Dv * cbe * cal) * xh[k] + (Dv * cbe * sal) * yh[k] + (Dv * sbe) * zh[k];
        a.pos_lead[3 * i + k] = x[k] - (Dr * cth * cphi) * xh[k] - (Dr * cth * sphi) * yh[k] + (Dr * sth) * zh[k];      // streamhelpers.py:419-430
        a.vel_lead[3 * i + k] = v[k] - (Dv * cbe * cal) * xh[k] - (Dv * cbe * sal) * yh[k] + (Dv * sbe) * zh[k];
    }
}

// jacfwd(release_func) (perturbative.py:281-296): jac[N,2,6,6], rows = (pos, vel) of lead / trail, columns = d/d(x, v)
__global__ void release_jacobian_kernel(const __grid_constant__ ssb_potential Pin, const ReleaseArgs a, double* jac) {
    __shared__ ssb_potential sP;
    stage_potential(&sP, &Pin);
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.N) return;
    const double* w = a.prog + 6 * i;
    const double xs[3] = {w[0], w[1], w[2]};
    const double t = a.t[i];
    double nr[4];
    release_draws(a, i, nr);
    double P, g[3];
    Sym3 Hs;
    Sym3x3 T3;
    pot_eval<WANT_HESS>(sP, xs, t, P, g, Hs);
    pot_third(sP, xs, t, T3);
    const double Hv[3][3] = {{Hs.xx, Hs.xy, Hs.xz}, {Hs.xy, Hs.yy, Hs.yz}, {Hs.xz, Hs.yz, Hs.zz}};
    D6 x[3], v[3], H[3][3], out[12];
    for (int k = 0; k < 3; ++k) { x[k] = D6(w[k]); x[k].d[k] = 1.0; v[k] = D6(w[3 + k]); v[k].d[3 + k] = 1.0; }
    for (int p = 0; p < 3; ++p) for (int q = 0; q < 3; ++q) {
        H[p][q] = D6(Hv[p][q]);
        for (int k = 0; k < 3; ++k) H[p][q].d[k] = third_at(T3, p, q, k);
    }
    release_math<D6>(x, v, H, a.G * a.Msat[i], a.kv, nr, out);
    double* J = jac + (size_t)i * 72;
    for (int c = 0; c < 3; ++c) for (int q = 0; q < 6; ++q) {
        J[0 * 36 + c * 6 + q] = out[c].d[q];             // lead position rows
        J[0 * 36 + (3 + c) * 6 + q] = out[6 + c].d[q];   // lead velocity rows
        J[1 * 36 + c * 6 + q] = out[3 + c].d[q];         // trail position rows
        J[1 * 36 + (3 + c) * 6 + q] = out[9 + c].d[q];   // trail velocity rows
    }
}

__global__ void potential_third_kernel(const __grid_constant__ ssb_potential Pin, int64_t n, const double* xyz, const double* t, double* third) {
    __shared__ ssb_potential sP;
    stage_potential(&sP, &Pin);
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double X[3] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    Sym3x3 T3;
    pot_third(sP, X, t[i], T3);
    for (int p = 0; p < 3; ++p) for (int q = 0; q < 3; ++q) for (int k = 0; k < 3; ++k) third[27 * i + 9 * p + 3 * q + k] = third_at(T3, p, q, k);
}


// fp64 peak probe: 8 independent DFMA chains per thread
__global__ void __launch_bounds__(256) dfma_probe_kernel(int iters, double* sink) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000000001, b = 1e-12;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, m, b); a1 = fma(a1, m, b); a2 = fma(a2, m, b); a3 = fma(a3, m, b);
            a4 = fma(a4, m, b); a5 = fma(a5, m, b); a6 = fma(a6, m, b); a7 = fma(a7, m, b);
        }
    }
    const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456) sink[0] = s;
}

// =============================================================================================
// host side: C ABI
// =============================================================================================
static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};
void ssb_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
int ssb_set_error(int code, const char* msg) { snprintf(g_err, sizeof(g_err), "%s", msg); return code; }
int ssb_cuda_check(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return SSB_ERR_CUDA;
}
#define CK(call) do { int _e = ssb_cuda_check((call), #call); if (_e) return _e; } while (0)
#define CKL(what) do { ssb_count_launch(); int _e = ssb_cuda_check(cudaGetLastError(), what); if (_e) return _e; } while (0)

int ssb_validate_potential(const ssb_potential* p) {
    if (!p) return ssb_set_error(SSB_ERR_ARG, "potential is NULL");
    if (p->n_comp < 0 || p->n_comp > SSB_MAX_COMP || p->n_track < 0 || p->n_track > SSB_MAX_TRACK || p->n_sh < 0 ||
        p->n_sh > SSB_MAX_SUBHALO_SETS || p->n_pset < 0 || p->n_pset > SSB_MAX_PSETS)
        return ssb_set_error(SSB_ERR_ARG, "potential: component/track/subhalo-set/perturber-set count out of range");
    for (int i = 0; i < p->n_pset; ++i) {
        const ssb_perturbers& s = p->pset[i];
        if (s.n < 0 || s.n_knots < 2 || (s.n > 0 && (!s.t || !s.y || !s.GM || !s.rs))) return ssb_set_error(SSB_ERR_ARG, "perturber set: NULL array or < 2 knots");
        if (s.profile < SSB_PROFILE_PLUMMER || s.profile > SSB_PROFILE_NFW) return ssb_set_error(SSB_ERR_UNSUPPORTED, "perturber set: unknown profile");
    }
    for (int i = 0; i < p->n_comp; ++i) {
        const ssb_component& c = p->comp[i];
        if (c.type < SSB_NFW || c.type > SSB_DEHNEN_BAR) return ssb_set_error(SSB_ERR_UNSUPPORTED, "potential: unknown component type");
        if (c.type == SSB_PERTURBERS && (c.sh < 0 || c.sh >= p->n_pset || c.track >= 0 || c.growth != 0))
            return ssb_set_error(SSB_ERR_ARG, "perturber-set component: missing set, or combined with a translation / growth factor");
        if (c.track >= p->n_track) return ssb_set_error(SSB_ERR_ARG, "potential: component references a missing track");
        if (c.growth < 0 || c.growth > p->n_track) return ssb_set_error(SSB_ERR_ARG, "potential: component references a missing growth track");
        if (c.growth > 0 && (c.type == SSB_UNIFORM_ACC || c.type == SSB_SUBHALOS)) return ssb_set_error(SSB_ERR_UNSUPPORTED, "growth factor on a force-only / subhalo component");
        if (c.type == SSB_UNIFORM_ACC && c.track < 0) return ssb_set_error(SSB_ERR_ARG, "uniform acceleration needs a velocity track");
        if (c.type == SSB_SUBHALOS && (c.sh < 0 || c.sh >= p->n_sh)) return ssb_set_error(SSB_ERR_ARG, "subhalo component references a missing set");
    }
    for (int i = 0; i < p->n_track; ++i) {
        const ssb_track& t = p->track[i];
        if (t.n < 2 || !t.t || !t.y) return ssb_set_error(SSB_ERR_ARG, "track: needs >= 2 knots and t, y pointers");
        if (t.kind == SSB_TRACK_CUBIC && !t.s) return ssb_set_error(SSB_ERR_ARG, "cubic track: slopes pointer is NULL (ssb_track_slopes_f64)");
        if (t.kind != SSB_TRACK_LINEAR && t.kind != SSB_TRACK_CUBIC) return ssb_set_error(SSB_ERR_UNSUPPORTED, "track: unknown kind");
    }
    for (int i = 0; i < p->n_sh; ++i) {
        const ssb_subhalos& s = p->sh[i];
        if (s.n < 0 || (s.n > 0 && (!s.m || !s.rs || !s.x0 || !s.v || !s.t0 || !s.tw))) return ssb_set_error(SSB_ERR_ARG, "subhalo set: NULL array");
        if (s.profile < SSB_PROFILE_PLUMMER || s.profile > SSB_PROFILE_NFW) return ssb_set_error(SSB_ERR_UNSUPPORTED, "subhalo set: unknown profile");
    }
    return 0;
}
int ssb_validate_ctrl(const ssb_ctrl& c) {
    if (c.solver != 5 && c.solver != 8) return ssb_set_error(SSB_ERR_UNSUPPORTED, "solver must be 5 (Dopri5) or 8 (Dopri8)");
    if (!(c.rtol >= 0) || !(c.atol >= 0) || !(c.dtmin >= 0) || c.max_steps < 0) return ssb_set_error(SSB_ERR_ARG, "ctrl: negative tolerance / dtmin / max_steps");
    return 0;
}
static CtrlDev to_dev(const ssb_ctrl& c) { CtrlDev d; d.rtol = c.rtol; d.atol = c.atol; d.dtmin = c.dtmin; d.dtmax = c.dtmax; d.max_steps = c.max_steps; return d; }
static inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

// jax.random.randint(PRNGKey(seed), (5,), 0, 1000) on the host (5 integers; main.py:220-221)
static inline uint32_t h_rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
static void h_threefry(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t* o0, uint32_t* o1) {
    static const int R[2][4] = {{13, 15, 26, 6}, {17, 29, 16, 24}};
    uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
    uint32_t x0 = c0 + ks[0], x1 = c1 + ks[1];
    for (int g = 0; g < 5; ++g) {
        for (int r = 0; r < 4; ++r) { x0 += x1; x1 = h_rotl32(x1, R[g & 1][r]); x1 ^= x0; }
        x0 += ks[(g + 1) % 3];
        x1 += ks[(g + 2) % 3] + (uint32_t)(g + 1);
    }
    *o0 = x0; *o1 = x1;
}
static void host_randint5(int64_t seed, int64_t* out) {
    const uint64_t u = (uint64_t)seed;
    const uint32_t ka = (uint32_t)(u >> 32), kb = (uint32_t)(u & 0xFFFFFFFFu);
    uint32_t a0, b0, a1, b1;
    h_threefry(ka, kb, 0, 2, &a0, &b0);          // jax.random.split(key): counts iota(4) -> halves (0,1),(2,3)
    h_threefry(ka, kb, 1, 3, &a1, &b1);
    const uint32_t k1[2] = {a0, a1}, k2[2] = {b0, b1};
    const uint64_t span = 1000, mult0 = (((uint64_t)1 << 32) % span), mult = (mult0 * mult0) % span;
    for (int j = 0; j < 5; ++j) {
        uint32_t h0, h1, l0, l1;
        h_threefry(k1[0], k1[1], (uint32_t)j, (uint32_t)(5 + j), &h0, &h1);
        h_threefry(k2[0], k2[1], (uint32_t)j, (uint32_t)(5 + j), &l0, &l1);
        const uint64_t hi = ((uint64_t)h0 << 32) | h1, lo = ((uint64_t)l0 << 32) | l1;
        out[j] = (int64_t)(((hi % span) * mult + (lo % span)) % span);
    }
}

extern "C" {

int ssb_abi_version(void) { return SSB_ABI_VERSION; }
unsigned long long ssb_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
const char* ssb_last_error(void) { return g_err; }

int ssb_potential_eval_f64(const ssb_potential* pot, int64_t n, const double* xyz, const double* t, double* phi, double* grad,
                           double* hess, void* stream) {
    if (int e = ssb_validate_potential(pot)) return e;
    if (n < 0 || (n > 0 && (!xyz || !t))) return ssb_set_error(SSB_ERR_ARG, "potential_eval: bad n / NULL xyz, t");
    if (n == 0) return 0;
    potential_eval_kernel<<<nblk(n, 128), 128, 0, (cudaStream_t)stream>>>(*pot, n, xyz, t, phi, grad, hess);
    CKL("potential_eval_kernel");
    return 0;
}

int ssb_subhalo_eval_f64(const ssb_subhalos* sh, int dradius, const double* xyz, double t, double* phi, double* grad, void* stream) {
    if (!sh || !xyz) return ssb_set_error(SSB_ERR_ARG, "subhalo_eval: NULL argument");
    if (sh->n <= 0) return 0;
    subhalo_eval_kernel<<<nblk(sh->n, 128), 128, 0, (cudaStream_t)stream>>>(*sh, dradius, xyz[0], xyz[1], xyz[2], t, phi, grad);
    CKL("subhalo_eval_kernel");
    return 0;
}

int ssb_track_slopes_f64(int64_t n, const double* t, const double* y, double* s, void* stream) {
    if (n < 2 || !t || !y || !s) return ssb_set_error(SSB_ERR_ARG, "track_slopes: need n >= 2 and non-NULL arrays");
    track_slopes_kernel<<<nblk(n, 128), 128, 0, (cudaStream_t)stream>>>(n, t, y, s);
    CKL("track_slopes_kernel");
    return 0;
}

int ssb_track_eval_f64(const ssb_track* tr, int64_t nq, const double* tq, double* out, double* dout, void* stream) {
    if (!tr || nq < 0 || (nq > 0 && (!tq || !out))) return ssb_set_error(SSB_ERR_ARG, "track_eval: NULL argument");
    if (nq == 0) return 0;
    track_eval_kernel<<<nblk(nq, 128), 128, 0, (cudaStream_t)stream>>>(*tr, nq, tq, out, dout);
    CKL("track_eval_kernel");
    return 0;
}

// SSB_ORBIT_NOEXTRAS=0: keep the general final-state kernel for programs that are exactly a fused signature (A/B)
static bool orbit_noextras_enabled() {
    const char* e = getenv("SSB_ORBIT_NOEXTRAS");
    return !(e && e[0] == '0');
}

int ssb_orbit_integrate_f64(const ssb_potential* pot, int64_t N, const double* w0, const double* t0, const double* t1, const double* ts,
                            int32_t M, int32_t ts_per_orbit, ssb_ctrl ctrl, double* ys, int32_t* status, int32_t* nsteps, void* stream) {
    if (int e = ssb_validate_potential(pot)) return e;
    if (int e = ssb_validate_ctrl(ctrl)) return e;
    if (N < 0 || M < 0) return ssb_set_error(SSB_ERR_ARG, "orbit_integrate: negative N or M");
    if (N == 0) return 0;
    if (!w0 || !t0 || !t1 || !status || !nsteps || (M > 0 && (!ts || !ys))) return ssb_set_error(SSB_ERR_ARG, "orbit_integrate: NULL array");
    OrbitArgs a;
    a.N = N; a.w0 = w0; a.t0 = t0; a.t1 = t1; a.ts = ts; a.M = M; a.ts_per_orbit = ts_per_orbit; a.ys = ys; a.status = status; a.nsteps = nsteps;
    a.c = to_dev(ctrl);
    const unsigned grid = nblk(N, SSB_ORBIT_THREADS);
    cudaStream_t st = (cudaStream_t)stream;
    const bool final_only = (M == 1 && ts_per_orbit && ts == t1);     // ts aliases t1: keep the final state, no dense output
    ssb_potential pc;
    const int sig = ssb_canonicalize(pot, &pc);
    // XS = number of "fast extras" (linear-track moving perturbers / frame acceleration) next to a fused MW signature, final-state mode
    int xs = (final_only && (sig == SIG_NHM || sig == SIG_NHHM)) ? ssb_fast_extras(&pc, sig == SIG_NHM ? 3 : 4) : 0;
    if (!xs && final_only && (sig == SIG_NHM || sig == SIG_NHHM) && ssb_fast_extra_cubic(&pc, sig == SIG_NHM ? 3 : 4)) xs = 3;
    if (!xs && final_only && (sig == SIG_NHM || sig == SIG_NHHM) && pc.n_comp == (sig == SIG_NHM ? 3 : 4) && orbit_noextras_enabled()) xs = 4;
    // MODE 0 (SaveAt with dense output): one step record of (14 + 3 stages) doubles per thread in dynamic shared memory (coop_dense)
#define SSB_LAUNCH_ORBIT_X(S, MD, SG, X) do { const size_t shm = (MD) == 0 ? sizeof(double) * ((14 + 3 * ((S) == 5 ? 7 : 14)) * SSB_ORBIT_THREADS + 208) : 0; \
        if (shm > 48 * 1024) CK(cudaFuncSetAttribute(orbit_kernel<S, MD, SG, X>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm)); \
        orbit_kernel<S, MD, SG, X><<<grid, SSB_ORBIT_THREADS, shm, st>>>(pc, a); } while (0)
#define SSB_LAUNCH_ORBIT(S, MD, SG) SSB_LAUNCH_ORBIT_X(S, MD, SG, 0)
    // saving mode for a program that is exactly a fused MW signature: the variant without any extras path in the step loop (as XS = 4 above)
#define SSB_LAUNCH_SAVE_NOX(S) do { if (sig == SIG_NHM) SSB_LAUNCH_ORBIT_X(S, 0, SIG_NHM, 4); else SSB_LAUNCH_ORBIT_X(S, 0, SIG_NHHM, 4); } while (0)
    const bool save_nox = !final_only && (sig == SIG_NHM || sig == SIG_NHHM) && pc.n_comp == (sig == SIG_NHM ? 3 : 4) && orbit_noextras_enabled();
#define SSB_LAUNCH_SIG(S, MD) do { switch (sig) { case SIG_N: SSB_LAUNCH_ORBIT(S, MD, SIG_N); break; case SIG_NHM: SSB_LAUNCH_ORBIT(S, MD, SIG_NHM); break; \
        case SIG_NHHM: SSB_LAUNCH_ORBIT(S, MD, SIG_NHHM); break; default: SSB_LAUNCH_ORBIT(S, MD, SIG_GENERIC); } } while (0)
#define SSB_LAUNCH_XS(S) do { if (sig == SIG_NHM) { if (xs == 1) orbit_kernel<S, 2, SIG_NHM, 1><<<grid, SSB_ORBIT_THREADS, 0, st>>>(pc, a); \
                                                      else if (xs == 2) orbit_kernel<S, 2, SIG_NHM, 2><<<grid, SSB_ORBIT_THREADS, 0, st>>>(pc, a); \
                                                      else if (xs == 3) orbit_kernel<S, 2, SIG_NHM, 3><<<grid, SSB_ORBIT_THREADS, 0, st>>>(pc, a); \
                                                      else orbit_kernel<S, 2, SIG_NHM, 4><<<grid, SSB_ORBIT_THREADS, 0, st>>>(pc, a); } \
        else { if (xs == 1) orbit_kernel<S, 2, SIG_NHHM, 1><<<grid, SSB_ORBIT_THREADS, 0, st>>>(pc, a); \
               else if (xs == 2) orbit_kernel<S, 2, SIG_NHHM, 2><<<grid, SSB_ORBIT_THREADS, 0, st>>>(pc, a); \
               else if (xs == 3) orbit_kernel<S, 2, SIG_NHHM, 3><<<grid, SSB_ORBIT_THREADS, 0, st>>>(pc, a); \
               else orbit_kernel<S, 2, SIG_NHHM, 4><<<grid, SSB_ORBIT_THREADS, 0, st>>>(pc, a); } } while (0)
    if (ctrl.solver == 5) { if (xs) SSB_LAUNCH_XS(5); else if (final_only) SSB_LAUNCH_SIG(5, 2); else if (save_nox) SSB_LAUNCH_SAVE_NOX(5); else SSB_LAUNCH_SIG(5, 0); }
    else { if (xs) SSB_LAUNCH_XS(8); else if (final_only) SSB_LAUNCH_SIG(8, 2); else if (save_nox) SSB_LAUNCH_SAVE_NOX(8); else SSB_LAUNCH_SIG(8, 0); }
    CKL("orbit_kernel");
    return 0;
}

size_t ssb_scratch_bytes(int32_t max_steps) { return sizeof(double) * (8 + (size_t)(max_steps > 0 ? max_steps : 1) * SSB_REC_STRIDE); }

static int dense_launch(const ssb_potential* pot, const double* w0, double t0, double t1, const double* t0p, const double* t1p,
                        const double* ts, int64_t M, const ssb_ctrl& ctrl, double* ys, int32_t* status, int32_t* nsteps, double* scratch,
                        cudaStream_t st, int64_t ts_begin = 0, int64_t ts_stride = 1, const double* tstop_p = nullptr) {
    const CtrlDev c = to_dev(ctrl);
    ssb_potential pc;
    const int sig = ssb_canonicalize(pot, &pc);
#define SSB_LAUNCH_DENSE(S, SG) dense_step_kernel<S, SG><<<1, 32, 0, st>>>(pc, w0, t0, t1, t0p, t1p, c, scratch, ctrl.max_steps, status, nsteps, tstop_p)
#define SSB_LAUNCH_DENSE_SIG(S) do { switch (sig) { case SIG_N: SSB_LAUNCH_DENSE(S, SIG_N); break; case SIG_NHM: SSB_LAUNCH_DENSE(S, SIG_NHM); break; \
        case SIG_NHHM: SSB_LAUNCH_DENSE(S, SIG_NHHM); break; default: SSB_LAUNCH_DENSE(S, SIG_GENERIC); } } while (0)
    // a program that is exactly a fused MW signature: the stepper without any extras path in its loop (XS = 4, as the orbit kernels) - this
    // single-thread solve is pure latency, and the call site's spills are on its critical path
#define SSB_LAUNCH_DENSE_NOX(S) do { if (sig == SIG_NHM) dense_step_kernel<S, SIG_NHM, 4><<<1, 32, 0, st>>>(pc, w0, t0, t1, t0p, t1p, c, scratch, ctrl.max_steps, status, nsteps, tstop_p); \
        else dense_step_kernel<S, SIG_NHHM, 4><<<1, 32, 0, st>>>(pc, w0, t0, t1, t0p, t1p, c, scratch, ctrl.max_steps, status, nsteps, tstop_p); } while (0)
    const bool nox = (sig == SIG_NHM || sig == SIG_NHHM) && pc.n_comp == (sig == SIG_NHM ? 3 : 4) && orbit_noextras_enabled();
    if (nox) { if (ctrl.solver == 5) SSB_LAUNCH_DENSE_NOX(5); else SSB_LAUNCH_DENSE_NOX(8); }
    else { if (ctrl.solver == 5) SSB_LAUNCH_DENSE_SIG(5); else SSB_LAUNCH_DENSE_SIG(8); }
    CKL("dense_step_kernel");
    if (M > 0) {
        if (ctrl.solver == 5) dense_eval_kernel<5><<<nblk(M, 128), 128, 0, st>>>(scratch, ts, M, ys, ts_begin, ts_stride);
        else dense_eval_kernel<8><<<nblk(M, 128), 128, 0, st>>>(scratch, ts, M, ys, ts_begin, ts_stride);
        CKL("dense_eval_kernel");
    }
    return 0;
}

int ssb_orbit_dense_f64(const ssb_potential* pot, const double* w0, double t0, double t1, const double* ts, int64_t M, ssb_ctrl ctrl,
                        double* ys, int32_t* status, int32_t* nsteps, void* scratch, size_t scratch_bytes, void* stream) {
    if (int e = ssb_validate_potential(pot)) return e;
    if (int e = ssb_validate_ctrl(ctrl)) return e;
    if (!w0 || M < 0 || (M > 0 && (!ts || !ys)) || !scratch) return ssb_set_error(SSB_ERR_ARG, "orbit_dense: NULL array");
    if (scratch_bytes < ssb_scratch_bytes(ctrl.max_steps)) return ssb_set_error(SSB_ERR_SCRATCH, "orbit_dense: scratch too small");
    cudaStream_t st = (cudaStream_t)stream;
    return dense_launch(pot, w0, t0, t1, nullptr, nullptr, ts, M, ctrl, ys, status, nsteps, (double*)scratch, st);
}

size_t ssb_record_bytes(int64_t N, int32_t rec_cap) { return sizeof(double) * (size_t)(N > 0 ? N : 1) * (8 + (size_t)(rec_cap > 0 ? rec_cap : 1) * SSB_REC_STRIDE); }

int ssb_orbit_record_f64(const ssb_potential* pot, int64_t N, const double* w0, const double* t0, const double* t1, ssb_ctrl ctrl, int32_t rec_cap,
                         void* recs, size_t rec_bytes, int32_t* status, int32_t* nsteps, void* stream) {
    if (int e = ssb_validate_potential(pot)) return e;
    if (int e = ssb_validate_ctrl(ctrl)) return e;
    if (N < 0 || rec_cap < 1) return ssb_set_error(SSB_ERR_ARG, "orbit_record: negative N or rec_cap < 1");
    if (N == 0) return 0;
    if (!w0 || !t0 || !t1 || !recs || !status || !nsteps) return ssb_set_error(SSB_ERR_ARG, "orbit_record: NULL array");
    if (rec_bytes < ssb_record_bytes(N, rec_cap)) return ssb_set_error(SSB_ERR_SCRATCH, "orbit_record: record buffer too small (ssb_record_bytes)");
    const CtrlDev c = to_dev(ctrl);
    ssb_potential pc;
    const int sig = ssb_canonicalize(pot, &pc);
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = nblk(N, SSB_ORBIT_THREADS);
#define SSB_LAUNCH_REC(S, SG) record_kernel<S, SG><<<grid, SSB_ORBIT_THREADS, 0, st>>>(pc, N, w0, t0, t1, c, (double*)recs, rec_cap, status, nsteps)
#define SSB_LAUNCH_REC_SIG(S) do { switch (sig) { case SIG_NHM: SSB_LAUNCH_REC(S, SIG_NHM); break; case SIG_NHHM: SSB_LAUNCH_REC(S, SIG_NHHM); break; \
        default: record_kernel<S, SIG_GENERIC><<<grid, SSB_ORBIT_THREADS, 0, st>>>(*pot, N, w0, t0, t1, c, (double*)recs, rec_cap, status, nsteps); } } while (0)
    if (ctrl.solver == 5) SSB_LAUNCH_REC_SIG(5); else SSB_LAUNCH_REC_SIG(8);
    CKL("record_kernel");
    return 0;
}

int ssb_orbit_trace_f64(const ssb_potential* pot, int64_t N, const double* w0, const double* t0, const double* t1, ssb_ctrl ctrl, int32_t trace_cap,
                        double* trace, double* yfin, int32_t* status, int32_t* nsteps, void* stream) {
    if (int e = ssb_validate_potential(pot)) return e;
    if (int e = ssb_validate_ctrl(ctrl)) return e;
    if (N < 0 || trace_cap < 1) return ssb_set_error(SSB_ERR_ARG, "orbit_trace: negative N or trace_cap < 1");
    if (N == 0) return 0;
    if (!w0 || !t0 || !t1 || !trace || !yfin || !status || !nsteps) return ssb_set_error(SSB_ERR_ARG, "orbit_trace: NULL array");
    const CtrlDev c = to_dev(ctrl);
    ssb_potential pc;
    const int sig = ssb_canonicalize(pot, &pc);
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = nblk(N, SSB_ORBIT_THREADS);
#define SSB_LAUNCH_TRACE(S, SG) trace_kernel<S, SG><<<grid, SSB_ORBIT_THREADS, 0, st>>>(pc, N, w0, t0, t1, c, trace, trace_cap, yfin, status, nsteps)
#define SSB_LAUNCH_TRACE_SIG(S) do { switch (sig) { case SIG_N: SSB_LAUNCH_TRACE(S, SIG_N); break; case SIG_NHM: SSB_LAUNCH_TRACE(S, SIG_NHM); break; \
        case SIG_NHHM: SSB_LAUNCH_TRACE(S, SIG_NHHM); break; default: SSB_LAUNCH_TRACE(S, SIG_GENERIC); } } while (0)
    if (ctrl.solver == 5) SSB_LAUNCH_TRACE_SIG(5); else SSB_LAUNCH_TRACE_SIG(8);
    CKL("trace_kernel");
    return 0;
}

int ssb_orbit_record_eval_f64(int32_t solver, int64_t N, const void* recs, int32_t rec_cap, const double* tq, int32_t per_orbit, double* ys, void* stream) {
    if (solver != 5 && solver != 8) return ssb_set_error(SSB_ERR_UNSUPPORTED, "solver must be 5 (Dopri5) or 8 (Dopri8)");
    if (N < 0 || rec_cap < 1) return ssb_set_error(SSB_ERR_ARG, "orbit_record_eval: negative N or rec_cap < 1");
    if (N == 0) return 0;
    if (!recs || !tq || !ys) return ssb_set_error(SSB_ERR_ARG, "orbit_record_eval: NULL array");
    cudaStream_t st = (cudaStream_t)stream;
    if (solver == 5) record_eval_kernel<5><<<nblk(N, 128), 128, 0, st>>>((const double*)recs, rec_cap, N, tq, per_orbit, ys);
    else record_eval_kernel<8><<<nblk(N, 128), 128, 0, st>>>((const double*)recs, rec_cap, N, tq, per_orbit, ys);
    CKL("record_eval_kernel");
    return 0;
}

int ssb_orbit_dense_eval_f64(int32_t solver, const void* scratch, const double* ts, int64_t M, double* ys, void* stream) {
    if (solver != 5 && solver != 8) return ssb_set_error(SSB_ERR_UNSUPPORTED, "solver must be 5 (Dopri5) or 8 (Dopri8)");
    if (!scratch || M < 0 || (M > 0 && (!ts || !ys))) return ssb_set_error(SSB_ERR_ARG, "orbit_dense_eval: NULL array");
    if (M == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (solver == 5) dense_eval_kernel<5><<<nblk(M, 128), 128, 0, st>>>((const double*)scratch, ts, M, ys);
    else dense_eval_kernel<8><<<nblk(M, 128), 128, 0, st>>>((const double*)scratch, ts, M, ys);
    CKL("dense_eval_kernel");
    return 0;
}

int ssb_release_spray_f64(const ssb_potential* pot, double G, int64_t N, const double* prog, const double* Msat, const int64_t* idx,
                          const double* t, int64_t seed, const double* kvals, const double* normals, double* pos_lead, double* pos_trail,
                          double* vel_lead, double* vel_trail, void* stream) {
    if (int e = ssb_validate_potential(pot)) return e;
    if (N < 0) return ssb_set_error(SSB_ERR_ARG, "release_spray: negative N");
    if (N == 0) return 0;
    if (!prog || !Msat || !t || !kvals || !pos_lead || !pos_trail || !vel_lead || !vel_trail) return ssb_set_error(SSB_ERR_ARG, "release_spray: NULL array");
    ReleaseArgs a;
    memset(&a, 0, sizeof(a));
    a.N = N; a.prog = prog; a.Msat = Msat; a.t = t; a.normals = normals; a.idx = idx; a.G = G;
    int64_t r5[5]; host_randint5(seed, r5);
    for (int q = 0; q < 4; ++q) a.r[q] = r5[q];
    memcpy(a.kv, kvals, sizeof(a.kv));
    a.pos_lead = pos_lead; a.pos_trail = pos_trail; a.vel_lead = vel_lead; a.vel_trail = vel_trail;
    release_kernel<<<nblk(N, 128), 128, 0, (cudaStream_t)stream>>>(*pot, a);
    CKL("release_kernel");
    return 0;
}

int ssb_release_chen25_f64(const ssb_potential* pot, double G, int64_t N, const double* prog, const double* Msat, const double* t,
                           const uint32_t* key, const double* mean, const double* factor, const double* normals, double* pos_lead,
                           double* pos_trail, double* vel_lead, double* vel_trail, void* stream) {
    if (int e = ssb_validate_potential(pot)) return e;
    if (N < 0) return ssb_set_error(SSB_ERR_ARG, "release_chen25: negative N");
    if (N == 0) return 0;
    if (!prog || !Msat || !t || !mean || !factor || (!key && !normals) || !pos_lead || !pos_trail || !vel_lead || !vel_trail)
        return ssb_set_error(SSB_ERR_ARG, "release_chen25: NULL array");
    Chen25Args a;
    memset(&a, 0, sizeof(a));
    a.N = N; a.prog = prog; a.Msat = Msat; a.t = t; a.normals = normals; a.G = G;
    if (key) { a.key[0] = key[0]; a.key[1] = key[1]; }
    memcpy(a.mean, mean, sizeof(a.mean)); memcpy(a.factor, factor, sizeof(a.factor));
    a.pos_lead = pos_lead; a.pos_trail = pos_trail; a.vel_lead = vel_lead; a.vel_trail = vel_trail;
    release_chen25_kernel<<<nblk(N, 128), 128, 0, (cudaStream_t)stream>>>(*pot, a);
    CKL("release_chen25_kernel");
    return 0;
}

int ssb_release_jacobian_f64(const ssb_potential* pot, double G, int64_t N, const double* prog, const double* Msat, const int64_t* idx,
                             const double* t, int64_t seed, const double* kvals, const double* normals, double* jac, void* stream) {
    if (int e = ssb_validate_potential(pot)) return e;
    if (N < 0) return ssb_set_error(SSB_ERR_ARG, "release_jacobian: negative N");
    if (N == 0) return 0;
    if (!prog || !Msat || !t || !kvals || !jac) return ssb_set_error(SSB_ERR_ARG, "release_jacobian: NULL array");
    ReleaseArgs a;
    memset(&a, 0, sizeof(a));
    a.N = N; a.prog = prog; a.Msat = Msat; a.t = t; a.normals = normals; a.idx = idx; a.G = G;
    int64_t r5[5]; host_randint5(seed, r5);
    for (int q = 0; q < 4; ++q) a.r[q] = r5[q];
    memcpy(a.kv, kvals, sizeof(a.kv));
    release_jacobian_kernel<<<nblk(N, 64), 64, 0, (cudaStream_t)stream>>>(*pot, a, jac);
    CKL("release_jacobian_kernel");
    return 0;
}

int ssb_potential_third_f64(const ssb_potential* pot, int64_t n, const double* xyz, const double* t, double* third, void* stream) {
    if (int e = ssb_validate_potential(pot)) return e;
    if (n < 0 || (n > 0 && (!xyz || !t || !third))) return ssb_set_error(SSB_ERR_ARG, "potential_third: bad n / NULL array");
    if (n == 0) return 0;
    potential_third_kernel<<<nblk(n, 128), 128, 0, (cudaStream_t)stream>>>(*pot, n, xyz, t, third);
    CKL("potential_third_kernel");
    return 0;
}

// scratch layout of gen_stream: [dense record | prog Nts*6 | w0_packed 2*Nts*6 | w0 2n*6 | t0 2n | t1 2n | ys 2n*6 | dense records of parts 1..3]
size_t ssb_stream_scratch_bytes(int64_t Nts, int32_t max_steps) {
    return 4 * ssb_scratch_bytes(max_steps) + sizeof(double) * (size_t)Nts * (6 + 12 + 12 + 2 + 2 + 12) + 256;      // 4 = SSB_STREAM_PARTS record buffers
}

// ---- multi-part pipeline of gen_stream ----
// The progenitor orbit is ONE serial solve (~0.75 ms for 3 Gyr) that every release waits for.  For large streams the particles are cut
// into up to SSB_STREAM_PARTS parts by stripping time (1/8, 2/8, 2/8, 3/8 of the shard: the earliest = longest orbits first):
//   caller's stream    : progenitor only until it has passed part 0's last stripping time (~0.1 ms) -> release -> orbits of part 0
//   internal stream k  : the progenitor again, until part k's last stripping time (its own record buffer) -> release -> orbits of part k
// All progenitor solves take the same steps (the early stop only ends the loop), so every particle sees the same bits as in the one-part
// pipeline.  The orbit kernels start after ~0.1 ms and are fed part by part faster than they drain (the progenitor reaches stripping
// fraction f after f x 0.75 ms), so the serial solve leaves the critical path - which is what a stream SPLIT OVER SEVERAL GPUs needs:
// at 125 k particles per GPU the orbits take ~1.3 ms, and a 0.75 ms serial prefix would cost a third of the step.
#ifndef SSB_STREAM_SPLIT_MIN
#define SSB_STREAM_SPLIT_MIN 32768          // particles per arm below which the one-part pipeline is used
#endif
#define SSB_STREAM_PARTS 4
struct StreamAux { bool ok; cudaStream_t s[SSB_STREAM_PARTS - 1]; cudaEvent_t e0, e[SSB_STREAM_PARTS - 1]; };
static StreamAux* stream_aux() {
    static thread_local StreamAux aux[64];
    static thread_local bool tried[64];
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    if (!tried[dev]) {
        tried[dev] = true;
        StreamAux& x = aux[dev];
        x.ok = cudaEventCreateWithFlags(&x.e0, cudaEventDisableTiming) == cudaSuccess;
        for (int k = 0; k < SSB_STREAM_PARTS - 1; ++k)
            x.ok = x.ok && cudaStreamCreateWithFlags(&x.s[k], cudaStreamNonBlocking) == cudaSuccess &&
                   cudaEventCreateWithFlags(&x.e[k], cudaEventDisableTiming) == cudaSuccess;
        if (!x.ok) cudaGetLastError();
    }
    return aux[dev].ok ? &aux[dev] : nullptr;
}
static bool stream_split_enabled() {
    const char* e = getenv("SSB_STREAM_SPLIT");
    return !(e && e[0] == '0');
}

// compact != 0 (internal, ssb_gen_stream_host on a shard): ts / Msat / normals hold only the shard's n_local stripping times (row k <->
// stripping index i_begin + k i_stride) followed, in ts, by the two ends of the progenitor interval: ts[n_local] = first, ts[n_local + 1] =
// last stripping time of the WHOLE stream.
static int gen_stream_impl(const ssb_potential* pot, const ssb_potential* pot_release, double G, int64_t Nts, const double* ts, const double* prog_w0,
                           const double* Msat, int64_t seed, const double* kvals, const double* normals, ssb_ctrl ctrl, int64_t i_begin, int64_t i_stride,
                           int64_t n_local, double* lead, double* trail, int32_t* status, int32_t* nsteps, void* scratch, size_t scratch_bytes, void* stream,
                           int compact);

int ssb_gen_stream_f64(const ssb_potential* pot, const ssb_potential* pot_release, double G, int64_t Nts, const double* ts, const double* prog_w0,
                       const double* Msat, int64_t seed, const double* kvals, const double* normals, ssb_ctrl ctrl, int64_t i_begin, int64_t i_stride,
                       int64_t n_local, double* lead, double* trail, int32_t* status, int32_t* nsteps, void* scratch, size_t scratch_bytes, void* stream) {
    return gen_stream_impl(pot, pot_release, G, Nts, ts, prog_w0, Msat, seed, kvals, normals, ctrl, i_begin, i_stride, n_local, lead, trail, status, nsteps,
                           scratch, scratch_bytes, stream, 0);
}
int ssb_gen_stream_compact(const ssb_potential* pot, const ssb_potential* pot_release, double G, int64_t Nts, const double* ts_c, const double* prog_w0,
                           const double* Msat_c, int64_t seed, const double* kvals, const double* normals_c, ssb_ctrl ctrl, int64_t i_begin, int64_t i_stride,
                           int64_t n_local, double* lead, double* trail, int32_t* status, int32_t* nsteps, void* scratch, size_t scratch_bytes, void* stream) {
    return gen_stream_impl(pot, pot_release, G, Nts, ts_c, prog_w0, Msat_c, seed, kvals, normals_c, ctrl, i_begin, i_stride, n_local, lead, trail, status,
                           nsteps, scratch, scratch_bytes, stream, 1);
}

static int gen_stream_impl(const ssb_potential* pot, const ssb_potential* pot_release, double G, int64_t Nts, const double* ts, const double* prog_w0,
                           const double* Msat, int64_t seed, const double* kvals, const double* normals, ssb_ctrl ctrl, int64_t i_begin, int64_t i_stride,
                           int64_t n_local, double* lead, double* trail, int32_t* status, int32_t* nsteps, void* scratch, size_t scratch_bytes, void* stream,
                           int compact) {
    if (int e = ssb_validate_potential(pot)) return e;
    if (int e = ssb_validate_potential(pot_release)) return e;
    if (int e = ssb_validate_ctrl(ctrl)) return e;
    if (Nts < 2 || !ts || !prog_w0 || !Msat || !kvals || !scratch) return ssb_set_error(SSB_ERR_ARG, "gen_stream: NULL array or Nts < 2");
    if (i_begin < 0 || i_stride < 1 || n_local < 0 || (n_local > 0 && i_begin + (n_local - 1) * i_stride > Nts - 2))
        return ssb_set_error(SSB_ERR_ARG, "gen_stream: particle selection outside [0, Nts-1)");
    const int64_t n = n_local;
    const int64_t Nlay = compact ? n + 2 : Nts;             // rows the scratch layout is sized for
    if (scratch_bytes < ssb_stream_scratch_bytes(Nlay, ctrl.max_steps)) return ssb_set_error(SSB_ERR_SCRATCH, "gen_stream: scratch too small");
    if (n > 0 && (!lead || !trail || !status || !nsteps)) return ssb_set_error(SSB_ERR_ARG, "gen_stream: NULL output");
    cudaStream_t st = (cudaStream_t)stream;
    double* dense = (double*)scratch;
    double* prog = dense + ssb_scratch_bytes(ctrl.max_steps) / sizeof(double);
    double* w0p = prog + 6 * Nlay;
    double* w0 = w0p + 12 * Nlay;
    double* t0 = w0 + 12 * Nlay;
    double* t1 = t0 + 2 * Nlay;
    double* ys = t1 + 2 * Nlay;
    double* dense2 = ys + 12 * Nlay;
    if (n == 0) return 0;
    // where this shard's stripping times live: the full arrays (row i_begin + k i_stride) or compact ones (row k)
    const double* ts_first = compact ? ts + n : ts;
    const double* ts_last = compact ? ts + n + 1 : ts + (Nts - 1);
    const int64_t tb = compact ? 0 : i_begin, tstr = compact ? 1 : i_stride;
    int64_t r5[5]; host_randint5(seed, r5);
    StreamAux* ax = (n >= SSB_STREAM_SPLIT_MIN && trail == lead + 6 * n && stream_split_enabled()) ? stream_aux() : nullptr;
    if (ax) {
        // part boundaries in units of this shard's stripping times (multiples of 128 releases = whole CTAs)
        auto r128 = [](int64_t v) { return ((v + 127) / 128) * 128; };
        const int64_t b[SSB_STREAM_PARTS + 1] = {0, r128(n / 8), r128(3 * n / 8), r128(5 * n / 8), n};
        const size_t dense_doubles = ssb_scratch_bytes(ctrl.max_steps) / sizeof(double);
        ReleaseArgs a;
        memset(&a, 0, sizeof(a));
        a.N = n; a.prog = prog; a.Msat = Msat; a.t = ts; a.normals = normals; a.idx = nullptr; a.G = G;
        a.sel_begin = tb; a.sel_stride = tstr;
        if (compact) { a.draw_begin = i_begin; a.draw_stride = i_stride; }
        for (int q = 0; q < 4; ++q) a.r[q] = r5[q];
        memcpy(a.kv, kvals, sizeof(a.kv));
        a.w0_packed = w0; a.t0_packed = t0; a.t1_packed = t1; a.t_end = ts_last;
        CK(cudaEventRecord(ax->e0, st));                                   // the inputs are ready in the caller's stream order
        // part 0 (shortest serial prefix, longest orbits) is enqueued first, on the caller's stream: its kernels reach the GPU while the
        // host is still enqueuing the other parts
        for (int k = 0; k < SSB_STREAM_PARTS; ++k) {
            const int64_t p0 = b[k], cnt = b[k + 1] - b[k];
            if (cnt <= 0) continue;
            cudaStream_t sk = k == 0 ? st : ax->s[k - 1];
            if (k > 0) CK(cudaStreamWaitEvent(sk, ax->e0, 0));
            double* rec = k == 0 ? dense : dense2 + (size_t)(k - 1) * dense_doubles;
            // progenitor until it has passed this part's last stripping time (the last part: the whole orbit), its dense output at the
            // part's stripping times, release, orbits (lead block, trail block)
            const double* tstop = (k == SSB_STREAM_PARTS - 1) ? nullptr : ts + (tb + (b[k + 1] - 1) * tstr);
            if (int e = dense_launch(pot, prog_w0, 0.0, 0.0, ts_first, ts_last, ts, cnt, ctrl, prog + 6 * p0, nullptr, nullptr, rec, sk, tb + p0 * tstr, tstr, tstop))
                return e;
            a.i0 = p0; a.cnt = cnt;
            release_kernel<<<nblk(cnt, 128), 128, 0, sk>>>(*pot_release, a);
            CKL("release_kernel");
            if (int e = ssb_orbit_integrate_f64(pot, cnt, w0 + 6 * p0, t0 + p0, t1 + p0, t1 + p0, 1, 1, ctrl, lead + 6 * p0, status + p0, nsteps + 3 * p0, sk)) return e;
            if (int e = ssb_orbit_integrate_f64(pot, cnt, w0 + 6 * (n + p0), t0 + n + p0, t1 + n + p0, t1 + n + p0, 1, 1, ctrl, lead + 6 * (n + p0),
                                                status + n + p0, nsteps + 3 * (n + p0), sk)) return e;
            if (k > 0) CK(cudaEventRecord(ax->e[k - 1], sk));
        }
        for (int k = 1; k < SSB_STREAM_PARTS; ++k)
            if (b[k + 1] - b[k] > 0) CK(cudaStreamWaitEvent(st, ax->e[k - 1], 0));      // join: the caller's stream continues after every part
        return 0;
    }
    // (1) progenitor orbit integrate_orbit(prog_w0, ts) with t0 = ts.min, t1 = ts.max (main.py:289, 152-153): ONE serial solve, then its
    //     dense output at THIS SHARD's stripping times only (ts[i_begin + k i_stride], k < n)
    if (int e = dense_launch(pot, prog_w0, 0.0, 0.0, ts_first, ts_last, ts, n, ctrl, prog, nullptr, nullptr, dense, st, tb, tstr)) return e;
    // (2) release at those stripping times (main.py:293-303), written straight into the orbit kernel's input layout
    //     (lead block, trail block) together with the per-orbit start / end times
    ReleaseArgs a;
    memset(&a, 0, sizeof(a));
    a.N = n; a.prog = prog; a.Msat = Msat; a.t = ts; a.normals = normals; a.idx = nullptr; a.G = G;
    a.sel_begin = tb; a.sel_stride = tstr;
        if (compact) { a.draw_begin = i_begin; a.draw_stride = i_stride; }
    for (int q = 0; q < 4; ++q) a.r[q] = r5[q];
    memcpy(a.kv, kvals, sizeof(a.kv));
    a.w0_packed = w0; a.t0_packed = t0; a.t1_packed = t1; a.t_end = ts_last;
    release_kernel<<<nblk(n, 128), 128, 0, st>>>(*pot_release, a);
    CKL("release_kernel");
    // (3) 2n independent solves from ts[i] to ts[-1], keep the final state (main.py:349-368); lead and trail that are adjacent in
    //     memory (one [2,n,6] buffer) receive the kernel's output directly
    if (trail == lead + 6 * n) return ssb_orbit_integrate_f64(pot, 2 * n, w0, t0, t1, t1, 1, 1, ctrl, lead, status, nsteps, stream);
    if (int e = ssb_orbit_integrate_f64(pot, 2 * n, w0, t0, t1, t1, 1, 1, ctrl, ys, status, nsteps, stream)) return e;
    CK(cudaMemcpyAsync(lead, ys, sizeof(double) * 6 * n, cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(trail, ys + 6 * n, sizeof(double) * 6 * n, cudaMemcpyDeviceToDevice, st));
    return 0;
}

int ssb_fp64_peak_probe(int iters, double* flops_per_s, void* stream) {
    if (!flops_per_s || iters <= 0) return ssb_set_error(SSB_ERR_ARG, "fp64_peak_probe: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double* sink = nullptr;
    CK(cudaMalloc(&sink, 8));
    const int blocks = sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    dfma_probe_kernel<<<blocks, threads, 0, st>>>(iters / 4 + 1, sink);
    CK(cudaEventRecord(e0, st));
    dfma_probe_kernel<<<blocks, threads, 0, st>>>(iters, sink);
    CK(cudaEventRecord(e1, st));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    *flops_per_s = 2.0 * 64.0 * (double)iters * blocks * threads / (ms * 1e-3);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
    return 0;
}

}  // extern "C"
