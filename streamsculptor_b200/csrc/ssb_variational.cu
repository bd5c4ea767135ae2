// K7: variational (tangent) equations along unperturbed orbits - the state-transition matrix dw/dw_init and, optionally, the
// second-order tensor d2w/dw_init^2 of every particle.
//
// Reference: examples/higher_order_variationalEqn.ipynb cell 3 (`second_order_field.term` wrapped in fields.CustomField,
// fields.py:362-377, integrated per particle by fields.integrate_field, fields.py:35-99): state [w(6), M(6,6), M2(6,6,6)],
//   dM/dt  = [dv_dw ; (da/dw) M],                       da/dw  = [-Hess Phi(x, t), 0]           (jacfwd of the acceleration)
//   dM2/dt = [d2v_dw2 ; (da/dw) M2 + (d2a/dw2) M M],    d2a/dw2 = -Phi_ijk on the position block (jacfwd twice)
// ONE diffrax solve per particle: the step-size controller's RMS error norm runs over all 42 (258) components.  It is also the
// forward sensitivity (JVP) of integrate_orbit with respect to w0.
//
// B200 mapping (warp-cooperative group per particle): every column of M and every (k <= l) pair of M2 is an independent
// second-order 3-vector system driven by the base orbit, q'' = T(x(t)) q [+ C(x; dx_k, dx_l)], exactly the shape of the orbit
// itself.  Lane 0 of a group integrates the base orbit, lanes 1..6 the six columns, lanes 7..27 (order 2) the 21 unique pairs
// (M2 is symmetric in (k,l); the duplicate entries enter the error norm with weight 2).  Group = 8 lanes (order 1: 4 particles
// per warp) or the whole warp (order 2).  Every lane keeps its Nystrom force stages in registers like K1; per stage the base
// position (and, for order 2, the two column positions a pair needs) travel by warp shuffles, every lane evaluates the tidal
// tensor at the base position in lock-step (SIMT: no extra issue slots), and the shared controller is a shuffle reduction.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "ssb_common.cuh"

using namespace ssb;

#define CKL(what) do { ssb_count_launch(); int _e = ssb_cuda_check(cudaGetLastError(), what); if (_e) return _e; } while (0)
#define SSB_VAR_THREADS 128
#define FULLMASK 0xffffffffu

struct VarArgs {
    int64_t N;
    const double *w0, *M0, *M20, *t0;
    double t1;
    CtrlDev c;
    double *wout, *Mout, *M2out;
    int32_t *status, *nsteps;
};

enum { ROLE_BASE = 0, ROLE_COL = 1, ROLE_PAIR = 2, ROLE_IDLE = 3 };

// gradient, Hessian (and third derivatives) of the total potential at the base position
template <int ORDER, int SIG>
__device__ __noinline__ void var_field_eval(const ssb_potential* P, const ssb_potential* Pc, double bx, double by, double bz, double t, double* g,
                                            Sym3* H, Sym3x3* T3) {
    const double X[3] = {bx, by, bz};
    double phi;
    if (SIG == SIG_GENERIC) {
        pot_eval<WANT_GRAD | WANT_HESS>(*P, X, t, phi, g, *H);
    } else {
        fused_eval<SIG, WANT_GRAD | WANT_HESS>(*Pc, X, g, *H);
        if (Pc->n_comp > SigInfo<SIG>::NF) {
            double g2[3];
            Sym3 H2;
            pot_eval<WANT_GRAD | WANT_HESS, false>(*P, X, t, phi, g2, H2, SigInfo<SIG>::NF);
            g[0] += g2[0]; g[1] += g2[1]; g[2] += g2[2];
            H->xx += H2.xx; H->yy += H2.yy; H->zz += H2.zz; H->xy += H2.xy; H->xz += H2.xz; H->yz += H2.yz;
        }
    }
    if (ORDER == 2) pot_third(*P, X, t, *T3);
}

template <int ORDER, int SIG>
struct VarForce {
    static constexpr int G = ORDER == 2 ? 32 : 8;
    const ssb_potential* P; const ssb_potential* Pc;
    double dir;
    int role, src_k, src_l;          // src_*: group-relative lanes holding columns k and l (pairs only)
    __device__ __forceinline__ void operator()(const double X[3], double tau, double A[3]) const {
        double B[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) B[k] = __shfl_sync(FULLMASK, X[k], 0, G);
        double g[3];
        Sym3 H;
        Sym3x3 T3;
        var_field_eval<ORDER, SIG>(P, Pc, B[0], B[1], B[2], tau * dir, g, &H, &T3);
        // tidal term  -Hess . X  (columns and pairs)
        double t0 = -(H.xx * X[0] + H.xy * X[1] + H.xz * X[2]);
        double t1 = -(H.xy * X[0] + H.yy * X[1] + H.yz * X[2]);
        double t2 = -(H.xz * X[0] + H.yz * X[1] + H.zz * X[2]);
        if (ORDER == 2) {
            double U[3], V[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) { U[k] = __shfl_sync(FULLMASK, X[k], src_k, G); V[k] = __shfl_sync(FULLMASK, X[k], src_l, G); }
            if (role == ROLE_PAIR) {
                // C_i = -sum_jp Phi_ijp U_j V_p
                const double a0 = T3.xxx * V[0] + T3.xxy * V[1] + T3.xxz * V[2];     // sum_p Phi_xxp V_p
                const double a1 = T3.xxy * V[0] + T3.xyy * V[1] + T3.xyz * V[2];     // Phi_xyp
                const double a2 = T3.xxz * V[0] + T3.xyz * V[1] + T3.xzz * V[2];     // Phi_xzp
                const double a3 = T3.xyy * V[0] + T3.yyy * V[1] + T3.yyz * V[2];     // Phi_yyp
                const double a4 = T3.xyz * V[0] + T3.yyz * V[1] + T3.yzz * V[2];     // Phi_yzp
                const double a5 = T3.xzz * V[0] + T3.yzz * V[1] + T3.zzz * V[2];     // Phi_zzp
                t0 -= a0 * U[0] + a1 * U[1] + a2 * U[2];
                t1 -= a1 * U[0] + a3 * U[1] + a4 * U[2];
                t2 -= a2 * U[0] + a4 * U[1] + a5 * U[2];
            }
        }
        const bool base = role == ROLE_BASE, idle = role == ROLE_IDLE;
        A[0] = idle ? 0.0 : (base ? -g[0] : t0);
        A[1] = idle ? 0.0 : (base ? -g[1] : t1);
        A[2] = idle ? 0.0 : (base ? -g[2] : t2);
    }
};

template <int G>
__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
    return v;
}
template <int G>
__device__ __forceinline__ bool group_any(bool f) {
    const unsigned b = __ballot_sync(FULLMASK, f);
    const int lane = threadIdx.x & 31;
    const unsigned gm = (G == 32) ? FULLMASK : (((1u << G) - 1u) << (lane & ~(G - 1)));
    return (b & gm) != 0u;
}

template <int SOLVER, int ORDER, int SIG>
__global__ void __launch_bounds__(SSB_VAR_THREADS, 3) variational_kernel(const __grid_constant__ ssb_potential Pin, const VarArgs a) {
    typedef Tab<SOLVER> T;
    constexpr int S = T::S;
    constexpr int G = ORDER == 2 ? 32 : 8;
    constexpr double NCOMP = ORDER == 2 ? 258.0 : 42.0;
    __shared__ ssb_potential sP;
    stage_potential(&sP, &Pin);
    logtab_init();
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t part = gtid / G;
    const int gl = (int)(gtid % G);
    const bool valid = part < a.N;
    const int64_t pi = valid ? part : 0;
    const CtrlDev c = a.c;
    // ---- role of this lane ----
    int role = ROLE_IDLE, ck = 0, cl = 0;
    if (gl == 0) role = ROLE_BASE;
    else if (gl <= 6) { role = ROLE_COL; ck = gl - 1; }
    else if (ORDER == 2 && gl <= 27) {
        role = ROLE_PAIR;
        int q = gl - 7;                       // unique pairs (k <= l) in row-major order
        ck = 0;
        while (q >= 6 - ck) { q -= 6 - ck; ck++; }
        cl = ck + q;
    }
    const double weight = role == ROLE_IDLE ? 0.0 : ((role == ROLE_PAIR && ck != cl) ? 2.0 : 1.0);
    const double t0_in = a.t0[pi], t1_in = a.t1;
    const double dir = (t0_in < t1_in) ? 1.0 : -1.0;
    const double T0 = t0_in * dir, T1 = t1_in * dir;
    VarForce<ORDER, SIG> force{&sP, &Pin, dir, role, 1 + ck, 1 + cl};
    double x[3], p[3], F[S][3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double xv = 0.0, vv = 0.0;
        if (role == ROLE_BASE) { xv = a.w0[6 * pi + k]; vv = a.w0[6 * pi + 3 + k]; }
        else if (role == ROLE_COL) {
            xv = a.M0 ? a.M0[36 * pi + 6 * k + ck] : (k == ck ? 1.0 : 0.0);
            vv = a.M0 ? a.M0[36 * pi + 6 * (k + 3) + ck] : (k + 3 == ck ? 1.0 : 0.0);
        } else if (role == ROLE_PAIR && a.M20) {
            xv = a.M20[216 * pi + 36 * k + 6 * ck + cl];
            vv = a.M20[216 * pi + 36 * (k + 3) + 6 * ck + cl];
        }
        x[k] = xv; p[k] = dir * vv;
    }
    int status = 0, n_steps = 0, n_acc = 0, n_rej = 0;
    bool at_dtmin = false;
    double tprev = T0, tnext = T0;
    // ---- Hairer-Norsett-Wanner initial step over the whole coupled state ----
    {
        force(x, T0, F[0]);
        double sx[3], sp[3], d0 = 0.0, d1 = 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            sx[k] = fma(c.rtol, fabs(x[k]), c.atol); sp[k] = fma(c.rtol, fabs(p[k]), c.atol);
            double q;
            q = x[k] / sx[k]; d0 = fma(q, q, d0); q = p[k] / sp[k]; d0 = fma(q, q, d0);
            q = p[k] / sx[k]; d1 = fma(q, q, d1); q = F[0][k] / sp[k]; d1 = fma(q, q, d1);
        }
        d0 = sqrt(group_sum<G>(weight * d0) / NCOMP); d1 = sqrt(group_sum<G>(weight * d1) / NCOMP);
        const double h0 = hnw_h0(d0, d1);
        double X1[3], F1[3], d2 = 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) X1[k] = fma(h0, p[k], x[k]);
        force(X1, T0 + h0, F1);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            double q;
            q = (fma(h0, F[0][k], p[k]) - p[k]) / sx[k]; d2 = fma(q, q, d2);
            q = (F1[k] - F[0][k]) / sp[k]; d2 = fma(q, q, d2);
        }
        d2 = sqrt(group_sum<G>(weight * d2) / NCOMP) / h0;
        double h = fmin(hnw_h1<T::ORDER>(h0, d1, d2), c.dtmax);
        at_dtmin = h <= c.dtmin;
        h = fmax(h, c.dtmin);
        tnext = fmin(T0 + h, T1);
    }
    // ---- main loop: control flow is uniform inside a group; finished groups keep executing (dt = 0, results discarded) so that
    //      the warp shuffles stay convergent ----
    for (;;) {
        bool active = valid && status == 0 && tprev < T1;
        if (active && n_steps >= c.max_steps) { status = 1; active = false; }
        if (!__any_sync(FULLMASK, active)) break;
        const double dt = active ? tnext - tprev : 0.0;
        double x1[3], p1[3], ex[3], ep[3];
        rk_stages<SOLVER>(force, x, p, tprev, dt, F);
        rk_candidate<SOLVER>(x, p, dt, F, x1, p1);
        force(x1, tprev + T::c(S - 1) * dt, F[S - 1]);
        rk_error<SOLVER>(p, dt, F, ex, ep);
        bool nanl = false, infl = false;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            nanl |= isnan(x1[k]) | isnan(p1[k]);
            infl |= !(isfinite(x1[k]) & isfinite(p1[k]));
        }
        const bool nan_cand = group_any<G>(nanl && role != ROLE_IDLE);
        const bool nonfinite = group_any<G>(infl && role != ROLE_IDLE);
        const double err = sqrt(group_sum<G>(weight * err_sq6(x, p, x1, p1, ex, ep, c.rtol, c.atol, nan_cand)) / NCOMP);
        double hn; bool bad;
        bool adm = at_dtmin;
        const bool keep = pid_update<T::ORDER>(err, dt, c, adm, hn, bad);
        if (!active) continue;
        at_dtmin = adm;
        n_steps++;
        if (bad) { status = 2; n_rej++; continue; }
        if (keep) {
            n_acc++;
            if (nonfinite) { status = 2; continue; }
#pragma unroll
            for (int k = 0; k < 3; ++k) { x[k] = x1[k]; p[k] = p1[k]; F[0][k] = F[S - 1][k]; }
            tprev = tnext;
        } else {
            n_rej++;
        }
        tprev = fmin(tprev, T1);
        double tn = tprev + hn;
        if (tn > T1 - 1e-10) tn = keep ? T1 : tprev + 0.5 * (T1 - tprev);
        tnext = tn;
    }
    if (!valid) return;
    const bool done = status == 0 && tprev >= T1 && T0 < T1;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double xo = done ? x[k] : inf, vo = done ? dir * p[k] : inf;
        if (role == ROLE_BASE) { a.wout[6 * part + k] = xo; a.wout[6 * part + 3 + k] = vo; }
        else if (role == ROLE_COL) { a.Mout[36 * part + 6 * k + ck] = xo; a.Mout[36 * part + 6 * (k + 3) + ck] = vo; }
        else if (role == ROLE_PAIR) {
            a.M2out[216 * part + 36 * k + 6 * ck + cl] = xo; a.M2out[216 * part + 36 * (k + 3) + 6 * ck + cl] = vo;
            a.M2out[216 * part + 36 * k + 6 * cl + ck] = xo; a.M2out[216 * part + 36 * (k + 3) + 6 * cl + ck] = vo;
        }
    }
    if (role == ROLE_BASE) {
        a.status[part] = status;
        a.nsteps[3 * part] = n_steps; a.nsteps[3 * part + 1] = n_acc; a.nsteps[3 * part + 2] = n_rej;
    }
}

// the field at one state y = [w(6), M(36), M2(216)] (unit tests / facade .term)
template <int ORDER>
__global__ void variational_term_kernel(const __grid_constant__ ssb_potential Pin, double t, const double* y, double* dy) {
    __shared__ ssb_potential sP;
    stage_potential(&sP, &Pin);
    if (threadIdx.x != 0) return;
    double g[3];
    Sym3 H;
    Sym3x3 T3;
    var_field_eval<ORDER, SIG_GENERIC>(&sP, &sP, y[0], y[1], y[2], t, g, &H, &T3);
    const double Hm[3][3] = {{H.xx, H.xy, H.xz}, {H.xy, H.yy, H.yz}, {H.xz, H.yz, H.zz}};
    for (int k = 0; k < 3; ++k) { dy[k] = y[3 + k]; dy[3 + k] = -g[k]; }
    const double* M = y + 6; double* dM = dy + 6;
    for (int k = 0; k < 6; ++k)
        for (int i = 0; i < 3; ++i) {
            dM[6 * i + k] = M[6 * (i + 3) + k];
            double acc = 0.0;
            for (int j = 0; j < 3; ++j) acc -= Hm[i][j] * M[6 * j + k];
            dM[6 * (i + 3) + k] = acc;
        }
    if (ORDER < 2) return;
    const double* M2 = y + 42; double* dM2 = dy + 42;
    for (int k = 0; k < 6; ++k)
        for (int l = 0; l < 6; ++l)
            for (int i = 0; i < 3; ++i) {
                dM2[36 * i + 6 * k + l] = M2[36 * (i + 3) + 6 * k + l];
                double acc = 0.0;
                for (int j = 0; j < 3; ++j) acc -= Hm[i][j] * M2[36 * j + 6 * k + l];
                for (int j = 0; j < 3; ++j) for (int q = 0; q < 3; ++q) acc -= third_at(T3, i, j, q) * M[6 * q + l] * M[6 * j + k];
                dM2[36 * (i + 3) + 6 * k + l] = acc;
            }
}

extern "C" {

int ssb_variational_f64(const ssb_potential* pot, int32_t order, int64_t N, const double* w0, const double* M0, const double* M20, const double* t0,
                        double t1, ssb_ctrl ctrl, double* wout, double* Mout, double* M2out, int32_t* status, int32_t* nsteps, void* stream) {
    if (int e = ssb_validate_potential(pot)) return e;
    if (int e = ssb_validate_ctrl(ctrl)) return e;
    if (order != 1 && order != 2) return ssb_set_error(SSB_ERR_UNSUPPORTED, "variational: order must be 1 or 2");
    if (N < 0) return ssb_set_error(SSB_ERR_ARG, "variational: negative N");
    if (N == 0) return 0;                               // empty batch is a no-op, like every batched entry point
    if (!w0 || !t0 || !wout || !Mout || !status || !nsteps || (order == 2 && !M2out)) return ssb_set_error(SSB_ERR_ARG, "variational: NULL array");
    VarArgs a;
    a.N = N; a.w0 = w0; a.M0 = M0; a.M20 = order == 2 ? M20 : nullptr; a.t0 = t0; a.t1 = t1;
    a.c.rtol = ctrl.rtol; a.c.atol = ctrl.atol; a.c.dtmin = ctrl.dtmin; a.c.dtmax = ctrl.dtmax; a.c.max_steps = ctrl.max_steps;
    a.wout = wout; a.Mout = Mout; a.M2out = M2out; a.status = status; a.nsteps = nsteps;
    ssb_potential pc;
    const int sig = ssb_canonicalize(pot, &pc);
    const int G = order == 2 ? 32 : 8;
    const unsigned grid = (unsigned)((N * G + SSB_VAR_THREADS - 1) / SSB_VAR_THREADS);
    cudaStream_t st = (cudaStream_t)stream;
#define SSB_LAUNCH_VAR(SV, OR, SG) variational_kernel<SV, OR, SG><<<grid, SSB_VAR_THREADS, 0, st>>>(pc, a)
#define SSB_LAUNCH_VAR_SIG(SV, OR) do { switch (sig) { case SIG_NHM: SSB_LAUNCH_VAR(SV, OR, SIG_NHM); break; case SIG_NHHM: SSB_LAUNCH_VAR(SV, OR, SIG_NHHM); break; \
        default: variational_kernel<SV, OR, SIG_GENERIC><<<grid, SSB_VAR_THREADS, 0, st>>>(*pot, a); } } while (0)
    if (ctrl.solver == 5) { if (order == 1) SSB_LAUNCH_VAR_SIG(5, 1); else SSB_LAUNCH_VAR_SIG(5, 2); }
    else { if (order == 1) SSB_LAUNCH_VAR_SIG(8, 1); else SSB_LAUNCH_VAR_SIG(8, 2); }
    CKL("variational_kernel");
    return 0;
}

int ssb_variational_term_f64(const ssb_potential* pot, int32_t order, double t, const double* y, double* dy, void* stream) {
    if (int e = ssb_validate_potential(pot)) return e;
    if (!y || !dy || (order != 1 && order != 2)) return ssb_set_error(SSB_ERR_ARG, "variational_term: NULL array / bad order");
    if (order == 1) variational_term_kernel<1><<<1, 32, 0, (cudaStream_t)stream>>>(*pot, t, y, dy);
    else variational_term_kernel<2><<<1, 32, 0, (cudaStream_t)stream>>>(*pot, t, y, dy);
    CKL("variational_term_kernel");
    return 0;
}

}  // extern "C"
