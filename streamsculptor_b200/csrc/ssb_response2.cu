// K4: second-order mass response  (fields.MassRadiusPerturbation_OTF_SecondOrder.term, fields.py:260-320, driven by
// GenerateMassRadiusPerturbation_Chen25.compute_perturbation_second_order_OTF, perturbative.py:757-772).
//
// State per particle: [w(6), D(N_sh,12), E(N_sh,6)] with E = (x2, v2) the second-order mass derivative:
//   x2' = v2,  v2' = T x2 + sum_jk (d2a/dx_k dx_j)_i x1_k x1_j + (da_pert/dx) x1      (fields.py:307-312)
// where T = da/dx = -Hess(Phi_base), d2a/dx2 = -d3 Phi_base, da_pert/dx = -Hess(Phi_subhalo).  One controller for all
// 6 + 18 N_sh components.  Same CTA-per-particle structure as K3; the mass block and the second-order block of a subhalo
// are coupled (x2 is driven by x1), so they are advanced together as ONE 6-component second-order system per thread
// (D = 6 instantiation of the Nystrom helpers), the radius block as a 3-component one.  Not tuned: correctness first.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "ssb_common.cuh"

using namespace ssb;

#define SSB_R2_THREADS 128
#define CK(call) do { int _e = ssb_cuda_check((call), #call); if (_e) return _e; } while (0)
#define CKL(what) do { ssb_count_launch(); int _e = ssb_cuda_check(cudaGetLastError(), what); if (_e) return _e; } while (0)

struct R2Args {
    int64_t N;
    const double *w0, *D0, *E0, *t0;
    double t1;
    const double* t1v;        // optional per-particle end times (ssb_second_order_response_ends_f64); NULL: the common end time t1
    CtrlDev c;
    double *wout, *Dout, *Eout;
    int32_t *status, *nsteps;
    double* scratch;            // [grid][2][18][n_sh]
    unsigned long long* counter;
};

template <int S>
struct Base2Shared {
    double X[S][3], T[S][6], T3[S][10], t[S];     // T = -Hess, T3 = -third derivatives (d2a/dx2), 10 unique components
};

template <int S>
__device__ __noinline__ double3 base2_force(const ssb_potential* P, Base2Shared<S>* sh, int stage, double x, double y, double z, double t) {
    const double X[3] = {x, y, z};
    double phi, g[3];
    Sym3 H;
    Sym3x3 T3;
    pot_eval<WANT_GRAD | WANT_HESS>(*P, X, t, phi, g, H);
    pot_third(*P, X, t, T3);
    sh->X[stage][0] = x; sh->X[stage][1] = y; sh->X[stage][2] = z;
    sh->T[stage][0] = -H.xx; sh->T[stage][1] = -H.yy; sh->T[stage][2] = -H.zz; sh->T[stage][3] = -H.xy; sh->T[stage][4] = -H.xz; sh->T[stage][5] = -H.yz;
    const double t3[10] = {T3.xxx, T3.yyy, T3.zzz, T3.xxy, T3.xxz, T3.xyy, T3.yyz, T3.xzz, T3.yzz, T3.xyz};
    for (int k = 0; k < 10; ++k) sh->T3[stage][k] = -t3[k];
    sh->t[stage] = t;
    return make_double3(-g[0], -g[1], -g[2]);
}
template <int S>
struct Base2Force {
    const ssb_potential* P; Base2Shared<S>* sh; double dir; int stage;
    __device__ __forceinline__ void operator()(const double X[3], double tau, double A[3]) {
        const double3 a = base2_force<S>(P, sh, stage, X[0], X[1], X[2], tau * dir);
        stage++;
        A[0] = a.x; A[1] = a.y; A[2] = a.z;
    }
};

struct Sub { double GM, rs, x0[3], v[3], t0, tw; int profile; };
__device__ __forceinline__ void load_sub(const ssb_subhalos& Sh, int j, Sub& s) {
    s.profile = Sh.profile; s.GM = Sh.G * Sh.m[j]; s.rs = Sh.rs[j]; s.t0 = Sh.t0[j]; s.tw = Sh.tw[j];
    for (int k = 0; k < 3; ++k) { s.x0[k] = Sh.x0[3 * j + k]; s.v[k] = Sh.v[3 * j + k]; }
}
__device__ __forceinline__ void matvec(const double* T, const double Q[3], double o[3]) {
    o[0] = T[0] * Q[0] + T[3] * Q[1] + T[4] * Q[2];
    o[1] = T[3] * Q[0] + T[1] * Q[1] + T[5] * Q[2];
    o[2] = T[4] * Q[0] + T[5] * Q[1] + T[2] * Q[2];
}

// (mass, second-order) pair: 6 components
template <int S>
struct PairForce {
    const Base2Shared<S>* sh; const Sub* s; int stage;
    __device__ __forceinline__ void at(int i, const double Q[6], double A[6]) const {
        const double* X = sh->X[i];
        const double dt = sh->t[i] - s->t0;
        double g[3] = {0, 0, 0}, hq[3] = {0, 0, 0};
        const double* Q1 = Q;
        const double* Q2 = Q + 3;
        if (fabs(dt) < s->tw) {
            double rel[3];
            for (int k = 0; k < 3; ++k) rel[k] = X[k] - fma(s->v[k], dt, s->x0[k]);
            const double r2 = fma(rel[0], rel[0], fma(rel[1], rel[1], rel[2] * rel[2]));
            double ph, q, w;
            profile_terms<WANT_GRAD | WANT_HESS>(s->profile, s->GM, s->rs, r2, ph, q, w);
            const double rq = rel[0] * Q1[0] + rel[1] * Q1[1] + rel[2] * Q1[2];
            for (int k = 0; k < 3; ++k) { g[k] = -q * rel[k]; hq[k] = -(q * Q1[k] + w * rel[k] * rq); }     // -Hess(Phi_sh) x1 (fields.py:305, 312)
        }
        double t1[3], t2[3];
        matvec(sh->T[i], Q1, t1);
        matvec(sh->T[i], Q2, t2);
        // sum_jk d2a[i][k][j] x1_k x1_j with the symmetric tensor stored as {xxx,yyy,zzz,xxy,xxz,xyy,yyz,xzz,yzz,xyz} (fields.py:310-311)
        const double* C = sh->T3[i];
        const double a = Q1[0], b = Q1[1], c = Q1[2];
        const double q0 = C[0] * a * a + C[5] * b * b + C[7] * c * c + 2.0 * (C[3] * a * b + C[4] * a * c + C[9] * b * c);
        const double q1 = C[3] * a * a + C[1] * b * b + C[8] * c * c + 2.0 * (C[5] * a * b + C[9] * a * c + C[6] * b * c);
        const double q2 = C[4] * a * a + C[6] * b * b + C[2] * c * c + 2.0 * (C[9] * a * b + C[7] * a * c + C[8] * b * c);
        A[0] = g[0] + t1[0]; A[1] = g[1] + t1[1]; A[2] = g[2] + t1[2];                                       // fields.py:307
        A[3] = t2[0] + q0 + hq[0]; A[4] = t2[1] + q1 + hq[1]; A[5] = t2[2] + q2 + hq[2];                     // fields.py:309-312
    }
    __device__ __forceinline__ void operator()(const double Q[6], double, double A[6]) { at(stage, Q, A); stage++; }
};
// radius block: 3 components (fields.py:314-315)
template <int S>
struct RadForce {
    const Base2Shared<S>* sh; const Sub* s; int stage;
    __device__ __forceinline__ void at(int i, const double Q[3], double A[3]) const {
        const double* X = sh->X[i];
        const double dt = sh->t[i] - s->t0;
        double g[3] = {0, 0, 0};
        if (fabs(dt) < s->tw) {
            double rel[3];
            for (int k = 0; k < 3; ++k) rel[k] = X[k] - fma(s->v[k], dt, s->x0[k]);
            const double r2 = fma(rel[0], rel[0], fma(rel[1], rel[1], rel[2] * rel[2]));
            double ph, q;
            profile_dradius(s->profile, s->GM, s->rs, r2, ph, q);
            for (int k = 0; k < 3; ++k) g[k] = -q * rel[k];
        }
        double t1[3];
        matvec(sh->T[i], Q, t1);
        A[0] = g[0] + t1[0]; A[1] = g[1] + t1[1]; A[2] = g[2] + t1[2];
    }
    __device__ __forceinline__ void operator()(const double Q[3], double, double A[3]) { at(stage, Q, A); stage++; }
};

__device__ __forceinline__ double block_sum2(double v, double* sred) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (l == 0) sred[w] = v;
    __syncthreads();
    double tot = 0.0;
    for (int i = 0; i < nw; ++i) tot += sred[i];
    return tot;
}

// state rows of the SoA buffer [18][n_sh]: pair item -> q = rows {0,1,2, 12,13,14}, p = rows {3,4,5, 15,16,17}; radius -> q = 6..8, p = 9..11
__device__ __forceinline__ int pair_qrow(int k) { return k < 3 ? k : 9 + k; }
__device__ __forceinline__ int pair_prow(int k) { return k < 3 ? 3 + k : 12 + k; }

// one sweep over all items.  MODE 0: step attempt (candidates -> nxt, squared scaled errors); MODE 1/2: HNW initial-step sums
template <int SOLVER, int MODE>
__device__ __forceinline__ void sweep2(const Base2Shared<Tab<SOLVER>::S>* sb, const ssb_subhalos& Sh, const double* cur, double* nxt, double dt,
                                       const CtrlDev& c, double h0, double& s0, double& s1, int& bad_local) {
    constexpr int S = Tab<SOLVER>::S;
    const int n = Sh.n;
    for (int idx = threadIdx.x; idx < 2 * n; idx += blockDim.x) {
        const bool rad = idx >= n;
        const int j = rad ? idx - n : idx;
        Sub sub; load_sub(Sh, j, sub);
        if (!rad) {
            double q[6], p[6];
            for (int k = 0; k < 6; ++k) { q[k] = cur[(size_t)pair_qrow(k) * n + j]; p[k] = cur[(size_t)pair_prow(k) * n + j]; }
            PairForce<S> f{sb, &sub, 1};
            if (MODE == 0) {
                double G[S][6], q1[6], p1[6], ex[6], ep[6];
                f.at(0, q, G[0]);
                rk_stages<SOLVER>(f, q, p, 0.0, dt, G);
                rk_candidate<SOLVER>(q, p, dt, G, q1, p1);
                f.at(S - 1, q1, G[S - 1]);
                rk_error<SOLVER>(p, dt, G, ex, ep);
                bool nanc = false;
                for (int k = 0; k < 6; ++k) { nanc |= isnan(q1[k]) | isnan(p1[k]); if (!isfinite(q1[k]) || !isfinite(p1[k])) bad_local = 1; }
                s0 += err_sq<6>(q, p, q1, p1, ex, ep, c.rtol, c.atol, nanc);
                for (int k = 0; k < 6; ++k) { nxt[(size_t)pair_qrow(k) * n + j] = q1[k]; nxt[(size_t)pair_prow(k) * n + j] = p1[k]; }
            } else {
                double G0[6], G1[6], qq[6];
                f.at(0, q, G0);
                if (MODE == 2) { for (int k = 0; k < 6; ++k) qq[k] = fma(h0, p[k], q[k]); f.at(1, qq, G1); }
                for (int k = 0; k < 6; ++k) {
                    const double sx = fma(c.rtol, fabs(q[k]), c.atol), sp = fma(c.rtol, fabs(p[k]), c.atol);
                    double r;
                    if (MODE == 1) { r = q[k] / sx; s0 = fma(r, r, s0); r = p[k] / sp; s0 = fma(r, r, s0); r = p[k] / sx; s1 = fma(r, r, s1); r = G0[k] / sp; s1 = fma(r, r, s1); }
                    else { r = (fma(h0, G0[k], p[k]) - p[k]) / sx; s0 = fma(r, r, s0); r = (G1[k] - G0[k]) / sp; s0 = fma(r, r, s0); }
                }
            }
        } else {
            double q[3], p[3];
            for (int k = 0; k < 3; ++k) { q[k] = cur[(size_t)(6 + k) * n + j]; p[k] = cur[(size_t)(9 + k) * n + j]; }
            RadForce<S> f{sb, &sub, 1};
            if (MODE == 0) {
                double G[S][3], q1[3], p1[3], ex[3], ep[3];
                f.at(0, q, G[0]);
                rk_stages<SOLVER>(f, q, p, 0.0, dt, G);
                rk_candidate<SOLVER>(q, p, dt, G, q1, p1);
                f.at(S - 1, q1, G[S - 1]);
                rk_error<SOLVER>(p, dt, G, ex, ep);
                bool nanc = false;
                for (int k = 0; k < 3; ++k) { nanc |= isnan(q1[k]) | isnan(p1[k]); if (!isfinite(q1[k]) || !isfinite(p1[k])) bad_local = 1; }
                s0 += err_sq<3>(q, p, q1, p1, ex, ep, c.rtol, c.atol, nanc);
                for (int k = 0; k < 3; ++k) { nxt[(size_t)(6 + k) * n + j] = q1[k]; nxt[(size_t)(9 + k) * n + j] = p1[k]; }
            } else {
                double G0[3], G1[3], qq[3];
                f.at(0, q, G0);
                if (MODE == 2) { for (int k = 0; k < 3; ++k) qq[k] = fma(h0, p[k], q[k]); f.at(1, qq, G1); }
                for (int k = 0; k < 3; ++k) {
                    const double sx = fma(c.rtol, fabs(q[k]), c.atol), sp = fma(c.rtol, fabs(p[k]), c.atol);
                    double r;
                    if (MODE == 1) { r = q[k] / sx; s0 = fma(r, r, s0); r = p[k] / sp; s0 = fma(r, r, s0); r = p[k] / sx; s1 = fma(r, r, s1); r = G0[k] / sp; s1 = fma(r, r, s1); }
                    else { r = (fma(h0, G0[k], p[k]) - p[k]) / sx; s0 = fma(r, r, s0); r = (G1[k] - G0[k]) / sp; s0 = fma(r, r, s0); }
                }
            }
        }
    }
}

template <int SOLVER>
__global__ void __launch_bounds__(SSB_R2_THREADS, 1) response2_kernel(const __grid_constant__ ssb_potential Pin, const ssb_subhalos Sh, const R2Args a) {
    typedef Tab<SOLVER> T;
    constexpr int S = T::S;
    __shared__ ssb_potential sP;
    __shared__ Base2Shared<S> sb;
    __shared__ double sred[32];
    __shared__ long long s_part;
    stage_potential(&sP, &Pin);
    const int tid = threadIdx.x, n_sh = Sh.n, ncomp = 6 + 18 * n_sh;
    double* buf0 = a.scratch + (size_t)blockIdx.x * 2 * 18 * n_sh;
    double* buf1 = buf0 + (size_t)18 * n_sh;
    const CtrlDev c = a.c;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_part = (long long)atomicAdd(a.counter, 1ULL);
        __syncthreads();
        const long long part = s_part;
        if (part >= a.N) break;
        const double t0_in = a.t0[part], t1_in = a.t1v ? a.t1v[part] : a.t1;
        const double dir = (t0_in < t1_in) ? 1.0 : -1.0;
        const double T0 = t0_in * dir, T1 = t1_in * dir;
        double* cur = buf0;
        double* nxt = buf1;
        for (int j = tid; j < n_sh; j += blockDim.x)
            for (int k = 0; k < 18; ++k) {
                double v = k < 12 ? (a.D0 ? a.D0[((size_t)part * n_sh + j) * 12 + k] : 0.0) : (a.E0 ? a.E0[((size_t)part * n_sh + j) * 6 + (k - 12)] : 0.0);
                const bool mom = (k >= 3 && k < 6) || (k >= 9 && k < 12) || k >= 15;
                cur[(size_t)k * n_sh + j] = mom ? v * dir : v;
            }
        double x[3] = {0, 0, 0}, p[3] = {0, 0, 0}, F[S][3];
        Base2Force<S> bforce{&sP, &sb, dir, 0};
        int status = 0, n_steps = 0, n_acc = 0, n_rej = 0, bad_local = 0;
        bool at_dtmin = false;
        double tprev = T0, tnext = T0;
        double d0s = 0.0, d1s = 0.0;
        if (tid == 0) {
            for (int k = 0; k < 3; ++k) { x[k] = a.w0[6 * part + k]; p[k] = dir * a.w0[6 * part + 3 + k]; }
            bforce.stage = 0;
            bforce(x, T0, F[0]);
            for (int k = 0; k < 3; ++k) {
                const double sx = fma(c.rtol, fabs(x[k]), c.atol), sp = fma(c.rtol, fabs(p[k]), c.atol);
                double q;
                q = x[k] / sx; d0s = fma(q, q, d0s); q = p[k] / sp; d0s = fma(q, q, d0s);
                q = p[k] / sx; d1s = fma(q, q, d1s); q = F[0][k] / sp; d1s = fma(q, q, d1s);
            }
        }
        __syncthreads();
        sweep2<SOLVER, 1>(&sb, Sh, cur, nxt, 0.0, c, 0.0, d0s, d1s, bad_local);
        const double d0 = sqrt(block_sum2(d0s, sred) / ncomp);
        const double d1 = sqrt(block_sum2(d1s, sred) / ncomp);
        const double h0 = hnw_h0(d0, d1);
        double d2s = 0.0, dummy = 0.0;
        if (tid == 0) {
            double X1[3], F1[3];
            for (int k = 0; k < 3; ++k) X1[k] = fma(h0, p[k], x[k]);
            bforce.stage = 1;
            bforce(X1, T0 + h0, F1);
            for (int k = 0; k < 3; ++k) {
                const double sx = fma(c.rtol, fabs(x[k]), c.atol), sp = fma(c.rtol, fabs(p[k]), c.atol);
                double q;
                q = (fma(h0, F[0][k], p[k]) - p[k]) / sx; d2s = fma(q, q, d2s);
                q = (F1[k] - F[0][k]) / sp; d2s = fma(q, q, d2s);
            }
        }
        __syncthreads();
        sweep2<SOLVER, 2>(&sb, Sh, cur, nxt, 0.0, c, h0, d2s, dummy, bad_local);
        const double d2 = sqrt(block_sum2(d2s, sred) / ncomp) / h0;
        {
            double h = fmin(hnw_h1<T::ORDER>(h0, d1, d2), c.dtmax);
            at_dtmin = h <= c.dtmin;
            h = fmax(h, c.dtmin);
            tnext = fmin(T0 + h, T1);
        }
        while (tprev < T1 && status == 0) {
            if (n_steps >= c.max_steps) { status = 1; break; }
            const double dt = tnext - tprev;
            double x1[3], p1[3], esq = 0.0;
            bad_local = 0;
            __syncthreads();
            if (tid == 0) {
                double ex[3], ep[3];
                bforce.stage = 1;
                rk_stages<SOLVER>(bforce, x, p, tprev, dt, F);
                rk_candidate<SOLVER>(x, p, dt, F, x1, p1);
                bforce.stage = S - 1;
                bforce(x1, tprev + T::c(S - 1) * dt, F[S - 1]);
                rk_error<SOLVER>(p, dt, F, ex, ep);
                bool nanc = false;
                for (int k = 0; k < 3; ++k) { nanc |= isnan(x1[k]) | isnan(p1[k]); if (!isfinite(x1[k]) || !isfinite(p1[k])) bad_local = 1; }
                esq = err_sq6(x, p, x1, p1, ex, ep, c.rtol, c.atol, nanc);
            }
            __syncthreads();
            sweep2<SOLVER, 0>(&sb, Sh, cur, nxt, dt, c, 0.0, esq, dummy, bad_local);
            const double err = sqrt(block_sum2(esq, sred) / ncomp);
            const int any_bad = __syncthreads_or(bad_local);
            double hn; bool bad;
            const bool keep = pid_update<T::ORDER>(err, dt, c, at_dtmin, hn, bad);
            n_steps++;
            if (bad) { status = 2; n_rej++; break; }
            if (keep) {
                n_acc++;
                if (any_bad) { status = 2; break; }
                double* tmp = cur; cur = nxt; nxt = tmp;
                if (tid == 0) {
                    for (int k = 0; k < 3; ++k) { x[k] = x1[k]; p[k] = p1[k]; F[0][k] = F[S - 1][k]; sb.X[0][k] = sb.X[S - 1][k]; }
                    for (int k = 0; k < 6; ++k) sb.T[0][k] = sb.T[S - 1][k];
                    for (int k = 0; k < 10; ++k) sb.T3[0][k] = sb.T3[S - 1][k];
                    sb.t[0] = sb.t[S - 1];
                }
                tprev = tnext;
            } else {
                n_rej++;
            }
            tprev = fmin(tprev, T1);
            double tn = tprev + hn;
            if (tn > T1 - 1e-10) tn = keep ? T1 : tprev + 0.5 * (T1 - tprev);
            tnext = tn;
        }
        __syncthreads();
        const bool ok = (status == 0) && (T0 < T1);
        const double inf = __longlong_as_double(0x7ff0000000000000LL);
        for (int j = tid; j < n_sh; j += blockDim.x)
            for (int k = 0; k < 18; ++k) {
                const bool mom = (k >= 3 && k < 6) || (k >= 9 && k < 12) || k >= 15;
                double v = cur[(size_t)k * n_sh + j];
                if (mom) v *= dir;
                if (k < 12) a.Dout[((size_t)part * n_sh + j) * 12 + k] = ok ? v : inf;
                else a.Eout[((size_t)part * n_sh + j) * 6 + (k - 12)] = ok ? v : inf;
            }
        if (tid == 0) {
            for (int k = 0; k < 3; ++k) { a.wout[6 * part + k] = ok ? x[k] : inf; a.wout[6 * part + 3 + k] = ok ? dir * p[k] : inf; }
            a.status[part] = status;
            a.nsteps[3 * part] = n_steps; a.nsteps[3 * part + 1] = n_acc; a.nsteps[3 * part + 2] = n_rej;
        }
    }
}

// RHS of the field at one state (fields.py:289-320): y = [w(6), D(n_sh,12), E(n_sh,6)] flattened in that order
__global__ void response2_term_kernel(const __grid_constant__ ssb_potential Pin, const ssb_subhalos Sh, double t, const double* y, double* dy) {
    __shared__ ssb_potential sP;
    __shared__ Base2Shared<1> sb;
    stage_potential(&sP, &Pin);
    if (threadIdx.x == 0) {
        const double3 acc = base2_force<1>(&sP, &sb, 0, y[0], y[1], y[2], t);
        if (blockIdx.x == 0) { dy[0] = y[3]; dy[1] = y[4]; dy[2] = y[5]; dy[3] = acc.x; dy[4] = acc.y; dy[5] = acc.z; }
    }
    __syncthreads();
    const int n = Sh.n;
    const double* D = y + 6;
    const double* E = y + 6 + 12 * (size_t)n;
    double* dD = dy + 6;
    double* dE = dy + 6 + 12 * (size_t)n;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        Sub sub; load_sub(Sh, j, sub);
        PairForce<1> pf{&sb, &sub, 0};
        RadForce<1> rf{&sb, &sub, 0};
        const double Q[6] = {D[12 * j], D[12 * j + 1], D[12 * j + 2], E[6 * j], E[6 * j + 1], E[6 * j + 2]};
        double A[6], B[3];
        pf.at(0, Q, A);
        const double Qr[3] = {D[12 * j + 6], D[12 * j + 7], D[12 * j + 8]};
        rf.at(0, Qr, B);
        for (int k = 0; k < 3; ++k) {
            dD[12 * j + k] = D[12 * j + 3 + k]; dD[12 * j + 3 + k] = A[k];
            dD[12 * j + 6 + k] = D[12 * j + 9 + k]; dD[12 * j + 9 + k] = B[k];
            dE[6 * j + k] = E[6 * j + 3 + k]; dE[6 * j + 3 + k] = A[3 + k];
        }
    }
}

#define SSB_R2_MAX_CTAS 160

extern "C" {

size_t ssb_second_order_scratch_bytes(int32_t n_sh) {
    return 256 + sizeof(double) * 2 * 18 * (size_t)(n_sh > 0 ? n_sh : 1) * SSB_R2_MAX_CTAS;
}

static int second_order_impl(const ssb_potential* pot_base, const ssb_subhalos* sh, int64_t N, const double* w0, const double* D0, const double* E0,
                             const double* t0, double t1, const double* t1v, ssb_ctrl ctrl, double* wout, double* Dout, double* Eout, int32_t* status,
                             int32_t* nsteps, void* scratch, size_t scratch_bytes, void* stream);
int ssb_second_order_response_f64(const ssb_potential* pot_base, const ssb_subhalos* sh, int64_t N, const double* w0, const double* D0, const double* E0,
                                  const double* t0, double t1, ssb_ctrl ctrl, double* wout, double* Dout, double* Eout, int32_t* status,
                                  int32_t* nsteps, void* scratch, size_t scratch_bytes, void* stream) {
    return second_order_impl(pot_base, sh, N, w0, D0, E0, t0, t1, nullptr, ctrl, wout, Dout, Eout, status, nsteps, scratch, scratch_bytes, stream);
}
int ssb_second_order_response_ends_f64(const ssb_potential* pot_base, const ssb_subhalos* sh, int64_t N, const double* w0, const double* D0, const double* E0,
                                       const double* t0, const double* t1, ssb_ctrl ctrl, double* wout, double* Dout, double* Eout, int32_t* status,
                                       int32_t* nsteps, void* scratch, size_t scratch_bytes, void* stream) {
    if (!t1) return ssb_set_error(SSB_ERR_ARG, "second_order_response_ends: NULL end times");
    return second_order_impl(pot_base, sh, N, w0, D0, E0, t0, 0.0, t1, ctrl, wout, Dout, Eout, status, nsteps, scratch, scratch_bytes, stream);
}
static int second_order_impl(const ssb_potential* pot_base, const ssb_subhalos* sh, int64_t N, const double* w0, const double* D0, const double* E0,
                             const double* t0, double t1, const double* t1v, ssb_ctrl ctrl, double* wout, double* Dout, double* Eout, int32_t* status,
                             int32_t* nsteps, void* scratch, size_t scratch_bytes, void* stream) {
    if (int e = ssb_validate_potential(pot_base)) return e;
    if (int e = ssb_validate_ctrl(ctrl)) return e;
    if (!sh || sh->n < 0 || sh->profile < SSB_PROFILE_PLUMMER || sh->profile > SSB_PROFILE_NFW) return ssb_set_error(SSB_ERR_ARG, "second_order_response: bad subhalo set");
    if (N < 0) return ssb_set_error(SSB_ERR_ARG, "second_order_response: negative N");
    if (N == 0) return 0;
    if (!w0 || !t0 || !wout || !status || !nsteps || !scratch || (sh->n > 0 && (!Dout || !Eout))) return ssb_set_error(SSB_ERR_ARG, "second_order_response: NULL array");
    if (scratch_bytes < ssb_second_order_scratch_bytes(sh->n)) return ssb_set_error(SSB_ERR_SCRATCH, "second_order_response: scratch too small");
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int grid = sms < SSB_R2_MAX_CTAS ? sms : SSB_R2_MAX_CTAS;
    if (N < grid) grid = (int)N;
    R2Args a;
    a.N = N; a.w0 = w0; a.D0 = D0; a.E0 = E0; a.t0 = t0; a.t1 = t1; a.t1v = t1v;
    a.c.rtol = ctrl.rtol; a.c.atol = ctrl.atol; a.c.dtmin = ctrl.dtmin; a.c.dtmax = ctrl.dtmax; a.c.max_steps = ctrl.max_steps;
    a.wout = wout; a.Dout = Dout; a.Eout = Eout; a.status = status; a.nsteps = nsteps;
    a.counter = (unsigned long long*)scratch;
    a.scratch = (double*)((char*)scratch + 256);
    CK(cudaMemsetAsync(scratch, 0, 256, st));
    if (ctrl.solver == 5) response2_kernel<5><<<grid, SSB_R2_THREADS, 0, st>>>(*pot_base, *sh, a);
    else response2_kernel<8><<<grid, SSB_R2_THREADS, 0, st>>>(*pot_base, *sh, a);
    CKL("response2_kernel");
    return 0;
}

int ssb_second_order_term_f64(const ssb_potential* pot_base, const ssb_subhalos* sh, double t, const double* y, double* dy, void* stream) {
    if (int e = ssb_validate_potential(pot_base)) return e;
    if (!sh || !y || !dy) return ssb_set_error(SSB_ERR_ARG, "second_order_term: NULL argument");
    int grid = (sh->n + 127) / 128; if (grid < 1) grid = 1;
    response2_term_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(*pot_base, *sh, t, y, dy);
    CKL("response2_term_kernel");
    return 0;
}

}  // extern "C"
