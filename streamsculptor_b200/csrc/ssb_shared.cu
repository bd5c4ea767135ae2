// K5: N tracers integrated as ONE ODE with a shared step-size controller.
//
// Reference: RestrictedNbody_generator.term (RestrictedNbody.py:93-106) integrated by fields.integrate_field
// (RestrictedNbody.py:131, fields.py:85-98): the whole (N,6) particle array is a single diffrax state, so there is one
// step sequence and the RMS error norm runs over all 6N components.  The field is the external potential plus the
// progenitor's monopole centred on its interpolated orbit = a potential program with a translating component.
//
// B200 mapping: one step ATTEMPT = one grid-wide kernel (thread per tracer, Nystrom stages in registers, block-reduced
// squared errors atomically added to a device accumulator) followed by a 1-thread controller kernel that decides
// accept/reject, flips the double buffer and writes the next (t, dt) into a device control block.  No host round trip per
// step: the host enqueues batches of attempts and polls a done flag between batches (an attempt at N >= 1e5 is >= tens of
// microseconds, a 1e7-tracer attempt ~13 ms).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "ssb_common.cuh"

using namespace ssb;

#define CK(call) do { int _e = ssb_cuda_check((call), #call); if (_e) return _e; } while (0)
#define CKL(what) do { ssb_count_launch(); int _e = ssb_cuda_check(cudaGetLastError(), what); if (_e) return _e; } while (0)

struct SharedCtl {            // device control block
    double tprev, tnext, T1, dir, acc, acc2, h0, d1;
    int which, at_dtmin, status, done, n_steps, n_acc, n_rej, bad;
};

__device__ __noinline__ double3 shared_accel(const ssb_potential* P, double x, double y, double z, double t) {
    const double X[3] = {x, y, z};
    double phi, g[3];
    Sym3 H;
    pot_eval<WANT_GRAD>(*P, X, t, phi, g, H);
    return make_double3(-g[0], -g[1], -g[2]);
}
struct SharedForce {
    const ssb_potential* P; double dir;
    __device__ __forceinline__ void operator()(const double X[3], double tau, double A[3]) const {
        const double3 a = shared_accel(P, X[0], X[1], X[2], tau * dir);
        A[0] = a.x; A[1] = a.y; A[2] = a.z;
    }
};

__device__ __forceinline__ void block_atomic_add(double v, double* target) {
    __shared__ double red[32];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    if (threadIdx.x == 0) { double tot = 0.0; for (int i = 0; i < nw; ++i) tot += red[i]; atomicAdd(target, tot); }
}

// state buffers: buf[which] = {x[3N] (AoS rows of 3), p[3N], F[3N]}
// init pass 1: p = dir*v, F0 = force(x, T0); sums d0^2 -> acc, d1^2 -> acc2
__global__ void shared_init1(const __grid_constant__ ssb_potential Pin, int64_t N, const double* w0, double* buf, SharedCtl* ctl, CtrlDev c) {
    __shared__ ssb_potential sP;
    stage_potential(&sP, &Pin);
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double d0 = 0.0, d1 = 0.0;
    if (i < N) {
        const double dir = ctl->dir, T0 = ctl->tprev;
        SharedForce f{&sP, dir};
        double x[3], p[3], F[3];
        for (int k = 0; k < 3; ++k) { x[k] = w0[6 * i + k]; p[k] = dir * w0[6 * i + 3 + k]; }
        f(x, T0, F);
        double* X = buf; double* Pm = buf + 3 * N; double* Fm = buf + 6 * N;
        for (int k = 0; k < 3; ++k) {
            X[3 * i + k] = x[k]; Pm[3 * i + k] = p[k]; Fm[3 * i + k] = F[k];
            const double sx = fma(c.rtol, fabs(x[k]), c.atol), sp = fma(c.rtol, fabs(p[k]), c.atol);
            double q;
            q = x[k] / sx; d0 = fma(q, q, d0); q = p[k] / sp; d0 = fma(q, q, d0);
            q = p[k] / sx; d1 = fma(q, q, d1); q = F[k] / sp; d1 = fma(q, q, d1);
        }
    }
    block_atomic_add(d0, &ctl->acc);
    block_atomic_add(d1, &ctl->acc2);
}
__global__ void shared_init_ctl1(int64_t N, SharedCtl* ctl) {
    const double d0 = sqrt(ctl->acc / (6.0 * N)), d1 = sqrt(ctl->acc2 / (6.0 * N));
    ctl->h0 = hnw_h0(d0, d1); ctl->d1 = d1; ctl->acc = 0.0; ctl->acc2 = 0.0;
}
// init pass 2: f1 = f(T0 + h0, y0 + h0 f0); sum ((f1 - f0)/sc)^2 -> acc
__global__ void shared_init2(const __grid_constant__ ssb_potential Pin, int64_t N, const double* buf, SharedCtl* ctl, CtrlDev c) {
    __shared__ ssb_potential sP;
    stage_potential(&sP, &Pin);
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double d2 = 0.0;
    if (i < N) {
        const double dir = ctl->dir, T0 = ctl->tprev, h0 = ctl->h0;
        SharedForce f{&sP, dir};
        const double* X = buf; const double* Pm = buf + 3 * N; const double* Fm = buf + 6 * N;
        double x[3], p[3], F0[3], X1[3], F1[3];
        for (int k = 0; k < 3; ++k) { x[k] = X[3 * i + k]; p[k] = Pm[3 * i + k]; F0[k] = Fm[3 * i + k]; X1[k] = fma(h0, p[k], x[k]); }
        f(X1, T0 + h0, F1);
        for (int k = 0; k < 3; ++k) {
            const double sx = fma(c.rtol, fabs(x[k]), c.atol), sp = fma(c.rtol, fabs(p[k]), c.atol);
            double q;
            q = (fma(h0, F0[k], p[k]) - p[k]) / sx; d2 = fma(q, q, d2);
            q = (F1[k] - F0[k]) / sp; d2 = fma(q, q, d2);
        }
    }
    block_atomic_add(d2, &ctl->acc);
}
template <int ORDER>
__global__ void shared_init_ctl2(int64_t N, SharedCtl* ctl, CtrlDev c) {
    const double d2 = sqrt(ctl->acc / (6.0 * N)) / ctl->h0;
    double h = fmin(hnw_h1<ORDER>(ctl->h0, ctl->d1, d2), c.dtmax);
    ctl->at_dtmin = h <= c.dtmin;
    h = fmax(h, c.dtmin);
    ctl->tnext = fmin(ctl->tprev + h, ctl->T1);
    ctl->acc = 0.0;
    ctl->done = !(ctl->tprev < ctl->T1);
}

// Force of the step-attempt kernel.  Every tracer is at the SAME time in every stage, so the time-dependent part of the program
// (track centres of translating components, frame accelerations) is evaluated once per stage and CTA into `frozen` and the
// tracers only subtract a centre; the static galaxy in front of the program (SIG != 0) is the fused inline signature of K1.
template <bool BARS>
__device__ __noinline__ double3 shared_accel_frozen(const ssb_potential* P, int first, double x, double y, double z, double t, const double* frozen,
                                                    const double* frozen_pc) {
    const double X[3] = {x, y, z};
    double phi, g[3];
    Sym3 H;
    pot_eval<WANT_GRAD, BARS>(*P, X, t, phi, g, H, first, frozen, frozen_pc);
    return make_double3(-g[0], -g[1], -g[2]);
}
// XM: 0 = general (extras through the interpreter call unless the inline Plummer applies), 1 = every extra is handled inline and there is no
// perturber set, 2 = the same with a frozen perturber set; in 1 and 2 the step loop contains NO call site (a never-taken call costs spills
// around it in every stage: -12 % on the orbit kernel, ssb_kernels.cu XS = 4)
template <int SIG, int XM>
struct SharedStageForce {
    static constexpr bool PS = XM == 2;
    const ssb_potential* P; const ssb_potential* Pc; double dir; const double* frozen; int stage; bool extra;
    const double* frozen_pc;      // centres of the perturber set at every stage time [S][3 SSB_PSET_FROZEN_MAX], or nullptr (no set / too large)
    int one;                      // > 0: index + 1 of an extra that is a Plummer sphere on a track (the progenitor of the restricted N-body field,
                                  // RestrictedNbody.py:93-106): its centre is frozen per stage, so the term is 20 inline instructions
    int nps;                      // > 0: the only other extra is a set of `nps` Plummer perturbers with frozen centres: inline loop, four
    const double* ps_par;         //      independent accumulators; ps_par = {GM[j], rs[j]} pairs in shared memory
    __device__ __forceinline__ void operator()(const double X[3], double tau, double A[3]) {
        const double* fz = frozen + stage * (6 * SSB_MAX_TRACK);
        const double* fp = frozen_pc ? frozen_pc + stage * (3 * SSB_PSET_FROZEN_MAX) : nullptr;
        if (SIG == SIG_GENERIC) {
            const double3 a = shared_accel_frozen<true>(P, 0, X[0], X[1], X[2], tau * dir, fz, fp);
            A[0] = a.x; A[1] = a.y; A[2] = a.z;
        } else {
            double g[3];
            fused_grad<SIG>(*Pc, X, g);
            if (one | (PS ? nps : 0)) {
                if (one) {                                      // same operations as pot_eval's SSB_PLUMMER case
                    const ssb_component& c = P->comp[one - 1];
                    const double* ctr = fz + 6 * c.track;
                    const double xs[3] = {X[0] - ctr[0], X[1] - ctr[1], X[2] - ctr[2]};
                    const double r2 = fma(xs[0], xs[0], fma(xs[1], xs[1], xs[2] * xs[2]));
                    double phi, q, w;
                    plummer_terms<WANT_GRAD>(c.p[0], c.p[1], r2, phi, q, w);
                    g[0] = fma(q, xs[0], g[0]); g[1] = fma(q, xs[1], g[1]); g[2] = fma(q, xs[2], g[2]);
                }
                if (PS && nps) {                                // sum over the perturbers: lane-independent work, four accumulator sets for ILP
                    double a0[3] = {0, 0, 0}, a1[3] = {0, 0, 0}, a2[3] = {0, 0, 0}, a3[3] = {0, 0, 0};
                    auto term = [&](int j, double (&acc)[3]) {
                        const double xs[3] = {X[0] - fp[3 * j], X[1] - fp[3 * j + 1], X[2] - fp[3 * j + 2]};
                        const double r2 = fma(xs[0], xs[0], fma(xs[1], xs[1], xs[2] * xs[2]));
                        double phi, q, w;
                        plummer_terms<WANT_GRAD>(ps_par[2 * j], ps_par[2 * j + 1], r2, phi, q, w);
                        acc[0] = fma(q, xs[0], acc[0]); acc[1] = fma(q, xs[1], acc[1]); acc[2] = fma(q, xs[2], acc[2]);
                    };
                    int j = 0;
                    for (; j + 4 <= nps; j += 4) { term(j, a0); term(j + 1, a1); term(j + 2, a2); term(j + 3, a3); }
                    for (; j < nps; ++j) term(j, a0);
#pragma unroll
                    for (int k = 0; k < 3; ++k) g[k] += (a0[k] + a1[k]) + (a2[k] + a3[k]);
                }
                A[0] = -g[0]; A[1] = -g[1]; A[2] = -g[2];
                stage++;
                return;
            }
            A[0] = -g[0]; A[1] = -g[1]; A[2] = -g[2];
            if (XM == 0 && extra) {
                const double3 a = shared_accel_frozen<false>(P, SigInfo<SIG>::NF, X[0], X[1], X[2], tau * dir, fz, fp);
                A[0] += a.x; A[1] += a.y; A[2] += a.z;
            }
        }
        stage++;
    }
};

// one step attempt for every tracer
#ifndef SSB_SHARED_MIN_BLOCKS
#define SSB_SHARED_MIN_BLOCKS 3          // CTAs per SM the register budget is set for (3: 168 registers, 2: 240 and no spills)
#endif
template <int SOLVER, int SIG, int XM>
__global__ void __launch_bounds__(128, SSB_SHARED_MIN_BLOCKS) shared_attempt(const __grid_constant__ ssb_potential Pin, int64_t N, double* buf0, double* buf1, SharedCtl* ctl, CtrlDev c) {
    typedef Tab<SOLVER> T;
    constexpr int S = T::S;
    __shared__ ssb_potential sP;
    __shared__ double s_frozen[S][6 * SSB_MAX_TRACK];
    __shared__ double s_ppar[2 * SSB_PSET_FROZEN_MAX];            // {G m, r_s} of the set's perturbers
    __shared__ double s_pc[S][3 * SSB_PSET_FROZEN_MAX];           // perturber-set centres at the stage times (BASELINE config 5: 100 moving perturbers)
    stage_potential(&sP, &Pin);
    logtab_init();
    if (ctl->done) return;
    const double tprev = ctl->tprev, dt = ctl->tnext - ctl->tprev, dir = ctl->dir;
    // tracks at the S stage times of this attempt (the same expressions rk_stages / the last-stage call below hand to the force)
    for (int q = threadIdx.x; q < S * sP.n_track; q += blockDim.x) {
        const int st = q / sP.n_track, k = q - st * sP.n_track;
        const double tau = tprev + T::c(st) * dt;
        double cv[3], dv[3];
        track_eval<true>(sP.track[k], tau * dir, cv, dv);
#pragma unroll
        for (int m = 0; m < 3; ++m) { s_frozen[st][6 * k + m] = cv[m]; s_frozen[st][6 * k + 3 + m] = dv[m]; }
    }
    const bool pc_frozen = sP.n_pset == 1 && sP.pset[0].n <= SSB_PSET_FROZEN_MAX;
    if (pc_frozen) {              // every centre of the set at every stage time, once per CTA: (stage, perturber) pairs dealt out over the threads
        const int np = sP.pset[0].n;
        for (int j = threadIdx.x; j < np; j += blockDim.x) { s_ppar[2 * j] = sP.pset[0].GM[j]; s_ppar[2 * j + 1] = sP.pset[0].rs[j]; }
        for (int q = threadIdx.x; q < S * np; q += blockDim.x) {
            const int st = q / np, j = q - st * np;
            pset_centres(sP.pset[0], (tprev + T::c(st) * dt) * dir, j, j + 1, &s_pc[st][3 * j]);
        }
    }
    __syncthreads();
    const double* cur = ctl->which ? buf1 : buf0;
    double* nxt = ctl->which ? buf0 : buf1;
    double esq = 0.0;
    int bad = 0;
    // extras the force functor evaluates inline: at most one Plummer sphere on a track and (PS) one frozen set of Plummer perturbers,
    // nothing else (shared_inline_extras on the host decides PS with the same rule)
    int one_plummer = 0, n_inline_ps = 0;
    if (SIG != SIG_GENERIC) {
        int n_pl = 0, i_pl = 0, n_set = 0, n_other = 0;
        for (int ic = SigInfo<SIG>::NF; ic < sP.n_comp; ++ic) {
            const ssb_component& cx = sP.comp[ic];
            if (cx.type == SSB_PLUMMER && cx.track >= 0 && cx.growth == 0) { n_pl++; i_pl = ic + 1; }
            else if (cx.type == SSB_PERTURBERS && pc_frozen && sP.pset[0].profile == SSB_PROFILE_PLUMMER) n_set++;
            else n_other++;
        }
        if (n_other == 0 && n_pl <= 1 && n_set <= (XM == 2 ? 1 : 0) && n_pl + n_set > 0) { one_plummer = n_pl ? i_pl : 0; n_inline_ps = n_set ? sP.pset[0].n : 0; }
    }
    // persistent CTAs: the prologue above (program, log table, 14 x tracks, perturber centres) and the atomic below are paid once per CTA
    // and attempt, not once per 128 tracers (1e7 tracers: 78 k CTAs before, each with a prologue as long as its step)
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        SharedStageForce<SIG, XM> f{&sP, &Pin, dir, &s_frozen[0][0], 1, Pin.n_comp > SigInfo<SIG>::NF, pc_frozen ? &s_pc[0][0] : nullptr, one_plummer, n_inline_ps, s_ppar};
        double x[3], p[3], F[S][3], x1[3], p1[3], ex[3], ep[3];
        for (int k = 0; k < 3; ++k) { x[k] = cur[3 * i + k]; p[k] = cur[3 * N + 3 * i + k]; F[0][k] = cur[6 * N + 3 * i + k]; }
        rk_stages<SOLVER>(f, x, p, tprev, dt, F);
        rk_candidate<SOLVER>(x, p, dt, F, x1, p1);
        f.stage = S - 1;
        f(x1, tprev + T::c(S - 1) * dt, F[S - 1]);
        rk_error<SOLVER>(p, dt, F, ex, ep);
        bool nanc = false;
        for (int k = 0; k < 3; ++k) { nanc |= isnan(x1[k]) | isnan(p1[k]); if (!isfinite(x1[k]) || !isfinite(p1[k])) bad = 1; }
        esq += err_sq6(x, p, x1, p1, ex, ep, c.rtol, c.atol, nanc);
        for (int k = 0; k < 3; ++k) { nxt[3 * i + k] = x1[k]; nxt[3 * N + 3 * i + k] = p1[k]; nxt[6 * N + 3 * i + k] = F[S - 1][k]; }
    }
    if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(&ctl->bad, 1);
    block_atomic_add(esq, &ctl->acc);
}
template <int ORDER>
__global__ void shared_control(int64_t N, SharedCtl* ctl, CtrlDev c) {
    if (ctl->done) return;
    if (ctl->n_steps >= c.max_steps) { ctl->status = 1; ctl->done = 1; return; }
    const double dt = ctl->tnext - ctl->tprev;
    const double err = sqrt(ctl->acc / (6.0 * N));
    bool at_dtmin = ctl->at_dtmin != 0, bad;
    double hn;
    const bool keep = pid_update<ORDER>(err, dt, c, at_dtmin, hn, bad);
    ctl->at_dtmin = at_dtmin;
    ctl->n_steps++;
    ctl->acc = 0.0;
    if (bad) { ctl->status = 2; ctl->n_rej++; ctl->done = 1; return; }
    if (keep) {
        ctl->n_acc++;
        if (ctl->bad) { ctl->status = 2; ctl->done = 1; return; }
        ctl->which ^= 1;
        ctl->tprev = ctl->tnext;
    } else {
        ctl->n_rej++;
        ctl->bad = 0;
    }
    ctl->tprev = fmin(ctl->tprev, ctl->T1);
    double tn = ctl->tprev + hn;
    if (tn > ctl->T1 - 1e-10) tn = keep ? ctl->T1 : ctl->tprev + 0.5 * (ctl->T1 - ctl->tprev);
    ctl->tnext = tn;
    if (!(ctl->tprev < ctl->T1)) ctl->done = 1;
}
__global__ void shared_finish(int64_t N, const double* buf0, const double* buf1, const SharedCtl* ctl, double* wout, int32_t* status, int32_t* nsteps) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const double* cur = ctl->which ? buf1 : buf0;
    const bool ok = ctl->status == 0 && !(ctl->tprev < ctl->T1) && ctl->n_acc > 0;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    if (i < N) for (int k = 0; k < 3; ++k) { wout[6 * i + k] = ok ? cur[3 * i + k] : inf; wout[6 * i + 3 + k] = ok ? ctl->dir * cur[3 * N + 3 * i + k] : inf; }
    if (i == 0) { status[0] = ctl->status; nsteps[0] = ctl->n_steps; nsteps[1] = ctl->n_acc; nsteps[2] = ctl->n_rej; }
}

// =====================================================================================================================
// K6: Nbody_field (fields.py:115-155): N live bodies with softened all-pairs gravity (+ external potential) as ONE ODE.
// Few-body problems (3-body tests, MW-LMC, ~100 live perturbers): pure latency, so ONE persistent CTA steps the system;
// thread b owns body b (strided for N > blockDim), stage positions are exchanged through shared memory, the Nystrom force
// stages live in an L1/L2-resident scratch, the shared controller is evaluated redundantly by every thread from a
// block-wide error sum.  SaveAt(ts) rows are interpolated inside the accepted step that covers them.
// =====================================================================================================================
#define SSB_NBODY_MAX 1024
#define SSB_NBODY_THREADS 256

__device__ __forceinline__ double block_sum_all(double v, double* red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    double tot = 0.0;
    for (int i = 0; i < (SSB_NBODY_THREADS >> 5); ++i) tot += red[i];
    return tot;
}

struct NbodyArgs {
    int N, M, has_ext;
    double G, eps2, t0, t1;
    const double* w0; const double* masses; const double* ts;
    double* scratch; double* ys; int32_t* status; int32_t* nsteps;
};

// accelerations of all bodies at stage positions sX (shared) -> Fout[3N] (global)
__device__ __noinline__ void nbody_force_all(const ssb_potential* sP, const NbodyArgs& a, const double* sX, const double* sM, double treal, double* Fout) {
    for (int b = threadIdx.x; b < a.N; b += blockDim.x) {
        const double X[3] = {sX[3 * b], sX[3 * b + 1], sX[3 * b + 2]};
        double ax = 0.0, ay = 0.0, az = 0.0;
        for (int j = 0; j < a.N; ++j) {
            if (j == b) continue;
            const double dx = sX[3 * j] - X[0], dy = sX[3 * j + 1] - X[1], dz = sX[3 * j + 2] - X[2];
            const double d2 = fma(dx, dx, fma(dy, dy, fma(dz, dz, a.eps2)));
            const double inv = 1.0 / sqrt(d2);
            const double w = a.G * sM[j] * inv * inv * inv;          // (G m_b m_j / d^2)(1/d) / m_b
            ax = fma(w, dx, ax); ay = fma(w, dy, ay); az = fma(w, dz, az);
        }
        if (a.has_ext) {
            double phi, g[3]; Sym3 H;
            pot_eval<WANT_GRAD>(*sP, X, treal, phi, g, H);
            ax -= g[0]; ay -= g[1]; az -= g[2];
        }
        Fout[3 * b] = ax; Fout[3 * b + 1] = ay; Fout[3 * b + 2] = az;
    }
}

template <int SOLVER>
__global__ void __launch_bounds__(SSB_NBODY_THREADS) nbody_kernel(const __grid_constant__ ssb_potential Pin, NbodyArgs a, CtrlDev c) {
    typedef Tab<SOLVER> T;
    constexpr int S = T::S;
    __shared__ ssb_potential sP;
    __shared__ double sX[3 * SSB_NBODY_MAX];
    __shared__ double sM[SSB_NBODY_MAX];
    __shared__ double red[SSB_NBODY_THREADS >> 5];
    stage_potential(&sP, &Pin);
    const int N = a.N, n3 = 3 * a.N;
    const double dir = (a.t0 < a.t1) ? 1.0 : -1.0, T0 = a.t0 * dir, T1 = a.t1 * dir;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    double* x = a.scratch; double* p = x + n3; double* x1 = p + n3; double* p1 = x1 + n3; double* F = p1 + n3;   // F[l] = F + l*n3
    for (int64_t i = threadIdx.x; i < (int64_t)a.M * N * 6; i += blockDim.x) a.ys[i] = inf;
    for (int b = threadIdx.x; b < N; b += blockDim.x) {
        sM[b] = a.masses[b];
        for (int k = 0; k < 3; ++k) { x[3 * b + k] = a.w0[6 * b + k]; p[3 * b + k] = dir * a.w0[6 * b + 3 + k]; sX[3 * b + k] = x[3 * b + k]; }
    }
    __syncthreads();
    nbody_force_all(&sP, a, sX, sM, T0 * dir, F);
    // ---- Hairer-Norsett-Wanner initial step over the whole 6N state ----
    double h;
    {
        double s0 = 0.0, s1 = 0.0;
        for (int b = threadIdx.x; b < N; b += blockDim.x)
            for (int k = 0; k < 3; ++k) {
                const double xv = x[3 * b + k], pv = p[3 * b + k], fv = F[3 * b + k];
                const double sx = fma(c.rtol, fabs(xv), c.atol), sp = fma(c.rtol, fabs(pv), c.atol);
                double q;
                q = xv / sx; s0 = fma(q, q, s0); q = pv / sp; s0 = fma(q, q, s0);
                q = pv / sx; s1 = fma(q, q, s1); q = fv / sp; s1 = fma(q, q, s1);
            }
        const double d0 = sqrt(block_sum_all(s0, red) / (6.0 * N)), d1 = sqrt(block_sum_all(s1, red) / (6.0 * N));
        const double h0 = hnw_h0(d0, d1);
        for (int b = threadIdx.x; b < N; b += blockDim.x)
            for (int k = 0; k < 3; ++k) sX[3 * b + k] = fma(h0, p[3 * b + k], x[3 * b + k]);
        __syncthreads();
        nbody_force_all(&sP, a, sX, sM, (T0 + h0) * dir, F + n3);
        double s2 = 0.0;
        for (int b = threadIdx.x; b < N; b += blockDim.x)
            for (int k = 0; k < 3; ++k) {
                const double xv = x[3 * b + k], pv = p[3 * b + k];
                const double sx = fma(c.rtol, fabs(xv), c.atol), sp = fma(c.rtol, fabs(pv), c.atol);
                double q;
                q = (fma(h0, F[3 * b + k], pv) - pv) / sx; s2 = fma(q, q, s2);
                q = (F[n3 + 3 * b + k] - F[3 * b + k]) / sp; s2 = fma(q, q, s2);
            }
        const double d2 = sqrt(block_sum_all(s2, red) / (6.0 * N)) / h0;
        h = fmin(hnw_h1<T::ORDER>(h0, d1, d2), c.dtmax);
    }
    bool at_dtmin = h <= c.dtmin;
    h = fmax(h, c.dtmin);
    double tprev = T0, tnext = fmin(T0 + h, T1);
    int n_steps = 0, n_acc = 0, n_rej = 0, status = 0, save_idx = 0;

    while (tprev < T1) {
        if (n_steps >= c.max_steps) { status = 1; break; }
        const double dt = tnext - tprev;
        for (int i = 1; i < S; ++i) {
            __syncthreads();                                             // everyone is done reading the previous sX
            for (int b = threadIdx.x; b < N; b += blockDim.x)
                for (int k = 0; k < 3; ++k) {
                    double ax = 0.0;
                    for (int l = 0; l < i; ++l) ax = fma(T::aa(i, l), F[l * n3 + 3 * b + k], ax);
                    const double X = fma(dt, fma(dt, ax, T::rs(i) * p[3 * b + k]), x[3 * b + k]);
                    sX[3 * b + k] = X;
                    if (i == S - 1) {                                    // last row == b: candidate state
                        double ap = 0.0;
                        for (int l = 0; l < i; ++l) ap = fma(T::a(i, l), F[l * n3 + 3 * b + k], ap);
                        x1[3 * b + k] = X; p1[3 * b + k] = fma(dt, ap, p[3 * b + k]);
                    }
                }
            __syncthreads();
            nbody_force_all(&sP, a, sX, sM, (tprev + T::c(i) * dt) * dir, F + i * n3);
        }
        // ---- error norm over all 6N components (scale uses y0 if ANY candidate component is NaN) ----
        int fl = 0;
        for (int b = threadIdx.x; b < N; b += blockDim.x)
            for (int k = 0; k < 3; ++k) {
                const double a1 = x1[3 * b + k], b1 = p1[3 * b + k];
                if (isnan(a1) || isnan(b1)) fl |= 1;
                if (!isfinite(a1) || !isfinite(b1)) fl |= 2;
            }
        const int anynan = __syncthreads_or(fl & 1), anybad = __syncthreads_or(fl & 2);
        double es = 0.0;
        for (int b = threadIdx.x; b < N; b += blockDim.x)
            for (int k = 0; k < 3; ++k) {
                double bx = 0.0, bp = 0.0;
                for (int l = 0; l < S; ++l) { const double f = F[l * n3 + 3 * b + k]; bx = fma(T::ea(l), f, bx); bp = fma(T::e(l), f, bp); }
                const double xv = x[3 * b + k], pv = p[3 * b + k];
                const double ex = dt * fma(dt, bx, T::esum() * pv), ep = dt * bp;
                const double xc = anynan ? xv : x1[3 * b + k], pc = anynan ? pv : p1[3 * b + k];
                const double sx = fma(c.rtol, fmax(fabs(xv), fabs(xc)), c.atol), sp = fma(c.rtol, fmax(fabs(pv), fabs(pc)), c.atol);
                const double qx = ex / sx, qp = ep / sp;
                es = fma(qx, qx, es); es = fma(qp, qp, es);
            }
        const double err = sqrt(block_sum_all(es, red) / (6.0 * N));
        double hn; bool bad;
        const bool keep = pid_update<T::ORDER>(err, dt, c, at_dtmin, hn, bad);
        n_steps++;
        if (bad) { status = 2; n_rej++; break; }
        if (keep) {
            n_acc++;
            if (anybad) { status = 2; break; }
            while (save_idx < a.M && a.ts[save_idx] * dir <= tnext) {
                const double tq = a.ts[save_idx] * dir, theta = (tq - tprev) / dt;
                for (int b = threadIdx.x; b < N; b += blockDim.x) {
                    double* o = a.ys + ((int64_t)save_idx * N + b) * 6;
                    if (tq == tnext) { for (int k = 0; k < 3; ++k) { o[k] = x1[3 * b + k]; o[3 + k] = dir * p1[3 * b + k]; } }
                    else {
                        double Fl[S][3], xb[3], pb[3], x1b[3], p1b[3], xo[3], po[3];
                        for (int k = 0; k < 3; ++k) { xb[k] = x[3 * b + k]; pb[k] = p[3 * b + k]; x1b[k] = x1[3 * b + k]; p1b[k] = p1[3 * b + k]; }
                        for (int l = 0; l < S; ++l) for (int k = 0; k < 3; ++k) Fl[l][k] = F[l * n3 + 3 * b + k];
                        rk_dense<SOLVER>(xb, pb, x1b, p1b, dt, Fl, theta, xo, po);
                        for (int k = 0; k < 3; ++k) { o[k] = xo[k]; o[3 + k] = dir * po[k]; }
                    }
                }
                save_idx++;
            }
            for (int b = threadIdx.x; b < N; b += blockDim.x)
                for (int k = 0; k < 3; ++k) { x[3 * b + k] = x1[3 * b + k]; p[3 * b + k] = p1[3 * b + k]; F[3 * b + k] = F[(S - 1) * n3 + 3 * b + k]; }
            tprev = tnext;
        } else {
            n_rej++;
        }
        tprev = fmin(tprev, T1);
        double tn = tprev + hn;
        if (tn > T1 - 1e-10) tn = keep ? T1 : tprev + 0.5 * (T1 - tprev);
        tnext = tn;
    }
    if (threadIdx.x == 0) { a.status[0] = status; a.nsteps[0] = n_steps; a.nsteps[1] = n_acc; a.nsteps[2] = n_rej; }
}

// Nbody_field.term at one state (unit tests / facade .term): dy[N,6]
__global__ void __launch_bounds__(SSB_NBODY_THREADS) nbody_term_kernel(const __grid_constant__ ssb_potential Pin, NbodyArgs a, double t, const double* y, double* dy) {
    __shared__ ssb_potential sP;
    __shared__ double sX[3 * SSB_NBODY_MAX];
    __shared__ double sM[SSB_NBODY_MAX];
    stage_potential(&sP, &Pin);
    for (int b = threadIdx.x; b < a.N; b += blockDim.x) { sM[b] = a.masses[b]; for (int k = 0; k < 3; ++k) sX[3 * b + k] = y[6 * b + k]; }
    __syncthreads();
    nbody_force_all(&sP, a, sX, sM, t, a.scratch);
    __syncthreads();
    for (int b = threadIdx.x; b < a.N; b += blockDim.x)
        for (int k = 0; k < 3; ++k) { dy[6 * b + k] = y[6 * b + 3 + k]; dy[6 * b + 3 + k] = a.scratch[3 * b + k]; }
}

extern "C" {

size_t ssb_shared_scratch_bytes(int64_t N) { return 512 + sizeof(double) * 18 * (size_t)(N > 0 ? N : 1); }

int ssb_shared_step_orbits_f64(const ssb_potential* pot, int64_t N, const double* w0, double t0, double t1, ssb_ctrl ctrl, double* wout,
                               int32_t* status, int32_t* nsteps, void* scratch, size_t scratch_bytes, void* stream) {
    if (int e = ssb_validate_potential(pot)) return e;
    if (int e = ssb_validate_ctrl(ctrl)) return e;
    if (N < 0) return ssb_set_error(SSB_ERR_ARG, "shared_step_orbits: negative N");
    if (N == 0) return 0;
    if (!w0 || !wout || !status || !nsteps || !scratch) return ssb_set_error(SSB_ERR_ARG, "shared_step_orbits: NULL array");
    if (scratch_bytes < ssb_shared_scratch_bytes(N)) return ssb_set_error(SSB_ERR_SCRATCH, "shared_step_orbits: scratch too small");
    cudaStream_t st = (cudaStream_t)stream;
    SharedCtl* ctl = (SharedCtl*)scratch;
    double* buf0 = (double*)((char*)scratch + 512);
    double* buf1 = buf0 + 9 * N;
    CtrlDev c; c.rtol = ctrl.rtol; c.atol = ctrl.atol; c.dtmin = ctrl.dtmin; c.dtmax = ctrl.dtmax; c.max_steps = ctrl.max_steps;
    SharedCtl h;
    memset(&h, 0, sizeof(h));
    h.dir = (t0 < t1) ? 1.0 : -1.0; h.tprev = t0 * h.dir; h.tnext = h.tprev; h.T1 = t1 * h.dir;
    CK(cudaMemcpyAsync(ctl, &h, sizeof(h), cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));                 // h lives on the host stack
    const unsigned grid = (unsigned)((N + 127) / 128);
    shared_init1<<<grid, 128, 0, st>>>(*pot, N, w0, buf0, ctl, c);
    shared_init_ctl1<<<1, 1, 0, st>>>(N, ctl);
    shared_init2<<<grid, 128, 0, st>>>(*pot, N, buf0, ctl, c);
    if (ctrl.solver == 5) shared_init_ctl2<5><<<1, 1, 0, st>>>(N, ctl, c); else shared_init_ctl2<8><<<1, 1, 0, st>>>(N, ctl, c);
    CKL("shared_init");
    // batches of attempts; poll the done flag between batches
    const int batch = N >= 1000000 ? 4 : 32;
    ssb_potential pc;
    const int sig = ssb_canonicalize(pot, &pc);          // fused static galaxy in front (as in K1), the rest through the interpreter
    // persistent grid for the attempts: SSB_SHARED_MIN_BLOCKS 128-thread CTAs per SM (44 KB of shared memory each), every CTA loops over tiles
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned grid_att = grid < (unsigned)(SSB_SHARED_MIN_BLOCKS * sms) ? grid : (unsigned)(SSB_SHARED_MIN_BLOCKS * sms);
    // which extras-handling variant of the attempt kernel (SharedStageForce): 0 general, 1 / 2 everything inline (without / with a frozen set of
    // Plummer perturbers) - the same rule as in the kernel
    int xm = 0;
    if (sig != SIG_GENERIC) {
        const int nf = sig == SIG_N ? 1 : sig == SIG_NHM ? 3 : 4;
        const bool set_ok = pot->n_pset == 1 && pot->pset[0].n <= SSB_PSET_FROZEN_MAX && pot->pset[0].profile == SSB_PROFILE_PLUMMER;
        int n_pl = 0, n_set = 0, n_other = 0;
        for (int ic = nf; ic < pc.n_comp; ++ic) {
            const ssb_component& cx = pc.comp[ic];
            if (cx.type == SSB_PLUMMER && cx.track >= 0 && cx.growth == 0) n_pl++;
            else if (cx.type == SSB_PERTURBERS && set_ok) n_set++;
            else n_other++;
        }
        if (n_other == 0 && n_pl <= 1 && n_set <= 1) xm = n_set ? 2 : 1;
    }
#define SSB_LAUNCH_ATT(S, SG) do { if (xm == 2) shared_attempt<S, SG, 2><<<grid_att, 128, 0, st>>>(sig == SG ? pc : *pot, N, buf0, buf1, ctl, c); \
        else if (xm == 1) shared_attempt<S, SG, 1><<<grid_att, 128, 0, st>>>(sig == SG ? pc : *pot, N, buf0, buf1, ctl, c); \
        else shared_attempt<S, SG, 0><<<grid_att, 128, 0, st>>>(sig == SG ? pc : *pot, N, buf0, buf1, ctl, c); } while (0)
#define SSB_LAUNCH_ATT_SIG(S) do { switch (sig) { case SIG_NHM: SSB_LAUNCH_ATT(S, SIG_NHM); break; case SIG_NHHM: SSB_LAUNCH_ATT(S, SIG_NHHM); break; \
        default: SSB_LAUNCH_ATT(S, SIG_GENERIC); } } while (0)
    for (int64_t launched = 0; launched <= (int64_t)ctrl.max_steps + batch;) {
        for (int b = 0; b < batch; ++b) {
            if (ctrl.solver == 5) { SSB_LAUNCH_ATT_SIG(5); shared_control<5><<<1, 1, 0, st>>>(N, ctl, c); }
            else { SSB_LAUNCH_ATT_SIG(8); shared_control<8><<<1, 1, 0, st>>>(N, ctl, c); }
        }
        launched += batch;
        CKL("shared_attempt");
        int done = 0;
        CK(cudaMemcpyAsync(&done, &ctl->done, sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (done) break;
    }
    shared_finish<<<grid, 128, 0, st>>>(N, buf0, buf1, ctl, wout, status, nsteps);
    CKL("shared_finish");
    return 0;
}

size_t ssb_nbody_scratch_bytes(int32_t N) { return sizeof(double) * 3 * (size_t)(N > 0 ? N : 1) * (4 + 14); }

static int nbody_fill(NbodyArgs& a, ssb_potential& P, const ssb_potential* ext, int32_t N, const double* masses, double G, double eps) {
    memset(&P, 0, sizeof(P));
    a.has_ext = 0;
    if (ext && ext->n_comp > 0) { if (int e = ssb_validate_potential(ext)) return e; P = *ext; a.has_ext = 1; }
    if (N <= 0 || N > SSB_NBODY_MAX) return ssb_set_error(SSB_ERR_ARG, "nbody: need 1 <= N <= 1024 live bodies");
    if (!masses) return ssb_set_error(SSB_ERR_ARG, "nbody: NULL masses");
    a.N = N; a.masses = masses; a.G = G; a.eps2 = eps * eps;
    return 0;
}

int ssb_nbody_integrate_f64(const ssb_potential* ext, int32_t N, const double* masses, double G, double eps, const double* w0, double t0, double t1,
                            const double* ts, int32_t M, ssb_ctrl ctrl, double* ys, int32_t* status, int32_t* nsteps, void* scratch,
                            size_t scratch_bytes, void* stream) {
    if (int e = ssb_validate_ctrl(ctrl)) return e;
    NbodyArgs a; ssb_potential P;
    if (int e = nbody_fill(a, P, ext, N, masses, G, eps)) return e;
    if (!w0 || !ts || M <= 0 || !ys || !status || !nsteps || !scratch) return ssb_set_error(SSB_ERR_ARG, "nbody: NULL array or M <= 0");
    if (scratch_bytes < ssb_nbody_scratch_bytes(N)) return ssb_set_error(SSB_ERR_SCRATCH, "nbody: scratch too small");
    a.M = M; a.t0 = t0; a.t1 = t1; a.w0 = w0; a.ts = ts; a.scratch = (double*)scratch; a.ys = ys; a.status = status; a.nsteps = nsteps;
    CtrlDev c; c.rtol = ctrl.rtol; c.atol = ctrl.atol; c.dtmin = ctrl.dtmin; c.dtmax = ctrl.dtmax; c.max_steps = ctrl.max_steps;
    cudaStream_t st = (cudaStream_t)stream;
    if (ctrl.solver == 5) nbody_kernel<5><<<1, SSB_NBODY_THREADS, 0, st>>>(P, a, c); else nbody_kernel<8><<<1, SSB_NBODY_THREADS, 0, st>>>(P, a, c);
    CKL("nbody_kernel");
    return 0;
}

int ssb_nbody_term_f64(const ssb_potential* ext, int32_t N, const double* masses, double G, double eps, double t, const double* y, double* dy,
                       void* scratch, size_t scratch_bytes, void* stream) {
    NbodyArgs a; ssb_potential P;
    if (int e = nbody_fill(a, P, ext, N, masses, G, eps)) return e;
    if (!y || !dy || !scratch || scratch_bytes < sizeof(double) * 3 * (size_t)N) return ssb_set_error(SSB_ERR_ARG, "nbody_term: NULL array / scratch");
    a.scratch = (double*)scratch;
    nbody_term_kernel<<<1, SSB_NBODY_THREADS, 0, (cudaStream_t)stream>>>(P, a, t, y, dy);
    CKL("nbody_term_kernel");
    return 0;
}

}  // extern "C"
