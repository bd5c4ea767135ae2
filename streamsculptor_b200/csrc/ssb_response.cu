// K3: first-order (mass, mass x radius) response of every particle to N_sh subhalos.
//
// Reference: GenerateMassRadiusPerturbation*.compute_perturbation_OTF (perturbative.py:101-135, 425-454, 726-755)
// integrating the field MassRadiusPerturbation_OTF.term (fields.py:175-206) with integrate_field (fields.py:35-99):
// per particle ONE coupled ODE with state [w(6), D(N_sh,12)]; all N_sh responses share the particle's step-size
// controller, whose RMS error norm runs over all 6 + 12 N_sh components.
//
// B200 mapping
//   * one CTA per particle (persistent CTAs pull particles from an atomic queue - spans differ per particle);
//   * the base orbit + tidal tensor are advanced ONCE per step attempt by one thread and broadcast through shared
//     memory (13 stage positions, 13 symmetric 3x3 tensors, 13 times);
//   * the 12 response components of a subhalo split into two independent second-order 3-vectors (mass block,
//     radius block: fields.py:195-202) = 2 N_sh "items"; threads sweep the items with the Nystrom form of the
//     same tableau (only force stages in registers), FSAL stage recomputed from the broadcast data;
//   * item state lives in an L2-resident ping-pong scratch (SoA, coalesced), candidates become current by a
//     pointer swap when the block-wide error reduction accepts the step;
//   * outside its time window a subhalo exerts no force (strict '<', potential.py:826); its tidal term is still
//     integrated, exactly as in the reference;
//   * subhalos are processed in the order of their window start t0 - t_window (sorted once per call on the device):
//     warps see coherent window states, and - when the perturbation ICs are zero (perturbative.py:712) and time runs
//     forwards - the subhalos whose window has not opened yet are still EXACTLY zero with zero error contribution, so
//     the sweep stops at the last "born" subhalo.  The reference evaluates them (vmapped lax.cond = select); the result
//     is bit-identical, the work is not.  Under the same conditions a subhalo whose window CLOSED before the particle was
//     released (t0 + t_window <= release time) never exerts a force on it: its response is exactly zero for the whole
//     integration.  In processing order those "dead" subhalos form a prefix (running maximum of the window ends, computed
//     with the sort), which the sweep of that particle skips as well.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "ssb_common.cuh"

using namespace ssb;

#ifndef SSB_RESP_THREADS
#define SSB_RESP_THREADS 128
#endif
#ifndef SSB_RESP_CTAS_PER_SM
#define SSB_RESP_CTAS_PER_SM 3      // 168 registers/thread: three 128-thread CTAs (three particles) per SM overlap base and item phases
#endif
#define SSB_RESP_MAX_SORT 4096
#define SSB_RESP_MAX_NP 16          // response_kernel_mp: particle slots per CTA (eight lanes each: four slots per warp)
#ifndef SSB_RESP_DEFAULT_NP
#define SSB_RESP_DEFAULT_NP 16      // measured (B200, 1000 subhalos): 1e5 particles 367 / 334 / 304 ms with 4 / 8 / 16 slots; small batches get fewer (resp_np)
#endif
#define CK(call) do { int _e = ssb_cuda_check((call), #call); if (_e) return _e; } while (0)
#define CKL(what) do { ssb_count_launch(); int _e = ssb_cuda_check(cudaGetLastError(), what); if (_e) return _e; } while (0)

struct RespArgs {
    int64_t N;
    const double *w0, *D0, *t0;
    double t1;
    CtrlDev c;
    double *wout, *Dout;
    int32_t *status, *nsteps;
    double* scratch;          // [grid][2][6][n_items]
    unsigned long long* counter;
    const double* sorted;     // [10][n_sh] subhalo parameters gathered in processing order: GM, rs, x0[3], v[3], t0, tw
    const double* start;      // [n_sh] window start t0 - tw, ascending
    const double* endmax;     // [n_sh] running maximum of the window ends t0 + tw in processing order (non-decreasing)
    const int* order;         // [n_sh] processing position -> user index
    int skip_unborn;          // 1: D0 == NULL (zero ICs): subhalos whose window has not opened are exactly zero
    int np;                   // response_kernel_mp: particles in flight per CTA (1..SSB_RESP_MAX_NP)
    // retired items (response_kernel_mp, see the kernel's header): per particle slot a log of step propagators [log_cap][36] and the
    // log position at which each subhalo retired [n_sh]
    int retire, log_cap;
    double* plog;             // [grid][SSB_RESP_MAX_NP][log_cap][36]
    int* rstep;               // [grid][SSB_RESP_MAX_NP][n_sh]
    // SaveAt(ts) for a single trajectory (backward progenitor response, perturbative.py:53-60): N == 1
    const double* ts_save; int M; double* wsave; double* Dsave;
};

template <int S>
struct BaseShared {
    double X[S][3];           // base stage positions
    double T[S][6];           // tidal tensors d2H/dq2 = -Hess(Phi_base): xx, yy, zz, xy, xz, yz
    double t[S];              // physical stage times
};

// base force + tidal tensor, recorded for the item sweep.  SIG != 0: the leading components are a fused static
// signature (parameters from the constant-bank copy Pc); the rest of the program goes through the interpreter.
template <int S, int SIG>
__device__ __noinline__ double3 base_force_call(const ssb_potential* P, const ssb_potential* Pc, BaseShared<S>* sh, int stage, double x, double y,
                                                double z, double t) {
    const double X[3] = {x, y, z};
    double phi, g[3];
    Sym3 H;
    if (SIG == SIG_GENERIC) {
        pot_eval<WANT_GRAD | WANT_HESS>(*P, X, t, phi, g, H);
    } else {
        fused_eval<SIG, WANT_GRAD | WANT_HESS>(*Pc, X, g, H);
        if (Pc->n_comp > SigInfo<SIG>::NF) {
            double g2[3];
            Sym3 H2;
            pot_eval<WANT_GRAD | WANT_HESS, false>(*P, X, t, phi, g2, H2, SigInfo<SIG>::NF);
            g[0] += g2[0]; g[1] += g2[1]; g[2] += g2[2];
            H.xx += H2.xx; H.yy += H2.yy; H.zz += H2.zz; H.xy += H2.xy; H.xz += H2.xz; H.yz += H2.yz;
        }
    }
    sh->X[stage][0] = x; sh->X[stage][1] = y; sh->X[stage][2] = z;
    sh->T[stage][0] = -H.xx; sh->T[stage][1] = -H.yy; sh->T[stage][2] = -H.zz;
    sh->T[stage][3] = -H.xy; sh->T[stage][4] = -H.xz; sh->T[stage][5] = -H.yz;
    sh->t[stage] = t;
    return make_double3(-g[0], -g[1], -g[2]);
}
template <int S, int SIG>
struct BaseForce {
    const ssb_potential* P; const ssb_potential* Pc; BaseShared<S>* sh; double dir; int stage;
    __device__ __forceinline__ void operator()(const double X[3], double tau, double A[3]) {
        const double3 a = base_force_call<S, SIG>(P, Pc, sh, stage, X[0], X[1], X[2], tau * dir);
        stage++;
        A[0] = a.x; A[1] = a.y; A[2] = a.z;
    }
};

// Main-loop variant executed by ALL lanes of warp 0 in lock-step: lane 0 advances the base orbit, lanes 1..6 ride along with the
// six unit vectors of the HOMOGENEOUS response system q'' = T(t) q.  Their step results are the columns of the 6x6 propagator
// Phi (and of the error map E) of this attempt, shared by every subhalo whose window is closed during the attempt - the sweep then
// costs two 6x6 mat-vecs per such item instead of 13 stages (see sweep_items).
template <int S, int SIG>
struct BaseWarpForce {
    const ssb_potential* P; const ssb_potential* Pc; BaseShared<S>* sh; double dir; int stage;
    __device__ __forceinline__ void operator()(const double X[3], double tau, double A[3]) {
        const double bx = __shfl_sync(0xffffffffu, X[0], 0), by = __shfl_sync(0xffffffffu, X[1], 0), bz = __shfl_sync(0xffffffffu, X[2], 0);
        // every lane evaluates the same point (and stores the same X, T, t record); lanes 1.. read the tidal tensor back from it
        const double3 a = base_force_call<S, SIG>(P, Pc, sh, stage, bx, by, bz, tau * dir);
        __syncwarp();
        const double* T = sh->T[stage];
        const bool base = (threadIdx.x & 31) == 0;
        A[0] = base ? a.x : (T[0] * X[0] + T[3] * X[1] + T[4] * X[2]);
        A[1] = base ? a.y : (T[3] * X[0] + T[1] * X[1] + T[5] * X[2]);
        A[2] = base ? a.z : (T[4] * X[0] + T[5] * X[1] + T[2] * X[2]);
        stage++;
    }
};

// one (subhalo, block) item: force = g_block(X_base(stage), t(stage)) + T(stage) . Q
struct ItemParams { double GM, rs, x0[3], v[3], t0, tw; int profile, blk; };

// PROFILE / BLK < 0: taken from the runtime fields (start-up code); >= 0: compile-time (the hot sweep)
template <int S, int PROFILE = -1, int BLK = -1>
struct ItemForce {
    const BaseShared<S>* sh; const ItemParams* ip; int stage;
    __device__ __forceinline__ void at(int i, const double Q[3], double A[3]) const {
        const int profile = PROFILE >= 0 ? PROFILE : ip->profile;
        const int blk = BLK >= 0 ? BLK : ip->blk;
        const double* X = sh->X[i];
        const double* T = sh->T[i];
        const double dt = sh->t[i] - ip->t0;
        double g0 = 0.0, g1 = 0.0, g2 = 0.0;
        if (PROFILE == SSB_PROFILE_HERNQUIST || PROFILE == SSB_PROFILE_PLUMMER) {
            // the hot sweeps: no branch around the window gate either (an item on the 13-stage path is inside its window at most stages) -
            // the force factor is computed and selected to zero outside the window; same operations inside it
            double rel[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) rel[k] = X[k] - fma(ip->v[k], dt, ip->x0[k]);
            const double r2 = fma(rel[0], rel[0], fma(rel[1], rel[1], rel[2] * rel[2]));
            double q;
            if (PROFILE == SSB_PROFILE_HERNQUIST) {
                // mass block (fields.py:191) and radius block (fields.py:200) share r, 1/(r + a) and G m/(r + a)^2: both factors from one copy of
                // the code and a select instead of two inlined branches per stage (the sweep's 13-stage body is what the instruction cache
                // cannot hold).  Same operations, same order as hernquist_terms / profile_dradius.
                const double ir = frsqrt(r2), r = r2 * ir, ira = frcp(r + ip->rs);
                const double t2 = ip->GM * ira * ira;
                const double qm = t2 * ir, qr = -2.0 * t2 * ira * ir;
                q = blk ? qr : qm;
            } else {
                const double sI = frsqrt(fma(ip->rs, ip->rs, r2)), s2 = sI * sI;
                const double qm = ip->GM * sI * s2, psi = ip->GM * ip->rs * sI * s2, qr = -3.0 * psi * s2;
                q = blk ? qr : qm;
            }
            const bool inside = fabs(dt) < ip->tw;                  // potential.py:826 / 846
            g0 = inside ? -q * rel[0] : 0.0; g1 = inside ? -q * rel[1] : 0.0; g2 = inside ? -q * rel[2] : 0.0;
        } else if (fabs(dt) < ip->tw) {                            // potential.py:826 / 846
            double rel[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) rel[k] = X[k] - fma(ip->v[k], dt, ip->x0[k]);
            const double r2 = fma(rel[0], rel[0], fma(rel[1], rel[1], rel[2] * rel[2]));
            double ph, q, w = 0;
            if (blk) profile_dradius(profile, ip->GM, ip->rs, r2, ph, q);                  // fields.py:200
            else profile_terms<WANT_GRAD>(profile, ip->GM, ip->rs, r2, ph, q, w);          // fields.py:191
            g0 = -q * rel[0]; g1 = -q * rel[1]; g2 = -q * rel[2];
        }
        A[0] = g0 + (T[0] * Q[0] + T[3] * Q[1] + T[4] * Q[2]);                              // fields.py:197 / 202
        A[1] = g1 + (T[3] * Q[0] + T[1] * Q[1] + T[5] * Q[2]);
        A[2] = g2 + (T[4] * Q[0] + T[5] * Q[1] + T[2] * Q[2]);
    }
    __device__ __forceinline__ void operator()(const double Q[3], double, double A[3]) { at(stage, Q, A); stage++; }
};

__device__ __forceinline__ double block_sum(double v, double* sred) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (l == 0) sred[w] = v;
    __syncthreads();
    double tot = 0.0;
    for (int i = 0; i < nw; ++i) tot += sred[i];
    return tot;
}

// item parameters from the gathered [10][n_sh] table (coalesced over j)
__device__ __forceinline__ void load_item_params(const double* __restrict__ tab, int profile, int j, int blk, int n_sh, ItemParams& ip) {
    ip.blk = blk;
    ip.profile = profile;
    ip.GM = __ldg(tab + j); ip.rs = __ldg(tab + n_sh + j);
#pragma unroll
    for (int k = 0; k < 3; ++k) { ip.x0[k] = __ldg(tab + (size_t)(2 + k) * n_sh + j); ip.v[k] = __ldg(tab + (size_t)(5 + k) * n_sh + j); }
    ip.t0 = __ldg(tab + (size_t)8 * n_sh + j); ip.tw = __ldg(tab + (size_t)9 * n_sh + j);
}

// ---- once per call: processing order (ascending window start) and the gathered parameter table ----
__global__ void response_sort_kernel(const ssb_subhalos Sh, int npad, int* order, double* start, double* endmax) {
    extern __shared__ unsigned char smem_raw[];
    double* key = reinterpret_cast<double*>(smem_raw);
    int* val = reinterpret_cast<int*>(key + npad);
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    for (int i = threadIdx.x; i < npad; i += blockDim.x) { key[i] = i < Sh.n ? Sh.t0[i] - Sh.tw[i] : inf; val[i] = i; }
    __syncthreads();
    for (int k = 2; k <= npad; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < npad; i += blockDim.x) {
                const int l = i ^ j;
                if (l > i) {
                    const bool up = (i & k) == 0;
                    const double a = key[i], b = key[l];
                    const bool sw = up ? (a > b || (a == b && val[i] > val[l])) : (a < b || (a == b && val[i] < val[l]));
                    if (sw) { key[i] = b; key[l] = a; const int t = val[i]; val[i] = val[l]; val[l] = t; }
                }
            }
            __syncthreads();
        }
    for (int i = threadIdx.x; i < Sh.n; i += blockDim.x) { order[i] = val[i]; start[i] = key[i]; }
    // running maximum of the window ends in processing order (inclusive scan with max, Hillis-Steele on the key array)
    __syncthreads();
    for (int i = threadIdx.x; i < npad; i += blockDim.x) key[i] = i < Sh.n ? Sh.t0[val[i]] + Sh.tw[val[i]] : inf;
    __syncthreads();
    for (int off = 1; off < npad; off <<= 1) {
        double mine[(SSB_RESP_MAX_SORT + 1023) / 1024];
        int q = 0;
        for (int i = threadIdx.x; i < npad; i += blockDim.x, ++q) mine[q] = i >= off ? fmax(key[i], key[i - off]) : key[i];
        __syncthreads();
        q = 0;
        for (int i = threadIdx.x; i < npad; i += blockDim.x, ++q) key[i] = mine[q];
        __syncthreads();
    }
    for (int i = threadIdx.x; i < Sh.n; i += blockDim.x) endmax[i] = key[i];
}
__global__ void response_identity_kernel(const ssb_subhalos Sh, int* order, double* start, double* endmax) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    if (i < Sh.n) { order[i] = i; start[i] = -inf; endmax[i] = inf; }          // every subhalo counts as born and alive
}
__global__ void response_gather_kernel(const ssb_subhalos Sh, const int* order, double* tab) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= Sh.n) return;
    const int o = order[j], n = Sh.n;
    tab[j] = Sh.G * Sh.m[o]; tab[n + j] = Sh.rs[o];
    for (int k = 0; k < 3; ++k) { tab[(size_t)(2 + k) * n + j] = Sh.x0[3 * o + k]; tab[(size_t)(5 + k) * n + j] = Sh.v[3 * o + k]; }
    tab[(size_t)8 * n + j] = Sh.t0[o]; tab[(size_t)9 * n + j] = Sh.tw[o];
}

// the item sweep of one step attempt: candidates into `nxt`, squared scaled errors into esq.  PROFILE is a template
// parameter so that the unrolled 13-stage body stays small (instruction cache).
//
// An item whose subhalo window is closed at every stage time of the attempt obeys the homogeneous linear system q'' = T(t) q, whose
// Runge-Kutta step is a LINEAR map of (q, p): candidate = Phi (q, p), error estimate = E (q, p), with the 6x6 matrices Phi, E of
// this attempt (PhiE, computed once by warp 0 alongside the base orbit).  Those items (~90 % at t_window = 150 Myr over 3 Gyr)
// cost 72 FMAs instead of 13 stages; algebraically identical to the staged evaluation, rounding differs at the 1e-16 level.
__device__ __forceinline__ void item_finish(const double (&q)[3], const double (&pp)[3], const double (&q1)[3], const double (&pp1)[3], const double (&ex)[3],
                                            const double (&ep)[3], const CtrlDev& c, double* __restrict__ nxt, int n_items, int it, double& esq, int& bad_local) {
    bool nan_cand = false, fin = true;                        // no short-circuit branches in the sweep's tail either
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        nan_cand |= isnan(q1[k]) | isnan(pp1[k]);
        fin &= isfinite(q1[k]) & isfinite(pp1[k]);
    }
    bad_local |= fin ? 0 : 1;
    esq += err_sq6(q, pp, q1, pp1, ex, ep, c.rtol, c.atol, nan_cand);
#pragma unroll
    for (int k = 0; k < 3; ++k) { nxt[(size_t)k * n_items + it] = q1[k]; nxt[(size_t)(3 + k) * n_items + it] = pp1[k]; }
}

template <int SOLVER, int PROFILE>
__device__ __forceinline__ void sweep_items(const BaseShared<Tab<SOLVER>::S>* sb, const double* __restrict__ PhiE, const double* __restrict__ tab, int n_sh,
                                            int n_items, int n_dead /* first position swept: dead or retired prefix */, int n_act, const double* __restrict__ cur, double* __restrict__ nxt, double dt,
                                            const CtrlDev& c, double& esq, int& bad_local, int rot = 0) {
    constexpr int S = Tab<SOLVER>::S;
    const double t_lo = fmin(sb->t[0], sb->t[S - 1]), t_hi = fmax(sb->t[0], sb->t[S - 1]);     // every stage time lies in [t_lo, t_hi]
    // processing positions [n_dead, n_act) of both blocks: idx -> (blk, j)
    const int n_span = n_act > n_dead ? n_act - n_dead : 0;
    const int n2 = 2 * n_span;
    const double* __restrict__ t0tab = tab + (size_t)8 * n_sh;
    const double* __restrict__ twtab = tab + (size_t)9 * n_sh;
    // software-pipelined: the state of the NEXT item (L2-resident scratch, ~1 us away) is requested before the current one is processed
    // `rot` rotates which warp takes which 32 items of a pass: the first positions of the range are closed windows waiting for their
    // chunk to retire (cheap propagator path), the rest open ones (13 stages) - without the rotation warp 0 would get the cheap ones of
    // every particle.  Callers derive it from the PARTICLE index, so a particle's arithmetic does not depend on where it runs.
    int idx = (int)((threadIdx.x + 32u * (unsigned)rot) % blockDim.x);
    double yn[6] = {0, 0, 0, 0, 0, 0}, t0n = 0.0, twn = 0.0;
    if (idx < n2) {
        const int blk = idx >= n_span, j = n_dead + idx - blk * n_span, it = blk * n_sh + j;
#pragma unroll
        for (int k = 0; k < 6; ++k) yn[k] = cur[(size_t)k * n_items + it];
        t0n = __ldg(t0tab + j); twn = __ldg(twtab + j);
    }
    while (idx < n2) {
        const int blk = idx >= n_span, j = n_dead + idx - blk * n_span, it = blk * n_sh + j;
        const double q[3] = {yn[0], yn[1], yn[2]}, pp[3] = {yn[3], yn[4], yn[5]};
        const double t0j = t0n, twj = twn;
        const int idn = idx + blockDim.x;
        if (idn < n2) {
            const int blkn = idn >= n_span, jn = n_dead + idn - blkn * n_span, itn = blkn * n_sh + jn;
#pragma unroll
            for (int k = 0; k < 6; ++k) yn[k] = cur[(size_t)k * n_items + itn];
            t0n = __ldg(t0tab + jn); twn = __ldg(twtab + jn);
        }
        double q1[3], pp1[3], ex[3], ep[3];
        const bool closed = (t_lo - t0j >= twj) || (t0j - t_hi >= twj);       // |t - t0| >= tw at every stage (potential.py:826)
        if (closed) {
            const double y[6] = {q[0], q[1], q[2], pp[0], pp[1], pp[2]};
            const double2* __restrict__ M2 = reinterpret_cast<const double2*>(PhiE);     // 16-byte shared loads: [12 rows][3] double2
            double o[6], e[6];
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                double so = 0.0, se = 0.0;
#pragma unroll
                for (int c2 = 0; c2 < 3; ++c2) {
                    const double2 m = M2[r * 3 + c2], me = M2[18 + r * 3 + c2];
                    so = fma(m.x, y[2 * c2], so); so = fma(m.y, y[2 * c2 + 1], so);
                    se = fma(me.x, y[2 * c2], se); se = fma(me.y, y[2 * c2 + 1], se);
                }
                o[r] = so; e[r] = se;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) { q1[k] = o[k]; pp1[k] = o[3 + k]; ex[k] = e[k]; ep[k] = e[3 + k]; }
        } else {
            ItemParams ip; load_item_params(tab, PROFILE, j, blk, n_sh, ip);
            ItemForce<S, PROFILE, -1> f{sb, &ip, 1};
            double G[S][3];
            f.at(0, q, G[0]);
            rk_stages<SOLVER>(f, q, pp, 0.0, dt, G);
            rk_candidate<SOLVER>(q, pp, dt, G, q1, pp1);
            if (SOLVER == 5) f.at(S - 1, q1, G[S - 1]);          // Dopri8: e_14 = (e^T A)_14 = 0, stage unused
            else { G[S - 1][0] = G[S - 1][1] = G[S - 1][2] = 0.0; }
            rk_error<SOLVER>(pp, dt, G, ex, ep);
        }
        item_finish(q, pp, q1, pp1, ex, ep, c, nxt, n_items, it, esq, bad_local);
        idx = idn;
    }
}

// dense output of every item at one save time inside an ACCEPTED step (stages recomputed from the pre-step state)
template <int SOLVER>
__device__ __noinline__ void save_items(const BaseShared<Tab<SOLVER>::S>* sb, const double* __restrict__ tab, int profile, int n_sh, int n_items,
                                        const double* __restrict__ cur, const int* __restrict__ order, double dt, double theta, double dir,
                                        double* __restrict__ Dsave_row) {
    constexpr int S = Tab<SOLVER>::S;
    for (int it = threadIdx.x; it < n_items; it += blockDim.x) {
        const int j = it % n_sh, blk = it / n_sh;
        ItemParams ip; load_item_params(tab, profile, j, blk, n_sh, ip);
        ItemForce<S> f{sb, &ip, 1};
        double q[3], pp[3], G[S][3], q1[3], pp1[3], qo[3], po[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) { q[k] = cur[(size_t)k * n_items + it]; pp[k] = cur[(size_t)(3 + k) * n_items + it]; }
        f.at(0, q, G[0]);
        rk_stages<SOLVER>(f, q, pp, 0.0, dt, G);
        rk_candidate<SOLVER>(q, pp, dt, G, q1, pp1);
        f.at(S - 1, q1, G[S - 1]);
        if (theta >= 1.0) {
#pragma unroll
            for (int k = 0; k < 3; ++k) { qo[k] = q1[k]; po[k] = pp1[k]; }
        } else {
            rk_dense<SOLVER>(q, pp, q1, pp1, dt, G, theta, qo, po);
        }
        double* o = Dsave_row + (size_t)order[j] * 12 + blk * 6;
#pragma unroll
        for (int k = 0; k < 3; ++k) { o[k] = qo[k]; o[3 + k] = dir * po[k]; }
    }
}

template <int SOLVER, int SIG, int PROFILE, bool SAVE>
__global__ void __launch_bounds__(SSB_RESP_THREADS, SSB_RESP_CTAS_PER_SM) response_kernel(const __grid_constant__ ssb_potential Pin, const ssb_subhalos Sh,
                                                                                          const RespArgs a) {
    typedef Tab<SOLVER> T;
    constexpr int S = T::S;
    __shared__ ssb_potential sP;
    __shared__ BaseShared<S> sb;
    __shared__ double sred[32];
    __shared__ __align__(16) double sPhiE[72];         // propagator Phi[6][6] and error map E[6][6] of the current attempt (homogeneous items)
    __shared__ int s_nact, s_ndead;
    __shared__ long long s_part;
    stage_potential(&sP, &Pin);
    logtab_init();
    const int tid = threadIdx.x;
    const int n_sh = Sh.n, n_items = 2 * n_sh, ncomp = 6 + 12 * n_sh;
    double* buf0 = a.scratch + (size_t)blockIdx.x * 2 * 6 * n_items;
    double* buf1 = buf0 + (size_t)6 * n_items;
    const CtrlDev c = a.c;

    for (;;) {
        __syncthreads();
        if (tid == 0) s_part = (long long)atomicAdd(a.counter, 1ULL);
        __syncthreads();
        const long long part = s_part;
        if (part >= a.N) break;
        const double t0_in = a.t0[part], t1_in = a.t1;
        const double dir = (t0_in < t1_in) ? 1.0 : -1.0;
        const double T0 = t0_in * dir, T1 = t1_in * dir;
        double* cur = buf0;
        double* nxt = buf1;
        const bool skip = a.skip_unborn && dir > 0.0;
        if (tid == 0) {            // subhalos whose window closed before the release of this particle: exact zeros throughout (see the header)
            int nd = 0;
            if (skip) { int lo = 0, hi = n_sh; while (lo < hi) { const int mid = (lo + hi) >> 1; if (a.endmax[mid] <= T0) lo = mid + 1; else hi = mid; } nd = lo & ~15; }
            s_ndead = nd;                                    // multiple of 16 items: the sweep's 128-byte segments stay aligned
        }
        // ---- load item state (SoA [6][n_items], item = blk * n_sh + processing position); momentum-like rows carry dir ----
        for (int it = tid; it < n_items; it += blockDim.x) {
            const int j = it % n_sh, blk = it / n_sh, o = a.order[j];
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                double v = a.D0 ? a.D0[((size_t)part * n_sh + o) * 12 + blk * 6 + k] : 0.0;
                if (k >= 3) v *= dir;
                cur[(size_t)k * n_items + it] = v;
                nxt[(size_t)k * n_items + it] = v;         // unborn subhalos are never written again: both buffers hold their zeros
            }
        }
        // base state lives in thread 0
        double x[3] = {0, 0, 0}, p[3] = {0, 0, 0}, F[S][3];
        BaseForce<S, SIG> bforce{&sP, &Pin, &sb, dir, 0};
        BaseWarpForce<S, SIG> wforce{&sP, &Pin, &sb, dir, 0};
        int status = 0, n_steps = 0, n_acc = 0, n_rej = 0;
        bool at_dtmin = false;
        double tprev = T0, tnext = T0;
        // ================= initial step (HNW over the whole coupled state) =================
        double d0s = 0.0, d1s = 0.0;
        if (tid == 0) {
#pragma unroll
            for (int k = 0; k < 3; ++k) { x[k] = a.w0[6 * part + k]; p[k] = dir * a.w0[6 * part + 3 + k]; }
            bforce.stage = 0;
            bforce(x, T0, F[0]);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double sx = fma(c.rtol, fabs(x[k]), c.atol), sp = fma(c.rtol, fabs(p[k]), c.atol);
                double q;
                q = x[k] / sx; d0s = fma(q, q, d0s); q = p[k] / sp; d0s = fma(q, q, d0s);
                q = p[k] / sx; d1s = fma(q, q, d1s); q = F[0][k] / sp; d1s = fma(q, q, d1s);
            }
        }
        __syncthreads();
        for (int it = tid; it < n_items; it += blockDim.x) {
            ItemParams ip; load_item_params(a.sorted, Sh.profile, it % n_sh, it / n_sh, n_sh, ip);
            ItemForce<S> f{&sb, &ip, 0};
            double q[3], pp[3], G[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) { q[k] = cur[(size_t)k * n_items + it]; pp[k] = cur[(size_t)(3 + k) * n_items + it]; }
            f.at(0, q, G);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double sx = fma(c.rtol, fabs(q[k]), c.atol), sp = fma(c.rtol, fabs(pp[k]), c.atol);
                double r;
                r = q[k] / sx; d0s = fma(r, r, d0s); r = pp[k] / sp; d0s = fma(r, r, d0s);
                r = pp[k] / sx; d1s = fma(r, r, d1s); r = G[k] / sp; d1s = fma(r, r, d1s);
            }
        }
        const double d0 = sqrt(block_sum(d0s, sred) / ncomp);
        const double d1 = sqrt(block_sum(d1s, sred) / ncomp);
        const double h0 = hnw_h0(d0, d1);
        double d2s = 0.0;
        double F1[3];
        if (tid == 0) {
            double X1[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) X1[k] = fma(h0, p[k], x[k]);
            bforce.stage = 1;
            bforce(X1, T0 + h0, F1);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double sx = fma(c.rtol, fabs(x[k]), c.atol), sp = fma(c.rtol, fabs(p[k]), c.atol);
                double q;
                q = (fma(h0, F[0][k], p[k]) - p[k]) / sx; d2s = fma(q, q, d2s);
                q = (F1[k] - F[0][k]) / sp; d2s = fma(q, q, d2s);
            }
        }
        __syncthreads();
        for (int it = tid; it < n_items; it += blockDim.x) {
            ItemParams ip; load_item_params(a.sorted, Sh.profile, it % n_sh, it / n_sh, n_sh, ip);
            ItemForce<S> f{&sb, &ip, 0};
            double q[3], pp[3], G0[3], G1[3], q1[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) { q[k] = cur[(size_t)k * n_items + it]; pp[k] = cur[(size_t)(3 + k) * n_items + it]; }
            f.at(0, q, G0);
#pragma unroll
            for (int k = 0; k < 3; ++k) q1[k] = fma(h0, pp[k], q[k]);
            f.at(1, q1, G1);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double sx = fma(c.rtol, fabs(q[k]), c.atol), sp = fma(c.rtol, fabs(pp[k]), c.atol);
                double r;
                r = (fma(h0, G0[k], pp[k]) - pp[k]) / sx; d2s = fma(r, r, d2s);
                r = (G1[k] - G0[k]) / sp; d2s = fma(r, r, d2s);
            }
        }
        const double d2 = sqrt(block_sum(d2s, sred) / ncomp) / h0;
        {
            double h = fmin(hnw_h1<T::ORDER>(h0, d1, d2), c.dtmax);
            at_dtmin = h <= c.dtmin;                 // every thread keeps an identical copy of the controller state
            h = fmax(h, c.dtmin);
            tnext = fmin(T0 + h, T1);
        }
        int n_act_run = 0;
        int save_idx = 0;
        if (SAVE) {                                   // rows never reached stay +inf (diffrax SaveAt semantics)
            const double inf0 = __longlong_as_double(0x7ff0000000000000LL);
            for (size_t q = tid; q < (size_t)a.M * n_sh * 12; q += blockDim.x) a.Dsave[q] = inf0;
            for (int q = tid; q < a.M * 6; q += blockDim.x) a.wsave[q] = inf0;
            __syncthreads();
        }
        // ================= main loop: all threads follow the same (uniform) control flow =================
        while (tprev < T1 && status == 0) {
            if (n_steps >= c.max_steps) { status = 1; break; }
            const double dt = tnext - tprev;
            double x1[3], p1[3];
            double esq = 0.0;
            int bad_local = 0;
            __syncthreads();
            if (tid < 32) {
                // warp 0: lane 0 = base orbit, lanes 1..6 = unit vectors of the homogeneous response system (propagator columns)
                const int col = tid - 1;
                if (tid > 0) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) { x[k] = (col == k) ? 1.0 : 0.0; p[k] = (col == 3 + k) ? 1.0 : 0.0; }
                    const double* T0m = sb.T[0];
                    F[0][0] = T0m[0] * x[0] + T0m[3] * x[1] + T0m[4] * x[2];
                    F[0][1] = T0m[3] * x[0] + T0m[1] * x[1] + T0m[5] * x[2];
                    F[0][2] = T0m[4] * x[0] + T0m[5] * x[1] + T0m[2] * x[2];
                }
                double ex[3], ep[3];
                wforce.stage = 1;
                // stage 0 of sb (X, T, t at tprev) is already in place: FSAL copy below / initial evaluation above
                rk_stages<SOLVER>(wforce, x, p, tprev, dt, F);
                rk_candidate<SOLVER>(x, p, dt, F, x1, p1);
                wforce.stage = S - 1;
                wforce(x1, tprev + T::c(S - 1) * dt, F[S - 1]);
                rk_error<SOLVER>(p, dt, F, ex, ep);
                if (tid == 0) {
                    bool nan_cand = false;
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        nan_cand |= isnan(x1[k]) | isnan(p1[k]);
                        if (!isfinite(x1[k]) || !isfinite(p1[k])) bad_local = 1;
                    }
                    esq = err_sq6(x, p, x1, p1, ex, ep, c.rtol, c.atol, nan_cand);
                    // number of subhalos whose window start lies before the end of this attempt (sorted ascending)
                    int na = n_sh;
                    if (skip) { int lo = 0, hi = n_sh; while (lo < hi) { const int mid = (lo + hi) >> 1; if (a.start[mid] < tnext) lo = mid + 1; else hi = mid; } na = lo; }
                    s_nact = na;
                } else if (col < 6) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        sPhiE[k * 6 + col] = x1[k]; sPhiE[(3 + k) * 6 + col] = p1[k];
                        sPhiE[36 + k * 6 + col] = ex[k]; sPhiE[36 + (3 + k) * 6 + col] = ep[k];
                    }
                }
            }
            __syncthreads();
            // monotone within a particle: a subhalo touched by a (possibly rejected) attempt has a candidate in `nxt` that must be
            // overwritten by every later attempt, even one that ends before its window opens
            n_act_run = max(n_act_run, s_nact);
            const int n_act = n_act_run;
            // ---- item sweep over the born subhalos: mass block, then radius block ----
            sweep_items<SOLVER, PROFILE>(&sb, sPhiE, a.sorted, n_sh, n_items, s_ndead, n_act, cur, nxt, dt, c, esq, bad_local);
            const double err = sqrt(block_sum(esq, sred) / ncomp);
            const int any_bad = __syncthreads_or(bad_local);
            double hn; bool bad;
            const bool keep = pid_update<T::ORDER>(err, dt, c, at_dtmin, hn, bad);
            n_steps++;
            if (bad) { status = 2; n_rej++; break; }
            if (keep) {
                n_acc++;
                if (any_bad) { status = 2; break; }
                if (SAVE) {
                    // every ts[save_idx] <= tnext is interpolated inside this accepted step (base by thread 0, items by all)
                    while (save_idx < a.M) {
                        const double tq = a.ts_save[save_idx] * dir;
                        if (!(tq <= tnext)) break;
                        const double theta = (tq == tnext) ? 1.0 : (tq - tprev) / dt;
                        if (tid == 0) {
                            double xo[3], po[3];
                            if (theta >= 1.0) { for (int k = 0; k < 3; ++k) { xo[k] = x1[k]; po[k] = p1[k]; } }
                            else rk_dense<SOLVER>(x, p, x1, p1, dt, F, theta, xo, po);
                            for (int k = 0; k < 3; ++k) { a.wsave[6 * save_idx + k] = xo[k]; a.wsave[6 * save_idx + 3 + k] = dir * po[k]; }
                        }
                        save_items<SOLVER>(&sb, a.sorted, Sh.profile, n_sh, n_items, cur, a.order, dt, theta, dir, a.Dsave + (size_t)save_idx * n_sh * 12);
                        save_idx++;
                    }
                }
                double* tmp = cur; cur = nxt; nxt = tmp;
                if (tid == 0) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) { x[k] = x1[k]; p[k] = p1[k]; F[0][k] = F[S - 1][k]; sb.X[0][k] = sb.X[S - 1][k]; }
#pragma unroll
                    for (int k = 0; k < 6; ++k) sb.T[0][k] = sb.T[S - 1][k];
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
        // ---- outputs: final state if the end was reached, +inf otherwise (diffrax SaveAt semantics) ----
        const bool ok = (status == 0) && (T0 < T1);
        const double inf = __longlong_as_double(0x7ff0000000000000LL);
        for (int it = tid; it < n_items; it += blockDim.x) {
            const int j = it % n_sh, blk = it / n_sh, o = a.order[j];
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                double v = cur[(size_t)k * n_items + it];
                if (k >= 3) v *= dir;
                a.Dout[((size_t)part * n_sh + o) * 12 + blk * 6 + k] = ok ? v : inf;
            }
        }
        if (tid == 0) {
#pragma unroll
            for (int k = 0; k < 3; ++k) { a.wout[6 * part + k] = ok ? x[k] : inf; a.wout[6 * part + 3 + k] = ok ? dir * p[k] : inf; }
            a.status[part] = status;
            a.nsteps[3 * part] = n_steps; a.nsteps[3 * part + 1] = n_acc; a.nsteps[3 * part + 2] = n_rej;
        }
    }
}


// =============================================================================================
// K3-mp: the same computation with SEVERAL particles in flight per CTA.
//
// ncu (profiles/r1_response_kernel_source.txt) shows what bounds response_kernel: the base-orbit phase of an attempt is a serial,
// latency-bound instruction stream of ONE warp (13 force + Hessian evaluations, ~4.7 k instructions at ~5 cycles each) during
// which the other warps of the CTA wait at the barrier - 45 % of every attempt.  SIMT makes that stream free to share: here
// warp 0 advances the base orbits (and propagator columns) of up to four particles AT ONCE, eight lanes per particle slot
// (lane 8q: base orbit of slot q, lanes 8q+1..8q+6: unit vectors), so the serial phase costs the same for np particles as
// for one; the item sweeps of the slots then run back to back on all threads.  Every slot keeps its own controller, scratch
// buffers, window bookkeeping and work-queue position; slots finish and refill independently.  Per-slot control state lives
// in shared memory and is advanced by thread 0 between two barriers.  With np = 1 the arithmetic (including the order of the
// error reduction) is that of response_kernel.
// =============================================================================================
//
// RETIRED ITEMS.  With zero perturbation ICs and forward time a subhalo is unborn (exact zeros), then open (window |t - t0| < t_window:
// 13-stage items), then closed FOR GOOD.  From then on its items obey the homogeneous system q'' = T(t) q: every accepted step maps
// them by the step's 6x6 propagator Phi_n, and their share of the error norm is sum (E_n y)^2 / scale^2.  Such items are RETIRED
// (in processing order they form a prefix: running maximum of the window ends, chunks of 16 positions) and never swept again:
//   * their state is frozen at retirement (y_c, kept in both ping-pong buffers) together with the position in the particle's LOG of
//     step propagators; at the end (or when the log is full) the suffix products R_i = Phi_N ... Phi_{i+1} are formed once and every
//     retired item gets y_final = R_i y_c - algebraically the step-by-step application, one 6x6 mat-vec instead of one per step;
//   * their error contribution needs only the 6x6 second-moment matrix C = sum y y^T of the retired set, advanced by C <- Phi C Phi^T
//     per accepted step: sum_items |E y|^2_k = (E C E^T)_kk.  The scale of a retired component is taken as atol (the exact
//     atol + rtol max(|y0|, |y1|) would need every item): the response to a unit-mass subhalo is ~1e-8 or smaller, so the relative
//     deviation of the norm is ~rtol |y| / atol ~ 1e-8; the kernel checks the bound rtol sqrt(max_k C_kk) <= 5e-4 atol at every attempt
//     and, should it ever fail (huge subhalo masses), restarts that particle with retirement switched off.
// ~80 % of the item-steps of a C4 run (1000 subhalos, t_window = 150 Myr over 3 Gyr) are retired ones: the sweep, and with it the
// item-state traffic (52 GB per launch before), shrinks to the open windows.  SSB_RESP_RETIRE=0 switches the path off (A/B, tests).
struct RespSlot {
    long long part;                 // particle index; -1: free (refill from the queue); -2: queue exhausted
    double dir, T0, T1, tprev, tnext;
    int status, n_steps, n_acc, n_rej;
    int at_dtmin, accepted, finishing, flip;     // finishing: 1 = write out and free the slot, 2 = (re)start the particle in `part`
    int n_act_run, n_dead, skip, pad;
    int n_ret, log_len, retire_on, no_retire, need_flush, pad2;
};

#define SSB_RESP_LOG_BLK 16         // log entries staged in shared memory per pass of the suffix-product chain

// R_i = L[len-1] ... L[i] for i = len-1 .. 0, in place (L: [len][36] row-major 6x6, global memory).  Uniform over the CTA: all threads
// move blocks of SSB_RESP_LOG_BLK entries through shared memory, warp 0 runs the chain.
__device__ __forceinline__ void suffix_products(double* __restrict__ L, int len, double* sBlk /*[BLK*36]*/, double* sR /*[2][36]*/) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid < 36) sR[tid] = (tid / 6 == tid % 6) ? 1.0 : 0.0;
    int cur = 0;
    for (int blk_end = len; blk_end > 0; blk_end -= SSB_RESP_LOG_BLK) {
        const int b0 = blk_end > SSB_RESP_LOG_BLK ? blk_end - SSB_RESP_LOG_BLK : 0, nb = blk_end - b0;
        __syncthreads();
        for (int e = tid; e < nb * 36; e += blockDim.x) sBlk[e] = L[(size_t)b0 * 36 + e];
        __syncthreads();
        if (wid == 0) {
            for (int i = nb - 1; i >= 0; --i) {
                const double* Rc = sR + cur * 36;
                double* Li = sBlk + i * 36;
                double acc[2] = {0.0, 0.0};
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int e = lane + 32 * u;
                    if (e < 36) {
                        const int r = e / 6, c = e % 6;
#pragma unroll
                        for (int k = 0; k < 6; ++k) acc[u] = fma(Rc[r * 6 + k], Li[k * 6 + c], acc[u]);
                    }
                }
                __syncwarp();
#pragma unroll
                for (int u = 0; u < 2; ++u) { const int e = lane + 32 * u; if (e < 36) { sR[(cur ^ 1) * 36 + e] = acc[u]; Li[e] = acc[u]; } }
                cur ^= 1;
                __syncwarp();
            }
        }
        __syncthreads();
        for (int e = tid; e < nb * 36; e += blockDim.x) L[(size_t)b0 * 36 + e] = sBlk[e];
    }
    __syncthreads();
}

// lanes of one slot follow the slot leader's base point; `sh` is the slot's stage record (or the trash record of idle lanes)
template <int S, int SIG>
struct BaseGroupForce {
    const ssb_potential* P; const ssb_potential* Pc; BaseShared<S>* sh; double dir; int stage; int src; bool base;
    __device__ __forceinline__ void operator()(const double X[3], double tau, double A[3]) {
        const double bx = __shfl_sync(0xffffffffu, X[0], src), by = __shfl_sync(0xffffffffu, X[1], src), bz = __shfl_sync(0xffffffffu, X[2], src);
        const double3 a = base_force_call<S, SIG>(P, Pc, sh, stage, bx, by, bz, tau * dir);
        __syncwarp();
        const double* T = sh->T[stage];
        A[0] = base ? a.x : (T[0] * X[0] + T[3] * X[1] + T[4] * X[2]);
        A[1] = base ? a.y : (T[3] * X[0] + T[1] * X[1] + T[5] * X[2]);
        A[2] = base ? a.z : (T[4] * X[0] + T[5] * X[1] + T[2] * X[2]);
        stage++;
    }
};

template <int SOLVER, int SIG, int PROFILE>
__global__ void __launch_bounds__(SSB_RESP_THREADS, SSB_RESP_CTAS_PER_SM) response_kernel_mp(const __grid_constant__ ssb_potential Pin, const ssb_subhalos Sh,
                                                                                             const RespArgs a) {
    typedef Tab<SOLVER> T;
    constexpr int S = T::S;
    constexpr int NPX = SSB_RESP_MAX_NP;
    __shared__ ssb_potential sP;
    // dynamic shared memory (resp_mp_smem_bytes): stage records [2 NPX] ([NPX + 4 w + g]: scratch record of lane group g of warp w while it
    // has no particle), propagator + error map [NPX][72], moment matrices sC, sT [NPX][36]
    extern __shared__ __align__(16) unsigned char s_dyn[];
    BaseShared<S>* const sb = reinterpret_cast<BaseShared<S>*>(s_dyn);
    double (*const sPhiE)[72] = reinterpret_cast<double (*)[72]>(s_dyn + sizeof(BaseShared<S>) * 2 * NPX);
    double (*const sC)[36] = reinterpret_cast<double (*)[36]>(s_dyn + sizeof(BaseShared<S>) * 2 * NPX + sizeof(double) * 72 * NPX);
    double (*const sT)[36] = sC + NPX;
    __shared__ double sred[32], sred_q[NPX][SSB_RESP_THREADS / 32];
    __shared__ RespSlot slot[NPX];
    __shared__ int s_nact[NPX], s_bad[NPX];
    __shared__ double s_besq[NPX];                            // squared scaled error of the base orbit's attempt
    __shared__ int s_service, s_live;
    __shared__ double sLogBlk[SSB_RESP_LOG_BLK * 36], sR[2 * 36];
    __shared__ int s_guard[NPX];
    stage_potential(&sP, &Pin);
    logtab_init();
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
    const int NP = a.np;
    const int n_sh = Sh.n, n_items = 2 * n_sh, ncomp = 6 + 12 * n_sh;
    const CtrlDev c = a.c;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    // warp 0: lane = 8 * slot + role; role 0 = base orbit, 1..6 = propagator columns, 7 = idle
    const int col = (lane & 7) - 1;
    const int my_slot = (4 * wid + (lane >> 3) < NP) ? 4 * wid + (lane >> 3) : NPX;        // NPX: this lane group carries no slot
    auto base_tid = [](int q) { return 32 * (q >> 2) + 8 * (q & 3); };             // thread that carries the base orbit of slot q
    double x[3] = {8.0, 0.0, 0.0}, p[3] = {0, 0, 0}, F[S][3], x1[3] = {8.0, 0.0, 0.0}, p1[3] = {0, 0, 0};
#pragma unroll
    for (int l = 0; l < S; ++l) F[l][0] = F[l][1] = F[l][2] = 0.0;
    if (tid < NPX) {
        RespSlot& s = slot[tid];
        s.part = tid < NP ? -1 : -2; s.dir = 1.0; s.T0 = s.T1 = s.tprev = s.tnext = 0.0; s.status = s.n_steps = s.n_acc = s.n_rej = 0;
        s.at_dtmin = s.accepted = s.finishing = s.flip = s.n_act_run = s.n_dead = s.skip = s.pad = 0;
        s.n_ret = s.log_len = s.retire_on = s.no_retire = s.need_flush = s.pad2 = 0;
    }
    if (tid < NPX) { s_nact[tid] = 0; s_bad[tid] = 0; s_besq[tid] = 0.0; s_guard[tid] = 0; }
    const double inv_atol2 = c.atol > 0.0 ? 1.0 / (c.atol * c.atol) : 0.0;
    const double guard_lim = 2.5e-7 * c.atol * c.atol;         // rtol^2 max_k C_kk <= (5e-4 atol)^2
    if (tid == 0) { s_service = 1; s_live = 1; }
    double* const cta_buf = a.scratch + (size_t)blockIdx.x * NPX * 2 * 6 * n_items;
    double* const cta_log = a.plog + (size_t)blockIdx.x * NPX * (size_t)a.log_cap * 36;
    int* const cta_rstep = a.rstep + (size_t)blockIdx.x * NPX * n_sh;

    for (;;) {
        __syncthreads();                                       // slot state written by thread 0 is visible
        // ---- retired items of the slots whose attempt was accepted (warp w serves slot w): log the step's propagator, advance the moment
        //      matrix C <- Phi C Phi^T, then retire the chunks of 16 positions whose windows are now closed for good ----
        for (int q = wid; q < NP; q += nw) {
            if (!(slot[q].part >= 0 && slot[q].accepted && slot[q].retire_on)) continue;
            int n_ret = slot[q].n_ret, len = slot[q].log_len;
            const double* PE = sPhiE[q];
            double* Cq = sC[q];
            if (n_ret > slot[q].n_dead) {
                double* Lq = cta_log + ((size_t)q * a.log_cap + len) * 36;
                for (int e = lane; e < 36; e += 32) Lq[e] = PE[e];
#pragma unroll
                for (int u = 0; u < 2; ++u) {                  // T = Phi C
                    const int e = lane + 32 * u;
                    if (e < 36) {
                        const int r = e / 6, cc = e % 6;
                        double acc = 0.0;
#pragma unroll
                        for (int k = 0; k < 6; ++k) acc = fma(PE[r * 6 + k], Cq[k * 6 + cc], acc);
                        sT[q][e] = acc;
                    }
                }
                __syncwarp();
                if (lane < 21) {                               // C = T Phi^T, upper triangle mirrored: exactly symmetric
                    int r = 0, rem = lane;
                    while (rem >= 6 - r) { rem -= 6 - r; ++r; }
                    const int cc = r + rem;
                    double acc = 0.0;
#pragma unroll
                    for (int k = 0; k < 6; ++k) acc = fma(sT[q][r * 6 + k], PE[cc * 6 + k], acc);
                    Cq[r * 6 + cc] = acc; Cq[cc * 6 + r] = acc;
                }
                __syncwarp();
                len++;
            }
            if (!slot[q].finishing) {
                const double tp = slot[q].tprev;
                double* cur = cta_buf + (size_t)(2 * q + slot[q].flip) * 6 * n_items;
                double* nxt = cta_buf + (size_t)(2 * q + (slot[q].flip ^ 1)) * 6 * n_items;
                while (n_ret + 16 <= n_sh && a.endmax[n_ret + 15] <= tp) {
                    const int j = n_ret + (lane & 15), blk = lane >> 4, it = blk * n_sh + j;
                    double y[6];
#pragma unroll
                    for (int k = 0; k < 6; ++k) { y[k] = cur[(size_t)k * n_items + it]; nxt[(size_t)k * n_items + it] = y[k]; }
                    if (blk == 0) cta_rstep[q * n_sh + j] = len;
#pragma unroll
                    for (int r = 0; r < 6; ++r)
#pragma unroll
                        for (int cc = r; cc < 6; ++cc) {
                            double v = y[r] * y[cc];
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                            if (lane == 0) { const double t = Cq[r * 6 + cc] + v; Cq[r * 6 + cc] = t; Cq[cc * 6 + r] = t; }
                        }
                    n_ret += 16;
                }
                __syncwarp();
            }
            __syncwarp();                                      // every lane has read the slot's counters (racecheck: read-before-write inside the warp)
            if (lane == 0) {
                slot[q].n_ret = n_ret; slot[q].log_len = len;
                if (len >= a.log_cap && !slot[q].finishing) { slot[q].need_flush = 1; s_service = 1; }
            }
        }
        __syncthreads();
        // ---- commit the attempts accepted in the previous round (FSAL): base lanes only ----
        if (my_slot < NPX && col == -1 && slot[my_slot].part >= 0 && slot[my_slot].accepted) {
            BaseShared<S>& r = sb[my_slot];
#pragma unroll
            for (int k = 0; k < 3; ++k) { x[k] = x1[k]; p[k] = p1[k]; F[0][k] = F[S - 1][k]; r.X[0][k] = r.X[S - 1][k]; }
#pragma unroll
            for (int k = 0; k < 6; ++k) r.T[0][k] = r.T[S - 1][k];
            r.t[0] = r.t[S - 1];
        }
        // ---- service: write out finished particles, refill their slots from the queue, start the new particles ----
        while (s_service) {                                    // uniform
            __syncthreads();
            for (int q = 0; q < NP; ++q) {
                const bool fin = slot[q].part >= 0 && slot[q].finishing == 1;
                const bool flush = slot[q].part >= 0 && slot[q].need_flush && !fin;
                if (!fin && !flush) continue;                  // uniform
                // retired items: bring the states frozen at retirement to the current time with the suffix products of the logged propagators
                const int n_dead = slot[q].n_dead, n_ret = slot[q].n_ret, len = slot[q].log_len;
                double* Lq = cta_log + (size_t)q * a.log_cap * 36;
                const int* rs = cta_rstep + q * n_sh;
                const bool apply = slot[q].retire_on && n_ret > n_dead && len > 0;
                if (apply) suffix_products(Lq, len, sLogBlk, sR);
                if (flush) {                                   // log full: apply now, restart the log at the current step
                    double* b0 = cta_buf + (size_t)(2 * q) * 6 * n_items;
                    double* b1 = b0 + (size_t)6 * n_items;
                    for (int idx = tid; idx < 2 * (n_ret - n_dead); idx += blockDim.x) {
                        const int blk = idx >= (n_ret - n_dead), j = n_dead + idx - blk * (n_ret - n_dead), it = blk * n_sh + j;
                        const int i = rs[j];
                        if (apply && i < len) {
                            double y[6], o6[6];
#pragma unroll
                            for (int k = 0; k < 6; ++k) y[k] = b0[(size_t)k * n_items + it];
                            const double* R = Lq + (size_t)i * 36;
#pragma unroll
                            for (int r = 0; r < 6; ++r) {
                                double acc = 0.0;
#pragma unroll
                                for (int k = 0; k < 6; ++k) acc = fma(R[r * 6 + k], y[k], acc);
                                o6[r] = acc;
                            }
#pragma unroll
                            for (int k = 0; k < 6; ++k) { b0[(size_t)k * n_items + it] = o6[k]; b1[(size_t)k * n_items + it] = o6[k]; }
                        }
                    }
                    __syncthreads();
                    for (int j = n_dead + tid; j < n_ret; j += blockDim.x) cta_rstep[q * n_sh + j] = 0;
                    if (tid == 0) { slot[q].log_len = 0; slot[q].need_flush = 0; }
                    continue;
                }
                // outputs: final state if the end was reached, +inf otherwise (diffrax SaveAt semantics)
                const long long part = slot[q].part;
                const double dir = slot[q].dir;
                const bool ok = (slot[q].status == 0) && (slot[q].T0 < slot[q].T1);
                const double* cur = cta_buf + (size_t)(2 * q + slot[q].flip) * 6 * n_items;
                for (int it = tid; it < n_items; it += blockDim.x) {
                    const int j = it % n_sh, blk = it / n_sh, o = a.order[j];
                    double v[6];
#pragma unroll
                    for (int k = 0; k < 6; ++k) v[k] = cur[(size_t)k * n_items + it];
                    if (apply && j >= n_dead && j < n_ret) {
                        const int i = rs[j];
                        if (i < len) {
                            const double* R = Lq + (size_t)i * 36;
                            double o6[6];
#pragma unroll
                            for (int r = 0; r < 6; ++r) {
                                double acc = 0.0;
#pragma unroll
                                for (int k = 0; k < 6; ++k) acc = fma(R[r * 6 + k], v[k], acc);
                                o6[r] = acc;
                            }
#pragma unroll
                            for (int k = 0; k < 6; ++k) v[k] = o6[k];
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        const double w = k >= 3 ? v[k] * dir : v[k];
                        a.Dout[((size_t)part * n_sh + o) * 12 + blk * 6 + k] = ok ? w : inf;
                    }
                }
                if (tid == base_tid(q)) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) { a.wout[6 * part + k] = ok ? x[k] : inf; a.wout[6 * part + 3 + k] = ok ? dir * p[k] : inf; }
                    a.status[part] = slot[q].status;
                    a.nsteps[3 * part] = slot[q].n_steps; a.nsteps[3 * part + 1] = slot[q].n_acc; a.nsteps[3 * part + 2] = slot[q].n_rej;
                }
            }
            __syncthreads();
            if (tid == 0) {
                s_service = 0;
                for (int q = 0; q < NP; ++q) {
                    RespSlot& s = slot[q];
                    if (s.part >= 0 && s.finishing == 1) { s.part = -1; s.finishing = 0; s.accepted = 0; }
                    if (s.part == -1) {
                        const long long nx = (long long)atomicAdd(a.counter, 1ULL);
                        s.part = nx < a.N ? nx : -2;
                        s.finishing = s.part >= 0 ? 2 : 0;        // 2: needs start-up
                        s.no_retire = 0;
                    }
                }
            }
            __syncthreads();
            for (int q = 0; q < NP; ++q) {
                if (!(slot[q].part >= 0 && slot[q].finishing == 2)) continue;
                // ================= start-up of slot q: state load + HNW initial step over the whole coupled state =================
                const long long part = slot[q].part;
                const double t0_in = a.t0[part], t1_in = a.t1;
                const double dir = (t0_in < t1_in) ? 1.0 : -1.0;
                const double T0 = t0_in * dir, T1 = t1_in * dir;
                double* cur = cta_buf + (size_t)(2 * q) * 6 * n_items;
                double* nxt = cur + (size_t)6 * n_items;
                for (int it = tid; it < n_items; it += blockDim.x) {
                    const int j = it % n_sh, blk = it / n_sh, o = a.order[j];
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        double v = a.D0 ? a.D0[((size_t)part * n_sh + o) * 12 + blk * 6 + k] : 0.0;
                        if (k >= 3) v *= dir;
                        cur[(size_t)k * n_items + it] = v;
                        nxt[(size_t)k * n_items + it] = v;         // unborn / dead subhalos are never written again: both buffers hold their zeros
                    }
                }
                BaseForce<S, SIG> bforce{&sP, &Pin, &sb[q], dir, 0};
                double d0s = 0.0, d1s = 0.0;
                if (tid == base_tid(q)) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) { x[k] = a.w0[6 * part + k]; p[k] = dir * a.w0[6 * part + 3 + k]; }
                    bforce.stage = 0;
                    bforce(x, T0, F[0]);
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const double sx = fma(c.rtol, fabs(x[k]), c.atol), sp = fma(c.rtol, fabs(p[k]), c.atol);
                        double r;
                        r = x[k] / sx; d0s = fma(r, r, d0s); r = p[k] / sp; d0s = fma(r, r, d0s);
                        r = p[k] / sx; d1s = fma(r, r, d1s); r = F[0][k] / sp; d1s = fma(r, r, d1s);
                    }
                }
                __syncthreads();
                for (int it = tid; it < n_items; it += blockDim.x) {
                    ItemParams ip; load_item_params(a.sorted, Sh.profile, it % n_sh, it / n_sh, n_sh, ip);
                    ItemForce<S> f{&sb[q], &ip, 0};
                    double qq[3], pp[3], G[3];
#pragma unroll
                    for (int k = 0; k < 3; ++k) { qq[k] = cur[(size_t)k * n_items + it]; pp[k] = cur[(size_t)(3 + k) * n_items + it]; }
                    f.at(0, qq, G);
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const double sx = fma(c.rtol, fabs(qq[k]), c.atol), sp = fma(c.rtol, fabs(pp[k]), c.atol);
                        double r;
                        r = qq[k] / sx; d0s = fma(r, r, d0s); r = pp[k] / sp; d0s = fma(r, r, d0s);
                        r = pp[k] / sx; d1s = fma(r, r, d1s); r = G[k] / sp; d1s = fma(r, r, d1s);
                    }
                }
                const double d0 = sqrt(block_sum(d0s, sred) / ncomp);
                const double d1 = sqrt(block_sum(d1s, sred) / ncomp);
                const double h0 = hnw_h0(d0, d1);
                double d2s = 0.0;
                if (tid == base_tid(q)) {
                    double X1[3], F1[3];
#pragma unroll
                    for (int k = 0; k < 3; ++k) X1[k] = fma(h0, p[k], x[k]);
                    bforce.stage = 1;
                    bforce(X1, T0 + h0, F1);
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const double sx = fma(c.rtol, fabs(x[k]), c.atol), sp = fma(c.rtol, fabs(p[k]), c.atol);
                        double r;
                        r = (fma(h0, F[0][k], p[k]) - p[k]) / sx; d2s = fma(r, r, d2s);
                        r = (F1[k] - F[0][k]) / sp; d2s = fma(r, r, d2s);
                    }
                }
                __syncthreads();
                for (int it = tid; it < n_items; it += blockDim.x) {
                    ItemParams ip; load_item_params(a.sorted, Sh.profile, it % n_sh, it / n_sh, n_sh, ip);
                    ItemForce<S> f{&sb[q], &ip, 0};
                    double qq[3], pp[3], G0[3], G1[3], q1[3];
#pragma unroll
                    for (int k = 0; k < 3; ++k) { qq[k] = cur[(size_t)k * n_items + it]; pp[k] = cur[(size_t)(3 + k) * n_items + it]; }
                    f.at(0, qq, G0);
#pragma unroll
                    for (int k = 0; k < 3; ++k) q1[k] = fma(h0, pp[k], qq[k]);
                    f.at(1, q1, G1);
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const double sx = fma(c.rtol, fabs(qq[k]), c.atol), sp = fma(c.rtol, fabs(pp[k]), c.atol);
                        double r;
                        r = (fma(h0, G0[k], pp[k]) - pp[k]) / sx; d2s = fma(r, r, d2s);
                        r = (G1[k] - G0[k]) / sp; d2s = fma(r, r, d2s);
                    }
                }
                const double d2 = sqrt(block_sum(d2s, sred) / ncomp) / h0;
                if (tid == 0) {
                    RespSlot& s = slot[q];
                    double h = fmin(hnw_h1<T::ORDER>(h0, d1, d2), c.dtmax);
                    s.at_dtmin = h <= c.dtmin;
                    h = fmax(h, c.dtmin);
                    s.dir = dir; s.T0 = T0; s.T1 = T1; s.tprev = T0; s.tnext = fmin(T0 + h, T1);
                    s.status = 0; s.n_steps = s.n_acc = s.n_rej = 0; s.accepted = 0; s.flip = 0; s.n_act_run = 0;
                    s.skip = (a.skip_unborn && dir > 0.0) ? 1 : 0;
                    int nd = 0;       // subhalos whose window closed before the release of this particle: exact zeros throughout (see the header)
                    if (s.skip) { int lo = 0, hi = n_sh; while (lo < hi) { const int mid = (lo + hi) >> 1; if (a.endmax[mid] <= T0) lo = mid + 1; else hi = mid; } nd = lo & ~15; }
                    s.n_dead = nd;
                    s.n_ret = nd; s.log_len = 0; s.need_flush = 0;
                    s.retire_on = (s.skip && a.retire && c.atol > 0.0 && !s.no_retire) ? 1 : 0;
                    for (int e = 0; e < 36; ++e) sC[q][e] = 0.0;
                    s_guard[q] = 0;
                    s.finishing = 0;
                    if (!(T0 < T1)) { s.finishing = 1; s_service = 1; }                    // nothing to integrate: rows stay +inf
                    else if (c.max_steps <= 0) { s.status = 1; s.finishing = 1; s_service = 1; }
                }
                __syncthreads();
            }
            if (tid == 0) {
                int live = 0;
                for (int q = 0; q < NP; ++q) live |= slot[q].part >= 0;
                s_live = live;
            }
            __syncthreads();
        }
        if (!s_live) break;                                    // uniform
        // ---- serial phase, shared by the slots: base orbits + propagator columns of this round's attempts (warp 0) ----
        if (4 * wid < NP) {                                    // every warp advances the base orbits of its (up to four) slots
            __syncwarp();                                      // the FSAL commit of the base lanes (stage-0 record) is visible to the column lanes
            // all eight lanes of a group evaluate the group leader's point and store the SAME record; a group without a particle
            // works on its own scratch record
            const bool act = my_slot < NPX && slot[my_slot].part >= 0;
            const int rec = act ? my_slot : NPX + 4 * wid + (lane >> 3);
            const double tp = act ? slot[rec].tprev : 0.0, dt = act ? slot[rec].tnext - tp : 1.0, dir = act ? slot[rec].dir : 1.0;
            if (col >= 0) {
#pragma unroll
                for (int k = 0; k < 3; ++k) { x[k] = (col == k) ? 1.0 : 0.0; p[k] = (col == 3 + k) ? 1.0 : 0.0; }
                const double* T0m = sb[rec].T[0];
                F[0][0] = T0m[0] * x[0] + T0m[3] * x[1] + T0m[4] * x[2];
                F[0][1] = T0m[3] * x[0] + T0m[1] * x[1] + T0m[5] * x[2];
                F[0][2] = T0m[4] * x[0] + T0m[5] * x[1] + T0m[2] * x[2];
            }
            double ex[3], ep[3];
            BaseGroupForce<S, SIG> gforce{&sP, &Pin, &sb[rec], dir, 1, lane & ~7, col == -1};
            // stage 0 of the record (X, T, t at tprev) is already in place: FSAL copy above / start-up evaluation
            rk_stages<SOLVER>(gforce, x, p, tp, dt, F);
            rk_candidate<SOLVER>(x, p, dt, F, x1, p1);
            gforce.stage = S - 1;
            gforce(x1, tp + T::c(S - 1) * dt, F[S - 1]);
            rk_error<SOLVER>(p, dt, F, ex, ep);
            if (act && col == -1) {
                bool nan_cand = false, finite = true;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    nan_cand |= isnan(x1[k]) | isnan(p1[k]);
                    finite &= isfinite(x1[k]) & isfinite(p1[k]);
                }
                if (!finite) s_bad[rec] = 1;
                s_besq[rec] = err_sq6(x, p, x1, p1, ex, ep, c.rtol, c.atol, nan_cand);
                // number of subhalos whose window start lies before the end of this attempt (sorted ascending)
                int na = n_sh;
                if (slot[rec].skip) {
                    const double tn = slot[rec].tnext;
                    int lo = 0, hi = n_sh;
                    while (lo < hi) { const int mid = (lo + hi) >> 1; if (a.start[mid] < tn) lo = mid + 1; else hi = mid; }
                    na = lo;
                }
                s_nact[rec] = na;
            } else if (act && col >= 0 && col < 6) {
                double* PE = sPhiE[rec];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    PE[k * 6 + col] = x1[k]; PE[(3 + k) * 6 + col] = p1[k];
                    PE[36 + k * 6 + col] = ex[k]; PE[36 + (3 + k) * 6 + col] = ep[k];
                }
            }
        }
        __syncthreads();
        // ---- item sweeps of the slots, one after the other, on all threads ----
#pragma unroll 1
        for (int q = 0; q < NP; ++q) {
            if (slot[q].part < 0) continue;                    // uniform
            // monotone within a particle: a subhalo touched by a (possibly rejected) attempt has a candidate in `nxt` that must be
            // overwritten by every later attempt, even one that ends before its window opens
            const int n_act = max(slot[q].n_act_run, s_nact[q]);
            const double* cur = cta_buf + (size_t)(2 * q + slot[q].flip) * 6 * n_items;
            double* nxt = cta_buf + (size_t)(2 * q + (slot[q].flip ^ 1)) * 6 * n_items;
            double esq = (tid == 0) ? s_besq[q] : 0.0;          // same summation order as response_kernel, whatever the slot
            int bad_local = 0;
            sweep_items<SOLVER, PROFILE>(&sb[q], sPhiE[q], a.sorted, n_sh, n_items, slot[q].n_ret, n_act, cur, nxt, slot[q].tnext - slot[q].tprev, c, esq,
                                         bad_local, slot[q].retire_on ? (int)(slot[q].part & 3) : 0);
            if (bad_local) s_bad[q] = 1;
            if (slot[q].n_ret > slot[q].n_dead && wid == (nw > 1 ? 1 : 0) && lane < 6) {
                // retired items: sum over them of (E y)_k^2 = (E C E^T)_kk, scale atol (see the header); always by the same lanes, whatever the
                // slot, so that a particle's error sum does not depend on where it runs
                const double* E = sPhiE[q] + 36;
                const double* Cq = sC[q];
                double v = 0.0;
#pragma unroll
                for (int cc = 0; cc < 6; ++cc) {
                    double u = 0.0;
#pragma unroll
                    for (int l = 0; l < 6; ++l) u = fma(E[lane * 6 + l], Cq[l * 6 + cc], u);
                    v = fma(u, E[lane * 6 + cc], v);
                }
                esq += fmax(v, 0.0) * inv_atol2;
                if (c.rtol * c.rtol * Cq[lane * 6 + lane] > guard_lim) s_guard[q] = 1;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) esq += __shfl_xor_sync(0xffffffffu, esq, o);
            if (lane == 0) sred_q[q][wid] = esq;
        }
        __syncthreads();
        // ---- controllers (thread q serves slot q): accept / reject, next step, completion ----
        if (tid < NP && slot[tid].part >= 0) {
            const int q = tid;
            RespSlot& s = slot[q];
            double tot = 0.0;
            for (int i = 0; i < nw; ++i) tot += sred_q[q][i];
            const double err = sqrt(tot / ncomp);
            const double dt = s.tnext - s.tprev;
            const int any_bad = s_bad[q];
            s_bad[q] = 0;
            if (s_guard[q]) {                              // the atol-only scale of the retired items is not good enough here: start this particle over
                s_guard[q] = 0;                            // with every item swept (exact scales)
                s.no_retire = 1; s.accepted = 0; s.finishing = 2; s_service = 1;
            } else {
                s.n_act_run = max(s.n_act_run, s_nact[q]);
                double hn; bool bad;
                bool at_dtmin = s.at_dtmin != 0;
                const bool keep = pid_update<T::ORDER>(err, dt, c, at_dtmin, hn, bad);
                s.at_dtmin = at_dtmin;
                s.n_steps++;
                s.accepted = 0;
                if (bad) { s.status = 2; s.n_rej++; }
                else if (keep) {
                    s.n_acc++;
                    if (any_bad) s.status = 2;
                    else { s.flip ^= 1; s.accepted = 1; s.tprev = s.tnext; }
                } else s.n_rej++;
                if (s.status == 0) {
                    s.tprev = fmin(s.tprev, s.T1);
                    double tn = s.tprev + hn;
                    if (tn > s.T1 - 1e-10) tn = keep ? s.T1 : s.tprev + 0.5 * (s.T1 - s.tprev);
                    s.tnext = tn;
                    if (!(s.tprev < s.T1)) s.finishing = 1;
                    else if (s.n_steps >= c.max_steps) { s.status = 1; s.finishing = 1; }
                } else s.finishing = 1;
                if (s.finishing) s_service = 1;
            }
        }
    }
}

// warp-autonomous form of the kernel: measured slower (51.9 vs 31.4 ms) and therefore not built by default;
// -DSSB_RESP_WA=1 (tools/build_variants.sh) adds it, SSB_RESP_KERNEL=wa selects it, the bit-identity test covers it (SSB_TEST_RESP_WA=1)
#ifndef SSB_RESP_WA
#define SSB_RESP_WA 0
#endif
#if SSB_RESP_WA
#include "ssb_response_wa.cuh"
#endif

// RHS of the coupled field at one state (fields.py:175-206): y = [w(6), D(n_sh,12)]
__global__ void response_term_kernel(const __grid_constant__ ssb_potential Pin, const ssb_subhalos Sh, double t, const double* y, double* dy) {
    __shared__ ssb_potential sP;
    __shared__ BaseShared<1> sb;
    stage_potential(&sP, &Pin);
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const double3 acc = base_force_call<1, SIG_GENERIC>(&sP, &Pin, &sb, 0, y[0], y[1], y[2], t);
        dy[0] = y[3]; dy[1] = y[4]; dy[2] = y[5]; dy[3] = acc.x; dy[4] = acc.y; dy[5] = acc.z;
    } else if (threadIdx.x == 0) {
        base_force_call<1, SIG_GENERIC>(&sP, &Pin, &sb, 0, y[0], y[1], y[2], t);
    }
    __syncthreads();
    const int n_sh = Sh.n;
    for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < 2 * n_sh; it += gridDim.x * blockDim.x) {
        const int j = it % n_sh, blk = it / n_sh;
        ItemParams ip;
        ip.blk = blk; ip.profile = Sh.profile; ip.GM = Sh.G * Sh.m[j]; ip.rs = Sh.rs[j]; ip.t0 = Sh.t0[j]; ip.tw = Sh.tw[j];
        for (int k = 0; k < 3; ++k) { ip.x0[k] = Sh.x0[3 * j + k]; ip.v[k] = Sh.v[3 * j + k]; }
        ItemForce<1> f{&sb, &ip, 0};
        const double* d = y + 6 + 12 * j + 6 * blk;
        double* o = dy + 6 + 12 * j + 6 * blk;
        const double Q[3] = {d[0], d[1], d[2]};
        double A[3];
        f.at(0, Q, A);
        o[0] = d[3]; o[1] = d[4]; o[2] = d[5]; o[3] = A[0]; o[4] = A[1]; o[5] = A[2];
    }
}

static int resp_grid(int64_t N, int np = 1) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t g = (int64_t)sms * SSB_RESP_CTAS_PER_SM, want = (N + np - 1) / np;
    return (int)(want < g ? want : g);
}

// particles in flight per CTA (response_kernel_mp).  SSB_RESP_NP in the environment overrides the default for A/B runs:
// 0 = the one-particle-per-CTA kernel, 1..4 = slots per CTA.
static int resp_np(int64_t N, int grid_cap) {
    if (const char* e = getenv("SSB_RESP_NP")) {                   // explicit choice: taken as is (tests force 2 and 4 slots on small batches)
        const int np = atoi(e);
        return np <= 0 ? 0 : (np > SSB_RESP_MAX_NP ? SSB_RESP_MAX_NP : np);
    }
    int np = SSB_RESP_DEFAULT_NP;
    while (np > 1 && N < (int64_t)grid_cap * np) np >>= 1;         // too few particles to fill every slot of every CTA: fewer slots, more CTAs
    return np;
}

extern "C" {

// scratch layout: [256 B header: work counter] [sorted table 10*n] [start n] [endmax n] [order n (int)] [ping-pong state per persistent CTA]
#define SSB_RESP_MAX_CTAS (160 * SSB_RESP_CTAS_PER_SM)
static size_t resp_table_bytes(int32_t n_sh) { const size_t n = (size_t)(n_sh > 0 ? n_sh : 1); return ((sizeof(double) * 12 * n + sizeof(int) * n + 255) / 256) * 256; }
#define SSB_RESP_LOG_CAP 512        // logged step propagators per particle slot before the log is applied and restarted
// per particle slot: ping-pong item state, the log of step propagators and the retirement positions (rounded to 256 bytes)
static size_t resp_state_bytes(int32_t n_sh) { return sizeof(double) * 2 * 6 * 2 * (size_t)(n_sh > 0 ? n_sh : 1); }
static size_t resp_log_bytes() { return sizeof(double) * 36 * (size_t)SSB_RESP_LOG_CAP; }
static size_t resp_rstep_bytes(int32_t n_sh) { return ((sizeof(int) * (size_t)(n_sh > 0 ? n_sh : 1) + 255) / 256) * 256; }
size_t ssb_response_scratch_bytes(int32_t n_sh) {
    const size_t per_slot = resp_state_bytes(n_sh) + resp_log_bytes() + resp_rstep_bytes(n_sh);
    return 256 + resp_table_bytes(n_sh) + per_slot * (size_t)SSB_RESP_MAX_NP * (size_t)SSB_RESP_MAX_CTAS;
}
static bool resp_retire_enabled() {
    const char* e = getenv("SSB_RESP_RETIRE");
    return !(e && e[0] == '0');
}

static int response_impl(const ssb_potential* pot_base, const ssb_subhalos* sh, int64_t N, const double* w0, const double* D0,
                         const double* t0, double t1, ssb_ctrl ctrl, double* wout, double* Dout, int32_t* status, int32_t* nsteps,
                         void* scratch, size_t scratch_bytes, void* stream, const double* ts_save, int M, double* wsave, double* Dsave) {
    if (int e = ssb_validate_potential(pot_base)) return e;
    if (int e = ssb_validate_ctrl(ctrl)) return e;
    if (!sh || sh->n < 0) return ssb_set_error(SSB_ERR_ARG, "linear_response: bad subhalo set");
    if (sh->profile < SSB_PROFILE_PLUMMER || sh->profile > SSB_PROFILE_NFW) return ssb_set_error(SSB_ERR_UNSUPPORTED, "linear_response: unknown profile");
    if (N < 0) return ssb_set_error(SSB_ERR_ARG, "linear_response: negative N");
    if (N == 0) return 0;
    if (!w0 || !t0 || !wout || !status || !nsteps || !scratch || (sh->n > 0 && !Dout)) return ssb_set_error(SSB_ERR_ARG, "linear_response: NULL array");
    if (scratch_bytes < ssb_response_scratch_bytes(sh->n)) return ssb_set_error(SSB_ERR_SCRATCH, "linear_response: scratch too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int grid_cap = resp_grid(INT64_MAX / 8);
    const int np = (M > 0) ? 0 : resp_np(N, grid_cap);
    const int grid = resp_grid(N, np > 0 ? np : 1);
    if (grid > SSB_RESP_MAX_CTAS) return ssb_set_error(SSB_ERR_SCRATCH, "linear_response: more SMs than the scratch layout assumes");
    RespArgs a;
    a.N = N; a.w0 = w0; a.D0 = D0; a.t0 = t0; a.t1 = t1; a.c.rtol = ctrl.rtol; a.c.atol = ctrl.atol; a.c.dtmin = ctrl.dtmin; a.c.dtmax = ctrl.dtmax;
    a.c.max_steps = ctrl.max_steps; a.wout = wout; a.Dout = Dout; a.status = status; a.nsteps = nsteps;
    a.counter = (unsigned long long*)scratch;
    double* tab = (double*)((char*)scratch + 256);
    double* start = tab + (size_t)10 * sh->n;
    double* endmax = start + sh->n;
    int* order = (int*)(endmax + sh->n);
    a.sorted = tab; a.start = start; a.endmax = endmax; a.order = order;
    a.scratch = (double*)((char*)scratch + 256 + resp_table_bytes(sh->n));
    {
        const size_t slots = (size_t)SSB_RESP_MAX_NP * (size_t)SSB_RESP_MAX_CTAS;
        a.plog = (double*)((char*)a.scratch + resp_state_bytes(sh->n) * slots);
        a.rstep = (int*)((char*)a.plog + resp_log_bytes() * slots);
        a.log_cap = SSB_RESP_LOG_CAP;
    }
    a.skip_unborn = (D0 == nullptr && M == 0) ? 1 : 0;
    a.retire = (a.skip_unborn && np > 0 && resp_retire_enabled()) ? 1 : 0;
    a.np = np;
    a.ts_save = ts_save; a.M = M; a.wsave = wsave; a.Dsave = Dsave;
    CK(cudaMemsetAsync(scratch, 0, 256, st));
    if (sh->n > 0) {
        if (sh->n <= SSB_RESP_MAX_SORT) {
            int npad = 1; while (npad < sh->n) npad <<= 1;
            response_sort_kernel<<<1, 1024, (size_t)npad * 12, st>>>(*sh, npad, order, start, endmax);
            CKL("response_sort_kernel");
        } else {                                   // identity order; every subhalo counts as born
            response_identity_kernel<<<(sh->n + 127) / 128, 128, 0, st>>>(*sh, order, start, endmax);
            CKL("response_identity_kernel");
        }
        response_gather_kernel<<<(sh->n + 127) / 128, 128, 0, st>>>(*sh, order, tab);
        CKL("response_gather_kernel");
    }
    ssb_potential pc;
    const int sig = ssb_canonicalize(pot_base, &pc);
    // base potentials of the response path: MW3 / Gala fused, everything else (incl. a lone NFW) through the interpreter
#define SSB_LAUNCH_RESP(S, SG, PR) response_kernel<S, SG, PR, false><<<grid, SSB_RESP_THREADS, 0, st>>>(sig == SG ? pc : *pot_base, *sh, a)
#define SSB_LAUNCH_RESP_SIG(S, PR) do { switch (sig) { case SIG_NHM: SSB_LAUNCH_RESP(S, SIG_NHM, PR); break; \
        case SIG_NHHM: SSB_LAUNCH_RESP(S, SIG_NHHM, PR); break; default: SSB_LAUNCH_RESP(S, SIG_GENERIC, PR); } } while (0)
#define SSB_LAUNCH_RESP_PR(S) do { switch (sh->profile) { case SSB_PROFILE_PLUMMER: SSB_LAUNCH_RESP_SIG(S, SSB_PROFILE_PLUMMER); break; \
        case SSB_PROFILE_HERNQUIST: SSB_LAUNCH_RESP_SIG(S, SSB_PROFILE_HERNQUIST); break; default: SSB_LAUNCH_RESP_SIG(S, SSB_PROFILE_NFW); } } while (0)
    if (M > 0) {          // single saved trajectory: interpreter base force, runtime profile inside save_items; one CTA
#define SSB_LAUNCH_SAVE(S, PR) response_kernel<S, SIG_GENERIC, PR, true><<<1, SSB_RESP_THREADS, 0, st>>>(*pot_base, *sh, a)
#define SSB_LAUNCH_SAVE_PR(S) do { switch (sh->profile) { case SSB_PROFILE_PLUMMER: SSB_LAUNCH_SAVE(S, SSB_PROFILE_PLUMMER); break; \
        case SSB_PROFILE_HERNQUIST: SSB_LAUNCH_SAVE(S, SSB_PROFILE_HERNQUIST); break; default: SSB_LAUNCH_SAVE(S, SSB_PROFILE_NFW); } } while (0)
        if (ctrl.solver == 5) SSB_LAUNCH_SAVE_PR(5); else SSB_LAUNCH_SAVE_PR(8);
    } else if (np > 0) {      // several particles in flight per CTA (shared serial phase)
        // dynamic shared memory of the multi-slot kernel: stage records, propagators, moment matrices of SSB_RESP_MAX_NP slots
#define SSB_LAUNCH_MP(S, SG, PR) do { const size_t shm = (sizeof(BaseShared<((S) == 5 ? 7 : 14)>) * 2 + sizeof(double) * (72 + 72)) * SSB_RESP_MAX_NP; \
        CK(cudaFuncSetAttribute(response_kernel_mp<S, SG, PR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm)); \
        response_kernel_mp<S, SG, PR><<<grid, SSB_RESP_THREADS, shm, st>>>(sig == SG ? pc : *pot_base, *sh, a); } while (0)
#define SSB_LAUNCH_MP_SIG(S, PR) do { switch (sig) { case SIG_NHM: SSB_LAUNCH_MP(S, SIG_NHM, PR); break; \
        case SIG_NHHM: SSB_LAUNCH_MP(S, SIG_NHHM, PR); break; default: SSB_LAUNCH_MP(S, SIG_GENERIC, PR); } } while (0)
#define SSB_LAUNCH_MP_PR(S) do { switch (sh->profile) { case SSB_PROFILE_PLUMMER: SSB_LAUNCH_MP_SIG(S, SSB_PROFILE_PLUMMER); break; \
        case SSB_PROFILE_HERNQUIST: SSB_LAUNCH_MP_SIG(S, SSB_PROFILE_HERNQUIST); break; default: SSB_LAUNCH_MP_SIG(S, SSB_PROFILE_NFW); } } while (0)
#if SSB_RESP_WA
        // the warp-autonomous form (ssb_response_wa.cuh): only on request, SSB_RESP_KERNEL=wa (A/B, tests)
#define SSB_LAUNCH_WA(S, SG, PR) do { const size_t shm = (sizeof(BaseShared<((S) == 5 ? 7 : 14)>) * 2 + sizeof(double) * (72 + 36)) * SSB_RESP_MAX_NP \
                                                        + sizeof(double) * 36 * (SSB_RESP_THREADS / 32); \
        CK(cudaFuncSetAttribute(response_kernel_wa<S, SG, PR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm)); \
        response_kernel_wa<S, SG, PR><<<grid, SSB_RESP_THREADS, shm, st>>>(sig == SG ? pc : *pot_base, *sh, a); } while (0)
#define SSB_LAUNCH_WA_SIG(S, PR) do { switch (sig) { case SIG_NHM: SSB_LAUNCH_WA(S, SIG_NHM, PR); break; \
        case SIG_NHHM: SSB_LAUNCH_WA(S, SIG_NHHM, PR); break; default: SSB_LAUNCH_WA(S, SIG_GENERIC, PR); } } while (0)
#define SSB_LAUNCH_WA_PR(S) do { switch (sh->profile) { case SSB_PROFILE_PLUMMER: SSB_LAUNCH_WA_SIG(S, SSB_PROFILE_PLUMMER); break; \
        case SSB_PROFILE_HERNQUIST: SSB_LAUNCH_WA_SIG(S, SSB_PROFILE_HERNQUIST); break; default: SSB_LAUNCH_WA_SIG(S, SSB_PROFILE_NFW); } } while (0)
        const char* kern = getenv("SSB_RESP_KERNEL");
        if (kern && kern[0] == 'w') { if (ctrl.solver == 5) SSB_LAUNCH_WA_PR(5); else SSB_LAUNCH_WA_PR(8); }
        else
#endif
        { if (ctrl.solver == 5) SSB_LAUNCH_MP_PR(5); else SSB_LAUNCH_MP_PR(8); }
    } else {
        if (ctrl.solver == 5) SSB_LAUNCH_RESP_PR(5); else SSB_LAUNCH_RESP_PR(8);
    }
    CKL("response_kernel");
    return 0;
}

int ssb_linear_response_f64(const ssb_potential* pot_base, const ssb_subhalos* sh, int64_t N, const double* w0, const double* D0,
                            const double* t0, double t1, ssb_ctrl ctrl, double* wout, double* Dout, int32_t* status, int32_t* nsteps,
                            void* scratch, size_t scratch_bytes, void* stream) {
    return response_impl(pot_base, sh, N, w0, D0, t0, t1, ctrl, wout, Dout, status, nsteps, scratch, scratch_bytes, stream, nullptr, 0, nullptr, nullptr);
}

size_t ssb_response_saveat_scratch_bytes(int32_t n_sh) {
    return ssb_response_scratch_bytes(n_sh) + sizeof(double) * (8 + 12 * (size_t)(n_sh > 0 ? n_sh : 1)) + 256;
}

int ssb_linear_response_saveat_f64(const ssb_potential* pot_base, const ssb_subhalos* sh, const double* w0, const double* D0, const double* t0,
                                   double t1, const double* ts, int32_t M, ssb_ctrl ctrl, double* ws, double* Ds, int32_t* status, int32_t* nsteps,
                                   void* scratch, size_t scratch_bytes, void* stream) {
    if (!sh) return ssb_set_error(SSB_ERR_ARG, "linear_response_saveat: bad subhalo set");
    if (M <= 0 || !ts || !ws || (sh->n > 0 && !Ds)) return ssb_set_error(SSB_ERR_ARG, "linear_response_saveat: needs M > 0 save times and outputs");
    if (scratch_bytes < ssb_response_saveat_scratch_bytes(sh->n)) return ssb_set_error(SSB_ERR_SCRATCH, "linear_response_saveat: scratch too small");
    // the final state the common path always writes goes to the tail of the scratch buffer
    const size_t head = ssb_response_scratch_bytes(sh->n);
    double* wfin = (double*)((char*)scratch + head);
    double* Dfin = wfin + 8;
    return response_impl(pot_base, sh, 1, w0, D0, t0, t1, ctrl, wfin, Dfin, status, nsteps, scratch, head, stream, ts, M, ws, Ds);
}

int ssb_response_term_f64(const ssb_potential* pot_base, const ssb_subhalos* sh, double t, const double* y, double* dy, void* stream) {
    if (int e = ssb_validate_potential(pot_base)) return e;
    if (!sh || !y || !dy) return ssb_set_error(SSB_ERR_ARG, "response_term: NULL argument");
    const int items = 2 * sh->n;
    int grid = (items + 127) / 128; if (grid < 1) grid = 1;
    response_term_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(*pot_base, *sh, t, y, dy);
    CKL("response_term_kernel");
    return 0;
}

}  // extern "C"
