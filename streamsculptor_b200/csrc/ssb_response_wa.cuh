// K3, warp-autonomous form (round 2): included by ssb_response.cu after response_kernel_mp, whose helpers and slot state it shares.
//
// response_kernel_mp runs its phases CTA-wide: every warp advances the base orbits of its slots, then ALL threads sweep the items of slot 0,
// slot 1, ..., then the controllers run - with a __syncthreads() between the phases.  ncu had a fifth of the warp-stall samples on those
// barriers (the warps of a CTA never have the same amount of work: open windows differ from particle to particle) and, because every warp
// of a CTA is in the same phase, the latency-bound serial phase of one warp was never hidden behind the FP64-dense sweep of another.
//
// Here a WARP is the worker.  It owns up to four particle slots (slot q belongs to warp q % 4, lane group q / 4), and does everything for them:
// base orbits + propagator columns (lanes 8g .. 8g+6 of group g, one instruction stream for the four groups), the item sweeps (32 lanes),
// the error norms (shuffles), the controllers (lane g), retirement, start-up and write-out.  Nothing in the main loop crosses the warp:
// there is no block-wide barrier after the tables are staged, and the twelve warps of an SM drift apart, so that serial phases, sweeps and
// the L2 latency of item states overlap.
//
// Same arithmetic, in the same order, as response_kernel / response_kernel_mp: a lane keeps the four partial sums that the threads lane,
// 32 + lane, 64 + lane, 96 + lane of a 128-thread CTA would hold (Acc4) and they are combined the way block_sum combines them, so the error
// norms - and with them every step size and result - are bit-identical (tests/test_gpu_parity.py checks exactly that).

struct Acc4 {
    double a[4];
    __device__ __forceinline__ void zero() { a[0] = a[1] = a[2] = a[3] = 0.0; }
    __device__ __forceinline__ double get(int w) const { return w == 0 ? a[0] : (w == 1 ? a[1] : (w == 2 ? a[2] : a[3])); }
    __device__ __forceinline__ void set(int w, double v) {
        a[0] = w == 0 ? v : a[0]; a[1] = w == 1 ? v : a[1]; a[2] = w == 2 ? v : a[2]; a[3] = w == 3 ? v : a[3];
    }
    // block_sum's order: shuffle tree inside each (virtual) warp, then warp 0 + warp 1 + warp 2 + warp 3
    __device__ __forceinline__ double total() const {
        double tot = 0.0;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            double v = a[w];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            tot += v;
        }
        return tot;
    }
};

#define SSB_RESP_WA_LOG_BLK 8

// suffix_products on one warp: R_i = L[len-1] ... L[i], in place; sBlk [SSB_RESP_WA_LOG_BLK * 36], sR [2][36] belong to the warp
__device__ __forceinline__ void suffix_products_warp(double* __restrict__ L, int len, double* sBlk, double* sR) {
    const int lane = threadIdx.x & 31;
    for (int e = lane; e < 36; e += 32) sR[e] = (e / 6 == e % 6) ? 1.0 : 0.0;
    int cur = 0;
    for (int blk_end = len; blk_end > 0; blk_end -= SSB_RESP_WA_LOG_BLK) {
        const int b0 = blk_end > SSB_RESP_WA_LOG_BLK ? blk_end - SSB_RESP_WA_LOG_BLK : 0, nb = blk_end - b0;
        __syncwarp();
        for (int e = lane; e < nb * 36; e += 32) sBlk[e] = L[(size_t)b0 * 36 + e];
        __syncwarp();
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
        for (int e = lane; e < nb * 36; e += 32) L[(size_t)b0 * 36 + e] = sBlk[e];
    }
    __syncwarp();
}

// sweep_items on one warp.  Position idx of the sweep range is taken by lane idx % 32; its squared error goes to the partial sum of the
// virtual thread that response_kernel_mp would have given it to (warp ((idx / 32) - rot) % 4 of the CTA).
template <int SOLVER, int PROFILE>
__device__ __forceinline__ void sweep_items_warp(const BaseShared<Tab<SOLVER>::S>* sb, const double* __restrict__ PhiE, const double* __restrict__ tab, int n_sh,
                                                 int n_items, int n_first, int n_act, const double* __restrict__ cur, double* __restrict__ nxt, double dt,
                                                 const CtrlDev& c, Acc4& acc, int& bad_local, int rot) {
    constexpr int S = Tab<SOLVER>::S;
    const double t_lo = fmin(sb->t[0], sb->t[S - 1]), t_hi = fmax(sb->t[0], sb->t[S - 1]);
    const int n_span = n_act > n_first ? n_act - n_first : 0;
    const int n2 = 2 * n_span;
    const double* __restrict__ t0tab = tab + (size_t)8 * n_sh;
    const double* __restrict__ twtab = tab + (size_t)9 * n_sh;
    int idx = threadIdx.x & 31;
    double yn[6] = {0, 0, 0, 0, 0, 0}, t0n = 0.0, twn = 0.0;
    if (idx < n2) {
        const int blk = idx >= n_span, j = n_first + idx - blk * n_span, it = blk * n_sh + j;
#pragma unroll
        for (int k = 0; k < 6; ++k) yn[k] = cur[(size_t)k * n_items + it];
        t0n = __ldg(t0tab + j); twn = __ldg(twtab + j);
    }
    while (idx < n2) {
        const int blk = idx >= n_span, j = n_first + idx - blk * n_span, it = blk * n_sh + j;
        const double q[3] = {yn[0], yn[1], yn[2]}, pp[3] = {yn[3], yn[4], yn[5]};
        const double t0j = t0n, twj = twn;
        const int idn = idx + 32;
        if (idn < n2) {                                   // the state of the lane's next item is requested before this one is processed
            const int blkn = idn >= n_span, jn = n_first + idn - blkn * n_span, itn = blkn * n_sh + jn;
#pragma unroll
            for (int k = 0; k < 6; ++k) yn[k] = cur[(size_t)k * n_items + itn];
            t0n = __ldg(t0tab + jn); twn = __ldg(twtab + jn);
        }
        double q1[3], pp1[3], ex[3], ep[3];
        const bool closed = (t_lo - t0j >= twj) || (t0j - t_hi >= twj);
        if (closed) {
            const double y[6] = {q[0], q[1], q[2], pp[0], pp[1], pp[2]};
            const double2* __restrict__ M2 = reinterpret_cast<const double2*>(PhiE);
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
            if (SOLVER == 5) f.at(S - 1, q1, G[S - 1]);
            else { G[S - 1][0] = G[S - 1][1] = G[S - 1][2] = 0.0; }
            rk_error<SOLVER>(pp, dt, G, ex, ep);
        }
        const int vw = ((idx >> 5) - rot) & 3;
        double esq = acc.get(vw);
        item_finish(q, pp, q1, pp1, ex, ep, c, nxt, n_items, it, esq, bad_local);
        acc.set(vw, esq);
        idx = idn;
    }
}

template <int SOLVER, int SIG, int PROFILE>
__global__ void __launch_bounds__(SSB_RESP_THREADS, SSB_RESP_CTAS_PER_SM) response_kernel_wa(const __grid_constant__ ssb_potential Pin, const ssb_subhalos Sh,
                                                                                             const RespArgs a) {
    typedef Tab<SOLVER> T;
    constexpr int S = T::S;
    constexpr int NPX = SSB_RESP_MAX_NP;
    constexpr int NW = SSB_RESP_THREADS / 32;
    __shared__ ssb_potential sP;
    // dynamic shared memory (as response_kernel_mp): stage records [2 NPX] ([NPX + 4 w + g]: scratch record of lane group g of warp w while
    // it has no particle), propagator + error map [NPX][72], moment matrices sC [NPX][36], one 6x6 temporary per warp
    extern __shared__ __align__(16) unsigned char s_dyn[];
    BaseShared<S>* const sb = reinterpret_cast<BaseShared<S>*>(s_dyn);
    double (*const sPhiE)[72] = reinterpret_cast<double (*)[72]>(s_dyn + sizeof(BaseShared<S>) * 2 * NPX);
    double (*const sC)[36] = reinterpret_cast<double (*)[36]>(s_dyn + sizeof(BaseShared<S>) * 2 * NPX + sizeof(double) * 72 * NPX);
    double (*const sTw)[36] = sC + NPX;
    __shared__ RespSlot slot[NPX];
    __shared__ int s_nact[NPX], s_bad[NPX], s_guard[NPX];
    __shared__ double s_besq[NPX];
    __shared__ int s_service[NW], s_live[NW];
    __shared__ double sLogBlk[NW][SSB_RESP_WA_LOG_BLK * 36], sR[NW][2 * 36];
    stage_potential(&sP, &Pin);
    logtab_init();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int NP = a.np;
    const int n_sh = Sh.n, n_items = 2 * n_sh, ncomp = 6 + 12 * n_sh;
    const CtrlDev c = a.c;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    // lane = 8 * group + role; role 0 = base orbit, 1..6 = propagator columns, 7 = idle.  Group g of warp w carries slot w + 4 g.
    const int col = (lane & 7) - 1;
    const int grp = lane >> 3;
    const int my_slot = (wid + NW * grp < NP) ? wid + NW * grp : NPX;
    double x[3] = {8.0, 0.0, 0.0}, p[3] = {0, 0, 0}, F[S][3], x1[3] = {8.0, 0.0, 0.0}, p1[3] = {0, 0, 0};
#pragma unroll
    for (int l = 0; l < S; ++l) F[l][0] = F[l][1] = F[l][2] = 0.0;
    if (lane < 4) {
        const int q = wid + NW * lane;
        RespSlot& s = slot[q];
        s.part = q < NP ? -1 : -2; s.dir = 1.0; s.T0 = s.T1 = s.tprev = s.tnext = 0.0; s.status = s.n_steps = s.n_acc = s.n_rej = 0;
        s.at_dtmin = s.accepted = s.finishing = s.flip = s.n_act_run = s.n_dead = s.skip = s.pad = 0;
        s.n_ret = s.log_len = s.retire_on = s.no_retire = s.need_flush = s.pad2 = 0;
        s_nact[q] = 0; s_bad[q] = 0; s_besq[q] = 0.0; s_guard[q] = 0;
    }
    const double inv_atol2 = c.atol > 0.0 ? 1.0 / (c.atol * c.atol) : 0.0;
    const double guard_lim = 2.5e-7 * c.atol * c.atol;
    if (lane == 0) { s_service[wid] = 1; s_live[wid] = 1; }
    double* const cta_buf = a.scratch + (size_t)blockIdx.x * NPX * 2 * 6 * n_items;
    double* const cta_log = a.plog + (size_t)blockIdx.x * NPX * (size_t)a.log_cap * 36;
    int* const cta_rstep = a.rstep + (size_t)blockIdx.x * NPX * n_sh;
    double* const myLogBlk = sLogBlk[wid];
    double* const myR = sR[wid];
    double* const myT = sTw[wid];

    for (;;) {
        __syncwarp();
        // ---- retired items of the slots whose attempt was accepted: log the step's propagator, advance the moment matrix C <- Phi C Phi^T,
        //      then retire the chunks of 16 positions whose windows are now closed for good ----
#pragma unroll 1
        for (int q = wid; q < NP; q += NW) {
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
                        myT[e] = acc;
                    }
                }
                __syncwarp();
                if (lane < 21) {                               // C = T Phi^T, upper triangle mirrored: exactly symmetric
                    int r = 0, rem = lane;
                    while (rem >= 6 - r) { rem -= 6 - r; ++r; }
                    const int cc = r + rem;
                    double acc = 0.0;
#pragma unroll
                    for (int k = 0; k < 6; ++k) acc = fma(myT[r * 6 + k], PE[cc * 6 + k], acc);
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
            if (lane == 0) {
                slot[q].n_ret = n_ret; slot[q].log_len = len;
                if (len >= a.log_cap && !slot[q].finishing) { slot[q].need_flush = 1; s_service[wid] = 1; }
            }
        }
        __syncwarp();
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
        while (s_service[wid]) {                               // uniform over the warp
            __syncwarp();
#pragma unroll 1
            for (int q = wid; q < NP; q += NW) {
                const bool fin = slot[q].part >= 0 && slot[q].finishing == 1;
                const bool flush = slot[q].part >= 0 && slot[q].need_flush && !fin;
                if (!fin && !flush) continue;
                const int n_dead = slot[q].n_dead, n_ret = slot[q].n_ret, len = slot[q].log_len;
                double* Lq = cta_log + (size_t)q * a.log_cap * 36;
                const int* rs = cta_rstep + q * n_sh;
                const bool apply = slot[q].retire_on && n_ret > n_dead && len > 0;
                if (apply) suffix_products_warp(Lq, len, myLogBlk, myR);
                if (flush) {                                   // log full: apply now, restart the log at the current step
                    double* b0 = cta_buf + (size_t)(2 * q) * 6 * n_items;
                    double* b1 = b0 + (size_t)6 * n_items;
                    for (int idx = lane; idx < 2 * (n_ret - n_dead); idx += 32) {
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
                    __syncwarp();
                    for (int j = n_dead + lane; j < n_ret; j += 32) cta_rstep[q * n_sh + j] = 0;
                    if (lane == 0) { slot[q].log_len = 0; slot[q].need_flush = 0; }
                    __syncwarp();
                    continue;
                }
                // outputs: final state if the end was reached, +inf otherwise (diffrax SaveAt semantics)
                const long long part = slot[q].part;
                const double dir = slot[q].dir;
                const bool ok = (slot[q].status == 0) && (slot[q].T0 < slot[q].T1);
                const double* cur = cta_buf + (size_t)(2 * q + slot[q].flip) * 6 * n_items;
                for (int it = lane; it < n_items; it += 32) {
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
                if (lane == 8 * (q / NW)) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) { a.wout[6 * part + k] = ok ? x[k] : inf; a.wout[6 * part + 3 + k] = ok ? dir * p[k] : inf; }
                    a.status[part] = slot[q].status;
                    a.nsteps[3 * part] = slot[q].n_steps; a.nsteps[3 * part + 1] = slot[q].n_acc; a.nsteps[3 * part + 2] = slot[q].n_rej;
                }
            }
            __syncwarp();
            if (lane == 0) {
                s_service[wid] = 0;
                for (int q = wid; q < NP; q += NW) {
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
            __syncwarp();
#pragma unroll 1
            for (int q = wid; q < NP; q += NW) {
                if (!(slot[q].part >= 0 && slot[q].finishing == 2)) continue;
                // ================= start-up of slot q: state load + HNW initial step over the whole coupled state =================
                const int bl = 8 * (q / NW);                   // base lane of the slot
                const long long part = slot[q].part;
                const double t0_in = a.t0[part], t1_in = a.t1;
                const double dir = (t0_in < t1_in) ? 1.0 : -1.0;
                const double T0 = t0_in * dir, T1 = t1_in * dir;
                double* cur = cta_buf + (size_t)(2 * q) * 6 * n_items;
                double* nxt = cur + (size_t)6 * n_items;
                for (int it = lane; it < n_items; it += 32) {
                    const int j = it % n_sh, blk = it / n_sh, o = a.order[j];
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        double v = a.D0 ? a.D0[((size_t)part * n_sh + o) * 12 + blk * 6 + k] : 0.0;
                        if (k >= 3) v *= dir;
                        cur[(size_t)k * n_items + it] = v;
                        nxt[(size_t)k * n_items + it] = v;
                    }
                }
                BaseForce<S, SIG> bforce{&sP, &Pin, &sb[q], dir, 0};
                double b0s = 0.0, b1s = 0.0;
                if (lane == bl) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) { x[k] = a.w0[6 * part + k]; p[k] = dir * a.w0[6 * part + 3 + k]; }
                    bforce.stage = 0;
                    bforce(x, T0, F[0]);
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const double sx = fma(c.rtol, fabs(x[k]), c.atol), sp = fma(c.rtol, fabs(p[k]), c.atol);
                        double r;
                        r = x[k] / sx; b0s = fma(r, r, b0s); r = p[k] / sp; b0s = fma(r, r, b0s);
                        r = p[k] / sx; b1s = fma(r, r, b1s); r = F[0][k] / sp; b1s = fma(r, r, b1s);
                    }
                }
                __syncwarp();
                // the base orbit's terms open the partial sum of (virtual) thread 0, as in response_kernel
                b0s = __shfl_sync(0xffffffffu, b0s, bl); b1s = __shfl_sync(0xffffffffu, b1s, bl);
                Acc4 d0a, d1a;
                d0a.zero(); d1a.zero();
                if (lane == 0) { d0a.a[0] = b0s; d1a.a[0] = b1s; }
                for (int it = lane; it < n_items; it += 32) {
                    ItemParams ip; load_item_params(a.sorted, Sh.profile, it % n_sh, it / n_sh, n_sh, ip);
                    ItemForce<S> f{&sb[q], &ip, 0};
                    double qq[3], pp[3], G[3];
#pragma unroll
                    for (int k = 0; k < 3; ++k) { qq[k] = cur[(size_t)k * n_items + it]; pp[k] = cur[(size_t)(3 + k) * n_items + it]; }
                    f.at(0, qq, G);
                    const int vw = (it >> 5) & 3;
                    double d0s = d0a.get(vw), d1s = d1a.get(vw);
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const double sx = fma(c.rtol, fabs(qq[k]), c.atol), sp = fma(c.rtol, fabs(pp[k]), c.atol);
                        double r;
                        r = qq[k] / sx; d0s = fma(r, r, d0s); r = pp[k] / sp; d0s = fma(r, r, d0s);
                        r = pp[k] / sx; d1s = fma(r, r, d1s); r = G[k] / sp; d1s = fma(r, r, d1s);
                    }
                    d0a.set(vw, d0s); d1a.set(vw, d1s);
                }
                const double d0 = sqrt(d0a.total() / ncomp);
                const double d1 = sqrt(d1a.total() / ncomp);
                const double h0 = hnw_h0(d0, d1);
                double b2s = 0.0;
                if (lane == bl) {
                    double X1[3], F1[3];
#pragma unroll
                    for (int k = 0; k < 3; ++k) X1[k] = fma(h0, p[k], x[k]);
                    bforce.stage = 1;
                    bforce(X1, T0 + h0, F1);
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const double sx = fma(c.rtol, fabs(x[k]), c.atol), sp = fma(c.rtol, fabs(p[k]), c.atol);
                        double r;
                        r = (fma(h0, F[0][k], p[k]) - p[k]) / sx; b2s = fma(r, r, b2s);
                        r = (F1[k] - F[0][k]) / sp; b2s = fma(r, r, b2s);
                    }
                }
                __syncwarp();
                b2s = __shfl_sync(0xffffffffu, b2s, bl);
                Acc4 d2a;
                d2a.zero();
                if (lane == 0) d2a.a[0] = b2s;
                for (int it = lane; it < n_items; it += 32) {
                    ItemParams ip; load_item_params(a.sorted, Sh.profile, it % n_sh, it / n_sh, n_sh, ip);
                    ItemForce<S> f{&sb[q], &ip, 0};
                    double qq[3], pp[3], G0[3], G1[3], q1[3];
#pragma unroll
                    for (int k = 0; k < 3; ++k) { qq[k] = cur[(size_t)k * n_items + it]; pp[k] = cur[(size_t)(3 + k) * n_items + it]; }
                    f.at(0, qq, G0);
#pragma unroll
                    for (int k = 0; k < 3; ++k) q1[k] = fma(h0, pp[k], qq[k]);
                    f.at(1, q1, G1);
                    const int vw = (it >> 5) & 3;
                    double d2s = d2a.get(vw);
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const double sx = fma(c.rtol, fabs(qq[k]), c.atol), sp = fma(c.rtol, fabs(pp[k]), c.atol);
                        double r;
                        r = (fma(h0, G0[k], pp[k]) - pp[k]) / sx; d2s = fma(r, r, d2s);
                        r = (G1[k] - G0[k]) / sp; d2s = fma(r, r, d2s);
                    }
                    d2a.set(vw, d2s);
                }
                const double d2 = sqrt(d2a.total() / ncomp) / h0;
                if (lane == 0) {
                    RespSlot& s = slot[q];
                    double h = fmin(hnw_h1<T::ORDER>(h0, d1, d2), c.dtmax);
                    s.at_dtmin = h <= c.dtmin;
                    h = fmax(h, c.dtmin);
                    s.dir = dir; s.T0 = T0; s.T1 = T1; s.tprev = T0; s.tnext = fmin(T0 + h, T1);
                    s.status = 0; s.n_steps = s.n_acc = s.n_rej = 0; s.accepted = 0; s.flip = 0; s.n_act_run = 0;
                    s.skip = (a.skip_unborn && dir > 0.0) ? 1 : 0;
                    int nd = 0;
                    if (s.skip) { int lo = 0, hi = n_sh; while (lo < hi) { const int mid = (lo + hi) >> 1; if (a.endmax[mid] <= T0) lo = mid + 1; else hi = mid; } nd = lo & ~15; }
                    s.n_dead = nd;
                    s.n_ret = nd; s.log_len = 0; s.need_flush = 0;
                    s.retire_on = (s.skip && a.retire && c.atol > 0.0 && !s.no_retire) ? 1 : 0;
                    for (int e = 0; e < 36; ++e) sC[q][e] = 0.0;
                    s_guard[q] = 0;
                    s.finishing = 0;
                    if (!(T0 < T1)) { s.finishing = 1; s_service[wid] = 1; }
                    else if (c.max_steps <= 0) { s.status = 1; s.finishing = 1; s_service[wid] = 1; }
                }
                __syncwarp();
            }
            if (lane == 0) {
                int live = 0;
                for (int q = wid; q < NP; q += NW) live |= slot[q].part >= 0;
                s_live[wid] = live;
            }
            __syncwarp();
        }
        if (!s_live[wid]) break;                               // uniform over the warp
        // ---- serial phase: base orbits + propagator columns of this round's attempts of the warp's slots, one instruction stream ----
        {
            const bool act = my_slot < NPX && slot[my_slot].part >= 0;
            const int rec = act ? my_slot : NPX + 4 * wid + grp;
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
        __syncwarp();
        // ---- item sweeps of the warp's slots, one after the other, on its 32 lanes ----
        double my_tot = 0.0;
#pragma unroll 1
        for (int q = wid; q < NP; q += NW) {
            if (slot[q].part < 0) continue;
            const int n_act = max(slot[q].n_act_run, s_nact[q]);
            const double* cur = cta_buf + (size_t)(2 * q + slot[q].flip) * 6 * n_items;
            double* nxt = cta_buf + (size_t)(2 * q + (slot[q].flip ^ 1)) * 6 * n_items;
            Acc4 acc;
            acc.zero();
            if (lane == 0) acc.a[0] = s_besq[q];
            int bad_local = 0;
            sweep_items_warp<SOLVER, PROFILE>(&sb[q], sPhiE[q], a.sorted, n_sh, n_items, slot[q].n_ret, n_act, cur, nxt, slot[q].tnext - slot[q].tprev, c, acc,
                                              bad_local, slot[q].retire_on ? (int)(slot[q].part & 3) : 0);
            if (bad_local) s_bad[q] = 1;
            if (slot[q].n_ret > slot[q].n_dead && lane < 6) {
                // retired items: sum over them of (E y)_k^2 = (E C E^T)_kk, scale atol; into the partial sums response_kernel_mp adds them to
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
                acc.a[1] += fmax(v, 0.0) * inv_atol2;
                if (c.rtol * c.rtol * Cq[lane * 6 + lane] > guard_lim) s_guard[q] = 1;
            }
            const double tot = acc.total();
            if (lane == q / NW) my_tot = tot;
        }
        __syncwarp();
        // ---- controllers (lane g serves the warp's slot of group g): accept / reject, next step, completion ----
        if (lane < 4 && wid + NW * lane < NP && slot[wid + NW * lane].part >= 0) {
            const int q = wid + NW * lane;
            RespSlot& s = slot[q];
            const double err = sqrt(my_tot / ncomp);
            const double dt = s.tnext - s.tprev;
            const int any_bad = s_bad[q];
            s_bad[q] = 0;
            if (s_guard[q]) {
                s_guard[q] = 0;
                s.no_retire = 1; s.accepted = 0; s.finishing = 2; s_service[wid] = 1;
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
                if (s.finishing) s_service[wid] = 1;
            }
        }
    }
}
