// Host-pointer entry points of libssb200 (include/ssb200.h, "*_host"): what a CPU-side plugin call looks like.
// Each call uploads its inputs (stream-ordered allocations from the CUDA memory pool, so repeated calls do not
// pay cudaMalloc), enqueues the same kernels as the device-pointer entry points, downloads the results and
// synchronises.  Pinned host buffers make the copies DMA at full PCIe rate; pageable buffers also work.
//
// Zero-copy outputs: when the final-state outputs of ssb_gen_stream_host / ssb_orbit_integrate_host are PINNED (page-locked,
// mapped) host buffers, the orbit kernel receives their device aliases and writes every warp's results straight into host
// memory as whole 256-byte runs while the other orbits are still integrating - the device-to-host transfer (64 MB for a
// 1e6-particle stream) disappears from the critical path.  SSB_HOST_ZEROCOPY=0 forces the staged copies (A/B measurement).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <vector>

#include "ssb_common.cuh"

#define CK(call) do { int _e = ssb_cuda_check((call), #call); if (_e) return _e; } while (0)

namespace {

struct Pool {                       // stream-ordered device allocations released at scope exit
    cudaStream_t st;
    std::vector<void*> ptrs;
    explicit Pool(cudaStream_t s) : st(s) {
        // keep freed blocks cached in the device's default pool between calls: once per DEVICE, safe under concurrent callers
        static std::atomic<bool> tuned[64];
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64 && !tuned[dev].exchange(true, std::memory_order_acq_rel)) {
            cudaMemPool_t mp;
            if (cudaDeviceGetDefaultMemPool(&mp, dev) == cudaSuccess) {
                uint64_t thr = UINT64_MAX;
                cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &thr);
            }
        }
    }
    ~Pool() { for (void* p : ptrs) cudaFreeAsync(p, st); }
    int alloc(void** out, size_t bytes) {
        *out = nullptr;
        if (bytes == 0) bytes = 8;
        int e = ssb_cuda_check(cudaMallocAsync(out, bytes, st), "cudaMallocAsync");
        if (!e) ptrs.push_back(*out);
        return e;
    }
    int up(const void* h, size_t bytes, const void** out) {
        void* d;
        if (int e = alloc(&d, bytes)) return e;
        if (bytes) { if (int e = ssb_cuda_check(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st), "H2D")) return e; }
        *out = d;
        return 0;
    }
};

int upload_subhalos(const ssb_subhalos* h, ssb_subhalos* d, Pool& pool) {
    *d = *h;
    const size_t n = (size_t)(h->n > 0 ? h->n : 0);
    if (n && (!h->m || !h->rs || !h->x0 || !h->v || !h->t0 || !h->tw)) return ssb_set_error(SSB_ERR_ARG, "subhalo set: NULL host array");
    if (int e = pool.up(h->m, 8 * n, (const void**)&d->m)) return e;
    if (int e = pool.up(h->rs, 8 * n, (const void**)&d->rs)) return e;
    if (int e = pool.up(h->x0, 24 * n, (const void**)&d->x0)) return e;
    if (int e = pool.up(h->v, 24 * n, (const void**)&d->v)) return e;
    if (int e = pool.up(h->t0, 8 * n, (const void**)&d->t0)) return e;
    if (int e = pool.up(h->tw, 8 * n, (const void**)&d->tw)) return e;
    return 0;
}

int upload_potential(const ssb_potential* h, ssb_potential* d, Pool& pool) {
    if (!h) return ssb_set_error(SSB_ERR_ARG, "potential is NULL");
    *d = *h;
    if (h->n_track < 0 || h->n_track > SSB_MAX_TRACK || h->n_sh < 0 || h->n_sh > SSB_MAX_SUBHALO_SETS)
        return ssb_set_error(SSB_ERR_ARG, "potential: track/subhalo-set count out of range");
    for (int i = 0; i < h->n_track; ++i) {
        const ssb_track& t = h->track[i];
        if (t.n < 2 || !t.t || !t.y) return ssb_set_error(SSB_ERR_ARG, "track: needs >= 2 knots and t, y host pointers");
        if (int e = pool.up(t.t, 8 * (size_t)t.n, (const void**)&d->track[i].t)) return e;
        if (int e = pool.up(t.y, 24 * (size_t)t.n, (const void**)&d->track[i].y)) return e;
        d->track[i].s = nullptr;
        if (t.kind == SSB_TRACK_CUBIC) {
            if (t.s) { if (int e = pool.up(t.s, 24 * (size_t)t.n, (const void**)&d->track[i].s)) return e; }
            else {
                void* s;
                if (int e = pool.alloc(&s, 24 * (size_t)t.n)) return e;
                if (int e = ssb_track_slopes_f64(t.n, d->track[i].t, d->track[i].y, (double*)s, pool.st)) return e;
                d->track[i].s = (const double*)s;
            }
        }
    }
    for (int i = 0; i < h->n_sh; ++i)
        if (int e = upload_subhalos(&h->sh[i], &d->sh[i], pool)) return e;
    if (h->n_pset < 0 || h->n_pset > SSB_MAX_PSETS) return ssb_set_error(SSB_ERR_ARG, "potential: perturber-set count out of range");
    for (int i = 0; i < h->n_pset; ++i) {
        const ssb_perturbers& s = h->pset[i];
        const size_t n = (size_t)(s.n > 0 ? s.n : 0), nk = (size_t)(s.n_knots > 0 ? s.n_knots : 0);
        if (n && (!s.t || !s.y || !s.GM || !s.rs)) return ssb_set_error(SSB_ERR_ARG, "perturber set: NULL host array");
        if (int e = pool.up(s.t, 8 * nk, (const void**)&d->pset[i].t)) return e;
        if (int e = pool.up(s.y, 24 * nk * n, (const void**)&d->pset[i].y)) return e;
        if (int e = pool.up(s.GM, 8 * n, (const void**)&d->pset[i].GM)) return e;
        if (int e = pool.up(s.rs, 8 * n, (const void**)&d->pset[i].rs)) return e;
    }
    return 0;
}

// per-thread pinned staging buffer (grown on demand, kept between calls): the strided slices a shard needs are gathered here on the
// host and uploaded with one DMA instead of uploading the whole arrays on every rank
double* staging(size_t doubles) {
    static thread_local double* buf = nullptr;
    static thread_local size_t cap = 0;
    if (doubles > cap) {
        if (buf) cudaFreeHost(buf);
        buf = nullptr; cap = 0;
        const size_t want = doubles + doubles / 2 + 1024;
        if (cudaHostAlloc((void**)&buf, want * sizeof(double), cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); buf = nullptr; return nullptr; }
        cap = want;
    }
    return buf;
}

// device alias of a pinned + mapped host range [h, h + bytes), or nullptr (pageable memory, device memory, zero-copy disabled)
void* pinned_alias(const void* h, size_t bytes) {
    if (!h || !bytes) return nullptr;
    const char* env = getenv("SSB_HOST_ZEROCOPY");
    if (env && env[0] == '0') return nullptr;
    cudaPointerAttributes lo, hi;
    if (cudaPointerGetAttributes(&lo, h) != cudaSuccess || cudaPointerGetAttributes(&hi, (const char*)h + bytes - 1) != cudaSuccess) {
        cudaGetLastError();          // older drivers report pageable memory as an error: clear it
        return nullptr;
    }
    if (lo.type != cudaMemoryTypeHost || hi.type != cudaMemoryTypeHost || !lo.devicePointer || !hi.devicePointer) return nullptr;
    if ((const char*)hi.devicePointer - (const char*)lo.devicePointer != (ptrdiff_t)(bytes - 1)) return nullptr;      // one contiguous mapping
    return lo.devicePointer;
}

int down(void* h, const void* d, size_t bytes, cudaStream_t st) {
    if (!bytes) return 0;
    return ssb_cuda_check(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, st), "D2H");
}

}  // namespace

extern "C" {

int ssb_orbit_integrate_host(const ssb_potential* pot_h, int64_t N, const double* w0, const double* t0, const double* t1, const double* ts,
                             int32_t M, int32_t ts_per_orbit, ssb_ctrl ctrl, double* ys, int32_t* status, int32_t* nsteps) {
    if (N < 0 || M < 0) return ssb_set_error(SSB_ERR_ARG, "orbit_integrate_host: negative N or M");
    if (N == 0) return 0;
    if (!w0 || !t0 || !t1 || !status || !nsteps || (M > 0 && (!ts || !ys))) return ssb_set_error(SSB_ERR_ARG, "orbit_integrate_host: NULL array");
    cudaStream_t st = cudaStreamPerThread;
    Pool pool(st);
    ssb_potential pd;
    if (int e = upload_potential(pot_h, &pd, pool)) return e;
    const void *dw0, *dt0, *dt1, *dts;
    if (int e = pool.up(w0, 48 * (size_t)N, &dw0)) return e;
    if (int e = pool.up(t0, 8 * (size_t)N, &dt0)) return e;
    if (int e = pool.up(t1, 8 * (size_t)N, &dt1)) return e;
    const bool final_only = (M == 1 && ts_per_orbit && ts == t1);       // same convention as the device entry point: ts aliases t1
    if (final_only) dts = dt1;
    else if (int e = pool.up(ts, 8 * (size_t)M * (ts_per_orbit ? (size_t)N : 1), &dts)) return e;
    // final-state mode with pinned outputs: the kernel writes host memory directly (see the header comment)
    void *dys = nullptr, *dstat = nullptr, *dns = nullptr;
    const bool zc = final_only && (dys = pinned_alias(ys, 48 * (size_t)N)) && (dstat = pinned_alias(status, 4 * (size_t)N)) &&
                    (dns = pinned_alias(nsteps, 12 * (size_t)N));
    if (!zc) {
        if (int e = pool.alloc(&dys, 48 * (size_t)N * M)) return e;
        if (int e = pool.alloc(&dstat, 4 * (size_t)N)) return e;
        if (int e = pool.alloc(&dns, 12 * (size_t)N)) return e;
    }
    if (int e = ssb_orbit_integrate_f64(&pd, N, (const double*)dw0, (const double*)dt0, (const double*)dt1, (const double*)dts, M, ts_per_orbit,
                                        ctrl, (double*)dys, (int32_t*)dstat, (int32_t*)dns, st)) return e;
    if (!zc) {
        if (int e = down(ys, dys, 48 * (size_t)N * M, st)) return e;
        if (int e = down(status, dstat, 4 * (size_t)N, st)) return e;
        if (int e = down(nsteps, dns, 12 * (size_t)N, st)) return e;
    }
    CK(cudaStreamSynchronize(st));
    return 0;
}

int ssb_gen_stream_host(const ssb_potential* pot_h, const ssb_potential* pot_release_h, double G, int64_t Nts, const double* ts,
                        const double* prog_w0, const double* Msat, int64_t seed, const double* kvals, const double* normals, ssb_ctrl ctrl,
                        int64_t i_begin, int64_t i_stride, int64_t n_local, double* lead, double* trail, int32_t* status, int32_t* nsteps) {
    if (Nts < 2 || !ts || !prog_w0 || !Msat || !kvals) return ssb_set_error(SSB_ERR_ARG, "gen_stream_host: NULL array or Nts < 2");
    if (n_local < 0) return ssb_set_error(SSB_ERR_ARG, "gen_stream_host: negative n_local");
    const size_t n = (size_t)n_local;
    if (n && (!lead || !trail || !status || !nsteps)) return ssb_set_error(SSB_ERR_ARG, "gen_stream_host: NULL output");
    cudaStream_t st = cudaStreamPerThread;
    Pool pool(st);
    ssb_potential pd, prd;
    if (int e = upload_potential(pot_h, &pd, pool)) return e;
    if (pot_release_h == pot_h) prd = pd;
    else if (int e = upload_potential(pot_release_h, &prd, pool)) return e;
    const void *dts, *dw0, *dms, *dnr = nullptr;
    if (int e = pool.up(prog_w0, 48, &dw0)) return e;
    // A shard (i_stride > 1) needs only its own n_local stripping times (+ the two ends of the progenitor interval): gather them on
    // the host and upload the compact arrays - h2d bytes per rank stay ~ 1/world of the arrays however many ranks share the stream.
    bool compact = i_stride > 1 && n > 0 && i_begin >= 0 && i_begin + (int64_t)(n - 1) * i_stride <= Nts - 2;
    double* stg = compact ? staging((n + 2) + n + (normals ? 4 * n : 0)) : nullptr;
    if (!stg) compact = false;
    if (compact) {
        double *c_ts = stg, *c_ms = stg + (n + 2), *c_nr = c_ms + n;
        for (size_t k = 0; k < n; ++k) {
            const size_t g = (size_t)i_begin + k * (size_t)i_stride;
            c_ts[k] = ts[g]; c_ms[k] = Msat[g];
            if (normals) { c_nr[4 * k] = normals[4 * g]; c_nr[4 * k + 1] = normals[4 * g + 1]; c_nr[4 * k + 2] = normals[4 * g + 2]; c_nr[4 * k + 3] = normals[4 * g + 3]; }
        }
        c_ts[n] = ts[0]; c_ts[n + 1] = ts[Nts - 1];
        if (int e = pool.up(c_ts, 8 * (n + 2), &dts)) return e;
        if (int e = pool.up(c_ms, 8 * n, &dms)) return e;
        if (normals) { if (int e = pool.up(c_nr, 32 * n, &dnr)) return e; }
    } else {
        if (int e = pool.up(ts, 8 * (size_t)Nts, &dts)) return e;
        if (int e = pool.up(Msat, 8 * (size_t)Nts, &dms)) return e;
        if (normals) { if (int e = pool.up(normals, 32 * (size_t)Nts, &dnr)) return e; }
    }
    // lead and trail adjacent in memory (one [2, n, 6] buffer): the orbit kernel writes them in place.  If that buffer and the
    // status / step-count outputs are pinned host memory, "in place" is the HOST buffer itself (zero-copy, see the header comment).
    void *dl = nullptr, *dtr, *dstat = nullptr, *dns = nullptr, *scr;
    const bool zc = n && trail == lead + 6 * n && (dl = pinned_alias(lead, 96 * n)) && (dstat = pinned_alias(status, 8 * n)) &&
                    (dns = pinned_alias(nsteps, 24 * n));
    if (!zc) {
        if (int e = pool.alloc(&dl, 96 * n)) return e;
        if (int e = pool.alloc(&dstat, 8 * n)) return e;
        if (int e = pool.alloc(&dns, 24 * n)) return e;
    }
    dtr = (char*)dl + 48 * n;
    const size_t sb = ssb_stream_scratch_bytes(compact ? (int64_t)n + 2 : Nts, ctrl.max_steps);
    if (int e = pool.alloc(&scr, sb)) return e;
    if (int e = (compact ? ssb_gen_stream_compact : ssb_gen_stream_f64)(&pd, &prd, G, Nts, (const double*)dts, (const double*)dw0, (const double*)dms, seed, kvals,
                                   (const double*)dnr, ctrl, i_begin, i_stride, n_local, (double*)dl, (double*)dtr, (int32_t*)dstat, (int32_t*)dns, scr, sb, st)) return e;
    if (!zc) {
        if (int e = down(lead, dl, 48 * n, st)) return e;
        if (int e = down(trail, dtr, 48 * n, st)) return e;
        if (int e = down(status, dstat, 8 * n, st)) return e;
        if (int e = down(nsteps, dns, 24 * n, st)) return e;
    }
    CK(cudaStreamSynchronize(st));
    return 0;
}

int ssb_linear_response_host(const ssb_potential* pot_base_h, const ssb_subhalos* sh_h, int64_t N, const double* w0, const double* D0,
                             const double* t0, double t1, ssb_ctrl ctrl, double* wout, double* Dout, int32_t* status, int32_t* nsteps) {
    if (!sh_h || N < 0) return ssb_set_error(SSB_ERR_ARG, "linear_response_host: bad argument");
    if (N == 0) return 0;
    if (!w0 || !t0 || !wout || !status || !nsteps || (sh_h->n > 0 && !Dout)) return ssb_set_error(SSB_ERR_ARG, "linear_response_host: NULL array");
    cudaStream_t st = cudaStreamPerThread;
    Pool pool(st);
    ssb_potential pd;
    ssb_subhalos sd;
    if (int e = upload_potential(pot_base_h, &pd, pool)) return e;
    if (int e = upload_subhalos(sh_h, &sd, pool)) return e;
    const size_t nsh = (size_t)sh_h->n;
    const void *dw0, *dD0 = nullptr, *dt0;
    if (int e = pool.up(w0, 48 * (size_t)N, &dw0)) return e;
    if (D0) { if (int e = pool.up(D0, 96 * (size_t)N * nsh, &dD0)) return e; }
    if (int e = pool.up(t0, 8 * (size_t)N, &dt0)) return e;
    void *dw, *dD, *dstat, *dns, *scr;
    if (int e = pool.alloc(&dw, 48 * (size_t)N)) return e;
    if (int e = pool.alloc(&dD, 96 * (size_t)N * nsh)) return e;
    if (int e = pool.alloc(&dstat, 4 * (size_t)N)) return e;
    if (int e = pool.alloc(&dns, 12 * (size_t)N)) return e;
    const size_t sb = ssb_response_scratch_bytes(sh_h->n);
    if (int e = pool.alloc(&scr, sb)) return e;
    if (int e = ssb_linear_response_f64(&pd, &sd, N, (const double*)dw0, (const double*)dD0, (const double*)dt0, t1, ctrl, (double*)dw, (double*)dD,
                                        (int32_t*)dstat, (int32_t*)dns, scr, sb, st)) return e;
    if (int e = down(wout, dw, 48 * (size_t)N, st)) return e;
    if (int e = down(Dout, dD, 96 * (size_t)N * nsh, st)) return e;
    if (int e = down(status, dstat, 4 * (size_t)N, st)) return e;
    if (int e = down(nsteps, dns, 12 * (size_t)N, st)) return e;
    CK(cudaStreamSynchronize(st));
    return 0;
}

}  // extern "C"
