// Shared declarations of the libssb200 translation units.
#ifndef SSB_COMMON_CUH
#define SSB_COMMON_CUH
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ssb200.h"
#include "ssb_potential.cuh"
#include "ssb_rk.cuh"

#ifndef SSB_ORBIT_THREADS
#define SSB_ORBIT_THREADS 128
#endif
#ifndef SSB_ORBIT_MIN_BLOCKS
#define SSB_ORBIT_MIN_BLOCKS 3
#endif
#ifndef SSB_SNAP_MIN_BLOCKS
#define SSB_SNAP_MIN_BLOCKS 3   // saving kernel (MODE 0): CTAs per SM the register budget is set for (A/B: 2 = 255 registers, no spills)
#endif
// bit 0: saving kernel (K1 MODE 0), bit 1: final-state kernel (MODE 2): one CTA-wide vote per step-loop iteration keeps the four warps of a
// CTA on the same iteration, so that they fetch the (larger-than-cache) loop body together
#ifndef SSB_ORBIT_CTA_ALIGN
#define SSB_ORBIT_CTA_ALIGN 1
#endif
#ifndef SSB_FUSED_INLINE
#define SSB_FUSED_INLINE 1
#endif
#ifndef SSB_SNAP_FUSED_INLINE
#define SSB_SNAP_FUSED_INLINE 1 // saving kernel (MODE 0): 0 = one out-of-line copy of the fused force (16 KB less code; measured 7 % slower: 24.1 vs 22.5 ms)
#endif
#define SSB_REC_STRIDE 64   // doubles per recorded step: ta, tb, x, p, x1, p1 (14) + up to 14 force stages (42)

namespace ssb {

// copy the by-value kernel parameter into shared memory so that non-inlined force calls can index it dynamically
__device__ __forceinline__ void stage_potential(ssb_potential* dst, const ssb_potential* src) {
    const int* s = reinterpret_cast<const int*>(src);
    int* d = reinterpret_cast<int*>(dst);
    for (int i = threadIdx.x; i < (int)(sizeof(ssb_potential) / sizeof(int)); i += blockDim.x) d[i] = s[i];
    __syncthreads();
    if ((int)threadIdx.x < dst->n_track) {             // segment-guess constants of every track, once per CTA (track_eval)
        ssb_track& T = dst->track[threadIdx.x];
        const double a0 = T.t[0], a1 = T.t[T.n - 1];
        T.t_first = a0;
        T.inv_dt = (a1 > a0) ? (double)(T.n - 1) / (a1 - a0) : 0.0;
    }
    if (dst->n_track > 0) __syncthreads();
}

}  // namespace ssb

// error helpers shared by the host-side translation units (defined in ssb_kernels.cu)
int ssb_set_error(int code, const char* msg);
int ssb_cuda_check(cudaError_t e, const char* what);
int ssb_validate_potential(const ssb_potential* p);
int ssb_validate_ctrl(const ssb_ctrl& c);
void ssb_count_launch();     // every kernel launch of the library bumps a process-wide counter (ssb_launch_count, read by bench.py's gpu_launches)
// ssb_gen_stream_f64 on COMPACT per-shard inputs (internal; used by ssb_gen_stream_host so that a rank uploads only its own stripping
// times): ts_c[n_local + 2] = the shard's stripping times, then the first and last stripping time of the whole stream; Msat_c[n_local],
// normals_c[n_local,4] (or NULL) likewise.  scratch >= ssb_stream_scratch_bytes(n_local + 2, max_steps).
extern "C" int ssb_gen_stream_compact(const ssb_potential* pot, const ssb_potential* pot_release, double G, int64_t Nts, const double* ts_c,
                                      const double* prog_w0, const double* Msat_c, int64_t seed, const double* kvals, const double* normals_c, ssb_ctrl ctrl,
                                      int64_t i_begin, int64_t i_stride, int64_t n_local, double* lead, double* trail, int32_t* status, int32_t* nsteps,
                                      void* scratch, size_t scratch_bytes, void* stream);
#endif
