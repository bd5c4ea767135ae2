// Branch-free fp64 reciprocal / rsqrt / log1p for the force evaluation (sm_100a).
//
// The RHS of the orbit ODE is ~4 sqrt + ~6 div + 1 log.  CUDA's IEEE-exact division, sqrt and log1p
// expand to 20-70 instructions each, with slow-path subroutine calls; on the FP64 pipe (64 DFMA/clk/SM) they, not the
// Runge-Kutta arithmetic, set the step rate.  These versions start from the MUFU 64-bit seed (rcp/rsqrt.approx.ftz.f64,
// ~23 good bits) and apply ONE cubically convergent correction (23 -> 69 bits), so they are accurate to ~1 ulp (not
// correctly rounded), need no branches and cost 4-6 DFMA-class instructions.  Inputs on the path are finite, positive
// and far from the subnormal range (kpc / Myr / Msun units); x = 0 yields inf/NaN exactly as the reference's formulas do.
#ifndef SSB_FASTMATH_CUH
#define SSB_FASTMATH_CUH
#include <cuda_runtime.h>

#include "ssb_logtab.h"

namespace ssb {

__device__ __forceinline__ double frcp(double x) {            // 1/x
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0);                          // 1 - x y          (|e| ~ 2^-23)
    return fma(y, fma(e, e, e), y);                            // y (1 + e + e^2)  (error ~ e^3)
}

__device__ __forceinline__ double frsqrt(double x) {           // x^(-1/2)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x * y, y, 1.0);                      // 1 - x y^2
    return fma(y, e * fma(0.375, e, 0.5), y);                  // y (1 + e/2 + 3 e^2/8)
}

// ln(1 + m) for m >= 0 (NFW: m = r / r_s).  w = 1 + m carries the rounding error of the sum, which the classic
// first-order correction (m - (w - 1)) / w removes; ln w = k ln2 + 2 atanh(s), s = (f - 1)/(f + 1), f in [sqrt(1/2), sqrt 2).
#ifndef SSB_LOG1P_INLINE
#define SSB_LOG1P_INLINE 1
#endif
#if SSB_LOG1P_INLINE
__device__ __forceinline__
#else
static __device__ __noinline__
#endif
double flog1p_pos(double m, double inv_w) {
    // inv_w ~ 1/(1 + m) multiplies only the rounding-error term of the sum (|corr| <= ulp(w)/2): a few good digits suffice, so
    // callers pass whatever reciprocal they already have (NFW: r_s / (r + r_s)).
    const double w = 1.0 + m;
    const double corr = m - (w - 1.0);                         // exact (Sterbenz-type) for m >= 0
    int hi = __double2hiint(w);
    int k = (hi >> 20) - 1023;
    hi = (hi & 0x000fffff) | 0x3ff00000;                       // mantissa -> [1, 2)
    double f = __hiloint2double(hi, __double2loint(w));
    if (f > 1.4142135623730951) { f *= 0.5; k += 1; }
    const double irw = frcp(f + 1.0);
    const double s = (f - 1.0) * irw;
    const double z = s * s;
    // atanh(s)/s = 1 + z/3 + z^2/5 + ... (|s| <= 0.1716: z^10/21 < 1e-17)
    double p = 1.0 / 19.0;
    p = fma(p, z, 1.0 / 17.0); p = fma(p, z, 1.0 / 15.0); p = fma(p, z, 1.0 / 13.0); p = fma(p, z, 1.0 / 11.0);
    p = fma(p, z, 1.0 / 9.0); p = fma(p, z, 1.0 / 7.0); p = fma(p, z, 1.0 / 5.0); p = fma(p, z, 1.0 / 3.0);
    const double s2 = s + s;
    const double lnf = fma(s2 * z, p, s2);
    const double kd = (double)k;
    // k ln2 split hi/lo so that small results keep full relative accuracy
    double r = fma(kd, 6.93147180369123816490e-01, lnf);
    r = fma(kd, 1.90821492927058770002e-10, r);
    return fma(corr, inv_w, r);
}
__device__ __forceinline__ double flog1p_pos(double m) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(1.0 + m));
    return flog1p_pos(m, y);
}

// Table-driven variant for the fused galaxy signatures (the hot force): ln w = k ln2 + L_i + ln(1 + r), r = fma(f, c_i, -1) exact,
// |r| <= 2^-8, {c_i, L_i} from a 256-entry table (tools/gen_logtab.py) held in shared memory - 14 FP64 instructions instead of ~26,
// max relative error 4e-16 (checked against mpmath over m in [1e-6, 1e3]).  Kernels that evaluate a fused signature call
// logtab_init() once per CTA.
__device__ __forceinline__ double2* ssb_logtab_ptr() {
    __shared__ double2 tab[256];
    return tab;
}
__device__ __forceinline__ void logtab_init() {
    double2* t = ssb_logtab_ptr();
    for (int i = threadIdx.x; i < 256; i += blockDim.x) t[i] = make_double2(ssb_logtab_g[2 * i], ssb_logtab_g[2 * i + 1]);
    __syncthreads();
}
__device__ __forceinline__ double flog1p_tab(double m, double inv_w) {
    const double w = 1.0 + m;
    const double corr = m - (w - 1.0);
    const int hi = __double2hiint(w);
    const int k = (hi >> 20) - 1023;
    const double f = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(w));
    const double2 cl = ssb_logtab_ptr()[(hi >> 12) & 0xff];
    const double r = fma(f, cl.x, -1.0);
    double q = fma(r, 1.0 / 7.0, -1.0 / 6.0);
    q = fma(q, r, 0.2); q = fma(q, r, -0.25); q = fma(q, r, 1.0 / 3.0); q = fma(q, r, -0.5); q = fma(q, r, 1.0);
    const double kd = (double)k;
    double res = fma(kd, 6.93147180369123816490e-01, cl.y);
    res = fma(r, q, res);
    res = fma(kd, 1.90821492927058770002e-10, res);
    return fma(corr, inv_w, res);
}

// x^(-1/ORDER) for the step-size controller (diffrax: factor = safety * (1/err)^(1/order))
template <int ORDER> __device__ __forceinline__ double inv_root(double x);
template <> __device__ __forceinline__ double inv_root<8>(double x) {
    const double a = x * frsqrt(x);        // x^(1/2)
    const double b = a * frsqrt(a);        // x^(1/4)
    const double r = frsqrt(b);            // x^(-1/8)
    // err == +inf (overflowing error estimate): diffrax's (1/inf)^(1/8) = 0 -> factor clipped to factormin, the step is rejected
    // and retried; rsqrt.approx(inf) = 0 would turn the refinement into inf * 0 = NaN (status 2) instead
    return x == 0.0 ? __longlong_as_double(0x7ff0000000000000LL) : (x == __longlong_as_double(0x7ff0000000000000LL) ? 0.0 : r);
}
template <> __device__ __forceinline__ double inv_root<5>(double x) { return exp(-0.2 * log(x)); }

}  // namespace ssb
#endif
