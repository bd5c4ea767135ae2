// Truncated multivariate Taylor arithmetic in (x, y, z) up to total order 3, and the two rotating-bar potentials built on it.
//
// The reference differentiates EVERY scalar potential by autodiff (main.py:37-65).  For the spherical / disk components the kernels carry
// closed forms; for the bars (BarPotential, potential.py:178-198: Long & Murali 1992 eq. 8a in a frame rotating with Omega; DehnenBarPotential,
// potential.py:200-222) closed-form second and third derivatives are pages of algebra, so they are evaluated the way the reference does it -
// by propagating derivatives through the scalar formula - with a jet of 4 / 10 / 20 Taylor coefficients (order 1 / 2 / 3).  These components
// live on the interpreter path (out-of-line calls), not in a fused signature; a bar costs about 25 NFW evaluations at order 2.
#pragma once
#include <cmath>
// the header also compiles as plain C++ (tests/test_host_cpu.py checks the jet algebra against the oracle's autodiff without a GPU)
#ifdef __CUDACC__
#define SSB_JET_FN __device__ __forceinline__
#define SSB_JET_CALL __device__ __noinline__
#else
#define SSB_JET_FN inline
#define SSB_JET_CALL inline
#endif

namespace ssb {

// monomials in graded order: 1 | x y z | xx xy xz yy yz zz | xxx xxy xxz xyy xyz xzz yyy yyz yzz zzz
constexpr int jet_n(int ord) { return (ord + 1) * (ord + 2) * (ord + 3) / 6; }
template <int ORD>
struct Jet {
    static constexpr int N = jet_n(ORD);
    double c[N];
    SSB_JET_FN static Jet constant(double v) { Jet r; for (int i = 0; i < N; ++i) r.c[i] = 0.0; r.c[0] = v; return r; }
    SSB_JET_FN static Jet variable(double v, int k) { Jet r = constant(v); r.c[1 + k] = 1.0; return r; }
};
template <int ORD> SSB_JET_FN Jet<ORD> operator+(const Jet<ORD>& a, const Jet<ORD>& b) { Jet<ORD> r; for (int i = 0; i < Jet<ORD>::N; ++i) r.c[i] = a.c[i] + b.c[i]; return r; }
template <int ORD> SSB_JET_FN Jet<ORD> operator-(const Jet<ORD>& a, const Jet<ORD>& b) { Jet<ORD> r; for (int i = 0; i < Jet<ORD>::N; ++i) r.c[i] = a.c[i] - b.c[i]; return r; }
template <int ORD> SSB_JET_FN Jet<ORD> operator+(const Jet<ORD>& a, double s) { Jet<ORD> r = a; r.c[0] += s; return r; }
template <int ORD> SSB_JET_FN Jet<ORD> operator*(const Jet<ORD>& a, double s) { Jet<ORD> r; for (int i = 0; i < Jet<ORD>::N; ++i) r.c[i] = a.c[i] * s; return r; }
// truncated product: every pair of monomials whose degrees add up to at most ORD.  Written out (generated from the exponent table: a
// constexpr index search inside unrolled loops takes nvcc tens of minutes per translation unit)
template <int ORD> SSB_JET_FN Jet<ORD> operator*(const Jet<ORD>& a, const Jet<ORD>& b) {
    Jet<ORD> r;
    r.c[0] = a.c[0] * b.c[0];
    r.c[1] = fma(a.c[1], b.c[0], a.c[0] * b.c[1]);
    r.c[2] = fma(a.c[2], b.c[0], a.c[0] * b.c[2]);
    r.c[3] = fma(a.c[3], b.c[0], a.c[0] * b.c[3]);
    if constexpr (ORD >= 2) {
        r.c[4] = fma(a.c[4], b.c[0], fma(a.c[1], b.c[1], a.c[0] * b.c[4]));
        r.c[5] = fma(a.c[5], b.c[0], fma(a.c[2], b.c[1], fma(a.c[1], b.c[2], a.c[0] * b.c[5])));
        r.c[6] = fma(a.c[6], b.c[0], fma(a.c[3], b.c[1], fma(a.c[1], b.c[3], a.c[0] * b.c[6])));
        r.c[7] = fma(a.c[7], b.c[0], fma(a.c[2], b.c[2], a.c[0] * b.c[7]));
        r.c[8] = fma(a.c[8], b.c[0], fma(a.c[3], b.c[2], fma(a.c[2], b.c[3], a.c[0] * b.c[8])));
        r.c[9] = fma(a.c[9], b.c[0], fma(a.c[3], b.c[3], a.c[0] * b.c[9]));
    }
    if constexpr (ORD >= 3) {
        r.c[10] = fma(a.c[10], b.c[0], fma(a.c[4], b.c[1], fma(a.c[1], b.c[4], a.c[0] * b.c[10])));
        r.c[11] = fma(a.c[11], b.c[0], fma(a.c[5], b.c[1], fma(a.c[4], b.c[2], fma(a.c[2], b.c[4], fma(a.c[1], b.c[5], a.c[0] * b.c[11])))));
        r.c[12] = fma(a.c[12], b.c[0], fma(a.c[6], b.c[1], fma(a.c[4], b.c[3], fma(a.c[3], b.c[4], fma(a.c[1], b.c[6], a.c[0] * b.c[12])))));
        r.c[13] = fma(a.c[13], b.c[0], fma(a.c[7], b.c[1], fma(a.c[5], b.c[2], fma(a.c[2], b.c[5], fma(a.c[1], b.c[7], a.c[0] * b.c[13])))));
        r.c[14] = fma(a.c[14], b.c[0], fma(a.c[8], b.c[1], fma(a.c[6], b.c[2], fma(a.c[5], b.c[3], fma(a.c[3], b.c[5], fma(a.c[2], b.c[6], fma(a.c[1], b.c[8], a.c[0] * b.c[14])))))));
        r.c[15] = fma(a.c[15], b.c[0], fma(a.c[9], b.c[1], fma(a.c[6], b.c[3], fma(a.c[3], b.c[6], fma(a.c[1], b.c[9], a.c[0] * b.c[15])))));
        r.c[16] = fma(a.c[16], b.c[0], fma(a.c[7], b.c[2], fma(a.c[2], b.c[7], a.c[0] * b.c[16])));
        r.c[17] = fma(a.c[17], b.c[0], fma(a.c[8], b.c[2], fma(a.c[7], b.c[3], fma(a.c[3], b.c[7], fma(a.c[2], b.c[8], a.c[0] * b.c[17])))));
        r.c[18] = fma(a.c[18], b.c[0], fma(a.c[9], b.c[2], fma(a.c[8], b.c[3], fma(a.c[3], b.c[8], fma(a.c[2], b.c[9], a.c[0] * b.c[18])))));
        r.c[19] = fma(a.c[19], b.c[0], fma(a.c[9], b.c[3], fma(a.c[3], b.c[9], a.c[0] * b.c[19])));
    }
    return r;
}
// g(f) from the Taylor coefficients g0 = g(f0), g1 = g'(f0), g2 = g''(f0)/2, g3 = g'''(f0)/6: g0 + g1 d + g2 d^2 + g3 d^3, d = f - f0
template <int ORD> SSB_JET_FN Jet<ORD> jet_compose(const Jet<ORD>& f, double g0, double g1, double g2, double g3) {
    Jet<ORD> d = f;
    d.c[0] = 0.0;
    Jet<ORD> r = d * g1;
    if (ORD >= 2) {
        const Jet<ORD> d2 = d * d;
        r = r + d2 * g2;
        if (ORD >= 3) r = r + (d2 * d) * g3;
    }
    r.c[0] = g0;
    return r;
}
template <int ORD> SSB_JET_FN Jet<ORD> jet_sqrt(const Jet<ORD>& f) {
    const double g = sqrt(f.c[0]), ig = 1.0 / g, ig2 = ig * ig;
    return jet_compose(f, g, 0.5 * ig, -0.125 * ig * ig2, 0.0625 * ig * ig2 * ig2);
}
template <int ORD> SSB_JET_FN Jet<ORD> jet_log(const Jet<ORD>& f) {
    const double i = 1.0 / f.c[0];
    return jet_compose(f, log(f.c[0]), i, -0.5 * i * i, i * i * i / 3.0);
}
template <int ORD> SSB_JET_FN Jet<ORD> jet_recip(const Jet<ORD>& f) {
    const double i = 1.0 / f.c[0];
    return jet_compose(f, i, -i * i, i * i * i, -i * i * i * i);
}

// BarPotential (potential.py:178-198).  p = {G m, a, b, c, Omega}; `gm` = G m times a growth factor, if any.
//   ang = -Omega t; (x', y') = Rz(ang) (x, y);  T± = sqrt((a ± x')^2 + y'^2 + (b + sqrt(c^2 + z^2))^2)
//   Phi = G m / (2 a) log((x' - a + T-) / (x' + a + T+))
template <int ORD>
SSB_JET_CALL void bar_jet(const double* __restrict__ p, double gm, double x, double y, double z, double t, double* __restrict__ out) {
    typedef Jet<ORD> J;
    const double a = p[1], b = p[2], c = p[3], ang = -p[4] * t;
    const double sn = sin(ang), cs = cos(ang);
    const J X = J::variable(x, 0), Y = J::variable(y, 1), Z = J::variable(z, 2);
    const J xr = X * cs - Y * sn, yr = X * sn + Y * cs;                     // potential.py:189-191
    const J zz = jet_sqrt(Z * Z + c * c) + b;
    const J com = yr * yr + zz * zz;
    const J xp = xr + a, xm = xr * (-1.0) + a;
    const J Tp = jet_sqrt(xp * xp + com), Tm = jet_sqrt(xm * xm + com);     // potential.py:193-194
    const J ratio = (xr + (-a) + Tm) * jet_recip(xr + a + Tp);
    const J phi = jet_log(ratio) * (gm / (2.0 * a));                        // potential.py:196
    for (int i = 0; i < J::N; ++i) out[i] = phi.c[i];
}

// DehnenBarPotential (potential.py:200-222).  p = {alpha, v0, R0, Rb, phib, Omega}; `alpha` = p[0] times a growth factor, if any.
//   Phi = alpha v0^2/3 (R0/Rb)^3 (R^2/r^2) U(r) cos(2 (phi - phib - Omega t)),  U = -(r/Rb)^-3 for r >= Rb, (r/Rb)^3 - 2 inside.
//   cos(2 (phi - beta)) R^2 = (x^2 - y^2) cos 2 beta + 2 x y sin 2 beta: no atan2 needed.
template <int ORD>
SSB_JET_CALL void dehnen_bar_jet(const double* __restrict__ p, double alpha, double x, double y, double z, double t, double* __restrict__ out) {
    typedef Jet<ORD> J;
    const double v0 = p[1], R0 = p[2], Rb = p[3], beta = p[4] + p[5] * t;
    const double s2 = sin(2.0 * beta), c2 = cos(2.0 * beta);
    const J X = J::variable(x, 0), Y = J::variable(y, 1), Z = J::variable(z, 2);
    const J r2 = X * X + Y * Y + Z * Z;
    const J r = jet_sqrt(r2);
    const J r3 = r * r2;
    const double rb3 = Rb * Rb * Rb;
    const J U = (r.c[0] >= Rb) ? jet_recip(r3) * (-rb3) : r3 * (1.0 / rb3) + (-2.0);      // potential.py:210-216
    const J ang = (X * X - Y * Y) * c2 + (X * Y) * (2.0 * s2);
    const double ratio = R0 / Rb;
    const double pref = alpha * (v0 * v0 / 3.0) * ratio * ratio * ratio;                 // potential.py:219
    const J phi = (ang * jet_recip(r2)) * U * pref;                                       // potential.py:220
    for (int i = 0; i < J::N; ++i) out[i] = phi.c[i];
}

}  // namespace ssb
