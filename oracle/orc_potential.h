// TEST INFRASTRUCTURE ONLY (see oracle/README.md) - never linked into the product library.
//
// Scalar potentials Phi(x,t) exactly as the reference writes them; all derivatives by AD
// (orc_ad.h).  Citations are to /root/reference/streamsculptor/.
#ifndef ORC_POTENTIAL_H
#define ORC_POTENTIAL_H
#include <algorithm>
#include <cstdint>
#include <limits>
#include <vector>

#include "orc_ad.h"

namespace orc {

enum CompType {
    C_NFW = 0,          // potential.py:74-84   p = {G*m, r_s, soft (0 today; 1e-3 in the notebook-era revision)}
    C_HERNQUIST = 1,    // potential.py:132-138 p = {G*m, r_s, soft}
    C_MIYAMOTO = 2,     // potential.py:66-72   p = {G*m, a, b}
    C_PLUMMER = 3,      // potential.py:124-130 p = {G*m, r_s}
    C_ISOCHRONE = 4,    // potential.py:114-122 p = {G*m, a}
    C_TRIAXNFW = 5,     // potential.py:86-97   p = {G*m, r_s, q1, q2, q3}
    C_UNIFORM_ACC = 6,  // potential.py:480-502 gradient = d velocity_func/dt ; track = velocity table
    C_SUBHALOS = 7,     // potential.py:802-904, 1161-1268 ; sh = index of the subhalo set
    C_BAR = 9,          // potential.py:178-198 Long & Murali bar rotating with Omega: p = {G*m, a, b, c, Omega}
    C_DEHNEN_BAR = 10,  // potential.py:200-222 p = {alpha, v0, R0, Rb, phib, Omega}
};
enum TrackKind { TK_LINEAR = 0, TK_CUBIC = 1 };
enum Profile { PR_PLUMMER = 0, PR_HERNQUIST = 1, PR_NFW = 2 };

// Interpolated track c(t) in R^dim.
//  TK_LINEAR: jax.scipy.interpolate.RegularGridInterpolator(method='linear', bounds_error=False,
//             fill_value=None) as used at potential.py:581-600 - piecewise linear, linearly
//             extrapolated from the end segments.  [3P: segment index = clip(searchsorted(left)-1)]
//  TK_CUBIC : interpax.Interpolator1D(method='cubic') as used at streamhelpers.py:520 and
//             perturbative.py:642 - C1 cubic Hermite, knot slopes = mean of the two adjacent secant
//             slopes (one-sided at the ends), NaN outside the knots.  [3P-from-memory: interpax 0.3.4]
struct Track {
    int kind = 0, n = 0, dim = 3;
    std::vector<double> t, y, s;   // s: knot slopes (cubic)
    void finalize() {
        if (kind != TK_CUBIC) return;
        s.assign((size_t)n * dim, 0.0);
        for (int c = 0; c < dim; ++c) {
            std::vector<double> sec(n - 1);
            for (int i = 0; i < n - 1; ++i) {
                double dx = t[i + 1] - t[i];
                double dxi = dx == 0 ? 0.0 : 1.0 / dx;
                sec[i] = dxi * (y[(size_t)(i + 1) * dim + c] - y[(size_t)i * dim + c]);
            }
            s[c] = sec[0];
            for (int i = 1; i < n - 1; ++i) s[(size_t)i * dim + c] = 0.5 * (sec[i - 1] + sec[i]);
            s[(size_t)(n - 1) * dim + c] = sec[n - 2];
        }
    }
    // value (out) and time-derivative (dout, may be null)
    void eval(double tq, double* out, double* dout) const {
        if (kind == TK_LINEAR) {
            int idx = int(std::lower_bound(t.begin(), t.end(), tq) - t.begin()) - 1;   // searchsorted 'left' - 1
            idx = std::min(std::max(idx, 0), n - 2);
            double h = t[idx + 1] - t[idx];
            double w = (tq - t[idx]) / h;
            for (int c = 0; c < dim; ++c) {
                double y0 = y[(size_t)idx * dim + c], y1 = y[(size_t)(idx + 1) * dim + c];
                out[c] = (1.0 - w) * y0 + w * y1;
                if (dout) dout[c] = (y1 - y0) / h;
            }
        } else {
            if (!(tq >= t[0] && tq <= t[n - 1])) {
                for (int c = 0; c < dim; ++c) { out[c] = std::numeric_limits<double>::quiet_NaN(); if (dout) dout[c] = out[c]; }
                return;
            }
            int i = int(std::upper_bound(t.begin(), t.end(), tq) - t.begin());            // searchsorted 'right'
            i = std::min(std::max(i, 1), n - 1);
            double dx = t[i] - t[i - 1];
            double dxi = dx == 0 ? 0.0 : 1.0 / dx;
            double u = (tq - t[i - 1]) * dxi;
            for (int c = 0; c < dim; ++c) {
                double f0 = y[(size_t)(i - 1) * dim + c], f1 = y[(size_t)i * dim + c];
                double m0 = s[(size_t)(i - 1) * dim + c] * dx, m1 = s[(size_t)i * dim + c] * dx;
                // cubic Hermite in monomial form: f0 + m0 u + (3(f1-f0) - 2 m0 - m1) u^2 + (2(f0-f1) + m0 + m1) u^3
                double c2 = 3.0 * (f1 - f0) - 2.0 * m0 - m1;
                double c3 = 2.0 * (f0 - f1) + m0 + m1;
                out[c] = f0 + u * (m0 + u * (c2 + u * c3));
                if (dout) dout[c] = (m0 + u * (2.0 * c2 + 3.0 * u * c3)) * dxi;
            }
        }
    }
};

struct SubhaloSet {      // potential.py:802-850 (Plummer), 1161-1213 (any func(m, r_s) - Hernquist in production)
    int n = 0, profile = 0, dradius = 0;
    double G = 0;
    std::vector<double> m, rs, x0, v, t0, tw;
};

struct Comp {
    int type = 0;
    double p[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int track = -1;   // >=0: evaluate at x - c(t) (TimeDepTranslatingPotential, potential.py:448-462)
    int sh = -1;
    int growth = -1;  // >=0: Phi * growth_func(t), growth_func = first column of that track (GrowingPotential, potential.py:464-477)
};

struct Program {
    std::vector<Comp> comps;
    std::vector<Track> tracks;
    std::vector<SubhaloSet> shs;
};

// ---- leaf potentials; T = coordinate scalar type, P = parameter scalar type (double or T) ----
// soft: 0 for the reference as it is today (potential.py:83 has no softening).  The notebook outputs D8 / S1 were produced by an earlier
// revision whose NFW radius was sqrt(r^2 + 0.001); the golden tests select it with soft = 1e-3 (adding 0.0 leaves today's arithmetic
// bit-identical).
template <class T, class P> inline T phi_nfw(const P& GM, const P& rs, const T* x, double soft = 0.0) {
    P v_h2 = -GM / rs;                                                    // potential.py:82
    T m = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + soft) / rs;      // potential.py:83
    return v_h2 * log(1.0 + m) / m;                                       // potential.py:84
}
template <class T, class P> inline T phi_hernquist(const P& GM, const P& rs, double soft, const T* x) {
    T r = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + soft);           // potential.py:137
    return -GM / (r + rs);                                                // potential.py:138
}
template <class T, class P> inline T phi_miyamoto(const P& GM, const P& a, const P& b, const T* x) {
    T R2 = x[0] * x[0] + x[1] * x[1];                                     // potential.py:71
    return -GM / sqrt(R2 + sq(sqrt(x[2] * x[2] + b * b) + a));            // potential.py:72
}
template <class T, class P> inline T phi_plummer(const P& GM, const P& rs, const T* x) {
    T r2 = x[0] * x[0] + x[1] * x[1] + x[2] * x[2];                       // potential.py:129
    return -GM / sqrt(r2 + rs * rs);                                      // potential.py:130
}
template <class T, class P> inline T phi_isochrone(const P& GM, const P& a, const T* x) {
    T r2 = x[0] * x[0] + x[1] * x[1] + x[2] * x[2];                       // potential.py:121 (r = norm)
    return -GM / (a + sqrt(r2 + a * a));                                  // potential.py:122
}
// BarPotential (potential.py:178-198): the point is rotated by ang = -Omega t about z into the bar's frame
template <class T> inline T phi_bar(double GM, double a, double b, double c, double Omega, const T* x, double t) {
    const double ang = -Omega * t, cs = std::cos(ang), sn = std::sin(ang);                 // potential.py:187-188
    T xr = x[0] * cs - x[1] * sn, yr = x[0] * sn + x[1] * cs;                             // potential.py:190: matmul(Rot_mat, xyz)
    T zz = sqrt(x[2] * x[2] + c * c) + b;
    T Tp = sqrt(sq(xr + a) + yr * yr + zz * zz);                                          // potential.py:192
    T Tm = sqrt(sq(a - xr) + yr * yr + zz * zz);                                          // potential.py:193
    return (GM / (2.0 * a)) * log((xr - a + Tm) / (xr + a + Tp));                         // potential.py:195
}
// DehnenBarPotential (potential.py:200-222).  cos(2 (phi - phib - Omega t)) R^2 is written as (x^2 - y^2) cos 2 beta + 2 x y sin 2 beta
// (beta = phib + Omega t): the same function without arctan2 of a dual number.
template <class T> inline T phi_dehnen_bar(double alpha, double v0, double R0, double Rb, double phib, double Omega, const T* x, double t) {
    const double beta = phib + Omega * t, c2 = std::cos(2.0 * beta), s2 = std::sin(2.0 * beta);
    T R2 = x[0] * x[0] + x[1] * x[1];
    T r2 = R2 + x[2] * x[2];
    T r = sqrt(r2);
    T q = r / Rb;
    T U = (val(r) >= Rb) ? -1.0 / (q * q * q) : q * q * q - 2.0;                           // potential.py:210-216
    const double pref = alpha * (v0 * v0 / 3.0) * (R0 / Rb) * (R0 / Rb) * (R0 / Rb);       // potential.py:219
    T ang = (x[0] * x[0] - x[1] * x[1]) * c2 + x[0] * x[1] * (2.0 * s2);
    return pref * (ang / r2) * U;                                                          // potential.py:220
}
template <class T, class P> inline T phi_profile(int profile, const P& GM, const P& rs, const T* x) {
    switch (profile) {
        case PR_PLUMMER: return phi_plummer<T, P>(GM, rs, x);
        case PR_HERNQUIST: return phi_hernquist<T, P>(GM, rs, 0.0, x);
        default: return phi_nfw<T, P>(GM, rs, x);
    }
}

// one subhalo of a SubhaloLine* set; window predicate strict '<' (potential.py:826, 846)
template <class T> inline T phi_subhalo(const SubhaloSet& S, int j, const T* x, double t) {
    if (!(std::fabs(t - S.t0[j]) < S.tw[j])) return T(0.0);
    T rel[3];
    for (int c = 0; c < 3; ++c) rel[c] = x[c] - (S.x0[3 * j + c] + S.v[3 * j + c] * (t - S.t0[j]));   // potential.py:819
    if (!S.dradius) return phi_profile<T, double>(S.profile, S.G * S.m[j], S.rs[j], rel);
    // d Phi / d r_s by differentiating the same formula w.r.t. its r_s argument (potential.py:862-864, 1226-1228)
    typedef Dual<T, 1> U;
    U relu[3];
    for (int c = 0; c < 3; ++c) { relu[c].v = rel[c]; }
    U rs = U::var(T(S.rs[j]), 0);
    U GM = U(0.0); GM.v = T(S.G * S.m[j]);
    U ph = phi_profile<U, U>(S.profile, GM, rs, relu);
    return ph.d[0];
}

// Sum of all conservative components (Potential_Combine.potential, potential.py:1284-1289).
// C_UNIFORM_ACC has no potential (potential.py:494-496): skipped here, added in gradient().
template <class T> inline T phi_total(const Program& P, const T* x, double t) {
    T acc(0.0);
    for (const Comp& c : P.comps) {
        if (c.type == C_UNIFORM_ACC) continue;
        T xs[3] = {x[0], x[1], x[2]};
        if (c.track >= 0) {
            double ctr[3]; P.tracks[c.track].eval(t, ctr, nullptr);
            for (int k = 0; k < 3; ++k) xs[k] = x[k] - ctr[k];                        // potential.py:460-462
        }
        double gf = 1.0;
        if (c.growth >= 0) { double gv[3]; P.tracks[c.growth].eval(t, gv, nullptr); gf = gv[0]; }       // potential.py:476-477: pot.potential(xyz, t) * growth_factor
        switch (c.type) {
            case C_NFW: acc += phi_nfw<T, double>(c.p[0], c.p[1], xs, c.p[2]) * gf; break;
            case C_HERNQUIST: acc += phi_hernquist<T, double>(c.p[0], c.p[1], c.p[2], xs) * gf; break;
            case C_MIYAMOTO: acc += phi_miyamoto<T, double>(c.p[0], c.p[1], c.p[2], xs) * gf; break;
            case C_PLUMMER: acc += phi_plummer<T, double>(c.p[0], c.p[1], xs) * gf; break;
            case C_ISOCHRONE: acc += phi_isochrone<T, double>(c.p[0], c.p[1], xs) * gf; break;
            case C_TRIAXNFW: {
                T xq[3] = {xs[0] / c.p[2], xs[1] / c.p[3], xs[2] / c.p[4]};           // potential.py:94
                acc += phi_nfw<T, double>(c.p[0], c.p[1], xq) * gf; break; }
            case C_BAR: acc += phi_bar<T>(c.p[0], c.p[1], c.p[2], c.p[3], c.p[4], xs, t) * gf; break;
            case C_DEHNEN_BAR: acc += phi_dehnen_bar<T>(c.p[0], c.p[1], c.p[2], c.p[3], c.p[4], c.p[5], xs, t) * gf; break;
            case C_SUBHALOS: {
                const SubhaloSet& S = P.shs[c.sh];
                for (int j = 0; j < S.n; ++j) acc += phi_subhalo<T>(S, j, xs, t);     // potential.py:830
                break; }
            default: break;
        }
    }
    return acc;
}

// gradient of the total field (main.py:37-40 through Potential_Combine.gradient_func, potential.py:1291-1296)
template <class T> inline void gradient(const Program& P, const T* x, double t, T* g) {
    typedef Dual<T, 3> D;
    D xd[3];
    for (int i = 0; i < 3; ++i) xd[i] = D::var(x[i], i);
    D ph = phi_total<D>(P, xd, t);
    for (int i = 0; i < 3; ++i) g[i] = ph.d[i];
    for (const Comp& c : P.comps)
        if (c.type == C_UNIFORM_ACC) {                       // potential.py:497-499: jacfwd(velocity_func)(t)
            double vv[3], dv[3]; P.tracks[c.track].eval(t, vv, dv);
            for (int i = 0; i < 3; ++i) g[i] = g[i] + dv[i];
        }
}

// Hessian of the total potential (main.py:59-65 jacfwd(gradient); fields.py:193 jacrev(gradient))
template <class T> inline void hessian(const Program& P, const T* x, double t, T H[3][3]) {
    typedef Dual<T, 3> D;
    D xd[3], g[3];
    for (int i = 0; i < 3; ++i) xd[i] = D::var(x[i], i);
    gradient<D>(P, xd, t, g);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) H[i][j] = g[i].d[j];
}

}  // namespace orc
#endif
