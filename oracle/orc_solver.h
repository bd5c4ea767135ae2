// TEST INFRASTRUCTURE ONLY (see oracle/README.md) - never linked into the product library.
//
// Restatement of the ODE driver every reference integration funnels into:
//   diffrax.diffeqsolve(ODETerm(f), Dopri5|Dopri8, t0, t1, dt0=None, y0, SaveAt(ts=ts),
//                       PIDController(rtol, atol, dtmin, dtmax, force_dtmin=True), max_steps, throw=False)
// called from /root/reference/streamsculptor/main.py:139-162 (integrate_orbit) and
// fields.py:85-98 (integrate_field).  diffrax==0.7.0 (reference requirements.txt:3) is a third-party
// dependency that is NOT present under /root/reference and cannot be installed offline, so this file
// restates its published algorithm:
//   * adaptive explicit Runge-Kutta loop with FSAL, accept/reject, clip-to-end (1e-10 in f64),
//   * PIDController defaults pcoeff=0, icoeff=1, dcoeff=0, safety=0.9, factormin=0.2, factormax=10,
//     norm = rms over ALL state components, keep = (err < 1) | at_dtmin (force_dtmin=True),
//     factor lower clip = 1 after an accepted step,
//   * Hairer-Norsett-Wanner initial step with exponent 1/order,
//   * t1 < t0 handled by mirroring time (direction = -1),
//   * SaveAt(ts): interpolate inside the accepted step that covers each ts; unsaved rows stay +inf.
// Dense output: Dopri5 = diffrax's 4th-order polynomial through (y0, y1, f0, f1, y_mid);
// Dopri8 = OUR 5th-order C1 continuous extension (tools/derive_dopri8_dense.py) because diffrax's
// coefficients are not reproducible here.  A save time equal to the step end returns y1 itself.
// PARITY STATUS: unpinned against diffrax at the 1e-10 level (no runnable reference); pinned by the
// reference's notebook goldens at their printed precision (tests/test_oracle_goldens.py): D2 to all 10
// printed digits, D3 (1000 saved rows of a 3 Gyr orbit) to all 9 printed digits when the first step is the
// dtmin-clipped chain the notebook evidently ran with (see DESIGN.md "Oracle pinning"), D5 to 2e-6.
#ifndef ORC_SOLVER_H
#define ORC_SOLVER_H
#include <cmath>
#include <limits>
#include <vector>

#include "orc_tableau.h"

namespace orc {

struct Ctrl {
    int solver = 8;           // 5 | 8
    double rtol = 1e-7, atol = 1e-7, dtmin = 0.3;
    double dtmax = std::numeric_limits<double>::infinity();   // None
    int max_steps = 10000;
    std::vector<double>* trace = nullptr;   // optional: one record {tprev, dt, err, keep} per step ATTEMPT (mirrored time), for the lock-step parity tests
};
struct Stats { int status = 0, n_steps = 0, n_acc = 0, n_rej = 0; };   // status: 0 ok, 1 max_steps, 2 non-finite

struct RK {
    int s, order;
    const double* c; const double* a; const double* b; const double* e;
    static RK get(int solver) {
        RK r;
        if (solver == 5) { r.s = 7; r.order = 5; r.c = orc_tab::d5_c; r.a = &orc_tab::d5_a[0][0]; r.b = orc_tab::d5_b; r.e = orc_tab::d5_e; }
        else { r.s = 14; r.order = 8; r.c = orc_tab::d8_c; r.a = &orc_tab::d8_a[0][0]; r.b = orc_tab::d8_b; r.e = orc_tab::d8_e; }
        return r;
    }
};

inline double rms(const double* x, const double* sc, int n) {     // diffrax rms_norm of x/sc
    double acc = 0;
    for (int i = 0; i < n; ++i) { double q = x[i] / sc[i]; acc += q * q; }
    return std::sqrt(acc / n);
}

// F: void operator()(double t, const double* y, double* dy)
// optional step recorder R: void operator()(double ta, double tb, const double* y0, const double* y1, const double* f /*[s][n]*/)
struct NoRec { void operator()(double, double, const double*, const double*, const double*) const {} };

template <class F, class R = NoRec>
Stats solve(F& f_user, int n, double t0_in, double t1_in, const double* y0_in, const double* ts_in, int M,
            const Ctrl& ctl, double* ys /*[M][n]*/, R rec = R()) {
    const RK rk = RK::get(ctl.solver);
    const int s = rk.s;
    const double inf = std::numeric_limits<double>::infinity();
    Stats st;
    for (int i = 0; i < M * n; ++i) ys[i] = inf;
    const double dir = (t0_in < t1_in) ? 1.0 : -1.0;                 // diffrax: direction = where(t0 < t1, 1, -1)
    const double t0 = t0_in * dir, t1 = t1_in * dir;
    auto f = [&](double t, const double* y, double* dy) {            // WrapTerm: f(t*dir, y)*dir
        f_user(t * dir, y, dy);
        if (dir < 0) for (int i = 0; i < n; ++i) dy[i] = -dy[i];
    };
    std::vector<double> y(y0_in, y0_in + n), ycand(n), yerr(n), sc(n), ystage(n), tmp(n), K((size_t)s * n);
    double* k0 = K.data();

    // ---- controller.init: initial step (dt0=None) ----
    f(t0, y.data(), k0);
    double h;
    {
        for (int i = 0; i < n; ++i) sc[i] = ctl.atol + std::fabs(y[i]) * ctl.rtol;
        double d0 = rms(y.data(), sc.data(), n), d1 = rms(k0, sc.data(), n);
        bool cond = (d0 < 1e-5) || (d1 < 1e-5);
        double h0 = cond ? 1e-6 : 0.01 * (d0 / d1);
        for (int i = 0; i < n; ++i) ystage[i] = y[i] + h0 * k0[i];
        f(t0 + h0, ystage.data(), tmp.data());
        for (int i = 0; i < n; ++i) tmp[i] = tmp[i] - k0[i];
        double d2 = rms(tmp.data(), sc.data(), n) / h0;
        double maxd = std::fmax(d1, d2);
        double h1 = (maxd <= 1e-15) ? std::fmax(1e-6, h0 * 1e-3) : std::pow(0.01 / maxd, 1.0 / rk.order);
        h = std::fmin(100.0 * h0, h1);
    }
    h = std::fmin(h, ctl.dtmax);
    bool at_dtmin = h <= ctl.dtmin;
    h = std::fmax(h, ctl.dtmin);
    double tprev = t0, tnext = std::fmin(t0 + h, t1);
    int save_idx = 0;

    while (tprev < t1) {
        if (st.n_steps >= ctl.max_steps) { st.status = 1; break; }
        const double dt = tnext - tprev;
        // ---- stages ----
        for (int i = 1; i < s; ++i) {
            const double* ai = rk.a + (size_t)i * s;
            for (int c = 0; c < n; ++c) {
                double acc = 0;
                for (int j = 0; j < i; ++j) if (ai[j] != 0.0) acc += ai[j] * K[(size_t)j * n + c];
                ystage[c] = y[c] + dt * acc;
            }
            f(tprev + rk.c[i] * dt, ystage.data(), &K[(size_t)i * n]);
            if (i == s - 1) ycand = ystage;                          // last stage row == b  (y1 = last stage value)
        }
        for (int c = 0; c < n; ++c) {
            double acc = 0;
            for (int j = 0; j < s; ++j) if (rk.e[j] != 0.0) acc += rk.e[j] * K[(size_t)j * n + c];
            yerr[c] = dt * acc;
        }
        // ---- PIDController.adapt_step_size ----
        bool nan_cand = false, finite = true;
        for (int c = 0; c < n; ++c) { if (std::isnan(ycand[c])) nan_cand = true; if (!std::isfinite(ycand[c])) finite = false; }
        for (int c = 0; c < n; ++c) {
            double yc = nan_cand ? y[c] : ycand[c];
            sc[c] = ctl.atol + std::fmax(std::fabs(y[c]), std::fabs(yc)) * ctl.rtol;
        }
        double err = rms(yerr.data(), sc.data(), n);
        bool keep = (err < 1.0) || at_dtmin;
        if (ctl.trace) { ctl.trace->push_back(tprev); ctl.trace->push_back(dt); ctl.trace->push_back(err); ctl.trace->push_back(keep ? 1.0 : 0.0); }
        double factor = 0.9 * std::pow(1.0 / err, 1.0 / rk.order);   // icoeff=1 only; err==0 -> inf -> clipped to 10
        double fmin = keep ? 1.0 : 0.2;
        if (std::isnan(factor)) { st.status = 2; st.n_steps++; st.n_rej++; break; }   // NaN field: the reference spins to max_steps, outputs stay +inf
        factor = std::fmin(std::fmax(factor, fmin), 10.0);
        double hn = dt * factor;
        hn = std::fmin(hn, ctl.dtmax);
        at_dtmin = hn <= ctl.dtmin;
        hn = std::fmax(hn, ctl.dtmin);
        st.n_steps++;
        if (keep) {
            if (!finite) { st.status = 2; st.n_acc++; break; }
            st.n_acc++;
            rec(tprev * dir, tnext * dir, y.data(), ycand.data(), K.data());
            // ---- SaveAt(ts): every ts[save_idx] <= tnext is interpolated inside this step ----
            while (save_idx < M && ts_in[save_idx] * dir <= tnext) {
                double tq = ts_in[save_idx] * dir;
                double* out = ys + (size_t)save_idx * n;
                double theta = (tq - tprev) / dt;
                if (tq == tnext) { for (int c = 0; c < n; ++c) out[c] = ycand[c]; }
                else if (ctl.solver == 5) {
                    for (int c = 0; c < n; ++c) {
                        double acc = 0;
                        for (int j = 0; j < s; ++j) acc += orc_tab::d5_cmid[j] * K[(size_t)j * n + c];
                        double ymid = y[c] + dt * acc;
                        double f0 = dt * K[c], f1 = dt * K[(size_t)(s - 1) * n + c], y0 = y[c], y1 = ycand[c];
                        double a = 2 * (f1 - f0) - 8 * (y1 + y0) + 16 * ymid;
                        double b = 5 * f0 - 3 * f1 + 18 * y0 + 14 * y1 - 32 * ymid;
                        double cc = f1 - 4 * f0 - 11 * y0 - 5 * y1 + 16 * ymid;
                        out[c] = (((a * theta + b) * theta + cc) * theta + f0) * theta + y0;
                    }
                } else {
                    double bw[14];
                    for (int j = 0; j < 14; ++j) {
                        double p = 0;
                        for (int k = 6; k >= 0; --k) p = p * theta + orc_tab::d8_dense[j][k];
                        bw[j] = p * theta;
                    }
                    for (int c = 0; c < n; ++c) {
                        double acc = 0;
                        for (int j = 0; j < 14; ++j) if (bw[j] != 0.0) acc += bw[j] * K[(size_t)j * n + c];
                        out[c] = y[c] + dt * acc;
                    }
                }
                save_idx++;
            }
            y = ycand;
            for (int c = 0; c < n; ++c) K[c] = K[(size_t)(s - 1) * n + c];    // FSAL
            tprev = tnext;
        } else {
            st.n_rej++;
        }
        tprev = std::fmin(tprev, t1);
        double tn = tprev + hn;
        if (tn > t1 - 1e-10) tn = keep ? t1 : tprev + 0.5 * (t1 - tprev);    // _clip_to_end (f64 tolerance)
        tnext = tn;
    }
    return st;
}

}  // namespace orc
#endif
