"""Adaptive runs: the CUDA path against the oracle (run with -m gpu on the B200 box).

north_star: "within 10 x the solver rtol/atol for adaptive runs".  tests/test_adaptive_criterion.py (CPU) measures what that can mean: the
oracle itself, after a 1-ulp change of the initial conditions or under FMA contraction, stays within 10 x tol of its own answer for
97 % / 89 % / 92 % of 3 Gyr orbits (Dopri8 1e-7, halo orbits 1e-7, 1e-10), because step sizes that follow a forced-small step are set by an
error estimate at the rounding-noise floor.  So this file asks for
  (A) LOCK STEP: the controller of the CUDA path makes the same decisions as the oracle's, attempt by attempt (step sizes to 1e-6, error
      estimates to 1e-3, accept/reject identical), up to the first attempt whose step size was set by a noise-floor estimate (or, rarely, an
      accept/reject decision within 1e-3 of the threshold / an amplified drift); every orbit's first divergence is classified, and every
      orbit that stays in lock step to the end agrees within 10 x tol - strictly, no fraction;
  (B) STATISTICS: as a member of the oracle's own 1-ulp ensemble the CUDA result is not an outlier (common.assert_adaptive_parity);
  (C) SHORT integrations (global error below tol): every orbit within 10 x tol.
"""
import warnings

import numpy as np
import pytest

from common import assert_adaptive_parity, halo_orbits, mw3_oracle, mw3_product, random_orbits, scaled_err, ulp_ensemble

TRUTH = dict(solver=8, rtol=1e-13, atol=1e-13, dtmin=1e-3, max_steps=400_000, threads=8)

pytestmark = pytest.mark.gpu


def _solver(s):
    import streamsculptor_b200 as ssc
    return ssc.Dopri8() if s == 8 else ssc.Dopri5()


def _gpu_traces(prod, w0, t0, t1, solver, tol, dtmin=0.3, max_steps=10_000):
    from streamsculptor_b200 import _runtime as rt
    ctrl = rt.make_ctrl(_solver(solver), tol, tol, dtmin, None, max_steps)
    n = len(w0)
    tr, yfin, st, ns = rt.orbit_trace(prod, rt.to_dev(w0), rt.to_dev(np.broadcast_to(t0, (n,)).copy()), rt.to_dev(np.full(n, float(t1))), ctrl, trace_cap=8192)
    return tr.cpu().numpy(), yfin.cpu().numpy(), st.cpu().numpy(), ns.cpu().numpy()


@pytest.mark.parametrize("solver,tol", [(5, 1e-7), (8, 1e-7), (8, 1e-10), (5, 1e-10)])
def test_controller_in_lock_step_with_the_oracle(cuda, solver, tol):
    orc, prod = mw3_oracle(), mw3_product()
    n = 96
    w0 = np.vstack([random_orbits(n // 2, seed=11), halo_orbits(n // 2, seed=21)])
    t0 = np.linspace(-3000, -200, n)
    tr_g, y_g, st_g, ns_g = _gpu_traces(prod, w0, t0, 0.0, solver, tol)
    assert (st_g == 0).all() and ns_g[:, 0].max() <= 8192
    n_lock, classes, worst_lock, unexplained = 0, dict(noise_floor=0, flip=0, drift=0), 0.0, []
    for i in range(n):
        a, y_o = orc.orbit_trace(w0[i], t0[i], 0.0, solver=solver, rtol=tol, atol=tol)
        b = tr_g[i, : ns_g[i, 0]]
        m = min(len(a), len(b))
        same_dt = np.abs(a[:m, 1] - b[:m, 1]) <= 1e-6 * np.abs(a[:m, 1])
        same_keep = a[:m, 3] == b[:m, 3]
        ok = same_dt & same_keep
        k = m if ok.all() else int(np.argmin(ok))
        # (A1) while in lock step the two error estimators are the same function of the same state: relative 1e-3 above the noise floor
        e_o, e_g = a[:k, 2], b[:k, 2]
        assert np.all(np.abs(e_g - e_o) <= 1e-3 * e_o + 1e-5), (i, k, np.abs(e_g - e_o).max())
        assert np.all(np.abs(a[:k, 0] - b[:k, 0]) <= 2e-6 * (1.0 + np.abs(a[:k, 0] - a[0, 0])))          # same step start times (sums of the step sizes)
        if k == m and len(a) == len(b):
            # same decisions to the end.  STRICT lock step (every step size within 1e-8): the results must agree within 10 x tol, every such
            # orbit; between 1e-8 and 1e-6 the step sizes drifted continuously (no discrete event) - counted as drift below
            if np.all(np.abs(a[:, 1] - b[:, 1]) <= 1e-8 * np.abs(a[:, 1])):
                n_lock += 1
                d = scaled_err(y_g[i][None], y_o[None], tol)[0]
                worst_lock = max(worst_lock, d)
                assert d <= 10.0, f"orbit {i} stayed in strict lock step for all {m} attempts but differs by {d:.2f} x tol"
            else:
                classes["drift"] += 1
            continue
        # (A2) classify the first divergence.  Attempt k's step size was chosen after attempt k-1 from its error estimate.
        assert 1 <= k < m, f"orbit {i}: diverges at attempt {k} of {len(a)} / {len(b)}"
        err_prev = a[k - 1, 2]
        if not same_keep[k] and same_dt[k] and abs(a[k, 2] - 1.0) < 1e-3:
            classes["flip"] += 1                      # accept/reject with the estimate within 1e-3 of the threshold
        elif err_prev < 1e-3:
            classes["noise_floor"] += 1               # estimate >= 3 orders below the tolerance: a difference of nearly equal numbers
        elif same_keep[k] and abs(a[k, 1] - b[k, 1]) <= 1e-3 * abs(a[k, 1]):
            classes["drift"] += 1                     # continuous drift of the step sizes that crossed the 1e-6 threshold, no discrete event
        else:
            unexplained.append((i, k, err_prev, a[k, 1], b[k, 1], a[k, 3], b[k, 3]))
    warnings.warn(f"Dopri{solver} tol={tol:g}: {n_lock}/{n} orbits in strict lock step to the end (worst {worst_lock:.3g} x tol); first divergence of the others: "
                  f"{classes}; unexplained {len(unexplained)}")
    assert not unexplained, unexplained[:5]


@pytest.mark.parametrize("orbits", ["plunging", "halo"])
@pytest.mark.parametrize("solver,tol", [(5, 1e-7), (8, 1e-7), (8, 1e-10)])
def test_final_states_as_a_member_of_the_oracle_ulp_ensemble(cuda, orbits, solver, tol):
    orc, prod = mw3_oracle(), mw3_product()
    n = 200
    w0 = random_orbits(n, seed=11) if orbits == "plunging" else halo_orbits(n, seed=21)
    t0 = np.linspace(-3000, -5, n)
    kw = dict(solver=solver, rtol=tol, atol=tol, threads=8)
    base = orc.integrate_orbits(w0, t0, 0.0, **kw)[0][:, 0]
    ens = ulp_ensemble(orc, w0, t0, 0.0, K=8, seed=3, **kw)
    truth = orc.integrate_orbits(w0, t0, 0.0, **TRUTH)[0][:, 0]
    sol = prod.integrate_orbit_batch_vmapped(w0=w0, ts=np.zeros((n, 1)), t0=t0, t1=0.0, solver=_solver(solver), rtol=tol, atol=tol)
    assert (np.asarray(sol.result) == 0).all()
    warnings.warn(assert_adaptive_parity(np.asarray(sol.ys)[:, 0], base, ens, truth, tol, f"CUDA vs oracle, {orbits} orbits, Dopri{solver} tol={tol:g}"))


@pytest.mark.parametrize("solver,tol", [(5, 1e-7), (8, 1e-7), (8, 1e-10), (5, 1e-10)])
def test_short_integrations_every_orbit_within_10x_tol(cuda, solver, tol):
    orc, prod = mw3_oracle(), mw3_product()
    w0 = np.vstack([random_orbits(150, seed=11), halo_orbits(150, seed=21)])
    for span in ((60.0,) if (solver, tol) == (8, 1e-10) else (60.0, 150.0, 300.0)):     # spans over which two roundings of the ORACLE stay within ~1 x tol
        ys_o, _, _ = orc.integrate_orbits(w0, -span, 0.0, solver=solver, rtol=tol, atol=tol)
        sol = prod.integrate_orbit_batch_vmapped(w0=w0, ts=np.zeros((300, 1)), t0=-span, t1=0.0, solver=_solver(solver), rtol=tol, atol=tol)
        d = scaled_err(np.asarray(sol.ys)[:, 0], ys_o[:, 0], tol)
        assert d.max() < 10.0, (span, d.max())
