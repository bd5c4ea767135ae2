"""CPU-only: what "within 10 x rtol/atol for adaptive runs" (BASELINE.json north_star) can mean, measured on the ALGORITHM alone.

The oracle is compared with (a) itself after moving every initial condition by 1 ulp and (b) another legal rounding of the same sources
(oracle/Makefile target `fma`: FMA contraction + AVX2 code generation).  No GPU, no second implementation: whatever these differ by is a
property of diffrax's controller in IEEE doubles - the embedded error estimate of a step that is much smaller than the tolerance needs
(the Hairer-Norsett-Wanner first step clipped to dtmin, the steps right after it while the x10 growth cap binds, the clipped last step)
is a difference of nearly equal numbers at the rounding level, and `factor = 0.9 err^(-1/order)` turns that noise into the next step size.
From there two runs follow different step sequences and differ by a fraction of the solver's own global error.  DESIGN.md section 4 quotes
the numbers printed here; tests/test_gpu_adaptive_parity.py applies the same statistics to the CUDA path."""
import warnings

import numpy as np

import oracle as O
from common import adaptive_parity_stats, assert_adaptive_parity, halo_orbits, mw3_oracle, random_orbits, ulp_ensemble

TRUTH = dict(solver=8, rtol=1e-13, atol=1e-13, dtmin=1e-3, max_steps=400_000, threads=8)
N = 120


def _first_divergence(a, b, rel=1e-6):
    n = min(len(a), len(b))
    bad = (np.abs(a[:n, 1] - b[:n, 1]) > rel * np.abs(a[:n, 1])) | (a[:n, 3] != b[:n, 3])
    return int(np.argmax(bad)) if bad.any() else (n if len(a) == len(b) else n)


def test_step_size_after_a_forced_small_step_is_rounding_noise():
    """Dopri8 at 1e-10: HNW proposes ~0.06 Myr, dtmin = 0.3 clips it, the 0.3 Myr step is accepted with an error estimate of ~1e-8 (true
    local error ~1e-20): pure rounding.  A 1-ulp change of the initial condition changes that estimate by tens of per cent and with it
    the second step - for (nearly) every orbit."""
    orc = mw3_oracle()
    w0, t0 = random_orbits(40, seed=11), np.linspace(-3000, -5, 40)
    w0p = w0 * (1.0 + 2.220446049250313e-16)
    early, errs = 0, []
    for i in range(40):
        a, _ = orc.orbit_trace(w0[i], t0[i], 0.0, solver=8, rtol=1e-10, atol=1e-10)
        b, _ = orc.orbit_trace(w0p[i], t0[i], 0.0, solver=8, rtol=1e-10, atol=1e-10)
        k = _first_divergence(a, b, rel=1e-4)
        if k <= 3 and k < min(len(a), len(b)):
            early += 1
            errs.append(a[k - 1, 2])
    assert early >= 36, early                       # measured 40/40: divergence at attempt 1..3
    assert max(errs) < 1e-5, max(errs)              # ... seeded by an estimate >= 5 orders of magnitude below the tolerance


def test_two_roundings_of_the_oracle():
    """The statistics the GPU parity test uses, applied to the oracle's FMA-contracted build: it must pass, and the fractions of orbits
    within 10 x tol it reaches (printed) are what ANY independent implementation can be expected to reach on these orbit sets."""
    lines = []
    for name, w0 in (("plunging (random_orbits)", random_orbits(N, seed=11)), ("well resolved (halo_orbits)", halo_orbits(N, seed=21))):
        t0 = np.linspace(-3000, -5, N)
        orc = mw3_oracle()
        truth = orc.integrate_orbits(w0, t0, 0.0, **TRUTH)[0][:, 0]
        for solver, tol in ((5, 1e-7), (8, 1e-7), (8, 1e-10)):
            kw = dict(solver=solver, rtol=tol, atol=tol, threads=8)
            base = orc.integrate_orbits(w0, t0, 0.0, **kw)[0][:, 0]
            ens = ulp_ensemble(orc, w0, t0, 0.0, K=6, seed=3, **kw)
            with O.variant("fma"):
                cand = mw3_oracle().integrate_orbits(w0, t0, 0.0, **kw)[0][:, 0]
            lines.append(assert_adaptive_parity(cand, base, ens, truth, tol, f"oracle[fma] vs oracle, {name}, Dopri{solver} tol={tol:g}"))
            d, dens, E = adaptive_parity_stats(cand, base, ens, truth, tol)
            lines[-1] += f"; global error of the solver itself: median {np.median(E):.3g}, max {E.max():.3g} x tol"
    warnings.warn("\n" + "\n".join(lines))


def test_short_integrations_agree_strictly():
    """When the solver's own global error stays below the tolerance (a few steps), two roundings agree within 10 x tol for EVERY orbit."""
    w0 = random_orbits(N, seed=11)
    orc = mw3_oracle()
    for solver, tol in ((5, 1e-7), (8, 1e-7), (8, 1e-10)):
        kw = dict(solver=solver, rtol=tol, atol=tol)
        base = orc.integrate_orbits(w0, -60.0, 0.0, **kw)[0][:, 0]
        with O.variant("fma"):
            cand = mw3_oracle().integrate_orbits(w0, -60.0, 0.0, **kw)[0][:, 0]
        d = np.abs(cand - base) / (tol * (1 + np.abs(base)))
        assert d.max() < 10.0, (solver, tol, d.max())


def test_dopri8_dense_output_error_on_the_headline_progenitor():
    """A18 / VERDICT item 9.  diffrax's own Dopri8 interpolation coefficients cannot be reproduced here, so Dopri8 SaveAt rows between step
    ends - the progenitor at every stripping time of a Dopri8 stream (main.py:289) - use our C1 5th-order continuous extension.  This bounds
    what that costs on the headline configuration (C2 progenitor, 3 Gyr, rtol = atol = 1e-7): the interpolation-only error of the release
    positions (dense row minus truth, minus the solver's own global error interpolated between the step ends) against the global error
    of the solve itself, in units of tol.  Measured: interpolation <= 41 x tol (median 0.4) while the solve is off by up to 474 x tol -
    and diffrax's exact Dopri5 quartic, the reference's default, interpolates WORSE on the same orbit (56 x tol, median 2.2)."""
    orc = mw3_oracle()
    w0 = orc.integrate_orbits([20.0, 0.0, 20.0, 0.0, 0.15, 0.0], 0.0, -3000.0)[0][0, 0]
    ts = np.linspace(-3000.0, 0.0, 20001)
    tol = 1e-7
    out = {}
    for solver in (5, 8):
        ys = orc.integrate_orbits(w0, -3000.0, 0.0, ts=ts, solver=solver, rtol=tol, atol=tol)[0][0]
        yt = orc.integrate_orbits(w0, -3000.0, 0.0, ts=ts, **{**TRUTH, "threads": 1})[0][0]
        tg, yg = orc.orbit_steps(w0, -3000.0, 0.0, solver=solver, rtol=tol, atol=tol)
        ytg = orc.integrate_orbits(w0, -3000.0, 0.0, ts=tg[1:], **{**TRUTH, "threads": 1})[0][0]
        sc = tol * (1.0 + np.abs(yt))
        glob = np.stack([np.interp(ts, tg[1:], (yg[1:] - ytg)[:, k]) for k in range(6)], axis=1)
        e_interp = (np.abs(ys - yt - glob) / sc).max(axis=1)
        e_glob = (np.abs(yg[1:] - ytg) / (tol * (1.0 + np.abs(ytg)))).max()
        out[solver] = (e_interp.max(), np.median(e_interp), e_glob)
    warnings.warn("dense-output error on the C2 progenitor, x tol (interpolation max, median | global error of the solve): "
                  + "; ".join(f"Dopri{s}: {a:.1f}, {b:.2f} | {g:.0f}" for s, (a, b, g) in out.items()))
    assert out[8][0] <= 100.0 and out[8][0] <= 0.25 * out[8][2]          # our Dopri8 extension: far inside the solver's own error
    assert out[8][0] <= 1.5 * out[5][0]                                  # and no worse than diffrax's Dopri5 quartic on the same orbit
