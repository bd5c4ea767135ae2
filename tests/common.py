"""Shared builders: the same physical setups expressed once for the oracle (oracle.Program) and once for the
product (streamsculptor_b200 classes), independently, so parameter-order mistakes cannot cancel."""
import numpy as np

import oracle as O


def mw3_oracle():
    return O.Program().hernquist(5e9, 1.0).miyamoto(6.8e10, 3.0, 0.28).nfw(5.4e11, 15.62)


def mw3_product():
    import streamsculptor_b200 as ssc
    P = ssc.potential
    return P.Potential_Combine([P.HernquistPotential(m=5e9, r_s=1.0, units=ssc.usys), P.MiyamotoNagaiDisk(m=6.8e10, a=3.0, b=0.28, units=ssc.usys),
                                P.NFWPotential(m=5.4e11, r_s=15.62, units=ssc.usys)], units=ssc.usys)


def lmc_track(n=200, t_lo=-3000.0, t_hi=0.0):
    """A smooth synthetic perturber track (SURVEY 8d C3 uses an orbit; any smooth table exercises the same code)."""
    t = np.linspace(t_lo, t_hi, n)
    y = np.stack([-1.0 + 0.02 * t + 30 * np.sin(t / 900.0), -41.0 - 0.01 * t + 10 * np.cos(t / 700.0), -28.0 + 0.015 * t], axis=1)
    return t, y


def random_orbits(n, seed=0):
    rng = np.random.default_rng(seed)
    x = rng.normal(size=(n, 3)) * np.array([12.0, 12.0, 6.0]) + np.array([10.0, 0.0, 5.0])
    v = rng.normal(size=(n, 3)) * 0.08
    return np.hstack([x, v])


def halo_orbits(n, seed=0):
    """Stream-progenitor-like orbits: r in [12, 30] kpc, mostly tangential velocities (pericentres of several kpc), so
    that fixed-step runs are well resolved and rounding differences are not amplified by near-centre passages."""
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1)[:, None]
    r = rng.uniform(12.0, 30.0, n)
    x = d * r[:, None]
    e = np.cross(d, rng.normal(size=(n, 3))); e /= np.linalg.norm(e, axis=1)[:, None]
    v = e * rng.uniform(0.14, 0.22, n)[:, None] + d * rng.normal(size=(n, 1)) * 0.03
    return np.hstack([x, v])


def subhalo_set(n, seed=1, t_lo=-3000.0, t_hi=0.0, tw=150.0):
    rng = np.random.default_rng(seed)
    M = 10 ** rng.uniform(5, 9, n)
    rs = 1.05 * np.sqrt(M / 1e8)
    x0 = rng.normal(size=(n, 3)) * 10.0
    v = rng.normal(size=(n, 3)) * 0.184
    t0 = rng.uniform(t_lo, t_hi, n)
    return dict(m=np.ones(n), M=M, rs=rs, x0=x0, v=v, t0=t0, tw=np.full(n, tw))


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = np.maximum(np.abs(b), 1e-300)
    return np.max(np.abs(a - b) / np.maximum(scale, np.abs(b).max() * 1e-6 + 1e-300))


def scaled_err(a, b, tol, ref=None):
    """max over components of |a-b| / (tol * (1 + |ref|)), per leading row."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    ref = b if ref is None else np.asarray(ref, dtype=np.float64)
    d = np.abs(a - b) / (tol * (1.0 + np.abs(ref)))
    return d.reshape(d.shape[0], -1).max(axis=1)


def orphan_chenab_prog_today():
    """Present-day phase-space position of the Orphan-Chenab progenitor exactly as examples/OrphanChenab_mw_lmc_example.ipynb cells 2-3
    compute it, without astropy / jax: (phi1, phi2) -> ICRS with the notebook's rotation matrix, position through the reference's
    JaxCoords.alpha_delta_to_simcart (JaxCoords.py:48-85), velocity through astropy's ICRS -> Galactocentric with its v4.0 defaults -
    the same rotation R and tilt H, solar motion (12.9, 245.6, 7.78) km/s - and astropy's unit conversions."""
    alpha_gc, delta_gc, eta, d_gc, z_sun = np.deg2rad(266.4051), np.deg2rad(-28.936175), np.deg2rad(58.5986320306), 8.122, 0.0208
    R1 = np.array([[np.cos(delta_gc), 0, np.sin(delta_gc)], [0, 1.0, 0], [-np.sin(delta_gc), 0, np.cos(delta_gc)]])
    R2 = np.array([[np.cos(alpha_gc), np.sin(alpha_gc), 0.0], [-np.sin(alpha_gc), np.cos(alpha_gc), 0.0], [0, 0, 1.0]])
    R3 = np.array([[1.0, 0, 0], [0, np.cos(eta), np.sin(eta)], [0, -np.sin(eta), np.cos(eta)]])
    R = R3 @ (R1 @ R2)
    th = np.arcsin(z_sun / d_gc)
    H = np.array([[np.cos(th), 0, np.sin(th)], [0, 1, 0], [-np.sin(th), 0, np.cos(th)]])
    R_oc = np.array([[-0.44761231, -0.08785756, -0.88990128], [-0.84246097, 0.37511331, 0.38671632], [0.29983786, 0.92280606, -0.2419219]])
    phi1, phi2 = np.deg2rad(6.34), np.deg2rad(-0.441)                      # column 1 of arXiv:1812.08192 (notebook cell 3)
    dist, vr, pmdec, pmra = 18.764, 107.573, 2.880, -3.681                 # kpc, km/s, mas/yr, mas/yr (pm_ra_cosdec)
    w = np.linalg.inv(R_oc) @ np.array([np.cos(phi1) * np.cos(phi2), np.sin(phi1) * np.cos(phi2), np.sin(phi2)])
    a, de = np.arctan2(w[1], w[0]), np.arcsin(w[2])
    a, de = np.deg2rad(np.rad2deg(a)), np.deg2rad(np.rad2deg(de))          # the notebook goes through degrees
    e_r = np.array([np.cos(a) * np.cos(de), np.sin(a) * np.cos(de), np.sin(de)])
    e_a = np.array([-np.sin(a), np.cos(a), 0.0])
    e_d = np.array([-np.cos(a) * np.sin(de), -np.sin(a) * np.sin(de), np.cos(de)])
    q = H @ (R @ (dist * e_r) - d_gc * np.array([1.0, 0, 0]))
    k_pm = 4.740470463533348                                               # km/s per (mas/yr x kpc)
    v = H @ (R @ (vr * e_r + k_pm * dist * (pmra * e_a + pmdec * e_d))) + np.array([12.9, 245.6, 7.78])
    return np.hstack([q, v * 1.0227121650537077e-3])                      # km/s -> kpc/Myr


def notebook_batch_ics():
    """The 1000 initial conditions of tests.ipynb cells 12 + 15: numpy's legacy generator (np.random.seed(4934202); frozen stream) around
    sol.evaluate(0.0) = w0 = [20, 15, 20, .08, .1, -.05]."""
    rs = np.random.RandomState(4934202)
    n = 1000
    ics = np.hstack([rs.normal(loc=0, scale=.01, size=(n, 3)), rs.normal(loc=0, scale=.005, size=(n, 3))])
    return ics + np.array([20.0, 15.0, 20.0, 0.08, 0.1, -0.05])


# ---- adaptive-run parity: what two correct implementations of the same adaptive solver can be asked to agree on ----
def ulp_ensemble(orc, w0, t0, t1, K=8, seed=0, **kw):
    """K oracle runs whose initial conditions are moved by +-1 ulp (relative 2.2e-16) per component: the spread of their results is what the
    ALGORITHM (diffrax's controller in IEEE doubles) does with rounding-level input noise, independent of any implementation."""
    rng = np.random.default_rng(seed)
    w0 = np.asarray(w0, dtype=np.float64)
    runs = []
    for _ in range(K):
        w0p = w0 * (1.0 + (rng.integers(0, 2, w0.shape) * 2 - 1) * 2.220446049250313e-16)
        ys = orc.integrate_orbits(w0p, t0, t1, **kw)[0]
        runs.append(ys if kw.get("ts") is not None else ys[:, 0])          # saved rows [N,M,6] or final states [N,6]
    return np.array(runs)


def adaptive_parity_stats(cand, base, ens, truth, tol):
    """Per-orbit numbers, all in units of tol * (1 + |truth|):
       d    |cand - base|                                   (the parity difference under test)
       dens |ens_k - base| per ensemble member              (what 1-ulp input noise does to the oracle itself)
       E    max over {base, ens} of |. - truth|             (the solver's own global error for that orbit at that tolerance)"""
    sc = tol * (1.0 + np.abs(truth))
    mx = lambda a: (np.abs(a) / sc).reshape(len(sc), -1).max(axis=1)
    d = mx(cand - base)
    dens = np.array([mx(e - base) for e in ens])
    E = np.maximum(mx(base - truth), np.array([mx(e - truth) for e in ens]).max(axis=0))
    return d, dens, E


def assert_adaptive_parity(cand, base, ens, truth, tol, what=""):
    """The candidate against the oracle's base run, judged with the oracle's own 1-ulp ensemble:
      (1) per orbit, |cand - base| <= 10 tol + 3 E_i (triangle inequality through the truth; E_i = that orbit's global error over the
          ensemble, i.e. what the SOLVER is off by at this tolerance) - for all orbits but at most 1.5 % (the ensemble's own leave-one-out
          exception rate: a changed step sequence occasionally lands on a heavy tail of the global-error distribution);
      (2) the candidate is as accurate as the oracle: median / 90th percentile / maximum of |cand - truth| are within 1.5 x the largest
          of the ensemble members' (+ 10 tol);
      (3) sanity: the fraction of orbits within 10 x tol of the base run is not more than 0.12 below the worst ensemble member's.  It is
          REPORTED, not matched: the CUDA path evaluates the embedded error estimate in Nystrom form, whose rounding floor is lower than
          the direct form's (oracle, diffrax), so after a forced-small first step its second step differs from the oracle's by more than
          two direct-form runs differ from each other - the step sequences then differ more, the results by the same global error."""
    finite = np.isfinite(base).reshape(len(base), -1).all(axis=1)
    assert np.array_equal(finite, np.isfinite(cand).reshape(len(cand), -1).all(axis=1)), what + ": inf pattern differs"
    cand, base, ens, truth = cand[finite], base[finite], ens[:, finite], truth[finite]
    d, dens, E = adaptive_parity_stats(cand, base, ens, truth, tol)
    n = len(d)
    f_c, f_e = np.mean(d <= 10.0), np.mean(dens <= 10.0, axis=1)
    exc = d > 10.0 + 3.0 * E
    sc = tol * (1.0 + np.abs(truth))
    mx = lambda a: (np.abs(a) / sc).reshape(n, -1).max(axis=1)
    d_t, e_t = mx(cand - truth), np.array([mx(e - truth) for e in ens] + [mx(base - truth)])
    pct = {q: (np.percentile(d_t, q), np.percentile(e_t, q, axis=1).max()) for q in (50, 90, 100)}
    msg = (f"{what}: within 10 x tol of the oracle: {f_c:.3f} (oracle 1-ulp ensemble {f_e.min():.3f}..{f_e.max():.3f}); max |cand - oracle| = {d.max():.3g} x tol "
           f"(ensemble {dens.max():.3g}); per-orbit bound 10 tol + 3 E_i exceeded by {int(exc.sum())}/{n}; error vs 1e-13 solution, candidate / ensemble: "
           + ", ".join(f"p{q} {a:.3g} / {b:.3g}" for q, (a, b) in pct.items()) + " x tol")
    assert exc.sum() <= max(1, int(0.015 * n)), msg
    for q, (a, b) in pct.items():
        assert a <= 1.5 * b + 10.0, msg
    assert f_c >= f_e.min() - 0.12, msg
    return msg
