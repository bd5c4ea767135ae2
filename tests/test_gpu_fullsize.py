"""BASELINE.json configurations at FULL size (1e6-particle streams, 1000-subhalo responses) on the B200, checked through
size-independent properties plus oracle parity on random sub-samples (the oracle cannot integrate 1e6 orbits in seconds):
  * every solve succeeds; the energy of every particle is conserved at solver accuracy (static potential);
  * forward-then-backward round trip returns to the release conditions;
  * a random sub-sample of the 1e6 particles equals the oracle's integration of the same initial conditions;
  * the sharded runs (rank r of W integrates i = r mod W) reproduce the unsharded stream bit for bit;
  * the response D is linear in the subhalo masses, invariant under a permutation of the subhalos, zero for zero mass.
"""
import numpy as np
import pytest

import oracle as O
from common import assert_adaptive_parity, ulp_ensemble, mw3_oracle, mw3_product, relerr, scaled_err

pytestmark = pytest.mark.gpu

PROG_TODAY = [20.0, 0.0, 20.0, 0.0, 0.15, 0.0]
TRUTH = dict(solver=8, rtol=1e-13, atol=1e-13, dtmin=1e-3, max_steps=400_000, threads=8)


def _energy(pot, w):
    import torch
    phi, = __import__("streamsculptor_b200")._runtime.potential_eval(pot, w[:, :3].contiguous(), torch.zeros(w.shape[0], dtype=torch.float64, device=w.device), ("phi",))
    return 0.5 * (w[:, 3:] ** 2).sum(1) + phi


def test_c2_million_particle_stream_properties(cuda):
    torch = cuda
    import streamsculptor_b200 as ssc
    from streamsculptor_b200 import _runtime as rt
    pot, orc = mw3_product(), mw3_oracle()
    back = pot.integrate_orbit(w0=PROG_TODAY, ts=np.array([0.0, -3000.0]), t0=0.0, t1=-3000.0).ys[-1]
    n_ts = 500_001
    ts = rt.to_dev(np.linspace(-3000.0, 0.0, n_ts))
    lead, trail, status, nsteps = pot.gen_stream_vmapped(ts=ts, prog_w0=rt.to_dev(back), Msat=1e4, seed_num=583, solver=ssc.Dopri8(), _return_stats=True)
    assert lead.shape == (n_ts - 1, 6) and trail.shape == (n_ts - 1, 6)
    assert int((status != 0).sum()) == 0 and bool(torch.isfinite(lead).all()) and bool(torch.isfinite(trail).all())
    steps = int(nsteps[..., 0].sum())
    assert 0.8e8 < steps < 1.3e8                                   # ~1e8 particle-steps
    # ---- release conditions of the same stream (device), energy conservation of all 1e6 particles ----
    pl, pt, vl, vt = pot.gen_stream_ics(ts=ts, prog_w0=rt.to_dev(back), Msat=1e4, seed_num=583, solver=ssc.Dopri8())
    w_l, w_t = torch.cat([pl, vl], 1)[:-1].contiguous(), torch.cat([pt, vt], 1)[:-1].contiguous()
    for w_start, w_end in ((w_l, lead), (w_t, trail)):
        e0, e1 = _energy(pot, w_start), _energy(pot, w_end)
        rel = ((e1 - e0).abs() / e0.abs()).cpu().numpy()
        assert rel.max() < 2e-5 and np.median(rel) < 2e-6          # rtol = atol = 1e-7 over <= 3 Gyr
    # ---- time reversal: integrate the final states back to their release times ----
    w_end = torch.cat([lead, trail]).contiguous()
    t_rel = torch.cat([ts[:-1], ts[:-1]]).contiguous()
    sol = pot.integrate_orbit_batch_vmapped(w0=w_end, ts=t_rel.reshape(-1, 1), t0=torch.zeros_like(t_rel), t1=t_rel, solver=ssc.Dopri8(), rtol=1e-9, atol=1e-9,
                                            dtmin=0.05)
    w_back = sol.ys[:, 0]
    w_start = torch.cat([w_l, w_t])
    d = (w_back - w_start).abs().cpu().numpy() / (1.0 + w_start.abs().cpu().numpy())
    assert np.percentile(d.max(1), 99) < 2e-4 and np.median(d.max(1)) < 2e-5   # the forward leg's own error (tol 1e-7), amplified by shear
    # ---- random sub-sample against the oracle (same release conditions, adaptive criterion of DESIGN.md section 4) ----
    rng = np.random.default_rng(0)
    sel = np.sort(rng.choice(n_ts - 1, 1500, replace=False))
    w0s = w_l[sel].cpu().numpy()
    t0s = ts[sel].cpu().numpy()
    yo, sto, _ = orc.integrate_orbits(w0s, t0s, 0.0, solver=8, rtol=1e-7, atol=1e-7, dtmin=0.3, threads=8)
    yt, _, _ = orc.integrate_orbits(w0s, t0s, 0.0, **TRUTH)
    assert not sto.any()
    ens = ulp_ensemble(orc, w0s, t0s, 0.0, K=6, seed=4, solver=8, rtol=1e-7, atol=1e-7, dtmin=0.3, threads=8)
    assert_adaptive_parity(lead[sel].cpu().numpy(), yo[:, 0], ens, yt[:, 0], 1e-7, what="C2 sub-sample")
    # ---- sharding: rank r of 4 integrates releases i = r (mod 4); interleaving the shares reproduces the stream exactly ----
    parts = [rt.gen_stream(pot, pot, pot._G, ts, rt.to_dev(back), rt.to_dev(np.full(n_ts, 1e4)), 583, ssc.main.DEFAULT_KVALS, None,
                           rt.make_ctrl(ssc.Dopri8(), 1e-7, 1e-7, 0.3, None, 10_000), i_begin=r, i_stride=4) for r in range(4)]
    for r in range(4):
        assert torch.equal(parts[r][0], lead[r::4]) and torch.equal(parts[r][1], trail[r::4])


def test_c3_million_particles_moving_perturber(cuda):
    """C3: MW3 + moving Plummer on a 1000-knot linear table; the inline fast-extras kernel must equal the generic interpreter and,
    on a sub-sample, the oracle."""
    torch = cuda
    import streamsculptor_b200 as ssc
    from streamsculptor_b200 import _runtime as rt
    P = ssc.potential
    mw3, orc = mw3_product(), mw3_oracle()
    tk = np.linspace(-3000.0, 0.0, 1000)
    lmc = mw3.integrate_orbit(w0=[-1.0, -41.0, -28.0, -0.058, -0.23, 0.23], ts=tk[::-1].copy(), t0=0.0, t1=-3000.0).ys[::-1, :3].copy()
    tv = np.linspace(-3000.0, 0.0, 64)
    vel = np.stack([1e-3 * np.sin(tv / 500.0), 2e-3 * np.cos(tv / 800.0), 1e-4 * tv / 3000.0], axis=1)
    plum = P.TimeDepTranslatingPotential(P.PlummerPotential(m=1.5e11, r_s=10.8, units=ssc.usys), ssc.LinearTrack(tk, lmc), units=ssc.usys)
    acc = P.UniformAcceleration(ssc.LinearTrack(tv, vel), units=ssc.usys)
    tr = orc.track(O.LINEAR, tk, lmc)
    orc.plummer(1.5e11, 10.8, track=tr).uniform_acc(tv, vel)
    c3 = P.Potential_Combine([mw3, plum, acc], units=ssc.usys)                                   # fused MW3 + 2 fast extras
    # same physics with a massless static Plummer appended: 3 extras -> the generic interpreter path
    c3_generic = P.Potential_Combine([mw3, plum, acc, P.PlummerPotential(m=0.0, r_s=1.0, units=ssc.usys)], units=ssc.usys)
    back = c3.integrate_orbit(w0=PROG_TODAY, ts=np.array([0.0, -3000.0]), t0=0.0, t1=-3000.0).ys[-1]
    n_ts = 500_001
    ts = rt.to_dev(np.linspace(-3000.0, 0.0, n_ts))
    lead, trail, status, nsteps = c3.gen_stream_vmapped(ts=ts, prog_w0=rt.to_dev(back), Msat=1e4, seed_num=583, solver=ssc.Dopri8(), _return_stats=True)
    assert int((status != 0).sum()) == 0 and bool(torch.isfinite(lead).all()) and bool(torch.isfinite(trail).all())
    # fixed step: fast path == interpreter path to rounding, on 20 000 particles of the same stream
    ts_s = ts[::50].contiguous()
    kw = dict(ts=ts_s, prog_w0=rt.to_dev(back), Msat=1e4, seed_num=583, solver=ssc.Dopri8(), dtmin=1.0, dtmax=1.0)
    lf, tf = c3.gen_stream_vmapped(**kw)
    lg, tg = c3_generic.gen_stream_vmapped(**kw)
    assert relerr(lf.cpu().numpy(), lg.cpu().numpy()) < 1e-10 and relerr(tf.cpu().numpy(), tg.cpu().numpy()) < 1e-10
    # sub-sample of the fixed-step stream against the oracle
    pl, pt, vl, vt = c3.gen_stream_ics(ts=ts_s, prog_w0=rt.to_dev(back), Msat=1e4, seed_num=583, solver=ssc.Dopri8(), dtmin=1.0, dtmax=1.0)
    sel = np.arange(0, ts_s.shape[0] - 1, 25)
    w0s = torch.cat([pl, vl], 1)[sel].cpu().numpy()
    yo, sto, _ = orc.integrate_orbits(w0s, ts_s[sel].cpu().numpy(), 0.0, solver=8, dtmin=1.0, dtmax=1.0, threads=8)
    assert not sto.any() and relerr(lf[sel].cpu().numpy(), yo[:, 0]) < 1e-10


def test_c4_thousand_subhalo_response_properties(cuda):
    """C4: 1000 Hernquist subhalos; fixed-step runs make the properties exact up to rounding."""
    import streamsculptor_b200 as ssc
    from streamsculptor_b200 import _runtime as rt
    from common import subhalo_set
    P = ssc.potential
    pot = mw3_product()
    back = pot.integrate_orbit(w0=PROG_TODAY, ts=np.array([0.0, -3000.0]), t0=0.0, t1=-3000.0).ys[-1]
    ts = np.linspace(-3000.0, 0.0, 1001)
    nr = np.random.Generator(np.random.PCG64(0)).standard_normal((len(ts), 4))
    pl, pt, vl, vt = pot.gen_stream_ics(ts=ts, prog_w0=back, Msat=1e4, seed_num=583, solver=ssc.Dopri8(), normals=nr)
    w0 = np.vstack([np.hstack([pl, vl])[:-1], np.hstack([pt, vt])[:-1]])              # 2000 particles
    t0 = np.concatenate([ts[:-1], ts[:-1]])
    n_sh = 1000
    sh = subhalo_set(n_sh, seed=7, tw=150.0)
    ctrl = rt.make_ctrl(ssc.Dopri8(), 1e-7, 1e-7, 2.0, 2.0, 10_000)                    # dtmin = dtmax: fixed 2 Myr steps

    def run(m, order=None):
        idx = np.arange(n_sh) if order is None else order
        pert = P.SubhaloLinePotentialCustom_fromFunc(func=P.HernquistPotential, m=m[idx], r_s=sh["rs"][idx], subhalo_x0=sh["x0"][idx], subhalo_v=sh["v"][idx],
                                                     subhalo_t0=sh["t0"][idx], t_window=150.0, units=ssc.usys)
        w, D, st, ns = rt.linear_response(pot, pert._arrays, rt.to_dev(w0), None, rt.to_dev(t0), 0.0, ctrl)
        assert int((st != 0).sum()) == 0
        return w.cpu().numpy(), D.cpu().numpy()

    m1 = np.ones(n_sh)
    w1, D1 = run(m1)
    assert D1.shape == (2000, n_sh, 12) and np.isfinite(D1).all()
    # linear in the subhalo masses (the response ODE is linear and homogeneous in m): exact power-of-two scaling
    m2 = m1.copy(); m2[::2] = 4.0; m2[1::3] = 0.0
    w2, D2 = run(m2)
    assert np.array_equal(w1, w2)
    assert np.array_equal(D2, D1 * m2[None, :, None])
    # permutation of the subhalos permutes D (every subhalo's response is independent of the others in a fixed-step run)
    perm = np.random.default_rng(3).permutation(n_sh)
    w3, D3 = run(m1, perm)
    assert np.array_equal(D3, D1[:, perm])
    # a particle released after every window closed has zero response; subhalos whose window opens after t = 0 never act
    late = sh["t0"] - 150.0 > 0.0
    assert not np.any(D1[:, late])
