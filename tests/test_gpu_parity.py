"""CUDA path vs the CPU oracle on the same seeded inputs (run with -m gpu on the B200 box).

Tolerances (BASELINE.json north_star): fixed-step runs <= 1e-10 relative; adaptive runs <= 10 x rtol/atol;
closed-form field evaluations <= 1e-11 relative against the oracle's autodiff of the reference's scalar formulas.
"""
import numpy as np
import pytest

import oracle as O
from common import assert_adaptive_parity, halo_orbits, lmc_track, mw3_oracle, mw3_product, random_orbits, relerr, scaled_err, subhalo_set, ulp_ensemble

TRUTH = dict(solver=8, rtol=1e-13, atol=1e-13, dtmin=1e-3, max_steps=400_000, threads=8)

pytestmark = pytest.mark.gpu


def _full_pair(cuda):
    """MW4 + translating Plummer on a linear track + uniform acceleration + Hernquist subhalos, built twice."""
    import streamsculptor_b200 as ssc
    P = ssc.potential
    t, y = lmc_track()
    tv = np.linspace(-3000, 0, 50)
    vel = np.stack([1e-3 * np.sin(tv / 500.0), 2e-3 * np.cos(tv / 800.0), 1e-4 * tv / 3000.0], axis=1)
    sh = subhalo_set(12, tw=400.0)
    orc = O.Program().miyamoto(6.8e10, 3.0, 0.28).hernquist(5e9, 1.0).hernquist(1.71e9, 0.07, soft=1e-4).nfw(5.4e11, 15.62)
    tr = orc.track(O.LINEAR, t, y)
    orc.plummer(1.5e11, 10.8, track=tr).isochrone(3e10, 2.0).triaxnfw(1e11, 12.0, 1.0, 0.9, 0.8)
    orc.uniform_acc(tv, vel)
    orc.subhalos(O.PR_HERNQUIST, sh["M"], sh["rs"], sh["x0"], sh["v"], sh["t0"], sh["tw"])
    prod = P.Potential_Combine([
        P.MiyamotoNagaiDisk(m=6.8e10, a=3.0, b=0.28, units=ssc.usys), P.HernquistPotential(m=5e9, r_s=1.0, units=ssc.usys),
        P.HernquistPotential(m=1.71e9, r_s=0.07, soft=1e-4, units=ssc.usys), P.NFWPotential(m=5.4e11, r_s=15.62, units=ssc.usys),
        P.TimeDepTranslatingPotential(P.PlummerPotential(m=1.5e11, r_s=10.8, units=ssc.usys), ssc.LinearTrack(t, y), units=ssc.usys),
        P.Isochrone(m=3e10, a=2.0, units=ssc.usys), P.TriaxialNFWPotential(m=1e11, r_s=12.0, q1=1.0, q2=0.9, q3=0.8, units=ssc.usys),
        P.UniformAcceleration(ssc.LinearTrack(tv, vel), units=ssc.usys),
        P.SubhaloLinePotentialCustom_fromFunc(func=P.HernquistPotential, m=sh["M"], r_s=sh["rs"], subhalo_x0=sh["x0"], subhalo_v=sh["v"],
                                              subhalo_t0=sh["t0"], t_window=sh["tw"], units=ssc.usys)], units=ssc.usys)
    return orc, prod


def test_field_closed_forms_match_oracle_autodiff(cuda):
    orc, prod = _full_pair(cuda)
    rng = np.random.default_rng(3)
    xyz = rng.normal(size=(257, 3)) * np.array([15, 15, 8.0])
    t = rng.uniform(-3200, 100, 257)
    g = prod.gradient(xyz, t)
    h = prod.jacobian_force(xyz, t)
    assert relerr(g, orc.gradient(xyz, t)) < 1e-11
    assert relerr(h, orc.hessian(xyz, t)) < 1e-10
    # potentials: everything except the force-only uniform acceleration
    import streamsculptor_b200 as ssc
    cons = ssc.potential.Potential_Combine([p for p in prod.potential_list if type(p).__name__ != "UniformAcceleration"], units=ssc.usys)
    assert relerr(cons.potential(xyz, t), orc.potential(xyz, t)) < 1e-12


def test_goldens_through_the_product_api(cuda):
    import streamsculptor_b200 as ssc
    nfw = ssc.potential.NFWPotential(m=1e12, r_s=20.0, units=ssc.usys)
    rhs = nfw.velocity_acceleration(0.0, np.array([1., 2., 3., .2, .3, .4]))          # golden D1 (tests.ipynb cell 58)
    assert np.allclose(rhs, [0.2, 0.3, 0.4, -0.0011937, -0.00238739, -0.00358109], rtol=0, atol=5e-9)
    sol = nfw.integrate_orbit(w0=[20., 15., 20., .08, .1, -.05], ts=np.array([0.0, 30.0, 3000.0]))   # golden D2 (tests.ipynb cell 17)
    assert np.allclose(sol.ys[1], [21.9793661, 17.67654451, 18.10530132, 0.051923, 0.07815599, -0.07552167], rtol=0, atol=6e-9)
    mw = ssc.potential.GalaMilkyWayPotential(units=ssc.usys)                                # golden D5 (StreamSubhaloExample cell 1)
    ic = mw.integrate_orbit(w0=[20.0, 0.0, 20, .0, .15, .0], ts=np.linspace(0, -3500, 1000), t0=0.0, t1=-3500).ys[-1]
    d5 = [-7.23164146, -7.96692572, -10.81840286, 0.19182623, -0.20351324, -0.01770436]
    # today's source (Hernquist softening removed since the notebook ran): measured 2e-6.  The notebook-era revision reproduces the printed
    # digits in the oracle (3e-8, tests/test_oracle_goldens.py); through the CUDA path that run differs by 1.6e-5 - another adaptive step
    # sequence over 3.5 Gyr (the solver's own error on this number is 8.7e-5), see DESIGN.md section 4 - so it is not asserted here
    assert np.allclose(ic, d5, rtol=0, atol=5e-6)


@pytest.mark.parametrize("solver", [5, 8])
def test_fixed_step_orbits_1e10(cuda, solver):
    import streamsculptor_b200 as ssc
    orc, prod = mw3_oracle(), mw3_product()
    w0 = halo_orbits(96, seed=5)       # well-resolved orbits: rounding differences are not amplified by near-centre passages
    t0 = np.linspace(-3000, -100, 96)
    for h in (0.5, 1.0):
        ys_o, st_o, ns_o = orc.integrate_orbits(w0, t0, 0.0, solver=solver, dtmin=h, dtmax=h)
        sol = prod.integrate_orbit_batch_vmapped(w0=w0, ts=np.zeros((96, 1)), t0=t0, t1=0.0, solver=ssc.Dopri8() if solver == 8 else ssc.Dopri5(),
                                                 dtmin=h, dtmax=h, max_steps=10_000)
        assert (np.asarray(sol.result) == 0).all()
        assert np.array_equal(sol.stats["num_steps"], ns_o[:, 0])
        assert relerr(sol.ys[:, 0], ys_o[:, 0]) < 1e-10


# adaptive final states: tests/test_gpu_adaptive_parity.py (lock-step controller traces, 1-ulp ensemble statistics, strict short runs)


def test_saved_snapshots_and_backward_integration(cuda):
    """integrate_orbit_batch_vmapped with per-orbit save times ts[N,M] (main.py:186-202), forwards and backwards, in a
    potential with every component type.  Fixed steps pin the in-kernel dense output at 1e-10; adaptive runs are compared
    at the accuracy of the interpolant."""
    import streamsculptor_b200 as ssc
    orc, prod = _full_pair(cuda)
    w0 = halo_orbits(40, seed=2)
    ts = np.linspace(-2500, -100, 17)[None, :] + np.linspace(0, 50, 40)[:, None]
    for solver in (5, 8):
        sv = ssc.Dopri8() if solver == 8 else ssc.Dopri5()
        for tsx in (ts, ts[:, ::-1].copy()):                      # forwards, then backwards (t1 < t0, ts decreasing)
            ys_f, _, ns_f = orc.integrate_orbits(w0, tsx[:, 0], tsx[:, -1], ts=tsx, solver=solver, dtmin=1.0, dtmax=1.0, threads=8)
            sol_f = prod.integrate_orbit_batch_vmapped(w0=w0, ts=tsx, t0=tsx[:, 0], t1=tsx[:, -1], solver=sv, dtmin=1.0, dtmax=1.0)
            assert np.array_equal(sol_f.ys[:, 0], w0)              # ts[0] == t0 returns y0 exactly
            assert np.array_equal(sol_f.stats["num_steps"], ns_f[:, 0])
            assert scaled_err(sol_f.ys, ys_f, 1e-10).max() < 1.0
            # adaptive: the subhalo windows switch forces on/off discontinuously (potential.py:826, no jump_ts), so the solver's
            # own error is ~1e3-1e4 x tol here and step sequences decorrelate; require equal accuracy against a 1e-13 solution
            ys_o, _, _ = orc.integrate_orbits(w0, tsx[:, 0], tsx[:, -1], ts=tsx, solver=solver, dtmin=0.05, threads=8)
            ys_t, _, _ = orc.integrate_orbits(w0, tsx[:, 0], tsx[:, -1], ts=tsx, **TRUTH)
            ens = ulp_ensemble(orc, w0, tsx[:, 0], tsx[:, -1], K=6, seed=1, ts=tsx, solver=solver, dtmin=0.05, threads=8)
            sol = prod.integrate_orbit_batch_vmapped(w0=w0, ts=tsx, t0=tsx[:, 0], t1=tsx[:, -1], solver=sv, dtmin=0.05)
            assert_adaptive_parity(np.asarray(sol.ys), ys_o, ens, ys_t, 1e-7, what=f"snapshots Dopri{solver}")


def test_dense_single_orbit_matches_oracle_saveat(cuda):
    import streamsculptor_b200 as ssc
    orc, prod = mw3_oracle(), mw3_product()
    w0 = [20.0, 0.0, 20.0, 0.0, 0.15, 0.0]
    ts = np.linspace(-3000.0, 0.0, 1001)
    for solver in (5, 8):
        sv = ssc.Dopri8() if solver == 8 else ssc.Dopri5()
        ys_o, _, ns_o = orc.integrate_orbits(w0, -3000.0, 0.0, ts=ts, solver=solver)
        sol = prod.integrate_orbit(w0=w0, ts=ts, solver=sv)
        assert abs(int(sol.stats["num_steps"]) - ns_o[0, 0]) <= 2
        assert scaled_err(sol.ys, ys_o[0], 1e-7).max() < 10.0
        dense = prod.integrate_orbit(w0=w0, ts=ts, dense=True, solver=sv)
        assert np.allclose(dense.evaluate(ts[37]), sol.ys[37], rtol=0, atol=1e-12)
        # fixed steps: identical sequences, so the interpolants themselves are compared (1e-10)
        ys_f, _, _ = orc.integrate_orbits(w0, -3000.0, 0.0, ts=ts, solver=solver, dtmin=2.0, dtmax=2.0)
        sol_f = prod.integrate_orbit(w0=w0, ts=ts, solver=sv, dtmin=2.0, dtmax=2.0)
        assert scaled_err(sol_f.ys, ys_f[0], 1e-10).max() < 1.0


def test_failure_semantics(cuda):
    import streamsculptor_b200 as ssc
    prod = mw3_product()
    w0 = random_orbits(4, seed=1)
    sol = prod.integrate_orbit_batch_vmapped(w0=w0, ts=np.zeros((4, 1)), t0=np.array([-3000.0, -3000.0, 0.0, -10.0]), t1=0.0, max_steps=5)
    res = np.asarray(sol.result)
    assert res[0] == 1 and res[1] == 1 and np.isinf(sol.ys[0]).all()          # max_steps reached -> rows stay +inf (main.py:136)
    assert res[2] == 0 and np.isinf(sol.ys[2]).all()                          # t0 == t1: the loop never runs, nothing is saved
    assert res[3] == 0 and np.isfinite(sol.ys[3]).all()
    with pytest.raises(NotImplementedError):
        ssc.potential.CustomPotential(potential_func=lambda x, t: 0.0, units=ssc.usys)
    with pytest.raises(NotImplementedError):
        prod.integrate_orbit(w0=w0[0], ts=np.array([0.0, 1.0]), solver="Tsit5")


def test_release_model_and_stream_c1(cuda):
    """BASELINE config C1: 2000-particle spray stream, 3 Gyr, Dopri8, static MW3 - full path vs the oracle."""
    import streamsculptor_b200 as ssc
    orc, prod = mw3_oracle(), mw3_product()
    prog_today = [20.0, 0.0, 20.0, 0.0, 0.15, 0.0]
    back, _, _ = orc.integrate_orbits(prog_today, 0.0, -3000.0)
    ts = np.linspace(-3000.0, 0.0, 1001)
    for normals in (None, np.random.Generator(np.random.PCG64(0)).standard_normal((1001, 4))):
        ics_o = orc.gen_stream_ics(ts, back[0, 0], 1e4, 583, solver=8, normals=normals)
        ics_p = prod.gen_stream_ics(ts=ts, prog_w0=back[0, 0], Msat=1e4, seed_num=583, solver=ssc.Dopri8(), normals=normals)
        for a, b in zip(ics_p, ics_o[:4]):
            assert scaled_err(a, b, 1e-7).max() < 10.0
        lead_o, trail_o, st_o, ns_o = orc.gen_stream(ts, back[0, 0], 1e4, 583, solver=8, normals=normals)
        lead, trail, status, nsteps = prod.gen_stream_vmapped(ts=ts, prog_w0=back[0, 0], Msat=1e4, seed_num=583, solver=ssc.Dopri8(),
                                                              normals=normals, _return_stats=True)
        assert lead.shape == (1000, 6) and trail.shape == (1000, 6) and (status == 0).all()
        prog_t, _, _ = orc.integrate_orbits(back[0, 0], -3000.0, 0.0, ts=ts, **TRUTH)
        pl, pt_, vl, vt = orc.release(prog_t[0], 1e4, np.arange(1001), ts, 583, normals=normals)
        tl, _, _ = orc.integrate_orbits(np.hstack([pl, vl])[:-1], ts[:-1], 0.0, **TRUTH)
        tt_, _, _ = orc.integrate_orbits(np.hstack([pt_, vt])[:-1], ts[:-1], 0.0, **TRUTH)
        # the particle solves as members of the oracle's 1-ulp ensemble (started from the ORACLE's release conditions; the CUDA release
        # conditions were compared above); tl / tt_ (a 1e-13 progenitor AND 1e-13 particle solves) bound the error of the whole pipeline
        for arm, got, base, truth_all in (("lead", lead, lead_o, tl), ("trail", trail, trail_o, tt_)):
            w_rel = np.hstack([ics_o[0], ics_o[2]])[:-1] if arm == "lead" else np.hstack([ics_o[1], ics_o[3]])[:-1]
            ens = ulp_ensemble(orc, w_rel, ts[:-1], 0.0, K=6, seed=2, solver=8, threads=8)
            truth = orc.integrate_orbits(w_rel, ts[:-1], 0.0, **TRUTH)[0][:, 0]
            assert_adaptive_parity(np.asarray(got), base, ens, truth, 1e-7, what=f"C1 {arm}")
            assert np.percentile(scaled_err(got, truth_all[:, 0], 1e-7), 99) < 3e4            # whole pipeline vs 1e-13 everywhere: the solver's global error
    # fixed-step C1 (dtmin = dtmax = 1 Myr): whole pipeline to 1e-10 relative
    nr = np.random.Generator(np.random.PCG64(0)).standard_normal((1001, 4))
    lo, to, _, _ = orc.gen_stream(ts, back[0, 0], 1e4, 583, solver=8, normals=nr, dtmin=1.0, dtmax=1.0)
    lf, tf = prod.gen_stream_vmapped(ts=ts, prog_w0=back[0, 0], Msat=1e4, seed_num=583, solver=ssc.Dopri8(), normals=nr, dtmin=1.0, dtmax=1.0)
    assert scaled_err(lf, lo, 1e-10).max() < 1.0 and scaled_err(tf, to, 1e-10).max() < 1.0
    # single release call, jax.random recipe, decaying mass array (StreamSubhaloExample cell 2)
    pl, pt, vl, vt = prod.release_model(x=back[0, 0, :3], v=back[0, 0, 3:], Msat=1e4, i=7, t=-3000.0, seed_num=583)
    ref = orc.release(back[0, 0][None], 1e4, [7], [-3000.0], 583)
    assert relerr(np.hstack([pl, pt, vl, vt]), np.hstack([r[0] for r in ref])) < 1e-10


def test_with_pert_streams(cuda):
    import streamsculptor_b200 as ssc
    P = ssc.potential
    sh = subhalo_set(5, seed=4, tw=250.0)
    orc_base, base = mw3_oracle(), mw3_product()
    orc_tot = mw3_oracle().subhalos(O.PR_PLUMMER, sh["M"], sh["rs"], sh["x0"], sh["v"], sh["t0"], sh["tw"])
    pert = P.SubhaloLinePotential(m=sh["M"], a=sh["rs"], subhalo_x0=sh["x0"], subhalo_v=sh["v"], subhalo_t0=sh["t0"], t_window=250.0, units=ssc.usys)
    ts = np.linspace(-2000.0, 0.0, 301)
    w0 = [-5.0, 12.0, 9.0, 0.12, 0.05, -0.06]
    nr = np.random.Generator(np.random.PCG64(5)).standard_normal((301, 4))
    # oracle version of gen_stream_vmapped_with_pert: progenitor + particles in total, release in base
    prog, _, _ = orc_tot.integrate_orbits(w0, -2000.0, 0.0, ts=ts, solver=5, dtmin=0.1)
    pl, pt, vl, vt = orc_base.release(prog[0], 1e4, np.arange(301), ts, 0, normals=nr)
    w_l, w_t = np.hstack([pl, vl])[:-1], np.hstack([pt, vt])[:-1]
    yl, _, _ = orc_tot.integrate_orbits(w_l, ts[:-1], 0.0, solver=5, dtmin=0.1)
    yt, _, _ = orc_tot.integrate_orbits(w_t, ts[:-1], 0.0, solver=5, dtmin=0.1)
    lead, trail = ssc.gen_stream_vmapped_with_pert(pot_base=base, pot_pert=pert, ts=ts, prog_w0=w0, Msat=1e4, seed_num=0, normals=nr)
    assert np.mean(scaled_err(lead, yl[:, 0], 1e-7) < 10.0) > 0.9 and np.mean(scaled_err(trail, yt[:, 0], 1e-7) < 10.0) > 0.9
    assert scaled_err(lead, yl[:, 0], 1e-5).max() < 10.0 and scaled_err(trail, yt[:, 0], 1e-5).max() < 10.0


def test_linear_response_matches_oracle(cuda):
    import streamsculptor_b200 as ssc
    from streamsculptor_b200 import _runtime as rt
    P = ssc.potential
    nsh = 24
    sh = subhalo_set(nsh, seed=9, t_lo=-1500.0)
    orc_base, base = mw3_oracle(), mw3_product()
    for prof_o, func in ((O.PR_HERNQUIST, P.HernquistPotential), (O.PR_PLUMMER, P.PlummerPotential)):
        orc_sh = O.Program().subhalos(prof_o, sh["m"], sh["rs"], sh["x0"], sh["v"], sh["t0"], sh["tw"])
        pert = P.SubhaloLinePotentialCustom_fromFunc(func=func, m=sh["m"], r_s=sh["rs"], subhalo_x0=sh["x0"], subhalo_v=sh["v"],
                                                     subhalo_t0=sh["t0"], t_window=sh["tw"], units=ssc.usys)
        # field term at a generic state (fields.py:175-206)
        rng = np.random.default_rng(0)
        y = np.concatenate([[8.0, -3.0, 4.0, 0.1, 0.12, -0.05], rng.normal(size=12 * nsh) * 1e-3])
        dy_o = O.response_term(orc_base, orc_sh, sh["t0"][3] + 20.0, y)
        dy = rt.response_term(base, pert._arrays, sh["t0"][3] + 20.0, y).cpu().numpy()
        assert relerr(dy, dy_o) < 1e-11
        # per-subhalo gradients incl. d/dr_s (perturbative.py:695-696)
        phi_o, g_o = orc_sh.per_sh([3.0, 2.0, 1.0], sh["t0"][5] + 10.0)
        assert relerr(pert.gradient_per_SH([3.0, 2.0, 1.0], sh["t0"][5] + 10.0), g_o) < 1e-11
        # full solves: zero ICs (Chen25 convention) and non-zero ICs, Dopri8 and Dopri5
        w0 = random_orbits(12, seed=21)
        t0 = np.linspace(-1800.0, -50.0, 12)
        D0 = rng.normal(size=(12, nsh, 12)) * 1e-4
        for solver, tol, d0 in ((8, 1e-7, None), (8, 1e-9, D0), (5, 1e-7, D0)):
            w_o, D_o, st_o, ns_o = O.linear_response(orc_base, orc_sh, w0, t0, 0.0, D0=d0, solver=solver, rtol=tol, atol=tol, dtmin=0.01)
            ctrl = rt.make_ctrl(ssc.Dopri8() if solver == 8 else ssc.Dopri5(), tol, tol, 0.01, None, 10_000)
            w, D, st, ns = rt.linear_response(base, pert._arrays, rt.to_dev(w0), None if d0 is None else rt.to_dev(d0), rt.to_dev(t0), 0.0, ctrl)
            w, D, st, ns = w.cpu().numpy(), D.cpu().numpy(), st.cpu().numpy(), ns.cpu().numpy()
            assert (st == 0).all() and (st_o == 0).all()
            w_t, D_t, _, _ = O.linear_response(orc_base, orc_sh, w0, t0, 0.0, D0=d0, solver=8, rtol=1e-13, atol=1e-13, dtmin=1e-3,
                                               max_steps=400_000, threads=8)
            pack = lambda ww, DD: np.hstack([ww, DD.reshape(12, -1)])
            ens, rng_e = [], np.random.default_rng(7)
            for _ in range(5):               # the oracle's own 1-ulp ensemble of the coupled solve
                w0p = w0 * (1.0 + (rng_e.integers(0, 2, w0.shape) * 2 - 1) * 2.220446049250313e-16)
                w_e, D_e, _, _ = O.linear_response(orc_base, orc_sh, w0p, t0, 0.0, D0=d0, solver=solver, rtol=tol, atol=tol, dtmin=0.01)
                ens.append(pack(w_e, D_e))
            assert_adaptive_parity(pack(w, D), pack(w_o, D_o), np.array(ens), pack(w_t, D_t), tol, what=f"response Dopri{solver} tol={tol}")
            # fixed steps: same sequence by construction -> 1e-10
            w_f, D_f, _, _ = O.linear_response(orc_base, orc_sh, w0, t0, 0.0, D0=d0, solver=solver, dtmin=2.0, dtmax=2.0)
            ctrl_f = rt.make_ctrl(ssc.Dopri8() if solver == 8 else ssc.Dopri5(), tol, tol, 2.0, 2.0, 10_000)
            wf, Df, _, nsf = rt.linear_response(base, pert._arrays, rt.to_dev(w0), None if d0 is None else rt.to_dev(d0), rt.to_dev(t0), 0.0, ctrl_f)
            assert scaled_err(wf.cpu().numpy(), w_f, 1e-10).max() < 1.0
            assert scaled_err(Df.cpu().numpy().reshape(12, -1), D_f.reshape(12, -1), 1e-10).max() < 1.0


def test_response_c4_scale_fixed_step(cuda):
    """BASELINE config C4 shape: 1000 Hernquist subhalos per particle (sorted / skipped / tiled inside the kernel).  Fixed steps
    make the step sequence identical by construction, so the whole [N, 1000, 12] response is compared at 1e-10."""
    import streamsculptor_b200 as ssc
    from streamsculptor_b200 import _runtime as rt
    P = ssc.potential
    nsh = 1000
    rng = np.random.Generator(np.random.PCG64(1234))
    M = 10 ** rng.uniform(5, 9, nsh); rs = 1.05 * np.sqrt(M / 1e8)
    sh = dict(x0=rng.normal(size=(nsh, 3)) * 12.0, v=rng.normal(size=(nsh, 3)) * 0.184, t0=rng.uniform(-1500.0, 100.0, nsh))
    base, orc_base = mw3_product(), mw3_oracle()
    pert = P.SubhaloLinePotentialCustom_fromFunc(func=P.HernquistPotential, m=np.ones(nsh), r_s=rs, subhalo_x0=sh["x0"], subhalo_v=sh["v"],
                                                 subhalo_t0=sh["t0"], t_window=150.0, units=ssc.usys)
    orc_sh = O.Program().subhalos(O.PR_HERNQUIST, np.ones(nsh), rs, sh["x0"], sh["v"], sh["t0"], 150.0)
    w0 = halo_orbits(6, seed=4)
    t0 = np.array([-1500.0, -1200.0, -900.0, -600.0, -300.0, -20.0])
    w_f, D_f, st_f, ns_f = O.linear_response(orc_base, orc_sh, w0, t0, 0.0, solver=8, dtmin=2.0, dtmax=2.0, threads=8)
    ctrl = rt.make_ctrl(ssc.Dopri8(), 1e-9, 1e-9, 2.0, 2.0, 10_000)
    w, D, st, ns = rt.linear_response(base, pert._arrays, rt.to_dev(w0), None, rt.to_dev(t0), 0.0, ctrl)
    assert (st.cpu().numpy() == 0).all() and np.array_equal(ns.cpu().numpy()[:, 0], ns_f[:, 0])
    assert scaled_err(w.cpu().numpy(), w_f, 1e-10).max() < 1.0
    D = D.cpu().numpy()
    assert np.abs(D - D_f).max() <= 1e-10 * np.abs(D_f).max()
    never = sh["t0"] + 150.0 < t0[-1]                        # windows that closed before the last particle was released
    assert (D[-1, never] == 0.0).all() and (D_f[-1, never] == 0.0).all()
    # non-zero ICs disable the unborn-subhalo shortcut: same answer as the oracle
    D0 = rng.normal(size=(6, nsh, 12)) * 1e-9
    w_g, D_g, _, _ = O.linear_response(orc_base, orc_sh, w0[:2], t0[:2], 0.0, D0=D0[:2], solver=5, dtmin=2.0, dtmax=2.0, threads=8)
    ctrl5 = rt.make_ctrl(ssc.Dopri5(), 1e-9, 1e-9, 2.0, 2.0, 10_000)
    w2, D2, _, _ = rt.linear_response(base, pert._arrays, rt.to_dev(w0[:2]), rt.to_dev(D0[:2]), rt.to_dev(t0[:2]), 0.0, ctrl5)
    assert np.abs(D2.cpu().numpy() - D_g).max() <= 1e-10 * np.abs(D_g).max()


def test_response_multi_slot_kernel_bit_identical(cuda):
    """response_kernel_mp keeps several particles in flight per CTA (shared base-orbit phase, csrc/ssb_response.cu).  Every particle's arithmetic
    is independent of the slot and CTA it lands in, so 1, 2 and 4 slots per CTA and the one-particle-per-CTA kernel must agree bit for bit -
    adaptive steps, zero and non-zero perturbation ICs, forward and backward integration, failures included."""
    import os
    import streamsculptor_b200 as ssc
    from streamsculptor_b200 import _runtime as rt
    P = ssc.potential
    nsh = 70
    sh = subhalo_set(nsh, seed=5, t_lo=-2000.0, tw=200.0)
    sh["tw"][::7] = 60.0                                        # unequal windows: the dead-prefix logic must use the running maximum of the ends
    base = mw3_product()
    pert = P.SubhaloLinePotentialCustom_fromFunc(func=P.HernquistPotential, m=sh["m"], r_s=sh["rs"], subhalo_x0=sh["x0"], subhalo_v=sh["v"],
                                                 subhalo_t0=sh["t0"], t_window=sh["tw"], units=ssc.usys)
    N = 37
    w0 = halo_orbits(N, seed=11)
    t0 = np.linspace(-2500.0, -5.0, N)
    t0[5] = 0.0                                                 # nothing to integrate: +inf rows
    rng = np.random.default_rng(2)
    D0 = rng.normal(size=(N, nsh, 12)) * 1e-6
    cases = [(ssc.Dopri8(), 1e-8, None, 0.0, 10_000), (ssc.Dopri5(), 1e-7, D0, 0.0, 10_000), (ssc.Dopri8(), 1e-9, None, -3000.0, 10_000),
             (ssc.Dopri8(), 1e-10, None, 0.0, 40)]             # last: max_steps reached by the long orbits
    try:
        for solver, tol, d0, t1, max_steps in cases:
            ctrl = rt.make_ctrl(solver, tol, tol, 0.01, None, max_steps)
            out = {}
            wa_out = {}                                         # the warp-autonomous form of the kernel (csrc/ssb_response_wa.cuh)
            for retire in (0, 1):
                os.environ["SSB_RESP_RETIRE"] = str(retire)
                for np_slots in (0, 1, 2, 4, 8, 16):
                    os.environ["SSB_RESP_NP"] = str(np_slots)
                    for kern in ("mp", "wa") if (np_slots and os.environ.get("SSB_TEST_RESP_WA") == "1") else ("mp",):   # wa: variant builds only
                        os.environ["SSB_RESP_KERNEL"] = kern
                        w, D, st, ns = rt.linear_response(base, pert._arrays, rt.to_dev(w0), None if d0 is None else rt.to_dev(d0), rt.to_dev(t0), t1, ctrl)
                        (out if kern == "mp" else wa_out)[retire, np_slots] = (w.cpu().numpy(), D.cpu().numpy(), st.cpu().numpy(), ns.cpu().numpy())
            for key, res in wa_out.items():                     # same arithmetic in the same order: the same bits, retired items or not
                for a, b in zip(res, out[key]):
                    assert np.array_equal(a, b), f"warp-autonomous kernel differs from the CTA-wide one at (retire, slots) = {key}"
            ref = out[0, 0]
            if max_steps == 40:
                assert (ref[2] == 1).any() and (ref[2] == 0).any()
            else:
                assert (ref[2] == 0).all()
            assert np.isinf(ref[0][5]).all() == (t1 == 0.0)
            for np_slots in (1, 2, 4, 8, 16):     # every item swept (SSB_RESP_RETIRE=0): the one-particle kernel's bits
                for a, b in zip(out[0, np_slots], ref):
                    assert np.array_equal(a, b), f"{np_slots} slots per CTA differ from the one-particle kernel"
            for np_slots in (2, 4, 8, 16):        # retired items (the default): independent of the slot count as well
                for a, b in zip(out[1, np_slots], out[1, 1]):
                    assert np.array_equal(a, b), f"{np_slots} slots per CTA differ from one slot (retired items)"
            # retired vs swept: the same step sequences up to the atol-only scale of the retired items in the error norm (relative 1e-8),
            # the same responses up to the rounding of the propagator products
            assert np.array_equal(out[1, 4][2], ref[2])
            fin = np.isfinite(ref[1])
            assert np.array_equal(np.isfinite(out[1, 4][1]), fin) and np.array_equal(np.isfinite(out[1, 4][0]), np.isfinite(ref[0]))
            ok = ref[2] == 0
            assert np.abs(out[1, 4][3][ok] - ref[3][ok]).max() <= (0 if (d0 is not None or t1 < 0) else 2)          # step counts
            assert scaled_err(out[1, 4][0][ok & np.isfinite(ref[0]).all(1)], ref[0][ok & np.isfinite(ref[0]).all(1)], tol).max() < 1e-2
            sel = ok & fin.reshape(N, -1).all(1)
            dD = np.abs(out[1, 4][1][sel] - ref[1][sel]).max()
            assert dD <= 1e-3 * tol * (1.0 + np.abs(ref[1][sel]).max()), dD
            if d0 is not None or t1 < 0:          # non-zero ICs / backward time: nothing is retired, bit-identical
                assert np.array_equal(out[1, 4][1], ref[1])
    finally:
        os.environ.pop("SSB_RESP_NP", None)
        os.environ.pop("SSB_RESP_RETIRE", None)
        os.environ.pop("SSB_RESP_KERNEL", None)
    # default slot count at a batch large enough to use it: same bits as the one-particle kernel
    N2 = 1200
    w0b = halo_orbits(N2, seed=12); t0b = np.linspace(-1000.0, -5.0, N2)
    ctrl = rt.make_ctrl(ssc.Dopri8(), 1e-7, 1e-7, 0.01, None, 10_000)
    wr, Dr, sr, nr_ = rt.linear_response(base, pert._arrays, rt.to_dev(w0b), None, rt.to_dev(t0b), 0.0, ctrl)      # the default configuration
    try:
        os.environ["SSB_RESP_RETIRE"] = "0"
        wa, Da, sa, na = rt.linear_response(base, pert._arrays, rt.to_dev(w0b), None, rt.to_dev(t0b), 0.0, ctrl)
        os.environ["SSB_RESP_NP"] = "0"
        wb, Db, sb_, nb = rt.linear_response(base, pert._arrays, rt.to_dev(w0b), None, rt.to_dev(t0b), 0.0, ctrl)
    finally:
        os.environ.pop("SSB_RESP_NP", None)
        os.environ.pop("SSB_RESP_RETIRE", None)
    assert bool((Da == Db).all()) and bool((wa == wb).all()) and bool((na == nb).all()) and int((sa != 0).sum()) == 0
    assert int((sr != 0).sum()) == 0 and int((nr_ - nb).abs().max()) <= 2
    assert float(((wr - wb).abs() / (1e-7 * (1 + wb.abs()))).max()) < 1e-2 and float((Dr - Db).abs().max()) <= 1e-10 * (1.0 + float(Db.abs().max()))


def test_response_generator_api(cuda):
    """GenerateMassRadiusPerturbation_Chen25.compute_perturbation_OTF shape contract (golden D7) + oracle parity."""
    import streamsculptor_b200 as ssc
    P, pt = ssc.potential, ssc.perturbative
    nsh, nts = 10, 40
    sh = subhalo_set(nsh, seed=13, t_lo=-900.0)
    base, orc_base = mw3_product(), mw3_oracle()
    ts = np.linspace(-1000.0, 0.0, nts)
    prog_w0 = [12.0, 3.0, -6.0, -0.05, 0.15, 0.03]
    rng = np.random.default_rng(2)
    prog = np.asarray(base.integrate_orbit(w0=prog_w0, ts=ts, solver=ssc.Dopri8()).ys)
    ics = [prog[:, :3] + rng.normal(size=(nts, 3)) * 0.05, prog[:, :3] - rng.normal(size=(nts, 3)) * 0.05,
           prog[:, 3:] + rng.normal(size=(nts, 3)) * 1e-3, prog[:, 3:] - rng.normal(size=(nts, 3)) * 1e-3]
    model = pt.BaseStreamModelChen25(pot_base=base, ts=ts, prog_w0=prog_w0, Msat=1e4, solver=ssc.Dopri8(), stream_ics=ics, prog_fwd=prog)
    pert = P.SubhaloLinePotentialCustom_fromFunc(func=P.HernquistPotential, m=sh["m"], r_s=sh["rs"], subhalo_x0=sh["x0"], subhalo_v=sh["v"],
                                                 subhalo_t0=sh["t0"], t_window=150.0, units=ssc.usys)
    gen = pt.GenerateMassRadiusPerturbation_Chen25(potential_base=base, potential_perturbation=pert, BaseStreamModel=model, units=ssc.usys)
    w, D = gen.compute_perturbation_OTF(cpu=False, solver=ssc.Dopri8(), rtol=1e-8, atol=1e-8, dtmin=0.01)
    assert w.shape == (2 * nts - 1, 6) and D.shape == (2 * nts - 1, nsh, 12)
    orc_sh = O.Program().subhalos(O.PR_HERNQUIST, sh["m"], sh["rs"], sh["x0"], sh["v"], sh["t0"], sh["tw"])
    bs = gen.base_stream
    w_o, D_o, st_o, _ = O.linear_response(orc_base, orc_sh, bs.streamICs[:-1], bs.ts[:-1], bs.ts[-1], solver=8, rtol=1e-8, atol=1e-8, dtmin=0.01)
    ok = np.isfinite(w_o).all(axis=1)
    assert np.array_equal(ok, np.isfinite(w).all(axis=1)) and ok.sum() == 2 * nts - 2     # the duplicated last release time has zero span
    assert np.mean(scaled_err(w[ok], w_o[ok], 1e-8) < 10.0) > 0.8 and scaled_err(w[ok], w_o[ok], 1e-6).max() < 10.0
    assert np.mean(scaled_err(D[ok].reshape(ok.sum(), -1), D_o[ok].reshape(ok.sum(), -1), 1e-8) < 10.0) > 0.8


def test_cubic_track_and_host_entry_points(cuda):
    import ctypes as C
    import streamsculptor_b200 as ssc
    from streamsculptor_b200 import _lib, _runtime as rt
    P = ssc.potential
    t, y = lmc_track(n=80)
    orc = mw3_oracle()
    tr = orc.track(O.CUBIC, t, y)
    orc.plummer(2e10, 3.0, track=tr)
    prod = P.Potential_Combine([mw3_product(), P.TimeDepTranslatingPotential(P.PlummerPotential(m=2e10, r_s=3.0, units=ssc.usys),
                                                                             ssc.CubicTrack(t, y), units=ssc.usys)], units=ssc.usys)
    tq = np.linspace(-2990.0, -10.0, 333)
    c_o, d_o = orc.track_eval(tr, tq)
    track = prod.potential_list[1]._track
    assert relerr(track(tq), c_o) < 1e-12 and relerr(track(tq, derivative=True), d_o) < 1e-10
    xyz = np.random.default_rng(1).normal(size=(50, 3)) * 10
    assert relerr(prod.gradient(xyz, tq[:50]), orc.gradient(xyz, tq[:50])) < 1e-11
    # host-pointer C-ABI entry (what a CPU-side plugin call binds): same numbers as the device-pointer path
    w0 = random_orbits(33, seed=8)
    t0 = np.full(33, -1500.0); t1 = np.zeros(33); ts = np.zeros((33, 1))
    pot_static = mw3_product()
    Pst, keep = rt.lower(pot_static)
    host = _lib.Potential.from_buffer_copy(Pst)          # static program: no device pointers inside
    ys = np.empty((33, 1, 6)); st = np.empty(33, np.int32); ns = np.empty((33, 3), np.int32)
    ctrl = rt.make_ctrl(ssc.Dopri8(), 1e-7, 1e-7, 0.3, None, 10_000)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    _lib.check(_lib.lib().ssb_orbit_integrate_host(C.byref(host), 33, p(w0), p(t0), p(t1), p(ts), 1, 1, ctrl, p(ys), p(st), p(ns)))
    sol = pot_static.integrate_orbit_batch_vmapped(w0=w0, ts=ts, t0=t0, t1=t1)
    assert np.array_equal(ys, sol.ys) and (st == 0).all()


def test_host_entry_points_zero_copy_outputs(cuda):
    """ssb_gen_stream_host / ssb_orbit_integrate_host with PINNED outputs (the orbit kernel writes host memory directly, csrc/ssb_host.cu)
    against pageable outputs (staged copies) and the device-pointer path: bit-identical, including partially filled warps."""
    import ctypes as C
    import torch
    import streamsculptor_b200 as ssc
    from streamsculptor_b200 import _lib, _runtime as rt
    pot = mw3_product()
    Pst, keep = rt.lower(pot)
    host = _lib.Potential.from_buffer_copy(Pst)
    ctrl = rt.make_ctrl(ssc.Dopri8(), 1e-7, 1e-7, 0.3, None, 10_000)
    L = _lib.lib()
    hp = lambda t: C.c_void_p(t.data_ptr())
    for nts in (78, 2, 1026):                      # 2 n = 154, 2, 2050 orbits: none a multiple of the warp size
        n = nts - 1
        ts = np.linspace(-1500.0, 0.0, nts)
        w0 = np.array([-3.0, 14.0, 8.0, 0.14, 0.02, -0.07])
        ms = np.full(nts, 1e4)
        kv = (C.c_double * 8)(*ssc.main.DEFAULT_KVALS)
        tt, tw, tm = torch.from_numpy(ts), torch.from_numpy(w0), torch.from_numpy(ms)
        res = []
        for pinned in (True, False):
            out = torch.full((2, n, 6), -7.0, dtype=torch.float64)
            st = torch.full((2, n), -7, dtype=torch.int32)
            ns = torch.full((2, n, 3), -7, dtype=torch.int32)
            if pinned:
                out, st, ns = out.pin_memory(), st.pin_memory(), ns.pin_memory()
            _lib.check(L.ssb_gen_stream_host(C.byref(host), C.byref(host), pot._G, nts, hp(tt), hp(tw), hp(tm), 583, kv, None, ctrl, 0, 1, n,
                                             hp(out[0]), hp(out[1]), hp(st), hp(ns)))
            res.append((out.clone(), st.clone(), ns.clone()))
        assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][2], res[1][2])
        assert (res[0][1] == 0).all() and (res[0][2][..., 0] == res[0][2][..., 1] + res[0][2][..., 2]).all()
        lead, trail = pot.gen_stream_vmapped(ts=ts, prog_w0=w0, Msat=1e4, seed_num=583, solver=ssc.Dopri8())
        assert np.array_equal(res[0][0][0].numpy(), np.asarray(lead)) and np.array_equal(res[0][0][1].numpy(), np.asarray(trail))
    # batch of orbits, final state only (ts aliases t1), pinned outputs
    N = 77
    w0 = torch.from_numpy(random_orbits(N, seed=8)); t0 = torch.full((N,), -1500.0, dtype=torch.float64); t1 = torch.zeros(N, dtype=torch.float64)
    ys = torch.full((N, 1, 6), -7.0, dtype=torch.float64).pin_memory()
    st = torch.full((N,), -7, dtype=torch.int32).pin_memory(); ns = torch.full((N, 3), -7, dtype=torch.int32).pin_memory()
    _lib.check(L.ssb_orbit_integrate_host(C.byref(host), N, hp(w0), hp(t0), hp(t1), hp(t1), 1, 1, ctrl, hp(ys), hp(st), hp(ns)))
    sol = pot.integrate_orbit_batch_vmapped(w0=w0.numpy(), ts=np.zeros((N, 1)), t0=t0.numpy(), t1=t1.numpy())
    assert (st == 0).all() and relerr(ys.numpy(), sol.ys) < 1e-12       # final-state kernel vs SaveAt kernel: same steps, t1 row
    assert (ns[:, 0] == ns[:, 1] + ns[:, 2]).all() and (ns[:, 0] > 10).all()


def test_two_part_stream_pipeline_bit_identical(cuda):
    """Large streams run gen_stream as two concurrent parts (early particles behind a cut-short progenitor solve, the rest behind the full one,
    csrc/ssb_kernels.cu).  Same bits as the one-part pipeline, also for a sharded selection and through the host entry point."""
    import os
    import streamsculptor_b200 as ssc
    from streamsculptor_b200 import _runtime as rt
    pot = mw3_product()
    nts = 40_001                                                # 40 000 particles per arm: above the split threshold
    ts = rt.to_dev(np.linspace(-2000.0, 0.0, nts))
    w0 = rt.to_dev(np.array([-3.0, 14.0, 8.0, 0.14, 0.02, -0.07]))
    ms = rt.to_dev(np.full(nts, 1e4))
    ctrl = rt.make_ctrl(ssc.Dopri8(), 1e-7, 1e-7, 0.3, None, 10_000)
    kv = ssc.main.DEFAULT_KVALS
    res = {}
    try:
        for mode in ("1", "0"):
            os.environ["SSB_STREAM_SPLIT"] = mode
            full = rt.gen_stream(pot, pot, pot._G, ts, w0, ms, 583, kv, None, ctrl)
            shard = rt.gen_stream(pot, pot, pot._G, ts, w0, ms, 583, kv, None, ctrl, i_begin=0, i_stride=1, n_local=36_000)
            res[mode] = [t.cpu().numpy().copy() for t in full] + [t.cpu().numpy().copy() for t in shard]
    finally:
        os.environ.pop("SSB_STREAM_SPLIT", None)
    for a, b in zip(res["1"], res["0"]):
        assert np.array_equal(a, b)
    assert (res["1"][2] == 0).all() and np.isfinite(res["1"][0]).all() and res["1"][0].shape == (nts - 1, 6)
    assert np.array_equal(res["1"][4], res["1"][0][:36_000])   # a prefix selection equals the prefix of the whole stream


def test_third_derivatives_and_release_jacobian(cuda):
    """A15: closed-form third derivatives and jacfwd(release_model) (perturbative.py:281-296) vs the oracle's nested autodiff."""
    import streamsculptor_b200 as ssc
    orc, prod = _full_pair(cuda)
    rng = np.random.default_rng(5)
    xyz = rng.normal(size=(64, 3)) * np.array([14, 14, 7.0])
    t = rng.uniform(-2900, -100, 64)
    assert relerr(prod.third_derivative(xyz, t), orc.third(xyz, t)) < 1e-9
    orc3, mw3 = mw3_oracle(), mw3_product()
    prog = halo_orbits(50, seed=8)
    ts = np.linspace(-2000.0, 0.0, 50)
    nr = np.random.Generator(np.random.PCG64(2)).standard_normal((50, 4))
    for normals in (nr, None):
        J = mw3.release_jacobian(prog, 1e4, np.arange(50), ts, 493, normals=normals)
        Jo = orc3.release(prog, 1e4, np.arange(50), ts, 493, normals=normals, jacobian=True)
        assert J.shape == (50, 2, 6, 6) and np.abs(J - Jo).max() < 1e-9 * np.abs(Jo).max()
    assert np.allclose(J[:, 0] + J[:, 1], 2 * np.eye(6), atol=1e-12)          # lead/trail are mirror images: J_lead + J_trail = 2 I (golden D8's structure)


def test_saved_response_trajectory_and_classic_generator(cuda):
    """A12-A15 classic path: BaseStreamModel + GenerateMassRadiusPerturbation (perturbative.py:24-296): backward progenitor
    response saved at every stripping time, release Jacobian, perturbation ICs, lead / trail responses."""
    import streamsculptor_b200 as ssc
    from streamsculptor_b200 import _runtime as rt
    P, pt = ssc.potential, ssc.perturbative
    nsh, nts = 8, 31
    sh = subhalo_set(nsh, seed=17, t_lo=-900.0)
    sh["x0"] = sh["x0"] * 0.3 + np.array([12.0, 3.0, -6.0])
    base, orc_base = mw3_product(), mw3_oracle()
    ts = np.linspace(-1000.0, 0.0, nts)
    prog_w0 = [12.0, 3.0, -6.0, -0.05, 0.15, 0.03]
    nr = np.random.Generator(np.random.PCG64(4)).standard_normal((nts, 4))
    pert = P.SubhaloLinePotential(m=sh["m"], a=sh["rs"], subhalo_x0=sh["x0"], subhalo_v=sh["v"], subhalo_t0=sh["t0"], t_window=150.0, units=ssc.usys)
    struct = P.SubhaloLinePotential_dRadius(m=sh["m"], a=sh["rs"], subhalo_x0=sh["x0"], subhalo_v=sh["v"], subhalo_t0=sh["t0"], t_window=150.0, units=ssc.usys)
    orc_sh = O.Program().subhalos(O.PR_PLUMMER, sh["m"], sh["rs"], sh["x0"], sh["v"], sh["t0"], 150.0)
    fixed = dict(rtol=1e-8, atol=1e-8, dtmin=2.0, dtmax=2.0)                  # identical step sequences -> 1e-10 comparisons
    model = pt.BaseStreamModel(potential_base=base, prog_w0=prog_w0, ts=ts, Msat=1e4, seednum=7, solver=ssc.Dopri8(), normals=nr, **fixed)
    # oracle pieces of the same constructor
    prog_o, _, _ = orc_base.integrate_orbits(prog_w0, ts[0], ts[-1], ts=ts, solver=8, **fixed)
    assert scaled_err(model.prog_loc_fwd, prog_o[0], 1e-10).max() < 1.0
    Jo = orc_base.release(prog_o[0], 1e4, np.arange(nts), ts, 7, normals=nr, jacobian=True)
    assert np.abs(model.dRel_dIC - Jo).max() < 1e-8
    gen = pt.GenerateMassRadiusPerturbation(potential_base=base, potential_perturbation=pert, potential_structural=struct, BaseStreamModel=model,
                                            units=ssc.usys, solver=ssc.Dopri8(), max_steps=5000, **fixed)
    ws_o, Ds_o, st_o, _ = O.linear_response_saveat(orc_base, orc_sh, prog_o[0, -1], ts[-1], ts[0], ts[::-1].copy(), solver=8, max_steps=5000, **fixed)
    assert st_o[0] == 0
    F_o = Ds_o[::-1]
    assert np.abs(gen.prog_fieldICs - F_o).max() <= 1e-10 * np.abs(F_o).max()
    assert scaled_err(gen.prog_base.ys[0], ws_o, 1e-10).max() < 1.0              # the base orbit integrated backwards alongside
    lead_ic = np.dstack([np.einsum('ijk,ilk->ilj', Jo[:, 0], F_o[:, :, :6]), np.einsum('ijk,ilk->ilj', Jo[:, 0], F_o[:, :, 6:])])
    assert np.abs(gen.perturbation_ICs_lead - lead_ic).max() <= 1e-8 * np.abs(lead_ic).max()
    (wl, Dl), (wt, Dt) = gen.compute_perturbation_OTF(cpu=False, solver=ssc.Dopri8(), **fixed)
    assert wl.shape == (nts - 1, 6) and Dl.shape == (nts - 1, nsh, 12) and Dt.shape == Dl.shape
    pl, pt_, vl, vt = orc_base.release(prog_o[0], 1e4, np.arange(nts), ts, 7, normals=nr)
    w_o, D_o, _, _ = O.linear_response(orc_base, orc_sh, np.hstack([pl, vl])[:-1], ts[:-1], 0.0, D0=gen.perturbation_ICs_lead[:-1], solver=8, **fixed)
    assert scaled_err(wl, w_o, 1e-10).max() < 1.0 and np.abs(Dl - D_o).max() <= 1e-9 * np.abs(D_o).max()
    # adaptive saved trajectory through the public integrate_field (fields.py:35-99), forwards
    fld = ssc.fields.MassRadiusPerturbation_OTF(gen)
    D0 = np.random.default_rng(1).normal(size=(nsh, 12)) * 1e-6
    sol = ssc.integrate_field(w0=[np.asarray(prog_w0), D0], ts=ts, field=fld, solver=ssc.Dopri5(), rtol=1e-8, atol=1e-8, dtmin=0.05, max_steps=5000)
    ws2, Ds2, _, _ = O.linear_response_saveat(orc_base, orc_sh, prog_w0, ts[0], ts[-1], ts, D0=D0, solver=5, rtol=1e-8, atol=1e-8, dtmin=0.05, max_steps=5000)
    assert np.array_equal(sol.ys[0][0], prog_w0) and np.array_equal(sol.ys[1][0], D0)
    assert scaled_err(sol.ys[0], ws2, 1e-8).max() < 10.0 and scaled_err(sol.ys[1].reshape(nts, -1), Ds2.reshape(nts, -1), 1e-8).max() < 10.0


def test_custom_base_generator_perturbation_ics(cuda):
    """GenerateMassRadiusPerturbation_CustomBase without explicit ICs (perturbative.py:375-454): the progenitor's backward response saved at
    every stripping time becomes the particles' perturbation ICs (the additive release has an identity Jacobian), then the responses."""
    import streamsculptor_b200 as ssc
    P, pt = ssc.potential, ssc.perturbative
    nsh, nts = 6, 25
    sh = subhalo_set(nsh, seed=23, t_lo=-700.0)
    sh["x0"] = sh["x0"] * 0.3 + np.array([12.0, 3.0, -6.0])
    base, orc_base = mw3_product(), mw3_oracle()
    ts = np.linspace(-800.0, 0.0, nts)
    prog_w0 = [12.0, 3.0, -6.0, -0.05, 0.15, 0.03]
    rng = np.random.default_rng(3)
    pos_rel, vel_rel = rng.normal(size=(nts, 3)) * 0.05, rng.normal(size=(nts, 3)) * 1e-3
    pert = P.SubhaloLinePotentialCustom_fromFunc(func=P.HernquistPotential, m=sh["m"], r_s=sh["rs"], subhalo_x0=sh["x0"], subhalo_v=sh["v"],
                                                 subhalo_t0=sh["t0"], t_window=150.0, units=ssc.usys)
    orc_sh = O.Program().subhalos(O.PR_HERNQUIST, sh["m"], sh["rs"], sh["x0"], sh["v"], sh["t0"], 150.0)
    fixed = dict(rtol=1e-8, atol=1e-8, dtmin=2.0, dtmax=2.0)
    model = pt.CustomBaseStreamModel(potential_base=base, prog_w0=prog_w0, ts=ts, pos_rel=pos_rel, vel_rel=vel_rel, solver=ssc.Dopri8(), units=ssc.usys,
                                     **fixed)
    gen = pt.GenerateMassRadiusPerturbation_CustomBase(potential_base=base, potential_perturbation=pert, BaseStreamModel=model, units=ssc.usys,
                                                       solver=ssc.Dopri8(), max_steps=5000, **fixed)
    prog_o, _, _ = orc_base.integrate_orbits(prog_w0, ts[0], ts[-1], ts=ts, solver=8, **fixed)
    ws_o, Ds_o, st_o, _ = O.linear_response_saveat(orc_base, orc_sh, prog_o[0, -1], ts[-1], ts[0], ts[::-1].copy(), solver=8, max_steps=5000, **fixed)
    F_o = Ds_o[::-1]
    assert st_o[0] == 0 and gen.perturbation_ICs.shape == (nts, nsh, 12)
    assert np.abs(gen.perturbation_ICs - F_o).max() <= 1e-9 * np.abs(F_o).max()
    w, D = gen.compute_perturbation_OTF(cpu=False, solver=ssc.Dopri8(), **fixed)
    ics = np.hstack([prog_o[0][:, :3] + pos_rel, prog_o[0][:, 3:] + vel_rel])
    w_o, D_o, _, _ = O.linear_response(orc_base, orc_sh, ics[:-1], ts[:-1], 0.0, D0=F_o[:-1], solver=8, **fixed)
    assert scaled_err(w, w_o, 1e-9).max() < 1.0 and np.abs(D - D_o).max() <= 1e-8 * np.abs(D_o).max()


def test_chen25_release_and_streams(cuda):
    """A9: release_model_Chen25 / gen_stream_ics_Chen25 / gen_stream_vmapped_Chen25 (streamhelpers.py:352-545) incl. the progenitor's
    own Plummer potential on an interpax-'cubic' track, vs the oracle.  Fixed steps -> 1e-10."""
    import streamsculptor_b200 as ssc
    from streamsculptor_b200 import streamhelpers as sh_
    P = ssc.potential
    base, orc = mw3_product(), mw3_oracle()
    ts = np.linspace(-1500.0, 0.0, 61)
    prog_w0 = [12.0, 3.0, -6.0, -0.05, 0.15, 0.03]
    fixed = dict(rtol=1e-8, atol=1e-8, dtmin=1.0, dtmax=1.0)
    mean, fac = sh_.CHEN25_MEAN, sh_.chen25_factor()
    assert np.allclose(fac @ fac.T, sh_.CHEN25_COV, atol=1e-10)
    prog_o, _, _ = orc.integrate_orbits(prog_w0, ts[0], ts[-1], ts=ts, solver=8, **fixed)
    for key, normals in ((1234, None), ((7, 99), None), (None, np.random.Generator(np.random.PCG64(9)).standard_normal((61, 6)))):
        ics, orb = ssc.gen_stream_ics_Chen25(pot_base=base, ts=ts, prog_w0=prog_w0, Msat=2e4, key=key, solver=ssc.Dopri8(), normals=normals, **fixed)
        ref = orc.release_chen25(prog_o[0], 2e4, ts, sh_.key_words(key), mean, fac, normals=normals)
        assert scaled_err(np.asarray(orb.ys), prog_o[0], 1e-10).max() < 1.0
        for a, b in zip(ics, ref):
            assert scaled_err(a, b, 1e-9).max() < 1.0
    # full stream with a live progenitor potential on the cubic track
    prog_pot = P.PlummerPotential(m=2e4, r_s=0.01, units=ssc.usys)
    lead, trail = ssc.gen_stream_vmapped_Chen25(base, ts, prog_w0, 2e4, 1234, solver=ssc.Dopri8(), prog_pot=prog_pot, **fixed)
    pl, pt, vl, vt = orc.release_chen25(prog_o[0], 2e4, ts, sh_.key_words(1234), mean, fac)
    tot = mw3_oracle()
    tr = tot.track(O.CUBIC, ts, prog_o[0][:, :3])
    tot.plummer(2e4, 0.01, track=tr)
    yl, _, _ = tot.integrate_orbits(np.hstack([pl, vl])[:-1], ts[:-1], 0.0, solver=8, **fixed)
    yt, _, _ = tot.integrate_orbits(np.hstack([pt, vt])[:-1], ts[:-1], 0.0, solver=8, **fixed)
    assert scaled_err(lead, yl[:, 0], 1e-9).max() < 1.0 and scaled_err(trail, yt[:, 0], 1e-9).max() < 1.0
    # BaseStreamModelChen25 draws its own release when none is supplied (perturbative.py:615-625)
    model = ssc.perturbative.BaseStreamModelChen25(pot_base=base, ts=ts, prog_w0=prog_w0, Msat=2e4, key=1234, solver=ssc.Dopri8(), **fixed)
    assert model.BaseModel.streamICs.shape == (2 * 61, 6) and np.all(np.diff(model.BaseModel.ts) >= 0)


def test_mw_lmc_potential_api(cuda):
    """A10: MW_LMC_Potential (potential.py:555-662) = MW3' + translating NFW LMC + uniform frame acceleration on 1000-knot
    LINEAR tables with linear extrapolation; force-only (no .potential).  Synthetic tables of the reference's shape."""
    import streamsculptor_b200 as ssc
    P = ssc.potential
    tk = np.linspace(-14000.0, 0.0, 1000)
    xyz = np.stack([-1.0 + 0.004 * tk + 60 * np.sin(tk / 4000.0), -41.0 - 0.006 * tk, -27.0 + 0.003 * tk + 20 * np.cos(tk / 3000.0)], axis=1)
    vel = np.stack([2e-3 * np.sin(tk / 2500.0), 1e-3 * np.cos(tk / 1800.0), 5e-8 * tk], axis=1)
    pot = P.MW_LMC_Potential(units=ssc.usys, t_lmc=tk, xyz_lmc=xyz, t_mw=tk, vel_mw=vel)
    orc = O.Program().hernquist(5e9, 1.0).miyamoto(5.0e10, 3.0, 0.3).nfw(5.4e11, 15.62)
    tr = orc.track(O.LINEAR, tk, xyz)
    orc.nfw(.85e11, (.85e11 / 1e11) ** 0.6 * 8.5, track=tr).uniform_acc(tk, vel)
    rng = np.random.default_rng(0)
    x = rng.normal(size=(100, 3)) * 20
    t = rng.uniform(-14500.0, 300.0, 100)                         # incl. extrapolation on both sides
    assert relerr(pot.gradient(x, t), orc.gradient(x, t)) < 1e-11
    assert relerr(pot.acceleration(x[0], t[0]), -orc.gradient(x[0], t[0])[0]) < 1e-11
    with pytest.raises(NotImplementedError):
        pot.potential(x[0], 0.0)
    w0 = halo_orbits(20, seed=3)
    ys_o, _, _ = orc.integrate_orbits(w0, -3000.0, 0.0, solver=8, dtmin=1.0, dtmax=1.0)
    sol = pot.integrate_orbit_batch_vmapped(w0=w0, ts=np.zeros((20, 1)), t0=-3000.0, t1=0.0, dtmin=1.0, dtmax=1.0)
    assert scaled_err(sol.ys[:, 0], ys_o[:, 0], 1e-10).max() < 1.0
    assert np.allclose(pot.LMC_center_spline(-7000.0), orc.track_eval(tr, [-7000.0])[0][0], rtol=1e-14)


def test_second_order_response(cuda):
    """A16: MassRadiusPerturbation_OTF_SecondOrder (fields.py:260-320) and compute_perturbation_second_order_OTF
    (perturbative.py:757-772) vs the oracle (third derivatives and per-subhalo Hessians by nested autodiff there)."""
    import streamsculptor_b200 as ssc
    from streamsculptor_b200 import _runtime as rt
    P = ssc.potential
    nsh = 7
    sh = subhalo_set(nsh, seed=23, t_lo=-900.0)
    sh["x0"] = sh["x0"] * 0.2 + np.array([12.0, 3.0, -6.0]); sh["rs"] = sh["rs"] + 0.3
    base, orc_base = mw3_product(), mw3_oracle()
    for prof_o, func in ((O.PR_HERNQUIST, P.HernquistPotential), (O.PR_PLUMMER, P.PlummerPotential)):
        orc_sh = O.Program().subhalos(prof_o, sh["M"], sh["rs"], sh["x0"], sh["v"], sh["t0"], sh["tw"])
        pert = P.SubhaloLinePotentialCustom_fromFunc(func=func, m=sh["M"], r_s=sh["rs"], subhalo_x0=sh["x0"], subhalo_v=sh["v"], subhalo_t0=sh["t0"],
                                                     t_window=sh["tw"], units=ssc.usys)
        rng = np.random.default_rng(3)
        y = np.concatenate([[11.0, 2.5, -5.0, 0.1, 0.12, -0.05], rng.normal(size=12 * nsh) * 1e-2, rng.normal(size=6 * nsh) * 1e-2])
        tq = sh["t0"][2] + 15.0
        assert relerr(rt.second_order_term(base, pert._arrays, tq, y).cpu().numpy(), O.second_order_term(orc_base, orc_sh, tq, y)) < 1e-9
        w0 = halo_orbits(5, seed=31); w0[:, :3] = np.array([12.0, 3.0, -6.0]) + np.random.default_rng(5).normal(size=(5, 3))
        t0 = np.linspace(-1000.0, -100.0, 5)
        for solver in (8, 5):
            w_o, D_o, E_o, st_o, ns_o = O.second_order_response(orc_base, orc_sh, w0, t0, 0.0, solver=solver, dtmin=1.0, dtmax=1.0)
            ctrl = rt.make_ctrl(ssc.Dopri8() if solver == 8 else ssc.Dopri5(), 1e-8, 1e-8, 1.0, 1.0, 10_000)
            w, D, E, st, ns = rt.second_order_response(base, pert._arrays, rt.to_dev(w0), None, None, rt.to_dev(t0), 0.0, ctrl)
            assert (st.cpu().numpy() == 0).all() and np.array_equal(ns.cpu().numpy()[:, 0], ns_o[:, 0])
            assert scaled_err(w.cpu().numpy(), w_o, 1e-10).max() < 1.0
            assert np.abs(D.cpu().numpy() - D_o).max() <= 1e-9 * np.abs(D_o).max() and np.abs(E.cpu().numpy() - E_o).max() <= 1e-9 * np.abs(E_o).max()
            assert np.abs(E_o).max() > 0
    # adaptive, through the public field API with non-zero ICs
    gen = type("G", (), {"potential_base_total": base, "subhalo_arrays": pert._arrays})()
    fld = ssc.fields.MassRadiusPerturbation_OTF_SecondOrder(gen)
    D0, E0 = rng.normal(size=(nsh, 12)) * 1e-8, rng.normal(size=(nsh, 6)) * 1e-8
    sol = ssc.integrate_field(w0=[w0[0], D0, E0], ts=np.array([-300.0, 0.0]), field=fld, solver=ssc.Dopri8(), rtol=1e-9, atol=1e-12, dtmin=0.01, max_steps=5000)
    w_o, D_o, E_o, _, _ = O.second_order_response(orc_base, orc_sh, w0[:1], -300.0, 0.0, D0=D0[None], E0=E0[None], solver=8, rtol=1e-9, atol=1e-12, dtmin=0.01)
    assert np.array_equal(sol.ys[2][0], E0) and np.abs(sol.ys[2][1] - E_o[0]).max() < 1e-4 * np.abs(E_o).max()
    assert np.abs(sol.ys[1][1] - D_o[0]).max() < 1e-4 * np.abs(D_o).max()


def _restricted_pair():
    """MW3 + Plummer progenitor on a cubic track, built for the oracle and for the product."""
    import streamsculptor_b200 as ssc
    P = ssc.potential
    orc0 = mw3_oracle()
    tk = np.linspace(-600.0, 0.0, 301)
    yk, _, _ = orc0.integrate_orbits(np.array([[20., 0, 20, 0, .15, 0]]), 0.0, -600.0, ts=tk[::-1].copy(), rtol=1e-12, atol=1e-12, dtmin=1e-3,
                                     max_steps=100_000)
    _restricted_pair.v0 = yk[0, -1, 3:].copy()
    yk = yk[0, ::-1, :3].copy()
    orc = mw3_oracle()
    tr = orc.track(O.CUBIC, tk, yk)
    orc.plummer(2e4, 0.01, track=tr)
    return orc, mw3_product(), ssc.CubicTrack(tk, yk), tk, yk


def test_restricted_nbody_shared_step(cuda):
    """A17: N tracers as ONE ODE with a shared controller (RestrictedNbody.py:93-106,131) vs the oracle."""
    import streamsculptor_b200 as ssc
    from streamsculptor_b200 import RestrictedNbody as RN
    orc, mw, track, tk, yk = _restricted_pair()
    rng = np.random.default_rng(5)
    N = 700
    w0 = np.hstack([yk[0] + rng.normal(size=(N, 3)) * 0.02, _restricted_pair.v0 + rng.normal(size=(N, 3)) * 5e-4])
    field = RN.RestrictedNbody_generator(potential=mw, progenitor_potential=ssc.potential.PlummerPotential, interp_prog=track, init_mass=2e4,
                                         init_rs=0.01, r_esc=0.05)
    # the field itself
    dy = field.term(-300.0, w0)
    g = orc.gradient(w0[:, :3], np.full(N, -300.0))
    assert relerr(dy[:, 3:], -g) < 1e-11 and np.array_equal(dy[:, :3], w0[:, 3:])
    # fixed step (dtmin = dtmax): 1e-10 relative
    for solver, sid in ((ssc.Dopri8(), 8), (ssc.Dopri5(), 5)):
        sol = ssc.integrate_field(w0=w0, ts=np.array([-600.0, -100.0]), solver=solver, field=field, dtmin=1.0, dtmax=1.0, max_steps=2000)
        yo, st, ns = O.shared_step_orbits(orc, w0, -600.0, -100.0, solver=sid, dtmin=1.0, dtmax=1.0, max_steps=2000)
        assert st == 0 and sol.ys.shape == (2, N, 6) and np.array_equal(sol.ys[0], w0)
        assert relerr(sol.ys[-1], yo[0]) < 1e-10
        assert int(sol.stats["num_steps"]) == ns[0]
    # adaptive, short span: both follow the same shared step sequence -> within 10 x tol and identical step counts
    sol = ssc.integrate_field(w0=w0, ts=np.array([-600.0, -550.0]), solver=ssc.Dopri8(), field=field, rtol=1e-8, atol=1e-8, dtmin=0.05, max_steps=5000)
    yo, st, ns = O.shared_step_orbits(orc, w0, -600.0, -550.0, solver=8, rtol=1e-8, atol=1e-8, dtmin=0.05, max_steps=5000)
    assert st == 0 and scaled_err(sol.ys[-1], yo[0], 1e-8).max() < 10.0
    assert [int(sol.stats[k]) for k in ("num_steps", "num_accepted_steps", "num_rejected_steps")] == list(ns)
    # adaptive, 600 Myr of bound tracer orbits: the RMS norm over 4200 components lets single tracers err by >> tol in BOTH
    # implementations (oracle vs a 1e-13 solution: ~5e4 x tol), so the criterion is DESIGN.md's: as accurate as the oracle
    sol = ssc.integrate_field(w0=w0, ts=np.array([-600.0, 0.0]), solver=ssc.Dopri8(), field=field, rtol=1e-8, atol=1e-8, dtmin=0.05, max_steps=5000)
    yo, st, ns = O.shared_step_orbits(orc, w0, -600.0, 0.0, solver=8, rtol=1e-8, atol=1e-8, dtmin=0.05, max_steps=5000)
    yt, _, _ = orc.integrate_orbits(w0, -600.0, 0.0, rtol=1e-13, atol=1e-13, dtmin=1e-4, max_steps=400_000, threads=8)
    # the two accept/reject sequences decorrelate after the first differing decision: compare the ACCEPTED step counts (5 %)
    assert st == 0 and abs(int(sol.stats["num_accepted_steps"]) - ns[1]) <= max(3, ns[1] // 20)
    e_gpu, e_orc = scaled_err(sol.ys[-1], yt[:, 0], 1e-8), scaled_err(yo[0], yt[:, 0], 1e-8)
    assert e_gpu.max() <= 1.5 * e_orc.max() + 10.0 and np.median(e_gpu) <= 1.5 * np.median(e_orc) + 10.0
    # backward + max_steps failure raises like diffrax throw=True
    with pytest.raises(RuntimeError):
        ssc.integrate_field(w0=w0, ts=np.array([-600.0, 0.0]), solver=ssc.Dopri8(), field=field, rtol=1e-10, atol=1e-10, dtmin=0.05, max_steps=5)
    solb = ssc.integrate_field(w0=sol.ys[-1], ts=np.array([0.0, -600.0]), t0=0.0, t1=-600.0, solver=ssc.Dopri8(), field=field, rtol=1e-10, atol=1e-10,
                               dtmin=0.01, max_steps=5000)
    assert np.abs(solb.ys[-1] - w0).max() < 2e-3          # reversibility at the solver's own accuracy (see above: ~5e4 x 1e-8)


def test_restricted_nbody_driver(cuda):
    """integrate_restricted_Nbody (RestrictedNbody.py:120-147): interrupt loop + monopole re-fit; the integration legs equal
    the oracle's shared-step solve with the fitted parameters."""
    import streamsculptor_b200 as ssc
    from streamsculptor_b200 import RestrictedNbody as RN
    orc0, mw, track, tk, yk = _restricted_pair()
    rng = np.random.default_rng(6)
    N = 256
    w0 = np.hstack([yk[0] + rng.normal(size=(N, 3)) * 0.01, _restricted_pair.v0 + rng.normal(size=(N, 3)) * 3e-4])
    field = RN.RestrictedNbody_generator(potential=mw, progenitor_potential=ssc.potential.PlummerPotential, interp_prog=track, init_mass=2e4,
                                         init_rs=0.01, r_esc=2.0)
    ts = np.array([-600.0, -300.0])
    states = RN.integrate_restricted_Nbody(w0=w0, ts=ts, interrupt_ts=np.array([-500.0, -400.0]), solver=ssc.Dopri8(), field=field, dtmin=0.5,
                                           dtmax=0.5, maxiter=3, max_steps=2000, mass_init=2e4, r_s_init=0.01)
    tstop, mass, rs, W = states
    assert np.array_equal(tstop, [-500.0, -400.0, -300.0]) and W.shape == (3, N, 6) and np.all(mass > 0) and np.all(rs > 0)
    # replay the legs with the oracle using the fitted parameters
    wc, tc = w0, -600.0
    for k in range(3):
        orc = mw3_oracle()
        tr = orc.track(O.CUBIC, tk, yk)
        orc.plummer(mass[k], rs[k], track=tr)
        yo, st, _ = O.shared_step_orbits(orc, wc, tc, tstop[k], solver=8, dtmin=0.5, dtmax=0.5, max_steps=2000)
        assert st == 0 and relerr(W[k], yo[0]) < 1e-10
        wc, tc = W[k], tstop[k]


def test_nbody_field(cuda):
    """A17: Nbody_field (fields.py:115-155) - softened all-pairs + external potential as ONE ODE, SaveAt(ts)."""
    import streamsculptor_b200 as ssc
    F = ssc.fields
    # 3-body in MW3 (tests.ipynb cells 7-11 style)
    m3 = np.array([1e9, 2e9, 3e9])
    w3 = np.array([[10., 0, 0, 0, .1, 0], [0, 10., 0, -.1, 0, 0.02], [-10., 0, 1, 0, -.1, 0]])
    f_ext = F.Nbody_field(ext_pot=mw3_product(), masses=m3, units=ssc.usys, eps=1e-3)
    f_iso = F.Nbody_field(ext_pot=None, masses=m3, units=ssc.usys, eps=0.05)
    assert relerr(f_ext.term(3.0, w3), O.nbody_term(mw3_oracle(), m3, 3.0, w3, eps=1e-3)) < 1e-12
    assert relerr(f_iso.term(0.0, w3), O.nbody_term(None, m3, 0.0, w3, eps=0.05)) < 1e-12
    ts = np.linspace(0.0, 400.0, 41)
    for solver, sid in ((ssc.Dopri8(), 8), (ssc.Dopri5(), 5)):
        sol = ssc.integrate_field(w0=w3, ts=ts, solver=solver, field=f_ext, dtmin=0.5, dtmax=0.5, max_steps=2000)
        yo, st, ns = O.nbody(mw3_oracle(), m3, w3, 0.0, 400.0, ts=ts, eps=1e-3, solver=sid, dtmin=0.5, dtmax=0.5, max_steps=2000)
        assert st == 0 and sol.ys.shape == (41, 3, 6)
        assert relerr(sol.ys, yo) < 1e-10
    sol = ssc.integrate_field(w0=w3, ts=ts, solver=ssc.Dopri8(), field=f_ext, rtol=1e-9, atol=1e-9, dtmin=0.05, max_steps=4000)
    yo, st, ns = O.nbody(mw3_oracle(), m3, w3, 0.0, 400.0, ts=ts, eps=1e-3, solver=8, rtol=1e-9, atol=1e-9, dtmin=0.05, max_steps=4000)
    assert st == 0 and scaled_err(sol.ys[-1], yo[-1], 1e-9).max() < 10.0 and abs(int(sol.stats["num_steps"]) - ns[0]) <= 2
    # 100 live perturbers (C5), isolated, backward in time, fixed step
    rng = np.random.default_rng(8)
    Nb = 100
    mb = 10 ** rng.uniform(7, 9, Nb)
    wb = np.hstack([rng.normal(size=(Nb, 3)) * 20.0, rng.normal(size=(Nb, 3)) * 0.05])
    fb = F.Nbody_field(ext_pot=mw3_product(), masses=mb, units=ssc.usys, eps=0.1)
    tsb = np.array([0.0, -50.0, -100.0])
    sol = ssc.integrate_field(w0=wb, ts=tsb, t0=0.0, t1=-100.0, solver=ssc.Dopri8(), field=fb, dtmin=0.25, dtmax=0.25, max_steps=1000)
    yo, st, _ = O.nbody(mw3_oracle(), mb, wb, 0.0, -100.0, ts=tsb, eps=0.1, solver=8, dtmin=0.25, dtmax=0.25, max_steps=1000)
    assert st == 0 and relerr(sol.ys, yo) < 1e-10


def _blockerr(A, B):
    """max over the (position | velocity)^rank unit blocks of max|A - B| / max|B| (entries of one block share units and scale)."""
    A, B = np.asarray(A), np.asarray(B)
    err = 0.0
    sl = (slice(0, 3), slice(3, 6))
    import itertools
    for idx in itertools.product(sl, repeat=A.ndim - 1):
        a, b = A[(slice(None),) + idx], B[(slice(None),) + idx]
        err = max(err, np.abs(a - b).max() / np.abs(b).max())
    return err


def test_variational_equations(cuda):
    """A16/N3: state-transition matrix and second-order tensor along unperturbed orbits (higher_order_variationalEqn.ipynb cell 3)
    vs the oracle (nested AD of the reference's scalar potentials), plus the physics check M == d w_f / d w_0 by finite differences."""
    import streamsculptor_b200 as ssc
    F = ssc.fields
    orc, prod = _full_pair(cuda)                 # every component type incl. tracks and subhalos (generic interpreter)
    mwo, mwp = mw3_oracle(), mw3_product()       # fused signature
    gala = ssc.potential.GalaMilkyWayPotential(units=ssc.usys)
    galo = O.Program().miyamoto(6.8e10, 3.0, 0.28).hernquist(5e9, 1.0).hernquist(1.71e9, 0.07).nfw(5.4e11, 15.62)
    rng = np.random.default_rng(11)
    # the field itself
    for order, n in ((1, 42), (2, 258)):
        y = np.concatenate([[8.0, -3.0, 5.0, 0.1, 0.15, -0.05], rng.normal(size=n - 6)])
        for po, pp in ((orc, prod), (mwo, mwp)):
            fld = F.variational_field(pp, order=order)
            coords = [y[:6], y[6:42].reshape(6, 6)] + ([y[42:].reshape(6, 6, 6)] if order == 2 else [])
            dy = np.concatenate([np.ravel(c) for c in fld.term(-700.0, coords)])
            assert relerr(dy, O.variational_term(po, -700.0, y, order=order)) < 1e-10
    w0 = halo_orbits(37, seed=4)
    t0 = rng.uniform(-1500.0, -300.0, 37)
    # fixed step, both solvers, both orders, fused and generic programs: 1e-10 relative
    for po, pp in ((mwo, mwp), (galo, gala), (orc, prod)):
        for solver, sid in ((ssc.Dopri8(), 8), (ssc.Dopri5(), 5)):
            for order in (1, 2):
                w, M, M2, st = F.integrate_variational_batch(pp, w0, t0, 0.0, order=order, solver=solver, dtmin=2.0, dtmax=2.0, max_steps=5000)
                wo, Mo, M2o, sto, nso = O.variational(po, w0, t0, 0.0, order=order, solver=sid, dtmin=2.0, dtmax=2.0, max_steps=5000, threads=8)
                assert not st.any() and not sto.any()
                assert relerr(w, wo) < 1e-10 and _blockerr(M, Mo) < 1e-10
                if order == 2:
                    assert _blockerr(M2, M2o) < 1e-10 and np.abs(M2 - M2.transpose(0, 1, 3, 2)).max() == 0.0
    # adaptive: shared controller over 42 / 258 components -> same accept/reject sequence on short spans, 10 x tol
    for order in (1, 2):
        w, M, M2, st = F.integrate_variational_batch(mwp, w0, -400.0, 0.0, order=order, rtol=1e-9, atol=1e-9, dtmin=0.05)
        wo, Mo, M2o, sto, nso = O.variational(mwo, w0, -400.0, 0.0, order=order, rtol=1e-9, atol=1e-9, dtmin=0.05, threads=8)
        assert not st.any()
        assert np.mean(scaled_err(M, Mo, 1e-9) < 10.0) >= 0.9 and np.mean(scaled_err(w, wo, 1e-9) < 10.0) >= 0.9
    # backward in time with non-trivial initial tensors, through integrate_field (single trajectory, tutorial usage)
    M0 = np.eye(6) + 0.01 * rng.normal(size=(6, 6))
    M20 = 0.01 * rng.normal(size=(6, 6, 6)); M20 = M20 + M20.transpose(0, 2, 1)
    sol = ssc.integrate_field(w0=[w0[0], M0, M20], ts=np.array([0.0, -600.0]), t0=0.0, t1=-600.0, field=F.variational_field(mwp, order=2),
                              solver=ssc.Dopri8(), dtmin=1.5, dtmax=1.5, max_steps=5000)
    wo, Mo, M2o, _, _ = O.variational(mwo, w0[:1], 0.0, -600.0, order=2, M0=M0[None], M20=M20[None], dtmin=1.5, dtmax=1.5, max_steps=5000)
    assert sol.ys[0].shape == (2, 6) and sol.ys[1].shape == (2, 6, 6) and sol.ys[2].shape == (2, 6, 6, 6)
    assert relerr(sol.ys[0][-1], wo[0]) < 1e-10 and _blockerr(sol.ys[1][-1:], Mo) < 1e-10 and _blockerr(sol.ys[2][-1:], M2o) < 1e-10
    # physics: M is the Jacobian of the flow map (forward-mode derivative of integrate_orbit w.r.t. w0)
    w, M, _, _ = F.integrate_variational_batch(mwp, w0[:4], -500.0, 0.0, order=1, rtol=1e-11, atol=1e-11, dtmin=0.01)
    eps = 1e-5
    for k in range(6):
        d = np.zeros(6); d[k] = eps
        yp = mwp.integrate_orbit_batch_vmapped(w0=w0[:4] + d, ts=np.array([-500.0, 0.0]), rtol=1e-12, atol=1e-12, dtmin=0.01, max_steps=100_000)
        ym = mwp.integrate_orbit_batch_vmapped(w0=w0[:4] - d, ts=np.array([-500.0, 0.0]), rtol=1e-12, atol=1e-12, dtmin=0.01, max_steps=100_000)
        fd = (np.asarray(yp.ys)[:, -1] - np.asarray(ym.ys)[:, -1]) / (2 * eps)
        assert np.abs(fd - M[:, :, k]).max() < 2e-5 * max(1.0, np.abs(M[:, :, k]).max())
    # failure semantics: max_steps -> +inf rows, status 1
    w, M, _, st = F.integrate_variational_batch(mwp, w0[:3], -500.0, 0.0, order=1, rtol=1e-10, atol=1e-10, dtmin=0.01, max_steps=4)
    assert (st == 1).all() and np.isinf(w).all() and np.isinf(M).all()


def test_progenitor_and_custom_subhalo_wrappers(cuda):
    """ProgenitorPotential (potential.py:140-153) and SubhaloLinePotential_Custom (potential.py:908-956) lower onto the same device
    components as their explicit forms."""
    import streamsculptor_b200 as ssc
    P = ssc.potential
    mw = mw3_product()
    sol = mw.integrate_orbit(w0=[20., 0, 20, 0, .15, 0], ts=np.array([0.0, -800.0]), t0=0.0, t1=-800.0, dense=True)
    prog = P.ProgenitorPotential(m=3e4, r_s=0.02, interp_func=sol, prog_pot=P.PlummerPotential, units=ssc.usys)
    x = np.array([[19.5, 1.0, 19.0], [5.0, -3.0, 2.0]])
    for t in (-10.0, -400.0):
        c = np.asarray(sol.evaluate(t))[:3]
        ref = P.PlummerPotential(m=3e4, r_s=0.02, units=ssc.usys)
        assert relerr(prog.gradient(x, t), ref.gradient(x - c, t)) < 1e-7          # cubic re-sampling of the dense track: 4097 knots
        assert relerr(prog.potential(x, t), ref.potential(x - c, t)) < 1e-7
    tot = P.Potential_Combine([mw, prog], units=ssc.usys)
    ys = tot.integrate_orbit(w0=[19.9, 0.1, 20.0, 0.0, 0.15, 0.0], ts=np.array([-800.0, 0.0])).ys
    assert np.isfinite(ys).all()
    sh = subhalo_set(6, seed=9, tw=300.0)
    one = P.HernquistPotential(m=2e7, r_s=0.4, units=ssc.usys)
    cust = P.SubhaloLinePotential_Custom(pot=one, subhalo_x0=sh["x0"], subhalo_v=sh["v"], subhalo_t0=sh["t0"], t_window=300.0, units=ssc.usys)
    expl = P.SubhaloLinePotentialCustom_fromFunc(func=P.HernquistPotential, m=np.full(6, 2e7), r_s=np.full(6, 0.4), subhalo_x0=sh["x0"], subhalo_v=sh["v"],
                                                 subhalo_t0=sh["t0"], t_window=300.0, units=ssc.usys)
    tq = float(sh["t0"][2]) + 20.0
    assert np.array_equal(cust.potential_per_SH(x[0], tq), expl.potential_per_SH(x[0], tq))
    assert np.array_equal(cust.gradient(x, np.full(2, tq)), expl.gradient(x, np.full(2, tq)))


def test_dense_stream_and_streakline(cuda):
    """gen_stream_vmapped_dense + eval_dense_stream (main.py:409-430, streamhelpers.py:23-53) and gen_streakline
    (streamhelpers.py:656-732) against the oracle."""
    import streamsculptor_b200 as ssc
    orc, prod = mw3_oracle(), mw3_product()
    prog_today = [20.0, 0.0, 20.0, 0.0, 0.15, 0.0]
    back, _, _ = orc.integrate_orbits(prog_today, 0.0, -2000.0)
    ts = np.linspace(-2000.0, 0.0, 201)
    nr = np.random.Generator(np.random.PCG64(2)).standard_normal((201, 4))
    kw = dict(ts=ts, prog_w0=back[0, 0], Msat=1e4, seed_num=583, normals=nr, dtmin=1.0, dtmax=1.0)          # fixed 1 Myr steps: exact comparison
    for solver, sid in ((ssc.Dopri8(), 8), (ssc.Dopri5(), 5)):
        ds = prod.gen_stream_vmapped_dense(solver=solver, rec_cap=2048, **kw)            # 2000 fixed steps for the oldest particle
        lead_f, trail_f = prod.gen_stream_vmapped(solver=solver, **kw)
        le, te = ssc.eval_dense_stream(0.0, ds)
        assert relerr(le.cpu().numpy(), lead_f) < 1e-12 and relerr(te.cpu().numpy(), trail_f) < 1e-12      # at ts[-1]: the final states
        # at an interior time: particles released before it are interpolated, the others are +inf
        t_eval = -733.3
        le, te = [a.cpu().numpy() for a in ssc.eval_dense_stream(t_eval, ds)]
        released = ts[:-1] <= t_eval
        assert np.isinf(le[~released]).all() and np.isfinite(le[released]).all()
        pl, pt, vl, vt, = orc.gen_stream_ics(ts, back[0, 0], 1e4, 583, solver=sid, normals=nr, dtmin=1.0, dtmax=1.0)[:4]
        w0l = np.hstack([pl, vl])[:-1][released]
        yo, sto, _ = orc.integrate_orbits(w0l, ts[:-1][released], 0.0, ts=np.array([t_eval, 0.0]), solver=sid, dtmin=1.0, dtmax=1.0)
        ok = ts[:-1][released] < t_eval
        assert relerr(le[released][ok], yo[ok, 0]) < 1e-10
        one = ssc.eval_dense_stream_id(time=np.array([-500.0, -100.0]), interp_func=ds, idx=10, lead=False)
        yo1, _, _ = orc.integrate_orbits(np.hstack([pt, vt])[10], ts[10], 0.0, ts=np.array([-500.0, -100.0]), solver=sid, dtmin=1.0, dtmax=1.0)
        assert relerr(one, yo1[0]) < 1e-10
    # too few record slots -> status 1 (like max_steps), never silent truncation
    ds_small = prod.gen_stream_vmapped_dense(solver=ssc.Dopri8(), rec_cap=16, **kw)
    assert int((ds_small.orbits.status == 1).sum()) > 0
    # ---- streakline ----
    Ns = 101
    Lc, vl_, Lf, vt_, tstrip = ssc.get_Streakline_ICs(prod, back[0, 0], 1e4, -2000.0, 0.0, Ns, solver=ssc.Dopri8(), rtol=1e-9, atol=1e-9)
    po, _, _ = orc.integrate_orbits(back[0, 0], -2000.0, 0.0, ts=tstrip, rtol=1e-9, atol=1e-9)
    x, v = po[0, :, :3], po[0, :, 3:]
    r = np.linalg.norm(x, axis=1); rh = x / r[:, None]
    H = orc.hessian(x, tstrip)
    om = np.linalg.norm(np.cross(x, v), axis=1) / r ** 2
    rt_ = (O.G_KPC_MYR_MSUN * 1e4 / (om ** 2 - np.einsum("ni,nij,nj->n", rh, H, rh))) ** (1 / 3)
    vh = v / np.linalg.norm(v, axis=1)[:, None]
    st_ = np.linalg.norm(np.cross(rh, vh), axis=1)
    Lc_o = x - rh * rt_[:, None]
    vl_o = (om * np.linalg.norm(Lc_o, axis=1) / st_)[:, None] * vh
    assert scaled_err(Lc, Lc_o, 1e-8).max() < 10.0 and scaled_err(vl_, vl_o, 1e-8).max() < 10.0
    lead, trail, tstrip2 = ssc.gen_streakline(prod, back[0, 0], 1e4, -2000.0, 0.0, Ns, solver=ssc.Dopri8(), rtol=1e-9, atol=1e-9)
    yo, _, _ = orc.integrate_orbits(np.hstack([Lc, vl_])[:-1], tstrip[:-1], 0.0, rtol=1e-9, atol=1e-9)
    assert lead.shape == (Ns, 6) and np.mean(scaled_err(lead[:-1], yo[:, 0], 1e-9) < 10.0) > 0.85
    assert np.isinf(lead[-1]).all()                                             # released at t1: zero-length solve, the SaveAt row stays +inf (diffrax)


def test_reference_printed_stream_through_the_product_api(cuda):
    """Golden OC (examples/OrphanChenab_mw_lmc_example.ipynb cells 3-6): the reference's own printed `stream_lead` of the Orphan-Chenab stream
    in the static GalaMilkyWayPotential, reproduced by the CUDA path through the public API with the notebook's calls - integrate_orbit back
    4 Gyr, gen_stream_vmapped(3001 stripping times, Msat=1e6, seed_num=9302, max_steps=1000), default Dopri5, rtol = atol = 1e-7.
    The oracle reproduces the 36 printed numbers to 4e-7 (tests/test_oracle_goldens.py); the CUDA path follows the same adaptive step
    sequences up to rounding-level decisions, hence the tolerance: 2e-2 kpc / 2e-4 kpc/Myr is two orders of magnitude
    below what a wrong PRNG draw, dispersion or potential parameter does to these rows (>= 1e-2 kpc/Myr-scale, kiloparsecs in position)."""
    import json
    import os
    import re
    import warnings
    import streamsculptor_b200 as ssc
    from common import orphan_chenab_prog_today
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "notebook_goldens.json")
    with open(path) as fh:
        oc = [g for g in json.load(fh) if g["id"] == "OC"][0]
    want = np.array([float(x) for x in re.findall(r"[-+]?\d+\.\d+e[-+]\d+", oc["output"])]).reshape(6, 6)
    mw = ssc.potential.GalaMilkyWayPotential(units=ssc.usys)
    ic = np.asarray(mw.integrate_orbit(w0=orphan_chenab_prog_today(), ts=np.array([0.0, -4000.0]), t0=0.0, t1=-4000.0).ys[-1])
    ts = np.hstack([np.linspace(-4000.0, -150.0, 3000), [0.0]])
    lead, trail = mw.gen_stream_vmapped(ts=ts, prog_w0=ic, Msat=1e6, seed_num=9302, max_steps=1000)
    lead = np.asarray(lead)
    assert lead.shape == (3000, 6) and np.isfinite(lead).all()
    got = lead[[0, 1, 2, -3, -2, -1]]
    dpos, dvel = np.abs(got[:, :3] - want[:, :3]).max(), np.abs(got[:, 3:] - want[:, 3:]).max()
    warnings.warn(f"CUDA stream vs the reference's printed stream: max |dx| = {dpos:.2e} kpc, max |dv| = {dvel:.2e} kpc/Myr")
    assert dpos < 5e-6 and dvel < 1e-8          # measured 3.2e-7 kpc / 9.0e-10 kpc/Myr (the printed digits); x10 margin


def test_reference_printed_batch_of_orbits_through_the_product_api(cuda):
    """Golden B1 (tests.ipynb cells 12, 15, 20, 22): the reference's printed final v_x of 1000 orbits, `integrate_orbit_batch_scan(w0=ics,
    ts=[0, 3000])` in NFW(1e12, 20) with the defaults (adaptive Dopri8, 1e-7), against the CUDA path through the same API.  The oracle
    reproduces all 1000 numbers to 5e-10 (tests/test_oracle_goldens.py).  Orbits whose rounding-level controller decisions coincide agree
    to the printed digits; the others differ by a fraction of the solver's own global error, 4e-8 here."""
    import json
    import os
    import re
    import warnings
    import streamsculptor_b200 as ssc
    from common import notebook_batch_ics
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "notebook_goldens.json")
    with open(path) as fh:
        b1 = [g for g in json.load(fh) if g["id"] == "B1"][0]
    want = np.array([float(x) for x in re.findall(r"[-+]?\d+\.\d+e[-+]\d+", b1["output"])])
    assert want.shape == (1000,)
    nfw = ssc.potential.NFWPotential(m=1e12, r_s=20.0, units=ssc.usys)
    sol = nfw.integrate_orbit_batch_vmapped(w0=notebook_batch_ics(), ts=np.full((1000, 1), 3000.0), t0=0.0, t1=3000.0)
    d = np.abs(np.asarray(sol.ys)[:, 0, 3] - want)
    warnings.warn(f"CUDA orbits vs the reference's printed values: {np.mean(d < 1e-9):.3f} within 1e-9, median {np.median(d):.1e}, max {d.max():.1e}")
    assert np.mean(d < 1e-9) >= 0.99 and d.max() < 5e-9         # measured: 1.000 within 1e-9, max 4.8e-10 (the 9 printed digits); x10 margin



def test_growing_potential_matches_oracle(cuda):
    """A10: GrowingPotential (potential.py:464-477) with a tabulated growth factor - field values against the oracle's autodiff of
    Phi * growth_func(t), fixed-step orbits to 1e-10, and the fused-galaxy recognition must skip growing components."""
    import streamsculptor_b200 as ssc
    P = ssc.potential
    tg = np.linspace(-3000.0, 0.0, 31)
    gf = 0.6 + 0.4 * (tg + 3000.0) / 3000.0 + 0.05 * np.sin(tg / 400.0)
    t, y = lmc_track()
    orc = O.Program().hernquist(5e9, 1.0).miyamoto(6.8e10, 3.0, 0.28).nfw(5.4e11, 15.62).growing(tg, gf)
    tr = orc.track(O.LINEAR, t, y)
    orc.plummer(1.5e11, 10.8, track=tr).growing(tg, gf, kind=O.CUBIC)
    prod = P.Potential_Combine([
        P.HernquistPotential(m=5e9, r_s=1.0, units=ssc.usys), P.MiyamotoNagaiDisk(m=6.8e10, a=3.0, b=0.28, units=ssc.usys),
        P.GrowingPotential(P.NFWPotential(m=5.4e11, r_s=15.62, units=ssc.usys), (tg, gf), units=ssc.usys),
        P.GrowingPotential(P.TimeDepTranslatingPotential(P.PlummerPotential(m=1.5e11, r_s=10.8, units=ssc.usys), ssc.LinearTrack(t, y), units=ssc.usys),
                           ssc.CubicTrack(tg, np.stack([gf, 0 * gf, 0 * gf], axis=1)), units=ssc.usys)], units=ssc.usys)
    rng = np.random.default_rng(5)
    xyz = rng.normal(size=(200, 3)) * np.array([15, 15, 8.0])
    tt = rng.uniform(-2990, -10, 200)
    assert relerr(prod.potential(xyz, tt), orc.potential(xyz, tt)) < 1e-12
    assert relerr(prod.gradient(xyz, tt), orc.gradient(xyz, tt)) < 1e-11
    assert relerr(prod.jacobian_force(xyz, tt), orc.hessian(xyz, tt)) < 1e-10
    assert relerr(prod.third_derivative(xyz, tt), orc.third(xyz, tt)) < 1e-9
    w0 = halo_orbits(64, seed=9)
    t0 = np.linspace(-2900, -100, 64)
    for solver in (5, 8):
        ys_o, _, ns_o = orc.integrate_orbits(w0, t0, 0.0, solver=solver, dtmin=1.0, dtmax=1.0, threads=8)
        sol = prod.integrate_orbit_batch_vmapped(w0=w0, ts=np.zeros((64, 1)), t0=t0, t1=0.0, solver=ssc.Dopri8() if solver == 8 else ssc.Dopri5(), dtmin=1.0, dtmax=1.0)
        assert np.array_equal(sol.stats["num_steps"], ns_o[:, 0]) and relerr(sol.ys[:, 0], ys_o[:, 0]) < 1e-10
    # the whole stream pipeline in a growing halo (fixed steps): release uses the grown Hessian as well
    ts = np.linspace(-2000.0, 0.0, 201)
    nr = np.random.Generator(np.random.PCG64(1)).standard_normal((201, 4))
    w_prog = orc.integrate_orbits([20.0, 0.0, 20.0, 0.0, 0.15, 0.0], 0.0, -2000.0, dtmin=1.0, dtmax=1.0)[0][0, 0]
    lo, to, _, _ = orc.gen_stream(ts, w_prog, 1e4, 583, solver=8, normals=nr, dtmin=1.0, dtmax=1.0)
    lf, tf = prod.gen_stream_vmapped(ts=ts, prog_w0=w_prog, Msat=1e4, seed_num=583, solver=ssc.Dopri8(), normals=nr, dtmin=1.0, dtmax=1.0)
    assert scaled_err(lf, lo, 1e-10).max() < 1.0 and scaled_err(tf, to, 1e-10).max() < 1.0


def test_perturber_set_matches_oracle(cuda):
    """N1 / BASELINE config 5: tracers in the field of many TABULATED MOVING perturbers.  The packed perturber set (ssb_perturbers) against
    the oracle's N translating components: field values, fixed-step orbits, the shared-step tracer solve (centres interpolated once per
    stage and CTA) and the tangent (variational) field; a Potential_Combine of the same moving spheres must lower to the same thing."""
    import streamsculptor_b200 as ssc
    from streamsculptor_b200 import _runtime as rt
    P = ssc.potential
    rng = np.random.default_rng(11)
    n, nk = 40, 200
    t = np.linspace(-1500.0, 0.0, nk)
    cen = (rng.normal(size=(1, n, 3)) * 25.0 + np.cumsum(rng.normal(size=(nk, n, 3)) * 0.6, axis=0))
    ms, rs = 10 ** rng.uniform(8, 10.5, n), rng.uniform(0.5, 4.0, n)
    orc = mw3_oracle()
    for i in range(n):
        orc.plummer(ms[i], rs[i], track=orc.track(O.LINEAR, t, cen[:, i]))
    mw = mw3_product()
    prod = P.Potential_Combine([mw, P.PerturberSetPotential(P.PlummerPotential, ms, rs, t, cen, units=ssc.usys)], units=ssc.usys)
    prod2 = P.Potential_Combine([mw] + [P.TimeDepTranslatingPotential(P.PlummerPotential(m=ms[i], r_s=rs[i], units=ssc.usys), ssc.LinearTrack(t, cen[:, i]),
                                                                      units=ssc.usys) for i in range(n)], units=ssc.usys)
    xyz = rng.normal(size=(300, 3)) * np.array([20, 20, 10.0])
    tt = rng.uniform(-1600, 50, 300)                  # incl. linear extrapolation beyond both ends of the table
    g_o = orc.gradient(xyz, tt)
    assert relerr(prod.gradient(xyz, tt), g_o) < 1e-11 and relerr(prod.potential(xyz, tt), orc.potential(xyz, tt)) < 1e-12
    assert relerr(prod.jacobian_force(xyz, tt), orc.hessian(xyz, tt)) < 1e-10
    assert relerr(prod.third_derivative(xyz[:40], tt[:40]), orc.third(xyz[:40], tt[:40])) < 1e-9
    assert np.array_equal(prod2.gradient(xyz, tt), prod.gradient(xyz, tt))          # the automatic packing gives the same program
    w0 = halo_orbits(64, seed=3)
    t0 = np.linspace(-1400, -100, 64)
    for solver in (5, 8):
        ys_o, _, ns_o = orc.integrate_orbits(w0, t0, 0.0, solver=solver, dtmin=1.0, dtmax=1.0, threads=8)
        sol = prod.integrate_orbit_batch_vmapped(w0=w0, ts=np.zeros((64, 1)), t0=t0, t1=0.0, solver=ssc.Dopri8() if solver == 8 else ssc.Dopri5(), dtmin=1.0, dtmax=1.0)
        assert np.array_equal(sol.stats["num_steps"], ns_o[:, 0]) and relerr(sol.ys[:, 0], ys_o[:, 0]) < 1e-10
    # shared-step tracers (RestrictedNbody.py:93-131): ONE ODE for all tracers, perturber centres frozen per stage
    w0s = halo_orbits(500, seed=4)
    ys_s, st_s, ns_s = O.shared_step_orbits(orc, w0s, -1000.0, 0.0, solver=8, rtol=1e-8, atol=1e-8, dtmin=2.0, dtmax=2.0, max_steps=2000)
    ctrl = rt.make_ctrl(ssc.Dopri8(), 1e-8, 1e-8, 2.0, 2.0, 2000)
    wout, st, ns = rt.shared_step_orbits(prod, rt.to_dev(w0s), -1000.0, 0.0, ctrl)
    assert int(st[0]) == 0 and int(ns[0]) == int(np.asarray(ns_s).reshape(-1)[0])
    assert relerr(wout.cpu().numpy(), np.asarray(ys_s).reshape(-1, 500, 6)[-1]) < 1e-10
    # tangent field (C5: "auxiliary / tangent ODEs"): state-transition matrices along orbits in the same potential, fixed steps
    w_v, M_v, _, st_v, _ = rt.variational(prod, 1, rt.to_dev(w0[:16]), None, None, rt.to_dev(t0[:16]), 0.0, rt.make_ctrl(ssc.Dopri8(), 1e-8, 1e-8, 2.0, 2.0, 5000))
    w_o, M_o, _, _, _ = O.variational(orc, w0[:16], t0[:16], 0.0, order=1, solver=8, rtol=1e-8, atol=1e-8, dtmin=2.0, dtmax=2.0, max_steps=5000)
    assert int((st_v != 0).sum()) == 0 and relerr(w_v.cpu().numpy(), w_o) < 1e-10
    assert np.abs(M_v.cpu().numpy() - M_o).max() <= 1e-9 * np.abs(M_o).max()


def test_custom_base_second_order_generator(cuda):
    """GenerateMassRadiusPerturbation_CustomBase_SecondOrder (perturbative.py:491-585): the progenitor's backward-integrated first- and
    second-order field at every stripping time (here: one backward solve per stripping time, per-particle end times of the second-order
    kernel) mapped into the particles' perturbation ICs, then the second-order responses.  Fixed steps: against the oracle's second-order
    solves started and ended at the same times."""
    import streamsculptor_b200 as ssc
    P, pt = ssc.potential, ssc.perturbative
    nsh, nts = 5, 13
    sh = subhalo_set(nsh, seed=29, t_lo=-500.0)
    sh["x0"] = sh["x0"] * 0.3 + np.array([12.0, 3.0, -6.0])
    sh["M"] = np.ones(nsh)
    base, orc_base = mw3_product(), mw3_oracle()
    ts = np.linspace(-600.0, 0.0, nts)
    prog_w0 = [12.0, 3.0, -6.0, -0.05, 0.15, 0.03]
    rng = np.random.default_rng(5)
    pos_rel, vel_rel = rng.normal(size=(nts, 3)) * 0.05, rng.normal(size=(nts, 3)) * 1e-3
    pert = P.SubhaloLinePotentialCustom_fromFunc(func=P.HernquistPotential, m=sh["m"], r_s=sh["rs"], subhalo_x0=sh["x0"], subhalo_v=sh["v"],
                                                 subhalo_t0=sh["t0"], t_window=150.0, units=ssc.usys)
    orc_sh = O.Program().subhalos(O.PR_HERNQUIST, sh["m"], sh["rs"], sh["x0"], sh["v"], sh["t0"], 150.0)
    fixed = dict(rtol=1e-8, atol=1e-8, dtmin=2.0, dtmax=2.0)
    model = pt.CustomBaseStreamModel(potential_base=base, prog_w0=prog_w0, ts=ts, pos_rel=pos_rel, vel_rel=vel_rel, solver=ssc.Dopri8(), units=ssc.usys, **fixed)
    gen = pt.GenerateMassRadiusPerturbation_CustomBase_SecondOrder(potential_base=base, potential_perturbation=pert, BaseStreamModel=model, units=ssc.usys,
                                                                   solver=ssc.Dopri8(), max_steps=5000, **fixed)
    assert gen.perturbation_ICs[0].shape == (nts, nsh, 12) and gen.perturbation_ICs[1].shape == (nts, nsh, 6)
    prog_o, _, _ = orc_base.integrate_orbits(prog_w0, ts[0], ts[-1], ts=ts, solver=8, **fixed)
    w_obs = prog_o[0, -1]
    for i in (0, 3, 7, nts - 2):                      # the progenitor's field at a few stripping times: oracle backward solves ending there
        _, D_o, E_o, st_o, _ = O.second_order_response(orc_base, orc_sh, w_obs, ts[-1], ts[i], solver=8, max_steps=5000, **fixed)
        assert st_o[0] == 0
        assert np.abs(gen.prog_fieldICs_first_order_mass[i] - D_o[0]).max() <= 1e-9 * np.abs(D_o).max() + 1e-300
        assert np.abs(gen.prog_fieldICs_second_order_mass[i] - E_o[0]).max() <= 1e-8 * np.abs(E_o).max() + 1e-300
    assert not gen.prog_fieldICs_first_order_mass[-1].any()           # observation time: the field starts from zero
    w, D, E = gen.compute_perturbation_OTF(cpu=False, solver=ssc.Dopri8(), **fixed)
    ics = np.hstack([prog_o[0][:, :3] + pos_rel, prog_o[0][:, 3:] + vel_rel])
    w_o, D_o, E_o, _, _ = O.second_order_response(orc_base, orc_sh, ics[:-1], ts[:-1], 0.0, D0=gen.perturbation_ICs[0][:-1], E0=gen.perturbation_ICs[1][:-1],
                                                  solver=8, **fixed)
    assert scaled_err(w, w_o, 1e-9).max() < 1.0
    assert np.abs(D - D_o).max() <= 1e-8 * np.abs(D_o).max() and np.abs(E - E_o).max() <= 1e-7 * np.abs(E_o).max()


def _small_driver_setup():
    import streamsculptor_b200 as ssc
    pot = mw3_product()
    prog_today = np.array([12.0, 3.0, -6.0, -0.05, 0.15, 0.03])

    def phi1(stream):                                   # stand-in for the observer-frame transform the user supplies (degrees along the orbit plane)
        stream = np.asarray(stream)
        return np.degrees(np.arctan2(stream[:, 1], stream[:, 0]))
    return ssc, pot, prog_today, phi1


@pytest.mark.gpu
def test_impact_generator(cuda):
    """ImpactGenerator (GenerateImpactParams.py:9-215): window means against masked passes over the stream, the sampled parameters inside
    their bounds, and the impact geometry of get_subhalo_ImpactParams (GenerateImpactParams.py:186-211): the subhalo sits at distance b from
    the patch, perpendicular to the patch's velocity; its velocity along the stream is the sampled parallel component."""
    ssc, pot, prog_today, phi1 = _small_driver_setup()
    from streamsculptor_b200 import GenerateImpactParams as G
    t_age, n_arm = 1500.0, 400
    IC = np.asarray(pot.integrate_orbit(w0=prog_today, t0=0.0, t1=-t_age, ts=np.array([-t_age])).ys[0])
    ts = np.hstack([np.linspace(-t_age, -1.0, n_arm), [0.0]])
    l, t = ssc.gen_stream_vmapped_Chen25(pot_base=pot, prog_w0=IC, ts=ts, key=3, Msat=3e4, atol=1e-7, rtol=1e-7, solver=ssc.Dopri8())
    stream = np.vstack([np.asarray(l), np.asarray(t)])
    ph = phi1(stream)
    strip = np.hstack([ts[:-1], ts[:-1]])
    lo, hi = np.percentile(ph, [10, 90])
    n = 64
    rs = np.linspace(0.1, 1.0, n)
    gen = ssc.ImpactGenerator(pot=pot, tobs=0.0, stream=stream, stream_phi1=ph, phi1_bounds=[lo, hi], tImpactBounds=[-t_age, 0.0], phi1window=1.0,
                              NumImpacts=n, bImpact_bounds=[0, 10.0 * rs], stripping_times=strip, phi1_exclude=[-0.5, 0.5], prog_today=prog_today, seednum=77)
    q = np.linspace(lo, hi, 9)
    m, tm = gen.get_particle_mean(q)
    for i, c in enumerate(q):
        sel = np.abs(ph - c) < 1.0
        assert sel.sum() > 0
        assert np.abs(m[i].cpu().numpy() - stream[sel].mean(axis=0)).max() < 1e-11 and abs(float(tm[i]) - strip[sel].mean()) < 1e-9
    out = gen.get_subhalo_ImpactParams()
    par, cart, patch = out["ImpactFrameParams"], out["CartesianImpactParams"], out["StreamPatch"]
    assert not out["status"].any() and np.isfinite(cart).all()
    assert (par["bImpact"] >= 0).all() and (par["bImpact"] <= 10.0 * rs).all()
    assert (par["tImpact"] >= -t_age).all() and (par["tImpact"] <= 0.0).all()
    assert ((par["phi1_samples"] >= lo) & (par["phi1_samples"] <= hi)).all() and not ((par["phi1_samples"] > -0.5) & (par["phi1_samples"] < 0.5)).any()
    d = cart[:, :3] - patch[:, :3]
    T = patch[:, 3:] / np.linalg.norm(patch[:, 3:], axis=1)[:, None]
    assert np.abs(np.linalg.norm(d, axis=1) - par["bImpact"]).max() < 1e-12
    assert np.abs((d * T).sum(1)).max() < 1e-12
    assert np.abs((cart[:, 3:] * T).sum(1) - G.jax_normal(gen.keys[0], n) * gen.sigma).max() < 1e-12
    # the patches really are the window means taken back to the impact times: the oracle's orbit from the same mean
    m0, _ = gen.get_particle_mean(par["phi1_samples"][:4])
    for i in range(4):
        y, st, _ = mw3_oracle().integrate_orbits(m0[i].cpu().numpy(), 0.0, par["tImpact"][i], solver=8, rtol=1e-7, atol=1e-7, dtmin=0.1)
        assert st[0] == 0 and scaled_err(patch[i], y[0], 1e-5).max() < 1.0
    # same seed, same draw
    gen2 = ssc.ImpactGenerator(pot=pot, tobs=0.0, stream=stream, stream_phi1=ph, phi1_bounds=[lo, hi], tImpactBounds=[-t_age, 0.0], phi1window=1.0,
                               NumImpacts=n, bImpact_bounds=[0, 10.0 * rs], stripping_times=strip, phi1_exclude=[-0.5, 0.5], prog_today=prog_today, seednum=77)
    assert np.array_equal(gen2.get_subhalo_ImpactParams()["CartesianImpactParams"], cart)


@pytest.mark.gpu
def test_production_driver_get_derivs(cuda, tmp_path):
    """get_derivs (generate_derivs.py:24-216) at a small size: the files it writes hold what the reference's hold, the derivatives equal a
    direct response solve over the same sampled subhaloes, and the device summary equals its numpy restatement."""
    ssc, pot, prog_today, phi1 = _small_driver_setup()
    from streamsculptor_b200.generate_derivs import get_derivs
    # where the small stream lies in phi1 (the user of the reference knows it from the data): bounds well inside it, windows of a few degrees
    IC = np.asarray(pot.integrate_orbit(w0=prog_today, t0=0.0, t1=-1500.0, ts=np.array([-1500.0])).ys[0])
    ts = np.hstack([np.linspace(-1500.0, -1.0, 150), [0.0]])
    l, t = ssc.gen_stream_vmapped_Chen25(pot_base=pot, prog_w0=IC, ts=ts, key=3, Msat=3e4, atol=1e-7, rtol=1e-7, solver=ssc.Dopri8(),
                                         prog_pot=ssc.potential.PlummerPotential(m=3e4, r_s=0.004, units=ssc.usys))
    ph_all = phi1(np.vstack([np.asarray(l), np.asarray(t)]))
    ph_prog = float(phi1(prog_today[None])[0])
    lo, hi = np.percentile(ph_all, [15, 85])
    assert lo < ph_prog - 0.5 and hi > ph_prog + 0.5
    window = 0.25 * (hi - lo)
    kw = dict(prog_wtoday=prog_today, t_age=1500.0, t_dissolve=-1.0, log10_min_mass=5.0, log10_max_mass=8.0, phi1_bounds=[lo, hi],
              phi1_exclude=[ph_prog - 0.5, ph_prog + 0.5], stream_seednum=3, key=21, Msat=3e4, r_s=0.004, target_num=24, phi1_function=phi1, pot=pot, N_batch=12,
              atol=1e-9, rtol=1e-9, phi1window=window, N_arm=150)
    get_derivs(path=str(tmp_path), save_iter_start=5, **kw)
    files = sorted(p.name for p in tmp_path.iterdir())
    assert files == ["5.npy", "6.npy"]
    rec = np.load(tmp_path / "5.npy", allow_pickle=True).item()
    assert set(rec) == {"pert_out", "r_s_root", "ImpactFrameParams"}
    w, D = np.asarray(rec["pert_out"][0]), np.asarray(rec["pert_out"][1])
    assert w.shape == (301, 6) and D.shape == (301, 12, 12)                 # lead and trail interleaved by release time
    # the last kept particle is released AT the final time: nothing to integrate, diffrax's SaveAt(ts) leaves its rows +inf (perturbative.py:733-749)
    assert np.isfinite(D[:-1]).all() and np.isinf(D[-1]).all() and np.abs(D[:-1]).max() > 0
    edges = np.linspace(ph_all.min() - 1e-9, ph_all.max() + 1e-9, 13)
    out = get_derivs(path=None, save=False, summaries=edges, **kw)
    assert len(out) == 2
    w2, D2 = out[0]["pert_out"]
    assert np.array_equal(out[0]["ImpactFrameParams"]["tImpact"], rec["ImpactFrameParams"]["tImpact"])          # same keys -> same draws
    assert np.array_equal(np.asarray(w2).reshape(w.shape), w) and np.array_equal(np.asarray(D2).reshape(D.shape), D)
    fin = np.isfinite(np.asarray(w2)).all(axis=1)
    ph = phi1(np.asarray(w2)[fin])
    disp = np.einsum("s,nsk->nk", out[0]["masses"], np.asarray(D2)[fin][:, :, :6])
    idx = np.searchsorted(edges, ph, side="left") - 1
    assert fin.sum() == 300
    for b in range(12):
        sel = idx == b
        want = disp[sel].mean(axis=0) if sel.any() else np.zeros(6)
        assert np.abs(out[0]["binned"][b] - want).max() <= 1e-12 * max(np.abs(disp).max(), 1e-300)
    assert not np.array_equal(out[1]["ImpactFrameParams"]["tImpact"], out[0]["ImpactFrameParams"]["tImpact"])   # a fresh key per batch
    # batches in flight on their own CUDA streams (the default) against one batch at a time: the same results, batch by batch
    kw4 = dict(kw, target_num=48)
    ser = get_derivs(path=None, save=False, pipeline=1, **kw4)
    par = get_derivs(path=None, save=False, pipeline=3, **kw4)
    assert len(ser) == len(par) == 4
    for a, b in zip(ser, par):
        assert np.array_equal(a["pert_out"][0], b["pert_out"][0]) and np.array_equal(a["pert_out"][1], b["pert_out"][1])
        assert np.array_equal(a["ImpactFrameParams"]["bImpact"], b["ImpactFrameParams"]["bImpact"])


@pytest.mark.gpu
def test_rotating_bars_match_oracle(cuda):
    """N4: BarPotential (potential.py:178-198) and DehnenBarPotential (potential.py:200-222).  The device evaluates their gradient, Hessian and
    third derivatives with Taylor jets through the scalar formula (csrc/ssb_jet.cuh); the oracle differentiates its own restatement with
    nested dual numbers.  Field values at random points and times, fixed-step orbits (1e-10), a translating growing bar, and a stream whose
    release needs the bar's Hessian."""
    import streamsculptor_b200 as ssc
    P = ssc.potential
    bar_kw = dict(m=1e10, a=3.5, b=0.5, c=0.6, Omega=0.04)
    deh_kw = dict(alpha=0.01, v0=0.22, R0=8.0, Rb=3.4, phib=0.4, Omega=0.05)
    mw, omw = mw3_product(), mw3_oracle
    for name, prod, orc in (
            ("bar", P.Potential_Combine([mw, P.BarPotential(units=ssc.usys, **bar_kw)], units=ssc.usys),
             omw().bar(bar_kw["m"], bar_kw["a"], bar_kw["b"], bar_kw["c"], bar_kw["Omega"])),
            ("dehnen", P.Potential_Combine([mw, P.DehnenBarPotential(units=ssc.usys, **deh_kw)], units=ssc.usys),
             omw().dehnen_bar(deh_kw["alpha"], deh_kw["v0"], deh_kw["R0"], deh_kw["Rb"], deh_kw["phib"], deh_kw["Omega"]))):
        rng = np.random.default_rng(11)
        xyz = rng.normal(size=(300, 3)) * np.array([6.0, 6.0, 2.0])            # inside and outside Rb, near and far from the bar's ends
        t = rng.uniform(-3000.0, 0.0, 300)
        assert relerr(prod.potential(xyz, t), orc.potential(xyz, t)) < 1e-12, name
        assert relerr(prod.gradient(xyz, t), orc.gradient(xyz, t)) < 1e-11, name
        assert relerr(prod.jacobian_force(xyz, t), orc.hessian(xyz, t)) < 1e-10, name
        assert relerr(prod.third_derivative(xyz[:64], t[:64]), orc.third(xyz[:64], t[:64])) < 1e-9, name
        # the bar alone, so that a wrong bar cannot hide behind the galaxy's larger numbers
        alone = prod.potential_list[1]
        o_alone = (O.Program().bar(**{k: bar_kw[k] for k in ("m", "a", "b", "c", "Omega")}) if name == "bar"
                   else O.Program().dehnen_bar(**deh_kw))
        assert relerr(alone.gradient(xyz, t), o_alone.gradient(xyz, t)) < 1e-10, name
        assert relerr(alone.jacobian_force(xyz, t), o_alone.hessian(xyz, t)) < 1e-9, name
        assert relerr(alone.third_derivative(xyz[:64], t[:64]), o_alone.third(xyz[:64], t[:64])) < 1e-8, name
        # fixed-step orbits through the rotating field
        w0 = halo_orbits(40, seed=21) * np.array([0.4, 0.4, 0.2, 1.0, 1.0, 0.5])
        for solver in (5, 8):
            sol = prod.integrate_orbit_batch_vmapped(w0=w0, ts=np.array([-400.0, -200.0, 0.0]), solver=ssc.Dopri8() if solver == 8 else ssc.Dopri5(),
                                                     rtol=1e-8, atol=1e-8, dtmin=1.0, dtmax=1.0)
            yo, st, _ = orc.integrate_orbits(w0, -400.0, 0.0, ts=np.array([-400.0, -200.0, 0.0]), solver=solver, rtol=1e-8, atol=1e-8, dtmin=1.0, dtmax=1.0)
            assert not st.any() and relerr(np.asarray(sol.ys), yo) < 1e-10, (name, solver)
    # a stream in MW3 + bar: release (Hessian of the bar) + orbits, fixed steps, against the oracle
    prod = P.Potential_Combine([mw, P.BarPotential(units=ssc.usys, **bar_kw)], units=ssc.usys)
    orc = omw().bar(bar_kw["m"], bar_kw["a"], bar_kw["b"], bar_kw["c"], bar_kw["Omega"])
    ts = np.linspace(-600.0, 0.0, 41)
    w0 = [8.0, 0.5, 3.0, -0.02, 0.2, 0.05]
    nr = np.random.Generator(np.random.PCG64(3)).standard_normal((41, 4))
    lead, trail = prod.gen_stream_vmapped(ts=ts, prog_w0=w0, Msat=1e4, seed_num=5, solver=ssc.Dopri8(), normals=nr, rtol=1e-8, atol=1e-8, dtmin=2.0, dtmax=2.0)
    lo, to, st, _ = orc.gen_stream(ts, w0, 1e4, 5, solver=8, normals=nr, rtol=1e-8, atol=1e-8, dtmin=2.0, dtmax=2.0)
    assert relerr(lead, lo) < 1e-9 and relerr(trail, to) < 1e-9


@pytest.mark.gpu
def test_lmc_potential_and_not_a_knot_track(cuda):
    """LMCPotential (potential.py:40-63): NFW on a cubic spline through the tabulated LMC orbit.  The track evaluates scipy's not-a-knot
    spline (value and derivative); field values and fixed-step orbits against the oracle built with the same Hermite data."""
    from scipy.interpolate import CubicSpline
    import streamsculptor_b200 as ssc
    P = ssc.potential
    t, y = lmc_track(n=41)                                             # 41 knots over 3 Gyr: the spline matters between them
    cs = CubicSpline(t, y, axis=0, bc_type="not-a-knot")
    tq = np.random.default_rng(1).uniform(t[0], t[-1], 300)
    trk = P.NotAKnotTrack(t, y)
    assert np.abs(trk(tq) - cs(tq)).max() <= 1e-12 * np.abs(y).max()
    assert np.abs(trk(tq, derivative=True) - cs(tq, 1)).max() <= 1e-11 * np.abs(cs(tq, 1)).max()
    lmc = P.LMCPotential({"m_NFW": 1.5e11, "r_s_NFW": 10.0}, {"t": t, "x": y[:, 0], "y": y[:, 1], "z": y[:, 2]}, units=ssc.usys)
    prod = P.Potential_Combine([mw3_product(), lmc], units=ssc.usys)
    orc = mw3_oracle()
    orc.nfw(1.5e11, 10.0, track=orc.track(O.CUBIC, t, y, slopes=cs(t, 1)))
    rng = np.random.default_rng(6)
    xyz = rng.normal(size=(200, 3)) * np.array([30.0, 30.0, 30.0])
    tt = rng.uniform(t[0], t[-1], 200)
    assert relerr(prod.potential(xyz, tt), orc.potential(xyz, tt)) < 1e-12
    assert relerr(prod.gradient(xyz, tt), orc.gradient(xyz, tt)) < 1e-11
    assert relerr(prod.jacobian_force(xyz, tt), orc.hessian(xyz, tt)) < 1e-10
    w0 = halo_orbits(30, seed=4)
    sol = prod.integrate_orbit_batch_vmapped(w0=w0, ts=np.array([-500.0, 0.0]), solver=ssc.Dopri8(), rtol=1e-8, atol=1e-8, dtmin=1.0, dtmax=1.0)
    yo, st, _ = orc.integrate_orbits(w0, -500.0, 0.0, ts=np.array([-500.0, 0.0]), solver=8, rtol=1e-8, atol=1e-8, dtmin=1.0, dtmax=1.0)
    assert not st.any() and relerr(np.asarray(sol.ys), yo) < 1e-10


@pytest.mark.gpu
def test_response_with_moving_progenitor_base_potential(cuda):
    """The production driver's base potential (perturbative.py:642-644): galaxy + the progenitor's Plummer sphere on a cubic track, plus a second
    moving sphere on a linear track (interpreter tail behind the fused galaxy in the serial base-orbit phase).  The multi-slot kernel against
    the one-particle kernel bit for bit, and fixed steps against the oracle.  (The progenitor is softer than the production one - r_s 0.05
    instead of 0.004 kpc - so that a 2 Myr fixed step resolves its core: inside a 0.004 kpc core the dynamical time is 0.7 Myr and a fixed-step
    comparison only measures how rounding differences are amplified.)"""
    import os
    import streamsculptor_b200 as ssc
    from streamsculptor_b200 import _runtime as rt
    P = ssc.potential
    mw = mw3_product()
    tk = np.linspace(-1000.0, 0.0, 201)
    yk = np.asarray(mw.integrate_orbit(w0=[12.0, 3.0, -6.0, -0.05, 0.15, 0.03], ts=tk[::-1].copy(), t0=0.0, t1=-1000.0).ys)[::-1].copy()
    tl, yl = lmc_track(n=60, t_lo=-1000.0)
    base = P.Potential_Combine([mw, P.TimeDepTranslatingPotential(P.PlummerPotential(m=3e4, r_s=0.05, units=ssc.usys), ssc.CubicTrack(tk, yk[:, :3].copy()), units=ssc.usys),
                                P.TimeDepTranslatingPotential(P.HernquistPotential(m=1e10, r_s=8.0, units=ssc.usys), ssc.LinearTrack(tl, yl), units=ssc.usys)],
                               units=ssc.usys)
    orc = mw3_oracle()
    orc.plummer(3e4, 0.05, track=orc.track(O.CUBIC, tk, yk[:, :3]))
    orc.hernquist(1e10, 8.0, track=orc.track(O.LINEAR, tl, yl))
    nsh = 24
    sh = subhalo_set(nsh, seed=9, t_lo=-900.0, tw=150.0)
    pert = P.SubhaloLinePotentialCustom_fromFunc(func=P.HernquistPotential, m=sh["m"], r_s=sh["rs"], subhalo_x0=sh["x0"], subhalo_v=sh["v"],
                                                 subhalo_t0=sh["t0"], t_window=sh["tw"], units=ssc.usys)
    orc_sh = O.Program().subhalos(O.PR_HERNQUIST, sh["m"], sh["rs"], sh["x0"], sh["v"], sh["t0"], sh["tw"])
    N = 45
    rng = np.random.default_rng(3)
    t0 = np.linspace(-950.0, -20.0, N)
    prog = np.stack([np.interp(t0, tk, yk[:, k]) for k in range(6)], axis=1)
    w0 = prog + np.hstack([rng.normal(size=(N, 3)) * 0.05, rng.normal(size=(N, 3)) * 1e-3])
    try:
        for solver, tol, fixed in ((ssc.Dopri8(), 1e-8, None), (ssc.Dopri5(), 1e-7, None), (ssc.Dopri8(), 1e-8, 2.0)):
            ctrl = rt.make_ctrl(solver, tol, tol, fixed or 0.01, fixed, 10_000)
            out = {}
            for retire in (0, 1):
                os.environ["SSB_RESP_RETIRE"] = str(retire)
                for np_slots in (0, 4, 16):
                    os.environ["SSB_RESP_NP"] = str(np_slots)
                    w, D, st, ns = rt.linear_response(base, pert._arrays, rt.to_dev(w0), None, rt.to_dev(t0), 0.0, ctrl)
                    out[retire, np_slots] = (w.cpu().numpy(), D.cpu().numpy(), st.cpu().numpy(), ns.cpu().numpy())
            assert not out[0, 0][2].any()
            for np_slots in (4, 16):
                for a, b in zip(out[0, np_slots], out[0, 0]):
                    assert np.array_equal(a, b), f"{np_slots} slots differ from the one-particle kernel with moving components in the base potential"
                for a, b in zip(out[1, np_slots], out[1, 4]):
                    assert np.array_equal(a, b)
            if fixed:
                w_o, D_o, st_o, _ = O.linear_response(orc, orc_sh, w0, t0, 0.0, solver=8, rtol=tol, atol=tol, dtmin=fixed, dtmax=fixed)
                assert not st_o.any() and scaled_err(out[0, 16][0], w_o, 1e-10).max() < 1.0
                assert np.abs(out[0, 16][1] - D_o).max() <= 1e-9 * np.abs(D_o).max()
    finally:
        os.environ.pop("SSB_RESP_NP", None)
        os.environ.pop("SSB_RESP_RETIRE", None)


@pytest.mark.gpu
def test_saving_kernel_ragged_save_lists(cuda):
    """Edge cases of the warp-cooperative dense output (K1 MODE 0): a batch size that is no multiple of a warp, save lists with MANY times
    inside one step (gallop + bisection), duplicate save times, save times equal to t0 and to t1, one save time only, orbits that hit max_steps
    half-way (later rows stay +inf, earlier rows are written) and zero-length orbits - fixed steps, against the oracle to 1e-10, and row by
    row the same +inf pattern."""
    import streamsculptor_b200 as ssc
    orc, prod = mw3_oracle(), mw3_product()
    N, M = 77, 41
    rng = np.random.default_rng(8)
    w0 = halo_orbits(N, seed=6)
    t0 = rng.uniform(-900.0, -600.0, N)
    t1 = rng.uniform(-100.0, 0.0, N)
    ts = np.sort(rng.uniform(0.0, 1.0, (N, M)), axis=1) * (t1 - t0)[:, None] + t0[:, None]
    ts[:, 0] = t0; ts[:, -1] = t1                                   # first row = y0 exactly, last row = the final state
    ts[::5, 10:30] = ts[::5, 10:11] + np.linspace(0.0, 3.0, 20)     # twenty save times inside two 2-Myr steps
    ts[::7, 5] = ts[::7, 4]                                         # duplicates
    ts = np.sort(ts, axis=1)
    t1z = t1.copy(); t1z[3] = t0[3]                                 # a zero-length orbit: nothing is saved
    for solver in (5, 8):
        sv = ssc.Dopri8() if solver == 8 else ssc.Dopri5()
        for max_steps in (10_000, 150):                             # 150 fixed steps of 2 Myr end before t1 for every orbit
            ys_o, st_o, ns_o = orc.integrate_orbits(w0, t0, t1z, ts=ts, solver=solver, dtmin=2.0, dtmax=2.0, max_steps=max_steps, threads=8)
            sol = prod.integrate_orbit_batch_vmapped(w0=w0, ts=ts, t0=t0, t1=t1z, solver=sv, dtmin=2.0, dtmax=2.0, max_steps=max_steps)
            ys = np.asarray(sol.ys)
            assert np.array_equal(np.asarray(sol.result), st_o) and np.array_equal(sol.stats["num_steps"], ns_o[:, 0])
            assert np.array_equal(np.isinf(ys), np.isinf(ys_o))
            fin = np.isfinite(ys_o)
            assert np.abs(ys[fin] - ys_o[fin]).max() <= 1e-10 * (1.0 + np.abs(ys_o[fin]).max())
            assert np.isinf(ys[3]).all()
            if max_steps == 150:
                assert (st_o[np.arange(N) != 3] == 1).all()
                assert np.isinf(ys[0, -1]).all() and np.isfinite(ys[0, 0]).all()        # cut short: the end is missing, the start is there
    # one save time only, per orbit, somewhere inside the span
    ts1 = (0.5 * (t0 + t1))[:, None]
    ys_o, _, _ = orc.integrate_orbits(w0, t0, t1, ts=ts1, solver=8, dtmin=2.0, dtmax=2.0, threads=8)
    sol = prod.integrate_orbit_batch_vmapped(w0=w0, ts=ts1, t0=t0, t1=t1, solver=ssc.Dopri8(), dtmin=2.0, dtmax=2.0)
    assert scaled_err(np.asarray(sol.ys), ys_o, 1e-10).max() < 1.0


@pytest.mark.gpu
def test_response_edge_sizes(cuda):
    """Sizes at the edges of the response kernel's bookkeeping: one subhalo, counts around the retirement chunk of 16 positions (15, 16, 17,
    33), a set larger than the in-kernel sort (4100 > 4096: identity order, nothing skipped or retired), one particle, and a particle count that
    leaves slots empty - fixed steps, against the oracle, retirement on and off."""
    import os
    import streamsculptor_b200 as ssc
    from streamsculptor_b200 import _runtime as rt
    P = ssc.potential
    base, orc_base = mw3_product(), mw3_oracle()
    ctrl = rt.make_ctrl(ssc.Dopri8(), 1e-8, 1e-8, 2.0, 2.0, 10_000)
    try:
        for nsh, N in ((1, 5), (15, 19), (16, 3), (17, 1), (33, 40), (4100, 2)):
            sh = subhalo_set(nsh, seed=nsh, t_lo=-700.0, tw=40.0)
            pert = P.SubhaloLinePotentialCustom_fromFunc(func=P.HernquistPotential, m=sh["m"], r_s=sh["rs"], subhalo_x0=sh["x0"], subhalo_v=sh["v"],
                                                         subhalo_t0=sh["t0"], t_window=sh["tw"], units=ssc.usys)
            orc_sh = O.Program().subhalos(O.PR_HERNQUIST, sh["m"], sh["rs"], sh["x0"], sh["v"], sh["t0"], sh["tw"])
            w0 = halo_orbits(N, seed=nsh + 1)
            t0 = np.linspace(-800.0, -30.0, N)
            w_o, D_o, st_o, ns_o = O.linear_response(orc_base, orc_sh, w0, t0, 0.0, solver=8, rtol=1e-8, atol=1e-8, dtmin=2.0, dtmax=2.0, threads=8)
            assert not st_o.any()
            for retire in (1, 0):
                os.environ["SSB_RESP_RETIRE"] = str(retire)
                w, D, st, ns = rt.linear_response(base, pert._arrays, rt.to_dev(w0), None, rt.to_dev(t0), 0.0, ctrl)
                assert not bool(st.any()) and np.array_equal(ns.cpu().numpy()[:, 0], ns_o[:, 0]), (nsh, N, retire)
                assert scaled_err(w.cpu().numpy(), w_o, 1e-10).max() < 1.0, (nsh, N, retire)
                assert np.abs(D.cpu().numpy() - D_o).max() <= 1e-9 * np.abs(D_o).max() + 1e-300, (nsh, N, retire)
    finally:
        os.environ.pop("SSB_RESP_RETIRE", None)
