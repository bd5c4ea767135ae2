"""C4 measurement: first-order response of a C1-like stream to N_sh Hernquist subhalos.
Usage: python tools/bench_response.py [n_particles] [n_sh] [tol]   (parity against the oracle lives in tests/test_gpu_parity.py)"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tools"))
import numpy as np, torch
import streamsculptor_b200 as ssc
from streamsculptor_b200 import _runtime as rt
from _workloads import mw3_product

n_p = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
n_sh = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
tol = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-6
pot = mw3_product()
P = ssc.potential
back = pot.integrate_orbit(w0=[20.0, 0.0, 20.0, 0.0, 0.15, 0.0], ts=np.array([0.0, -3000.0]), t0=0.0, t1=-3000.0).ys[-1]
ts = np.linspace(-3000.0, 0.0, n_p // 2 + 1)
nr = np.random.Generator(np.random.PCG64(0)).standard_normal((len(ts), 4))
pl, pt, vl, vt = pot.gen_stream_ics(ts=ts, prog_w0=back, Msat=1e4, seed_num=583, solver=ssc.Dopri8(), normals=nr)
w0 = np.vstack([np.hstack([pl, vl])[:-1], np.hstack([pt, vt])[:-1]])
t0 = np.concatenate([ts[:-1], ts[:-1]])
# subhalos: impact points on the stream at random times (generate_derivs.py:155; GenerateImpactParams.py:24)
rng = np.random.Generator(np.random.PCG64(1234))
M = 10 ** rng.uniform(5, 9, n_sh); rs = 1.05 * np.sqrt(M / 1e8)
t_imp = rng.uniform(-3000.0, 0.0, n_sh)
prog_at = pot.integrate_orbit(w0=back, ts=np.sort(t_imp), t0=-3000.0, t1=0.0).ys[np.argsort(np.argsort(t_imp))]
b = rng.uniform(0, 10 * rs)
d = rng.normal(size=(n_sh, 3)); d /= np.linalg.norm(d, axis=1)[:, None]
x0 = prog_at[:, :3] + b[:, None] * d
v = rng.normal(size=(n_sh, 3)) * 0.184
pert = P.SubhaloLinePotentialCustom_fromFunc(func=P.HernquistPotential, m=np.ones(n_sh), r_s=rs, subhalo_x0=x0, subhalo_v=v, subhalo_t0=t_imp,
                                             t_window=150.0, units=ssc.usys)
if len(sys.argv) > 4 and sys.argv[4] == "prog":          # the production driver's base potential: galaxy + the progenitor's moving Plummer (perturbative.py:642-644)
    tk = np.linspace(-3000.0, 0.0, 2001)
    yk = pot.integrate_orbit(w0=back, ts=tk, t0=-3000.0, t1=0.0).ys
    rs_prog = float(sys.argv[5]) if len(sys.argv) > 5 else 0.004
    mk = ssc.LinearTrack if (len(sys.argv) > 6 and sys.argv[6] == "linear") else ssc.CubicTrack
    pot = P.Potential_Combine([pot, P.TimeDepTranslatingPotential(P.PlummerPotential(m=1e4, r_s=rs_prog, units=ssc.usys), mk(tk, yk[:, :3].copy()),
                                                                  units=ssc.usys)], units=ssc.usys)
    print(f"base potential: MW3 + moving Plummer progenitor (r_s = {rs_prog}) on a {mk.__name__}")
ctrl = rt.make_ctrl(ssc.Dopri8(), tol, tol, 0.01, None, 10_000)
w0_d, t0_d = rt.to_dev(w0), rt.to_dev(t0)
for it in range(3):
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    a.record()
    w, D, st, ns = rt.linear_response(pot, pert._arrays, w0_d, None, t0_d, 0.0, ctrl)
    e.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(e)
steps = int(ns[:, 0].sum().item())
print(f"max |D| = {float(D[torch.isfinite(D)].abs().max()):.3e}; steps acc/rej {int(ns[:, 1].sum())}/{int(ns[:, 2].sum())}")
print(f"C4: {len(w0)} particles x {n_sh} subhalos, Dopri8 tol={tol}: {ms:.1f} ms, particle-steps {steps}, pair-steps/s {steps * n_sh / ms * 1e3:.3e}, "
      f"status!=0: {int((st != 0).sum().item())}, mean steps {steps / len(w0):.1f}, flop/s (2.85 kflop/pair-step) {2850.0 * steps * n_sh / ms * 1e3 / 1e12:.2f} TF")
