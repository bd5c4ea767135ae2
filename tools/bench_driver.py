"""Production driver (generate_derivs.get_derivs) end to end: batches of sampled subhalo impacts on one stream model, wall-clock per batch with
1, 2 and 4 batches in flight.  The base potential includes the progenitor's own compact Plummer sphere (as the reference's driver builds it),
so every batch has a few particles that take thousands of steps: one batch at a time is bound by them, overlapping batches is not.
Usage: python tools/bench_driver.py [N_arm] [N_batch] [n_batches] [r_s]"""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tools"))
import numpy as np, torch
import streamsculptor_b200 as ssc
from streamsculptor_b200.generate_derivs import get_derivs
from _workloads import mw3_product

n_arm = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
n_batch = int(sys.argv[2]) if len(sys.argv) > 2 else 500
n_it = int(sys.argv[3]) if len(sys.argv) > 3 else 6
r_s = float(sys.argv[4]) if len(sys.argv) > 4 else 0.004
pot = mw3_product()
prog_today = np.array([20.0, 0.0, 20.0, 0.0, 0.15, 0.0])


def phi1(stream):               # stand-in for the observer-frame transform: degrees along the orbital plane of this progenitor
    s = np.asarray(stream)
    return np.degrees(np.arctan2(s[:, 1], (s[:, 0] + s[:, 2]) / np.sqrt(2.0)))


# where this stream lies in phi1 (a user knows it from the data): bounds inside it, windows of a tenth of its length
IC = np.asarray(pot.integrate_orbit(w0=prog_today, t0=0.0, t1=-3000.0, ts=np.array([-3000.0])).ys[0])
ts_ = np.hstack([np.linspace(-3000.0, -1.0, n_arm), [0.0]])
l_, t_ = ssc.gen_stream_vmapped_Chen25(pot_base=pot, prog_w0=IC, ts=ts_, key=3, Msat=1e4, atol=1e-7, rtol=1e-7, solver=ssc.Dopri8(),
                                       prog_pot=ssc.potential.PlummerPotential(m=1e4, r_s=r_s, units=ssc.usys))
ph_all = phi1(np.vstack([np.asarray(l_), np.asarray(t_)]))
ph_prog = float(phi1(prog_today[None])[0])
lo, hi = np.percentile(ph_all, [10, 90])
kw = dict(prog_wtoday=prog_today, t_age=3000.0, t_dissolve=-1.0, log10_min_mass=5.0, log10_max_mass=8.0, phi1_bounds=[lo, hi], phi1_exclude=[ph_prog - 0.5, ph_prog + 0.5],
          stream_seednum=3, key=21, Msat=1e4, r_s=r_s, target_num=n_it * n_batch, phi1_function=phi1, pot=pot, N_batch=n_batch, atol=1e-11, rtol=1e-11,
          phi1window=0.1 * (hi - lo), N_arm=n_arm, path=None, save=False)
get_derivs(**dict(kw, target_num=n_batch, pipeline=1))          # warm-up (library load, allocator)
for depth in (1, 2, 4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = get_derivs(pipeline=depth, **kw)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    D = out[0]["pert_out"][1]
    print(f"get_derivs: {n_it} batches of {n_batch} subhalos x {2 * n_arm + 1} particles, rtol = atol = 1e-11, progenitor r_s = {r_s}: pipeline={depth}: "
          f"{dt / n_it * 1e3:.1f} ms per batch (wall, host sampling + device solve + D2H of {D.nbytes / 1e6:.0f} MB per batch), max |D| = {np.abs(D[np.isfinite(D)]).max():.2e}")
