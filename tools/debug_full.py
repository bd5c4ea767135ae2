import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
import oracle as O
import streamsculptor_b200 as ssc
from common import *
import test_gpu_parity as T
orc, prod = T._full_pair(None)
w0 = halo_orbits(40, seed=2)
ts = np.linspace(-2500, -100, 17)[None, :] + np.linspace(0, 50, 40)[:, None]
q = lambda v: np.percentile(v, [50, 90, 100]).round(2)
for solver in (5, 8):
    sv = ssc.Dopri8() if solver == 8 else ssc.Dopri5()
    ys_o, _, ns_o = orc.integrate_orbits(w0, ts[:, 0], ts[:, -1], ts=ts, solver=solver, dtmin=0.05, threads=8)
    ys_t, _, _ = orc.integrate_orbits(w0, ts[:, 0], ts[:, -1], ts=ts, **T.TRUTH)
    sol = prod.integrate_orbit_batch_vmapped(w0=w0, ts=ts, t0=ts[:, 0], t1=ts[:, -1], solver=sv, dtmin=0.05)
    print(solver, "same count", np.mean(sol.stats["num_steps"] == ns_o[:, 0]), "steps", ns_o[:3], sol.stats["num_steps"][:3])
    print("  d_ab", q(scaled_err(sol.ys, ys_o, 1e-7)), "d_bt", q(scaled_err(ys_o, ys_t, 1e-7)), "d_at", q(scaled_err(sol.ys, ys_t, 1e-7)))
    print("  last row only: d_ab", q(scaled_err(sol.ys[:, -1], ys_o[:, -1], 1e-7)), "d_bt", q(scaled_err(ys_o[:, -1], ys_t[:, -1], 1e-7)))
