import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
import oracle as O
import streamsculptor_b200 as ssc
from common import mw3_oracle, mw3_product, random_orbits
orc, prod = mw3_oracle(), mw3_product()
w0 = random_orbits(200, seed=11)
t0 = np.linspace(-3000, -5, 200)
for solver, tol in ((5, 1e-7), (8, 1e-7), (8, 1e-10)):
    ys_o, st_o, ns_o = orc.integrate_orbits(w0, t0, 0.0, solver=solver, rtol=tol, atol=tol, threads=8)
    ys_t, _, _ = orc.integrate_orbits(w0, t0, 0.0, solver=8, rtol=1e-13, atol=1e-13, dtmin=1e-3, max_steps=200000, threads=8)
    sol = prod.integrate_orbit_batch_vmapped(w0=w0, ts=np.zeros((200, 1)), t0=t0, t1=0.0, solver=ssc.Dopri8() if solver == 8 else ssc.Dopri5(), rtol=tol, atol=tol)
    sc = tol * (1 + np.abs(ys_t[:, 0]))
    d_ab = (np.abs(sol.ys[:, 0] - ys_o[:, 0]) / sc).max(1)
    d_bt = (np.abs(ys_o[:, 0] - ys_t[:, 0]) / sc).max(1)
    d_at = (np.abs(sol.ys[:, 0] - ys_t[:, 0]) / sc).max(1)
    same = sol.stats["num_steps"] == ns_o[:, 0]
    q = lambda v: np.percentile(v, [50, 90, 99, 100]).round(3)
    print(f"solver {solver} tol {tol}: same-count {same.mean():.3f}  d_ab {q(d_ab)}  d_bt(oracle err) {q(d_bt)}  d_at(gpu err) {q(d_at)}  ratio ab/bt pct {q(d_ab/np.maximum(d_bt,1e-3))}")
    print("    worst orbits:", np.argsort(d_ab)[-5:], d_ab[np.argsort(d_ab)[-5:]].round(2), d_bt[np.argsort(d_ab)[-5:]].round(2), "rmin", np.linalg.norm(w0[np.argsort(d_ab)[-5:], :3], axis=1).round(1))
