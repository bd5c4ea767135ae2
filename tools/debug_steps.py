"""Debug helper: compare the accepted-step sequence of the CUDA stepper (K0 recorder) with the oracle's."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import oracle as O
import streamsculptor_b200 as ssc
from streamsculptor_b200 import _runtime as rt
from common import mw3_oracle, mw3_product, random_orbits
np.set_printoptions(precision=15, linewidth=200)
orc, prod = mw3_oracle(), mw3_product()
w0 = [20.0, 0.0, 20.0, 0.0, 0.15, 0.0]
for solver, tol in ((5, 1e-7), (8, 1e-7), (8, 1e-10)):
    tg, yg = orc.orbit_steps(w0, -3000.0, 0.0, solver=solver, rtol=tol, atol=tol)
    ctrl = rt.make_ctrl(ssc.Dopri8() if solver == 8 else ssc.Dopri5(), tol, tol, 0.3, None, 10000)
    ts = rt.to_dev(np.array([0.0]))
    ys, st, ns, scratch = rt.orbit_dense(prod, rt.to_dev(w0), -3000.0, 0.0, ts, ctrl)
    sc = scratch.cpu().numpy()
    n = int(sc[0]); rec = sc[8:8 + 64 * n].reshape(n, 64)
    tb = rec[:, 1]
    m = min(n, len(tg) - 1)
    d = np.abs(tb[:m] - tg[1:m + 1])
    first = np.argmax(d > 1e-9) if (d > 1e-9).any() else -1
    print(f"solver {solver} tol {tol}: gpu n_acc {n}, oracle n_acc {len(tg)-1}, max |dt_boundary| {d.max():.3e}, first diverge idx {first}")
    print("   final diff", np.abs(ys.cpu().numpy()[0] - yg[-1]).max(), " y1 diffs along steps:", np.abs(rec[:m, 8:11] - yg[1:m+1, :3]).max(axis=1)[[0, 1, 2, 5, 10, m // 2, m - 1]])
    if first >= 0:
        lo = max(0, first - 2)
        print("   gpu tb", tb[lo:first + 3]); print("   orc tb", tg[1 + lo:first + 4])
