import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
import oracle as O
import streamsculptor_b200 as ssc
from streamsculptor_b200 import _runtime as rt
from common import *
import test_gpu_parity as T
np.set_printoptions(precision=15, linewidth=220)
orc, prod = T._full_pair(None)
w0s = halo_orbits(40, seed=2)
for idx in (0, 1):
    w0 = w0s[idx]
    for solver in (8,):
        tg, yg = orc.orbit_steps(w0, -2500.0, -100.0, solver=solver, dtmin=0.05)
        ctrl = rt.make_ctrl(ssc.Dopri8(), 1e-7, 1e-7, 0.05, None, 10000)
        ys, st, ns, scratch = rt.orbit_dense(prod, rt.to_dev(w0), -2500.0, -100.0, rt.to_dev(np.array([-100.0])), ctrl)
        sc = scratch.cpu().numpy(); n = int(sc[0]); rec = sc[8:8 + 64 * n].reshape(n, 64)
        tb = rec[:, 1]; m = min(n, len(tg) - 1)
        d = np.abs(tb[:m] - tg[1:m + 1])
        first = int(np.argmax(d > 1e-6)) if (d > 1e-6).any() else -1
        print(f"orbit {idx}: gpu n_acc {n} (steps {ns.cpu().numpy()}), oracle n_acc {len(tg)-1}; first boundary diff>1e-6 at {first}")
        lo = max(0, first - 4)
        print("  gpu tb ", tb[lo:first + 3]); print("  orc tb ", tg[1 + lo:first + 4]); print("  diff   ", d[lo:first+3])
        print("  state diff at those steps", np.abs(rec[lo:first+3, 8:11] - yg[1+lo:first+4, :3]).max(axis=1))
