import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import oracle as O
import streamsculptor_b200 as ssc
from streamsculptor_b200 import RestrictedNbody as RN
import test_gpu_parity as T
from common import scaled_err
orc, mw, track, tk, yk = T._restricted_pair()
rng = np.random.default_rng(5)
N = 700
w0 = np.hstack([yk[0] + rng.normal(size=(N, 3)) * 0.02, T._restricted_pair.v0 + rng.normal(size=(N, 3)) * 5e-4])
field = RN.RestrictedNbody_generator(potential=mw, progenitor_potential=ssc.potential.PlummerPotential, interp_prog=track, init_mass=2e4, init_rs=0.01, r_esc=0.05)
for t1 in (-550.0, -400.0, 0.0):
    for tol in (1e-8, 1e-10):
        sol = ssc.integrate_field(w0=w0, ts=np.array([-600.0, t1]), solver=ssc.Dopri8(), field=field, rtol=tol, atol=tol, dtmin=0.05, max_steps=20000)
        yo, st, ns = O.shared_step_orbits(orc, w0, -600.0, t1, solver=8, rtol=tol, atol=tol, dtmin=0.05, max_steps=20000)
        yt, _, _ = orc.integrate_orbits(w0, -600.0, t1, rtol=1e-13, atol=1e-13, dtmin=1e-4, max_steps=400000, threads=8)
        print(t1, tol, {k: int(v) for k, v in sol.stats.items()}, ns, "gpu-orc", scaled_err(sol.ys[-1], yo[0], tol).max(), "orc-truth", scaled_err(yo[0], yt[:, 0], tol).max(),
              "gpu-truth", scaled_err(sol.ys[-1], yt[:, 0], tol).max())
