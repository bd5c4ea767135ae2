"""K1 A/B microbenchmark: C2 stream (1e6 particles, Dopri8 1e-7) through gen_stream_vmapped; CUDA events, best of 5.
Usage: SSB_LIB_PATH=build/variants/x.so python tools/bench_k1.py [n_particles] [solver]"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tools"))
import numpy as np, torch
import streamsculptor_b200 as ssc
from streamsculptor_b200 import _runtime as rt
from _workloads import mw3_product

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
solver = ssc.Dopri5() if (len(sys.argv) > 2 and sys.argv[2] == "5") else ssc.Dopri8()
mw3 = mw3_product()
base = mw3
back = mw3.integrate_orbit(w0=[20.0, 0.0, 20.0, 0.0, 0.15, 0.0], ts=np.array([0.0, -3000.0]), t0=0.0, t1=-3000.0).ys[-1]
if len(sys.argv) > 3 and sys.argv[3] in ("c3", "c3cubic"):      # C3: MW3 + translating Plummer on a 1000-knot linear (or cubic) track (its own orbit in MW3)
    P = ssc.potential
    tk = np.linspace(-3000.0, 0.0, 1000)
    lmc = base.integrate_orbit(w0=[-1.0, -41.0, -28.0, -0.058, -0.23, 0.23], ts=tk[::-1].copy(), t0=0.0, t1=-3000.0).ys[::-1, :3].copy()
    mw3 = P.Potential_Combine([base, P.TimeDepTranslatingPotential(P.PlummerPotential(m=1.5e11, r_s=10.8, units=ssc.usys), (ssc.CubicTrack if sys.argv[3] == "c3cubic" else ssc.LinearTrack)(tk, lmc),
                                                                    units=ssc.usys)], units=ssc.usys)
ts = rt.to_dev(np.linspace(-3000.0, 0.0, n // 2 + 1))
pw = rt.to_dev(back)
fn = lambda: mw3.gen_stream_vmapped(ts=ts, prog_w0=pw, Msat=1e4, seed_num=583, solver=solver, _return_stats=True)
for _ in range(3):
    out = fn()
best = 1e30
for _ in range(5):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); out = fn(); b.record(); torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b))
steps = int(out[3][..., 0].sum().item())
chk = float(out[0].sum().item() + out[1].sum().item())
print(f"{os.environ.get('SSB_LIB_PATH', 'default')}: {best:.3f} ms, {steps} steps, {steps / best * 1e3:.4e} particle-steps/s, checksum {chk:.12e}")
