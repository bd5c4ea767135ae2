"""C2 with M saved snapshots per particle (integrate_orbit_batch_vmapped, ts[N,M]): output-heavy variant of the stream hot path.
Usage: python tools/bench_snapshots.py [n_particles] [M]"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tools"))
import numpy as np, torch
import streamsculptor_b200 as ssc
from streamsculptor_b200 import _runtime as rt
from _workloads import mw3_product

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
M = int(sys.argv[2]) if len(sys.argv) > 2 else 64
mw3 = mw3_product()
back = mw3.gen_stream_ics  # noqa
# the bench.py stream (C2): progenitor today [20, 0, 20, 0, 0.15, 0] integrated back 3 Gyr
w_then = rt.to_dev(np.asarray(mw3.integrate_orbit(w0=[20.0, 0.0, 20.0, 0.0, 0.15, 0.0], ts=np.array([0.0, -3000.0]), t0=0.0, t1=-3000.0, solver=ssc.Dopri8()).ys[-1]))
ts = np.linspace(-3000.0, 0.0, n // 2 + 1)
pl, pt, vl, vt = mw3.gen_stream_ics(ts=rt.to_dev(ts), prog_w0=w_then, Msat=1e4, seed_num=583, solver=ssc.Dopri8())
w0 = torch.cat([torch.cat([pl, vl], 1)[:-1], torch.cat([pt, vt], 1)[:-1]]).contiguous()
t0 = rt.to_dev(np.concatenate([ts[:-1], ts[:-1]]))
frac = torch.linspace(0.0, 1.0, M, dtype=torch.float64, device=w0.device)
tsN = (t0[:, None] + (0.0 - t0)[:, None] * frac[None, :]).contiguous()
tsN[:, -1] = 0.0
ctrl = rt.make_ctrl(ssc.Dopri8(), 1e-7, 1e-7, 0.3, None, 10_000)
t1 = torch.zeros_like(t0)
best = 1e30
for it in range(4):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ys, st, ns = rt.orbit_integrate(mw3, w0, t0, t1, tsN, ctrl, ts_per_orbit=1)
    b.record(); torch.cuda.synchronize()
    if it: best = min(best, a.elapsed_time(b))
steps = int(ns[:, 0].sum().item())
gb = ys.numel() * 8 / 1e9
print(f"snapshots: {w0.shape[0]} particles x {M}: {best:.2f} ms, {steps / best * 1e3:.3e} particle-steps/s, {gb:.2f} GB saved -> {gb / best * 1e3:.0f} GB/s, "
      f"failed {int((st != 0).sum().item())}, finite {bool(torch.isfinite(ys).all())}, checksum {float(ys.sum().item()):.10e}")
