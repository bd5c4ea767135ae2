"""Throughput of the BASELINE.json configs other than the bench.py headline (C2): C1, C3, C2 with 64 saved snapshots.
Usage: python tools/bench_configs.py [n_particles]   (prints one line per config; CUDA events, 3 warm-ups, best of 3)"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tools"))
import numpy as np, torch
import streamsculptor_b200 as ssc
from streamsculptor_b200 import _runtime as rt
from _workloads import mw3_product

n_part = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
P = ssc.potential
mw3 = mw3_product()
back = mw3.integrate_orbit(w0=[20.0, 0.0, 20.0, 0.0, 0.15, 0.0], ts=np.array([0.0, -3000.0]), t0=0.0, t1=-3000.0).ys[-1]


def timed(fn, reps=3):
    for _ in range(2):
        out = fn()
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best, out


def stream(pot, n, solver, label, **kw):
    ts = rt.to_dev(np.linspace(-3000.0, 0.0, n // 2 + 1))
    ms, out = timed(lambda: pot.gen_stream_vmapped(ts=ts, prog_w0=rt.to_dev(back), Msat=1e4, seed_num=583, solver=solver, _return_stats=True, **kw))
    steps = int(out[3][..., 0].sum().item())
    print(f"{label}: {n} particles, {ms:.2f} ms, {steps} particle-steps, {steps / ms * 1e3:.3e} particle-steps/s, failed {int((out[2] != 0).sum().item())}")


stream(mw3, 2000, ssc.Dopri8(), "C1  MW3 Dopri8 2000 particles")
stream(mw3, n_part, ssc.Dopri8(), "C2  MW3 Dopri8 final state")
stream(mw3, n_part, ssc.Dopri5(), "C2' MW3 Dopri5 final state (gen_stream default solver)")
gala = P.GalaMilkyWayPotential(units=ssc.usys)
stream(gala, n_part, ssc.Dopri8(), "C2g GalaMilkyWayPotential Dopri8")
# C3: MW3 + translating Plummer (m=1.5e11, r_s=10.8) on a 1000-knot linear table = its own orbit in MW3
tk = np.linspace(-3000.0, 0.0, 1000)
lmc_back = mw3.integrate_orbit(w0=[-1.0, -41.0, -28.0, -0.058, -0.23, 0.23], ts=tk[::-1].copy(), t0=0.0, t1=-3000.0).ys[::-1, :3].copy()
c3 = P.Potential_Combine([mw3, P.TimeDepTranslatingPotential(P.PlummerPotential(m=1.5e11, r_s=10.8, units=ssc.usys), ssc.LinearTrack(tk, lmc_back), units=ssc.usys)],
                         units=ssc.usys)
stream(c3, n_part, ssc.Dopri8(), "C3  MW3 + moving Plummer (1000-knot track) Dopri8")
# C2 with M = 64 saved snapshots per particle (integrate_orbit_batch_vmapped, ts[N,M]): output-heavy
n64 = min(n_part, 1_000_000)
ts = np.linspace(-3000.0, 0.0, n64 // 2 + 1)
pl, pt, vl, vt = mw3.gen_stream_ics(ts=rt.to_dev(ts), prog_w0=rt.to_dev(back), Msat=1e4, seed_num=583, solver=ssc.Dopri8())
w0 = torch.cat([torch.cat([pl, vl], 1)[:-1], torch.cat([pt, vt], 1)[:-1]]).contiguous()
t0 = rt.to_dev(np.concatenate([ts[:-1], ts[:-1]]))
frac = torch.linspace(0.0, 1.0, 64, dtype=torch.float64, device=w0.device)
tsN = (t0[:, None] + (0.0 - t0)[:, None] * frac[None, :]).contiguous()
tsN[:, -1] = 0.0
ms, sol = timed(lambda: mw3.integrate_orbit_batch_vmapped(w0=w0, ts=tsN, solver=ssc.Dopri8()))
steps = int(np.asarray(sol.stats["num_steps"]).sum())
gb = sol.ys.numel() * 8 / 1e9
print(f"C2-M64 saved snapshots: {w0.shape[0]} particles x 64, {ms:.2f} ms, {steps / ms * 1e3:.3e} particle-steps/s, output {gb:.2f} GB -> {gb / ms * 1e3:.0f} GB/s written")
