"""C4 on N GPUs (weak scaling: n_p particles per GPU x n_sh subhalos, particles interleaved over ranks, only final states and
response summaries all-gathered).  Launch: python -m torch.distributed.run --nproc-per-node N tools/bench_response_multi.py [n_p] [n_sh] [tol]"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tools"))
import numpy as np, torch
import streamsculptor_b200 as ssc
from streamsculptor_b200 import _runtime as rt, parallel as par
from _workloads import mw3_product

n_p = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
n_sh = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
tol = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-11
rank, world = par.init_from_env("nccl") if int(os.environ.get("WORLD_SIZE", "1")) > 1 else (0, 1)
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
pot = mw3_product()
P = ssc.potential
N = n_p * world
back = pot.integrate_orbit(w0=[20.0, 0.0, 20.0, 0.0, 0.15, 0.0], ts=np.array([0.0, -3000.0]), t0=0.0, t1=-3000.0).ys[-1]
ts = np.linspace(-3000.0, 0.0, N // 2 + 1)
nr = np.random.Generator(np.random.PCG64(0)).standard_normal((len(ts), 4))
pl, pt, vl, vt = pot.gen_stream_ics(ts=ts, prog_w0=back, Msat=1e4, seed_num=583, solver=ssc.Dopri8(), normals=nr)
w0 = np.vstack([np.hstack([pl, vl])[:-1], np.hstack([pt, vt])[:-1]])
t0 = np.concatenate([ts[:-1], ts[:-1]])
rng = np.random.Generator(np.random.PCG64(1234))
M = 10 ** rng.uniform(5, 9, n_sh); rs = 1.05 * np.sqrt(M / 1e8)
t_imp = rng.uniform(-3000.0, 0.0, n_sh)
prog_at = pot.integrate_orbit(w0=back, ts=np.sort(t_imp), t0=-3000.0, t1=0.0).ys[np.argsort(np.argsort(t_imp))]
b = rng.uniform(0, 10 * rs)
d = rng.normal(size=(n_sh, 3)); d /= np.linalg.norm(d, axis=1)[:, None]
pert = P.SubhaloLinePotentialCustom_fromFunc(func=P.HernquistPotential, m=np.ones(n_sh), r_s=rs, subhalo_x0=prog_at[:, :3] + b[:, None] * d,
                                             subhalo_v=rng.normal(size=(n_sh, 3)) * 0.184, subhalo_t0=t_imp, t_window=150.0, units=ssc.usys)
ctrl = rt.make_ctrl(ssc.Dopri8(), tol, tol, 0.01, None, 10_000)
w0_d, t0_d = rt.to_dev(w0), rt.to_dev(t0)
best = 1e30
for it in range(4):
    torch.cuda.synchronize()
    if world > 1: par.dist().barrier()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    w, dstream, D_l, sel = par.linear_response_sharded(pot, pert._arrays, w0_d, t0_d, 0.0, ctrl, rank, world, M)
    e.record(); torch.cuda.synchronize()
    tt = torch.tensor([a.elapsed_time(e)], device="cuda")
    if world > 1: par.dist().all_reduce(tt, op=par.dist().ReduceOp.MAX)
    if it > 0: best = min(best, float(tt.item()))
if rank == 0:
    print(f"C4 x{world} GPUs: {N} particles x {n_sh} subhalos, Dopri8 tol={tol:g}: {best:.1f} ms (max over ranks), |dstream|max {float(dstream.abs().max()):.3e}, "
          f"finite {bool(torch.isfinite(dstream).all())}")
if world > 1: par.dist().destroy_process_group()
