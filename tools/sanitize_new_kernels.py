"""Small invocations of the kernels touched in rounds 1-2, sized for compute-sanitizer (memcheck / racecheck):
response_kernel_mp with 4 and 16 particle slots per CTA and retired items, the zero-copy host outputs of the orbit kernel, the shared-step
attempt kernel with once-per-stage track evaluation, the saving orbit kernel (warp-cooperative dense output), the four-part stream pipeline,
a perturber set and a growth factor.  Usage: compute-sanitizer --tool racecheck python tools/sanitize_new_kernels.py"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tools"))
import ctypes as C
import numpy as np, torch
import streamsculptor_b200 as ssc
from streamsculptor_b200 import _lib, _runtime as rt, RestrictedNbody as RN
from _workloads import mw3_product, halo_orbits, subhalo_set

P = ssc.potential
base = mw3_product()
# (1) K3-mp, 4 slots per CTA, zero and non-zero perturbation ICs
os.environ["SSB_RESP_NP"] = "4"
nsh = 40
sh = subhalo_set(nsh, seed=5, t_lo=-300.0, tw=60.0)
pert = P.SubhaloLinePotentialCustom_fromFunc(func=P.HernquistPotential, m=sh["m"], r_s=sh["rs"], subhalo_x0=sh["x0"], subhalo_v=sh["v"],
                                             subhalo_t0=sh["t0"], t_window=sh["tw"], units=ssc.usys)
N = 10
w0, t0 = halo_orbits(N, seed=11), np.linspace(-300.0, -5.0, N)
D0 = np.random.default_rng(2).normal(size=(N, nsh, 12)) * 1e-6
for solver, d0 in ((ssc.Dopri8(), None), (ssc.Dopri5(), D0)):
    ctrl = rt.make_ctrl(solver, 1e-7, 1e-7, 0.01, None, 10_000)
    w, D, st, ns = rt.linear_response(base, pert._arrays, rt.to_dev(w0), None if d0 is None else rt.to_dev(d0), rt.to_dev(t0), 0.0, ctrl)
    assert int((st != 0).sum().item()) == 0
print("response_kernel_mp ok", ns[:, 0].cpu().numpy())
# (2) zero-copy host outputs
Pst, keep = rt.lower(base)
host = _lib.Potential.from_buffer_copy(Pst)
nts = 78; n = nts - 1
ts = torch.from_numpy(np.linspace(-300.0, 0.0, nts)); pw = torch.from_numpy(np.array([-3.0, 14.0, 8.0, 0.14, 0.02, -0.07])); ms = torch.full((nts,), 1e4, dtype=torch.float64)
out = torch.empty((2, n, 6), dtype=torch.float64).pin_memory(); st = torch.empty((2, n), dtype=torch.int32).pin_memory(); nsb = torch.empty((2, n, 3), dtype=torch.int32).pin_memory()
hp = lambda t: C.c_void_p(t.data_ptr())
ctrl = rt.make_ctrl(ssc.Dopri8(), 1e-7, 1e-7, 0.3, None, 10_000)
_lib.check(_lib.lib().ssb_gen_stream_host(C.byref(host), C.byref(host), base._G, nts, hp(ts), hp(pw), hp(ms), 583, (C.c_double * 8)(*ssc.main.DEFAULT_KVALS), None,
                                          ctrl, 0, 1, n, hp(out[0]), hp(out[1]), hp(st), hp(nsb)))
assert (st == 0).all() and torch.isfinite(out).all()
print("zero-copy host outputs ok")
# (3) shared-step tracers: fused galaxy + progenitor on a cubic track evaluated once per stage
tk = np.linspace(-60.0, 0.0, 31)
yk = base.integrate_orbit(w0=[20.0, 0.0, 20.0, 0.0, 0.15, 0.0], ts=tk[::-1].copy(), t0=0.0, t1=-60.0).ys[::-1].copy()
field = RN.RestrictedNbody_generator(potential=base, progenitor_potential=P.PlummerPotential, interp_prog=ssc.CubicTrack(tk, yk[:, :3].copy()), init_mass=2e4,
                                     init_rs=0.01, r_esc=0.05)
rng = np.random.default_rng(5)
wt = np.hstack([yk[0, :3] + rng.normal(size=(300, 3)) * 0.02, yk[0, 3:] + rng.normal(size=(300, 3)) * 5e-4])
sol = ssc.integrate_field(w0=wt, ts=np.array([-60.0, -40.0]), solver=ssc.Dopri8(), field=field, rtol=1e-8, atol=1e-8, dtmin=0.05, max_steps=500)
assert np.isfinite(sol.ys).all()
print("shared-step kernels ok", int(sol.stats["num_steps"]))

# ---- round 2 ----
# (4) K3-mp with 16 slots per CTA and retired items (long spans: windows close for good, chunks of 16 retire, logs are applied at the end)
os.environ["SSB_RESP_NP"] = "16"
nsh = 48
sh = subhalo_set(nsh, seed=7, t_lo=-900.0, tw=40.0)
pert = P.SubhaloLinePotentialCustom_fromFunc(func=P.HernquistPotential, m=sh["m"], r_s=sh["rs"], subhalo_x0=sh["x0"], subhalo_v=sh["v"],
                                             subhalo_t0=sh["t0"], t_window=sh["tw"], units=ssc.usys)
N = 40
w0, t0 = halo_orbits(N, seed=13), np.linspace(-1000.0, -5.0, N)
ctrl = rt.make_ctrl(ssc.Dopri8(), 1e-7, 1e-7, 0.01, None, 10_000)
w, D, st, ns = rt.linear_response(base, pert._arrays, rt.to_dev(w0), None, rt.to_dev(t0), 0.0, ctrl)
assert int((st != 0).sum().item()) == 0 and bool(torch.isfinite(D).all())
os.environ.pop("SSB_RESP_NP")
print("response_kernel_mp, 16 slots + retired items ok")
# (5) saving orbit kernel: per-orbit save times, warp-cooperative dense output (both solvers)
wq = halo_orbits(200, seed=3)
for solver in (ssc.Dopri8(), ssc.Dopri5()):
    tsq = np.sort(np.random.default_rng(1).uniform(-400.0, 0.0, size=(200, 9)), axis=1)
    tsq[:, 0] = -400.0
    sol = base.integrate_orbit_batch_vmapped(w0=wq, ts=tsq, solver=solver, rtol=1e-7, atol=1e-7, dtmin=0.3)
    assert np.isfinite(np.asarray(sol.ys)).all()
print("saving orbit kernel ok")
# (6) four-part stream pipeline (>= 32768 particles per arm), perturber set and growth factor in the orbit kernel
ts4 = np.linspace(-400.0, 0.0, 33001)
lead, trail = base.gen_stream_vmapped(ts=ts4, prog_w0=[-3.0, 14.0, 8.0, 0.14, 0.02, -0.07], Msat=1e4, seed_num=5, solver=ssc.Dopri8())
assert np.isfinite(lead).all() and np.isfinite(trail).all()
tkp = np.linspace(-400.0, 0.0, 41)
cen = np.random.default_rng(4).normal(size=(41, 20, 3)) * 30.0
pset = P.PerturberSetPotential(P.PlummerPotential, np.full(20, 1e8), np.full(20, 0.5), tkp, cen, units=ssc.usys)
grow = P.GrowingPotential(P.PlummerPotential(m=1e10, r_s=2.0, units=ssc.usys), (tkp, 1.0 + 1e-3 * (tkp + 400.0)), units=ssc.usys)
pot6 = P.Potential_Combine([base, pset, grow], units=ssc.usys)
sol = pot6.integrate_orbit_batch_vmapped(w0=wq, ts=np.array([-400.0, 0.0]), solver=ssc.Dopri8(), rtol=1e-7, atol=1e-7, dtmin=0.3)
assert np.isfinite(np.asarray(sol.ys)).all()
print("four-part pipeline, perturber set, growth factor ok")
# (7) shared-step kernel with both inline extras: the progenitor's moving Plummer and a frozen set of Plummer perturbers (kernel variant PS)
field7 = RN.RestrictedNbody_generator(potential=P.Potential_Combine([base, pset], units=ssc.usys), progenitor_potential=P.PlummerPotential,
                                      interp_prog=ssc.CubicTrack(tk, yk[:, :3].copy()), init_mass=2e4, init_rs=0.01, r_esc=0.05)
sol = ssc.integrate_field(w0=wt, ts=np.array([-60.0, -40.0]), solver=ssc.Dopri8(), field=field7, rtol=1e-8, atol=1e-8, dtmin=0.05, max_steps=500)
assert np.isfinite(sol.ys).all()
print("shared-step kernel with inline perturber set ok", int(sol.stats["num_steps"]))
