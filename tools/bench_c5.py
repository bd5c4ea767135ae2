"""C5 measurements (BASELINE.json config 5, stretch): restricted N-body tracers as ONE ODE with a shared controller (K5), variational /
tangent ODEs along every orbit (K7) and a live softened N-body field of 100 bodies (K6).  CUDA events, best of 3.
Usage: python tools/bench_c5.py [n_tracers] [n_variational]"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tools"))
import numpy as np, torch
import streamsculptor_b200 as ssc
from streamsculptor_b200 import RestrictedNbody as RN
from _workloads import mw3_product, halo_orbits

n_tr = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
n_var = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_000_000
mw = mw3_product()


def timed(fn, reps=3):
    out = fn()
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best, out


# ---- K5: N tracers around a Plummer progenitor on a cubic track, ONE ODE (RestrictedNbody.py:93-106) ----
tk = np.linspace(-600.0, 0.0, 301)
yk = mw.integrate_orbit(w0=[20.0, 0.0, 20.0, 0.0, 0.15, 0.0], ts=tk[::-1].copy(), t0=0.0, t1=-600.0, rtol=1e-10, atol=1e-10, dtmin=1e-3,
                        max_steps=100_000).ys[::-1].copy()
field = RN.RestrictedNbody_generator(potential=mw, progenitor_potential=ssc.potential.PlummerPotential, interp_prog=ssc.CubicTrack(tk, yk[:, :3].copy()),
                                     init_mass=2e4, init_rs=0.01, r_esc=0.05)
g = torch.Generator(device="cuda").manual_seed(5)
w0 = torch.cat([torch.randn((n_tr, 3), generator=g, device="cuda", dtype=torch.float64) * 0.02, torch.randn((n_tr, 3), generator=g, device="cuda",
                dtype=torch.float64) * 5e-4], 1) + torch.as_tensor(yk[0], device="cuda")
ms, sol = timed(lambda: ssc.integrate_field(w0=w0, ts=np.array([-600.0, -400.0]), solver=ssc.Dopri8(), field=field, rtol=1e-8, atol=1e-8, dtmin=0.05,
                                            max_steps=5000))
ns = int(sol.stats["num_steps"])
print(f"C5 restricted N-body (K5): {n_tr} tracers as one ODE, 200 Myr, Dopri8 1e-8: {ms:.1f} ms, {ns} shared steps, "
      f"{n_tr * ns / ms * 1e3:.3e} tracer-steps/s, {ms / ns * 1e3:.1f} us per shared step")

# ---- K7: state-transition matrix (order 1) and second-order tensor (order 2) along independent orbits ----
wv = halo_orbits(n_var, seed=3)
wv_d = torch.as_tensor(wv, device="cuda")
for order, n in ((1, n_var), (2, max(n_var // 10, 1))):
    ms, out = timed(lambda: ssc.fields.integrate_variational_batch(mw, wv_d[:n], -1000.0, 0.0, order=order, solver=ssc.Dopri8(), rtol=1e-7, atol=1e-7,
                                                                   dtmin=0.05, max_steps=10_000))
    failed = int((out[3] != 0).sum().item())
    print(f"C5 variational order {order} (K7): {n} orbits x 1 Gyr, Dopri8 1e-7: {ms:.1f} ms ({n / ms * 1e3:.3e} orbits/s), failed {failed}")

# ---- K6: 100 live softened bodies in MW3 as ONE ODE, 64 saved rows (fields.py:115-155) ----
rng = np.random.default_rng(7)
wb = halo_orbits(100, seed=9)
nb = ssc.fields.Nbody_field(ext_pot=mw, masses=10 ** rng.uniform(6, 9, 100), units=ssc.usys, eps=0.05)
ms, sol = timed(lambda: ssc.integrate_field(w0=wb, ts=np.linspace(-1000.0, 0.0, 64), solver=ssc.Dopri8(), field=nb, rtol=1e-8, atol=1e-8, dtmin=0.01,
                                            max_steps=20_000))
print(f"C5 live N-body field (K6): 100 bodies, 1 Gyr, Dopri8 1e-8, 64 rows: {ms:.1f} ms, {int(sol.stats['num_steps'])} steps")

# ---- C5 end to end: 100 moving perturbers (tables from the live N-body pre-integration above) + restricted N-body tracers + tangent ODEs ----
nk = 256
tk5 = np.linspace(-600.0, 0.0, nk)
solp = ssc.integrate_field(w0=wb, ts=tk5, t0=-600.0, t1=0.0, solver=ssc.Dopri8(), field=nb, rtol=1e-8, atol=1e-8, dtmin=0.01, max_steps=20_000)
cen = np.asarray(solp.ys)[:, :, :3]                                    # [nk, 100, 3] perturber centres
pset = ssc.potential.PerturberSetPotential(ssc.potential.PlummerPotential, nb.masses, np.full(100, 0.5), tk5, cen, units=ssc.usys)
pot5 = ssc.potential.Potential_Combine([mw, pset], units=ssc.usys)
field5 = RN.RestrictedNbody_generator(potential=pot5, progenitor_potential=ssc.potential.PlummerPotential, interp_prog=ssc.CubicTrack(tk, yk[:, :3].copy()),
                                      init_mass=2e4, init_rs=0.01, r_esc=0.05)
ms, sol = timed(lambda: ssc.integrate_field(w0=w0, ts=np.array([-600.0, -400.0]), solver=ssc.Dopri8(), field=field5, rtol=1e-8, atol=1e-8, dtmin=0.05,
                                            max_steps=5000))
ns = int(sol.stats["num_steps"])
print(f"C5 end to end (K5 + perturber set): {n_tr} tracers + 100 tabulated moving perturbers as one ODE, 200 Myr, Dopri8 1e-8: {ms:.1f} ms, {ns} shared steps, "
      f"{n_tr * ns / ms * 1e3:.3e} tracer-steps/s, {ms / ns * 1e3:.1f} us per shared step, {n_tr * ns * 100 / ms * 1e3:.3e} tracer-perturber pair-steps/s")
nv5 = max(n_var // 10, 1)
ms, out = timed(lambda: ssc.fields.integrate_variational_batch(pot5, wv_d[:nv5], -600.0, 0.0, order=1, solver=ssc.Dopri8(), rtol=1e-7, atol=1e-7, dtmin=0.05,
                                                               max_steps=10_000))
print(f"C5 tangent ODEs in the same potential (K7 order 1): {nv5} orbits x 600 Myr: {ms:.1f} ms ({nv5 / ms * 1e3:.3e} orbits/s), failed {int((out[3] != 0).sum().item())}")
