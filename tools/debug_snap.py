import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
import oracle as O
import streamsculptor_b200 as ssc
from common import *
P = ssc.potential
t, y = lmc_track()
tv = np.linspace(-3000, 0, 50)
vel = np.stack([1e-3 * np.sin(tv / 500.0), 2e-3 * np.cos(tv / 800.0), 1e-4 * tv / 3000.0], axis=1)
sh = subhalo_set(12, tw=400.0)
def build(which):
    orc = mw3_oracle(); lst = [mw3_product()]
    if 'tr' in which:
        tr = orc.track(O.LINEAR, t, y); orc.plummer(1.5e11, 10.8, track=tr)
        lst.append(P.TimeDepTranslatingPotential(P.PlummerPotential(m=1.5e11, r_s=10.8, units=ssc.usys), ssc.LinearTrack(t, y), units=ssc.usys))
    if 'ua' in which:
        orc.uniform_acc(tv, vel); lst.append(P.UniformAcceleration(ssc.LinearTrack(tv, vel), units=ssc.usys))
    if 'sh' in which:
        orc.subhalos(O.PR_HERNQUIST, sh["M"], sh["rs"], sh["x0"], sh["v"], sh["t0"], sh["tw"])
        lst.append(P.SubhaloLinePotentialCustom_fromFunc(func=P.HernquistPotential, m=sh["M"], r_s=sh["rs"], subhalo_x0=sh["x0"], subhalo_v=sh["v"], subhalo_t0=sh["t0"], t_window=sh["tw"], units=ssc.usys))
    return orc, P.Potential_Combine(lst, units=ssc.usys)
w0 = halo_orbits(40, seed=2)
ts = np.linspace(-2500, -100, 17)[None, :] + np.linspace(0, 50, 40)[:, None]
for which in ([], ['tr'], ['ua'], ['sh'], ['tr','ua','sh']):
    orc, prod = build(which)
    for solver in (5, 8):
        sv = ssc.Dopri8() if solver == 8 else ssc.Dopri5()
        for name, tsx in (("fwd", ts), ("bwd", ts[:, ::-1].copy())):
            ys_f, _, ns_f = orc.integrate_orbits(w0, tsx[:, 0], tsx[:, -1], ts=tsx, solver=solver, dtmin=1.0, dtmax=1.0, threads=8)
            sol_f = prod.integrate_orbit_batch_vmapped(w0=w0, ts=tsx, t0=tsx[:, 0], t1=tsx[:, -1], solver=sv, dtmin=1.0, dtmax=1.0)
            e = scaled_err(sol_f.ys, ys_f, 1e-10)
            print(which, solver, name, "max %.3g" % e.max(), "n>1:", (e > 1).sum(), "last-row max %.3g" % scaled_err(sol_f.ys[:, -1], ys_f[:, -1], 1e-10).max())
