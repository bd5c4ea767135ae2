"""streamhelpers.py of the reference: stream generation with perturbers (streamhelpers.py:56-198)."""
import numpy as np

from . import _runtime as rt
from .main import DEFAULT_KVALS
from .potential import Potential_Combine
from .solvers import Dopri5
from .units import usys


def custom_release_model(pos_prog=None, vel_prog=None, pos_rel=None, vel_rel=None):   # streamhelpers.py:186-198
    return np.asarray(pos_prog) + np.asarray(pos_rel), np.asarray(vel_prog) + np.asarray(vel_rel)


def gen_stream_ics_pert(pot_base=None, pot_pert=None, ts=None, prog_w0=None, Msat=None, seed_num=None, solver=Dopri5(scan_kind='bounded'),
                        kval_arr=1.0, max_steps=10_000, rtol=1e-7, atol=1e-7, dtmin=0.1, normals=None):
    """Progenitor orbit in base+pert, release in the smooth base potential (streamhelpers.py:56-84)."""
    dev_in = rt.is_dev(ts)
    tt = rt.torch()
    pot_total = Potential_Combine(potential_list=[pot_base, pot_pert], units=usys)
    ts_d, w0_d, Ms, kv, nr = pot_base._stream_inputs(ts, prog_w0, Msat, kval_arr, normals)
    ctrl = rt.make_ctrl(solver, rtol, atol, dtmin, None, max_steps)
    ws, _, _, _ = rt.orbit_dense(pot_total, w0_d, float(ts_d.min()), float(ts_d.max()), ts_d, ctrl)
    idx = tt.arange(ts_d.shape[0], dtype=tt.int64, device=ts_d.device)
    outs = rt.release_spray(pot_base, pot_base._G, ws, Ms, idx, ts_d, 0 if seed_num is None else int(seed_num), kv, nr)
    return tuple(rt.out(o, dev_in) for o in outs)


def gen_stream_vmapped_with_pert(pot_base=None, pot_pert=None, ts=None, prog_w0=None, Msat=None, seed_num=None,
                                 solver=Dopri5(scan_kind='bounded'), kval_arr=1.0, max_steps=10_000, rtol=1e-7, atol=1e-7, dtmin=0.1,
                                 normals=None):
    """Perturbed stream: progenitor and particles in base+pert, release in base (streamhelpers.py:89-116)."""
    dev_in = rt.is_dev(ts)
    pot_total = Potential_Combine(potential_list=[pot_base, pot_pert], units=usys)
    ts_d, w0_d, Ms, kv, nr = pot_base._stream_inputs(ts, prog_w0, Msat, kval_arr, normals)
    ctrl = rt.make_ctrl(solver, rtol, atol, dtmin, None, max_steps)
    lead, trail, _, _ = rt.gen_stream(pot_total, pot_base, pot_base._G, ts_d, w0_d, Ms, 0 if seed_num is None else int(seed_num), kv, nr, ctrl)
    return rt.out(lead, dev_in), rt.out(trail, dev_in)


gen_stream_scan_with_pert = gen_stream_vmapped_with_pert      # streamhelpers.py:150-182: sequential schedule of the same computation


def gen_stream_vmapped_with_pert_fixed_prog(pot_base=None, pot_pert=None, ts=None, prog_w0=None, Msat=None, seed_num=None,
                                            solver=Dopri5(scan_kind='bounded'), kval_arr=1.0, max_steps=10_000, rtol=1e-7, atol=1e-7,
                                            dtmin=0.1, normals=None):
    """Progenitor orbit and release in the base potential, particles in base+pert (streamhelpers.py:118-146)."""
    dev_in = rt.is_dev(ts)
    tt = rt.torch()
    pot_total = Potential_Combine(potential_list=[pot_base, pot_pert], units=usys)
    pl, pt, vl, vt = pot_base.gen_stream_ics(ts=rt.to_dev(ts), prog_w0=prog_w0, Msat=Msat, seed_num=seed_num, solver=solver, kval_arr=kval_arr,
                                             max_steps=max_steps, rtol=rtol, atol=atol, dtmin=dtmin, normals=normals)
    ts_d = rt.to_dev(ts).reshape(-1)
    n = ts_d.shape[0] - 1
    w0 = tt.cat([tt.cat([pl, vl], dim=1)[:n], tt.cat([pt, vt], dim=1)[:n]]).contiguous()
    t0 = tt.cat([ts_d[:n], ts_d[:n]]).contiguous()
    t1 = ts_d[-1].reshape(1).expand(2 * n).contiguous()
    ctrl = rt.make_ctrl(solver, rtol, atol, dtmin, None, max_steps)
    ys, _, _ = rt.orbit_integrate(pot_total, w0, t0, t1, t1.reshape(-1, 1).contiguous(), ctrl, ts_per_orbit=1)
    return rt.out(ys[:n, 0], dev_in), rt.out(ys[n:, 0], dev_in)


# ---- Chen+25 release model and the stream generators built on it (streamhelpers.py:352-651) ----------------------
CHEN25_MEAN = np.array([1.6, -30, 0, 1, 20, 0], dtype=np.float64)                    # streamhelpers.py:362
CHEN25_COV = np.array([[0.1225, 0, 0, 0, -4.9, 0], [0, 529, 0, 0, 0, 0], [0, 0, 144, 0, 0, 0], [0, 0, 0, 0, 0, 0],
                       [-4.9, 0, 0, 0, 400, 0], [0, 0, 0, 0, 0, 484]], dtype=np.float64)  # streamhelpers.py:363-370


def chen25_factor():
    """jax.random.multivariate_normal(..., method='svd'): factor = u * sqrt(s); sample = mean + factor @ z.
    (The column signs of u are a LAPACK convention; a flipped column flips the sign of one standard normal, i.e. gives another
    equally valid draw.  Pass `normals=` for bit-reproducible inputs.)"""
    u, s, _ = np.linalg.svd(CHEN25_COV)
    return u * np.sqrt(s)[None, :]


def key_words(key):
    """A jax PRNG key as two uint32 words: an int seed means jax.random.PRNGKey(seed); a length-2 array is used as is."""
    if key is None:
        return None
    if np.ndim(key) == 0:
        s = int(key) & 0xFFFFFFFFFFFFFFFF
        return (s >> 32) & 0xFFFFFFFF, s & 0xFFFFFFFF
    k = np.asarray(key).reshape(-1)
    return int(k[0]) & 0xFFFFFFFF, int(k[1]) & 0xFFFFFFFF


def _chen25_release(pot_base, ws, Ms, ts_d, key, normals):
    nr = None if normals is None else rt.to_dev(normals).reshape(ws.shape[0], 6)
    return rt.release_chen25(pot_base, pot_base._G, ws, Ms, ts_d, key_words(key), CHEN25_MEAN, chen25_factor(), nr)


def gen_stream_ics_Chen25(pot_base=None, ts=None, prog_w0=None, Msat=None, key=None, solver=Dopri5(scan_kind='bounded'), rtol=1e-7, atol=1e-7,
                          dtmin=0.3, dtmax=None, max_steps=10_000, normals=None, _pot_orbit=None):
    """([pos_lead, pos_trail, vel_lead, vel_trail], progenitor Solution) (streamhelpers.py:434-459)."""
    from .main import Solution
    dev_in = rt.is_dev(ts)
    ts_d, w0_d, Ms, _, _ = pot_base._stream_inputs(ts, prog_w0, Msat, 1.0, None)
    ctrl = rt.make_ctrl(solver, rtol, atol, dtmin, dtmax, max_steps)
    ws, status, nsteps, _ = rt.orbit_dense(pot_base if _pot_orbit is None else _pot_orbit, w0_d, float(ts_d.min()), float(ts_d.max()), ts_d, ctrl)
    outs = _chen25_release(pot_base, ws, Ms, ts_d, key, normals)
    return [rt.out(o, dev_in) for o in outs], Solution(rt.out(ts_d, dev_in), rt.out(ws, dev_in), rt.out(status[0], dev_in), nsteps)


def gen_stream_ics_pert_Chen25(pot_base=None, pot_pert=None, ts=None, prog_w0=None, Msat=None, key=None, solver=Dopri5(scan_kind='bounded'),
                               rtol=1e-7, atol=1e-7, dtmin=0.3, dtmax=None, max_steps=10_000, normals=None):
    """Progenitor in base+pert, release computed in the smooth base potential (streamhelpers.py:461-489)."""
    pot_total = Potential_Combine(potential_list=[pot_base, pot_pert], units=usys)
    return gen_stream_ics_Chen25(pot_base=pot_base, ts=ts, prog_w0=prog_w0, Msat=Msat, key=key, solver=solver, rtol=rtol, atol=atol, dtmin=dtmin,
                                 dtmax=dtmax, max_steps=max_steps, normals=normals, _pot_orbit=pot_total)


def _chen25_integrate(pot_list, prog_pot, ics, orb, ts, solver, rtol, atol, dtmin, dtmax, max_steps, throw=False):
    """Particles from ts[i] to ts[-1] in pot_list + the progenitor's own potential on its cubic track (streamhelpers.py:520-545)."""
    from .potential import CubicTrack, TimeDepTranslatingPotential
    tt = rt.torch()
    dev_in = rt.is_dev(ts)
    ts_d = rt.to_dev(ts).reshape(-1)
    pots = list(pot_list)
    if prog_pot is not None and getattr(prog_pot, "m", 0.0) != 0.0:      # m = 0 is the reference's "no progenitor" default: contributes exactly 0
        track = CubicTrack(np.asarray(orb.ts.cpu() if hasattr(orb.ts, "cpu") else orb.ts), np.asarray(orb.ys.cpu() if hasattr(orb.ys, "cpu") else orb.ys)[:, :3])
        pots.append(TimeDepTranslatingPotential(pot=prog_pot, center_spl=track, units=usys))
    pot_tot = Potential_Combine(potential_list=pots, units=usys)
    pl, pt, vl, vt = [rt.to_dev(a) for a in ics]
    n = ts_d.shape[0] - 1
    w0 = tt.cat([tt.cat([pl, vl], dim=1)[:n], tt.cat([pt, vt], dim=1)[:n]]).contiguous()
    t0 = tt.cat([ts_d[:n], ts_d[:n]]).contiguous()
    t1 = ts_d[-1].reshape(1).expand(2 * n).contiguous()
    ctrl = rt.make_ctrl(solver, rtol, atol, dtmin, dtmax, max_steps)
    ys, status, _ = rt.orbit_integrate(pot_tot, w0, t0, t1, t1.reshape(-1, 1), ctrl, ts_per_orbit=1)
    if throw and bool((status != 0).any()):
        raise RuntimeError("Chen25 stream: an orbit failed (max_steps reached or non-finite state)")
    return rt.out(ys[:n, 0], dev_in), rt.out(ys[n:, 0], dev_in)


def gen_stream_vmapped_Chen25(pot_base, ts, prog_w0, Msat, key, solver=Dopri5(scan_kind='bounded'), rtol=1e-7, atol=1e-7, dtmin=0.3, dtmax=None,
                              max_steps=10_000, throw=False, prog_pot=None, normals=None):
    """streamhelpers.py:491-545: Chen25 release, particles in pot_base + the progenitor's own (moving) potential."""
    ics, orb = gen_stream_ics_Chen25(pot_base=pot_base, ts=ts, prog_w0=prog_w0, Msat=Msat, key=key, solver=solver, rtol=rtol, atol=atol, dtmin=dtmin,
                                     dtmax=dtmax, max_steps=max_steps, normals=normals)
    return _chen25_integrate([pot_base], prog_pot, ics, orb, ts, solver, rtol, atol, dtmin, dtmax, max_steps, throw)


def gen_stream_vmapped_with_pert_Chen25(pot_base=None, pot_pert=None, prog_pot=None, ts=None, prog_w0=None, Msat=None, key=None,
                                        solver=Dopri5(scan_kind='bounded'), max_steps=10_000, rtol=1e-7, atol=1e-7, dtmin=0.1, normals=None):
    """streamhelpers.py:546-598: progenitor and particles in base + pert."""
    ics, orb = gen_stream_ics_pert_Chen25(pot_base=pot_base, pot_pert=pot_pert, ts=ts, prog_w0=prog_w0, Msat=Msat, key=key, solver=solver,
                                          max_steps=max_steps, rtol=rtol, atol=atol, dtmin=dtmin, normals=normals)
    return _chen25_integrate([pot_base, pot_pert], prog_pot, ics, orb, ts, solver, rtol, atol, dtmin, None, max_steps)


def gen_stream_vmapped_with_pert_Chen25_fixed_prog(pot_base=None, pot_pert=None, prog_pot=None, ts=None, prog_w0=None, Msat=None, key=None,
                                                   solver=Dopri5(scan_kind='bounded'), max_steps=10_000, rtol=1e-7, atol=1e-7, dtmin=0.1, normals=None):
    """streamhelpers.py:601-651: unperturbed progenitor orbit, particles in base + pert."""
    ics, orb = gen_stream_ics_Chen25(pot_base=pot_base, ts=ts, prog_w0=prog_w0, Msat=Msat, key=key, solver=solver, max_steps=max_steps, rtol=rtol,
                                     atol=atol, dtmin=dtmin, normals=normals)
    return _chen25_integrate([pot_base, pot_pert], prog_pot, ics, orb, ts, solver, rtol, atol, dtmin, None, max_steps)
