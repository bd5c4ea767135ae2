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
