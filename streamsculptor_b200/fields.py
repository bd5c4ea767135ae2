"""fields.py of the reference (/root/reference/streamsculptor/fields.py): integration along non-Hamiltonian fields.

On the B200 hot path a "field" is one of a closed set the CUDA kernels implement:
  hamiltonian_field             -> K1 orbit kernel        (fields.py:101-113)
  MassRadiusPerturbation_OTF    -> K3 response kernel     (fields.py:159-206)
  MassRadiusPerturbation_OTF_SecondOrder -> K4           (fields.py:260-320)
  Nbody_field                   -> K6 one-CTA N-body kernel (fields.py:115-155)
  RestrictedNbody_generator     -> K5 shared-step tracer kernels (RestrictedNbody.py:93-106)
Arbitrary Python `term` functions (CustomField, ...) raise NotImplementedError: no CPU fallback.
"""
import numpy as np

from . import _runtime as rt
from .main import Solution
from .solvers import Dopri8


class hamiltonian_field:
    def __init__(self, pot):
        self.pot = pot

    def term(self, t, xv, args=None):
        return self.pot.velocity_acceleration(t, xv, args)


class Nbody_field:
    """Softened all-pairs self gravity + external potential; state (N,6) is ONE ODE (fields.py:115-155)."""

    def __init__(self, ext_pot=None, masses=None, units=None, eps=1e-3):
        from .units import resolve_G, usys
        self.ext_pot = ext_pot                       # None == the reference's zero-mass Plummer placeholder (fields.py:127-128)
        self.masses = np.ascontiguousarray(np.asarray(masses, dtype=np.float64).reshape(-1))
        self._G = resolve_G(usys if units is None else units)
        self.eps = float(eps)
        self._masses_dev = None

    def masses_dev(self):
        if self._masses_dev is None:
            self._masses_dev = rt.to_dev(self.masses)
        return self._masses_dev

    def term(self, t, xv, args=None):
        dev_in = rt.is_dev(xv)
        return rt.out(rt.nbody_term(self.ext_pot, self.masses_dev(), self._G, self.eps, t, xv), dev_in)


class MassRadiusPerturbation_OTF:
    """coords = [w(6), D(nSH,12)] with D rows [dx/deps(3), dv/deps(3), d2x/dtheta deps(3), d2v/dtheta deps(3)]."""

    def __init__(self, perturbation_generator):
        self.pertgen = perturbation_generator

    def term(self, t, coords, args=None):
        w, D = np.asarray(coords[0], dtype=np.float64), np.asarray(coords[1], dtype=np.float64)
        y = np.concatenate([w.reshape(6), D.reshape(-1)])
        dy = rt.response_term(self.pertgen.potential_base_total, self.pertgen.subhalo_arrays, t, y).cpu().numpy()
        return [dy[:6], dy[6:].reshape(D.shape)]


class MassRadiusPerturbation_OTF_SecondOrder:
    """coords = [w(6), D(nSH,12), E(nSH,6)], E = second-order mass derivative (x2, v2) (fields.py:260-320)."""

    def __init__(self, perturbation_generator):
        self.pertgen = perturbation_generator

    def term(self, t, coords, args=None):
        w, D, E = [np.asarray(c, dtype=np.float64) for c in coords]
        y = np.concatenate([w.reshape(6), D.reshape(-1), E.reshape(-1)])
        dy = rt.second_order_term(self.pertgen.potential_base_total, self.pertgen.subhalo_arrays, t, y).cpu().numpy()
        n = D.shape[0]
        return [dy[:6], dy[6:6 + 12 * n].reshape(n, 12), dy[6 + 12 * n:].reshape(n, 6)]


def _interval(ts, t0, t1, backwards_int):
    """fields.py:58-82: t0 == t1 (default 0.0) means 'derive the interval from ts'."""
    if t0 != t1:
        return float(t0), float(t1)
    lo, hi = float(np.min(ts)), float(np.max(ts))
    return (hi, lo) if backwards_int else (lo, hi)


def integrate_field(w0=None, ts=None, dense=False, solver=Dopri8(scan_kind='bounded'), field=None, args=None, rtol=1e-7, atol=1e-7,
                    dtmin=0.05, dtmax=None, max_steps=1_000, jump_ts=None, backwards_int=False, t0=0.0, t1=0.0, step_ts=None):
    """Integrate a trajectory on a field (fields.py:35-99).  Raises on solver failure like the reference
    (diffrax default throw=True)."""
    if dense or jump_ts is not None or step_ts is not None:
        raise NotImplementedError("dense / jump_ts / step_ts are not on the B200 hot path")
    ts_h = np.asarray(ts.cpu() if hasattr(ts, "cpu") else ts, dtype=np.float64).reshape(-1)
    a, b = _interval(ts_h, t0, t1, backwards_int)
    if isinstance(field, hamiltonian_field):
        sol = field.pot.integrate_orbit(w0=w0, ts=ts, solver=solver, rtol=rtol, atol=atol, dtmin=dtmin, dtmax=dtmax, max_steps=max_steps,
                                        t0=a, t1=b, throw=True)
        return sol
    if isinstance(field, MassRadiusPerturbation_OTF):
        pg = field.pertgen
        ctrl = rt.make_ctrl(solver, rtol, atol, dtmin, dtmax, max_steps)
        w = rt.to_dev(w0[0]).reshape(6)
        D0h = np.asarray(w0[1].cpu() if hasattr(w0[1], "cpu") else w0[1], dtype=np.float64).reshape(pg.subhalo_arrays.n, 12)
        D0 = rt.to_dev(D0h) if np.any(D0h) else None
        ws, Ds, status, nsteps = rt.linear_response_saveat(pg.potential_base_total, pg.subhalo_arrays, w, D0, rt.to_dev([a]), b, rt.to_dev(ts_h), ctrl)
        if int(status[0]) != 0:     # diffrax default throw=True (fields.py:85-98)
            raise RuntimeError("integrate_field failed: " + ("max_steps reached" if int(status[0]) == 1 else "non-finite state"))
        return Solution(ts_h, [ws.cpu().numpy(), Ds.cpu().numpy()], status[0].cpu().numpy(), nsteps[0])
    if isinstance(field, MassRadiusPerturbation_OTF_SecondOrder):
        pg = field.pertgen
        if len(ts_h) > 2 or (len(ts_h) == 2 and ts_h[0] != a):
            raise NotImplementedError("the second-order kernel keeps the final state only (ts = [t_start, t_end], perturbative.py:764)")
        ctrl = rt.make_ctrl(solver, rtol, atol, dtmin, dtmax, max_steps)
        n = pg.subhalo_arrays.n
        w = rt.to_dev(w0[0]).reshape(1, 6)
        D0, E0 = rt.to_dev(w0[1]).reshape(1, n, 12), rt.to_dev(w0[2]).reshape(1, n, 6)
        wout, Dout, Eout, status, nsteps = rt.second_order_response(pg.potential_base_total, pg.subhalo_arrays, w, D0, E0, rt.to_dev([a]), b, ctrl)
        if int(status[0]) != 0:
            raise RuntimeError("integrate_field failed: " + ("max_steps reached" if int(status[0]) == 1 else "non-finite state"))
        first = len(ts_h) == 2
        stack = lambda y0, y1: np.concatenate([y0.cpu().numpy(), y1.cpu().numpy()]) if first else y1.cpu().numpy()
        return Solution(ts_h, [stack(w, wout), stack(D0, Dout), stack(E0, Eout)], status[0].cpu().numpy(), nsteps[0])
    if isinstance(field, Nbody_field):
        dev_in = rt.is_dev(w0)
        ctrl = rt.make_ctrl(solver, rtol, atol, dtmin, dtmax, max_steps)
        w = rt.to_dev(w0).reshape(-1, 6)
        if w.shape[0] != len(field.masses):
            raise ValueError("Nbody_field: w0 must be (N,6) with N = len(masses)")
        ys, status, nsteps = rt.nbody_integrate(field.ext_pot, field.masses_dev(), field._G, field.eps, w, a, b, rt.to_dev(ts_h), ctrl)
        if int(status[0]) != 0:
            raise RuntimeError("integrate_field failed: " + ("max_steps reached" if int(status[0]) == 1 else "non-finite state"))
        return Solution(ts_h, rt.out(ys, dev_in), status[0].cpu().numpy(), nsteps)
    from .RestrictedNbody import RestrictedNbody_generator
    if isinstance(field, RestrictedNbody_generator):
        dev_in = rt.is_dev(w0)
        if len(ts_h) > 2 or (len(ts_h) == 2 and ts_h[0] != a) or ts_h[-1] != b:
            raise NotImplementedError("the shared-step tracer kernel keeps the final state only (ts = [t_start, t_end], RestrictedNbody.py:131)")
        ctrl = rt.make_ctrl(solver, rtol, atol, dtmin, dtmax, max_steps)
        w = rt.to_dev(w0).reshape(-1, 6)
        wout, status, nsteps = rt.shared_step_orbits(field.potential_total, w, a, b, ctrl)
        if int(status[0]) != 0:
            raise RuntimeError("integrate_field failed: " + ("max_steps reached" if int(status[0]) == 1 else "non-finite state"))
        t = rt.torch()
        ys = t.stack([w, wout]) if len(ts_h) == 2 else wout[None]
        return Solution(ts_h, rt.out(ys, dev_in), status[0].cpu().numpy(), nsteps)
    raise NotImplementedError(f"field {type(field).__name__} is not implemented on the device (closed set: hamiltonian_field, Nbody_field, "
                              "RestrictedNbody_generator, MassRadiusPerturbation_OTF, MassRadiusPerturbation_OTF_SecondOrder)")


def _unsupported(name):
    class _X:
        def __init__(self, *a, **k):
            raise NotImplementedError(f"{name} is outside the B200 hot path (SURVEY.md section 8f)")
    _X.__name__ = name
    return _X


MassRadiusPerturbation_Interp = _unsupported("MassRadiusPerturbation_Interp")
MW_LMC_field = _unsupported("MW_LMC_field")
CustomField = _unsupported("CustomField")
