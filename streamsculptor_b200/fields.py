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


class variational_field:
    """Variational (tangent) equations along an unperturbed orbit - the built-in equivalent of the `second_order_field` the reference's
    tutorial wraps in CustomField (examples/higher_order_variationalEqn.ipynb cell 3): coords = [w(6), dw/dw_init (6,6)] (order 1) or
    [w, dw/dw_init, d2w/dw_init^2 (6,6,6)] (order 2).  The tidal tensor and third derivatives are closed forms on the device."""

    def __init__(self, pot, order=2):
        if order not in (1, 2):
            raise ValueError("order must be 1 or 2")
        self.pot, self.order = pot, int(order)

    def _flat(self, coords):
        parts = [np.asarray(c.cpu() if hasattr(c, "cpu") else c, dtype=np.float64).reshape(-1) for c in coords[:1 + self.order]]
        if [len(p) for p in parts] != [6, 36, 216][:1 + self.order]:
            raise ValueError("variational_field: coords must be [w(6), M(6,6)" + (", M2(6,6,6)]" if self.order == 2 else "]"))
        return np.concatenate(parts)

    def term(self, t, coords, args=None):
        dy = rt.variational_term(self.pot, self.order, t, self._flat(coords)).cpu().numpy()
        out = [dy[:6], dy[6:42].reshape(6, 6)]
        if self.order == 2:
            out.append(dy[42:].reshape(6, 6, 6))
        return out


def integrate_variational_batch(pot, w0, t0, t1, order=1, M0=None, M20=None, solver=Dopri8(scan_kind='bounded'), rtol=1e-7, atol=1e-7, dtmin=0.05,
                                dtmax=None, max_steps=10_000):
    """vmap of integrate_field(w0=[w_i, I, 0], t0=t0_i, t1=t1, ts=[t1], field=variational_field(pot, order)) over particles - what the
    tutorial's scan does per arm (higher_order_variationalEqn.ipynb cells 10-11).  Returns (w[N,6], M[N,6,6], M2[N,6,6,6] | None, status[N]).
    M is also the forward-mode Jacobian d integrate_orbit / d w0 (main.py:160)."""
    dev_in = rt.is_dev(w0)
    w = rt.to_dev(w0).reshape(-1, 6)
    N = w.shape[0]
    t0d = rt.to_dev(np.broadcast_to(np.asarray(t0.cpu() if hasattr(t0, "cpu") else t0, dtype=np.float64), (N,)).copy())
    M0d = None if M0 is None else rt.to_dev(M0).reshape(N, 6, 6)
    M20d = None if M20 is None else rt.to_dev(M20).reshape(N, 6, 6, 6)
    ctrl = rt.make_ctrl(solver, rtol, atol, dtmin, dtmax, max_steps)
    wout, Mout, M2out, status, nsteps = rt.variational(pot, order, w, M0d, M20d, t0d, float(t1), ctrl)
    return rt.out(wout, dev_in), rt.out(Mout, dev_in), (None if M2out is None else rt.out(M2out, dev_in)), rt.out(status, dev_in)


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
    if isinstance(field, variational_field):
        if len(ts_h) > 2 or (len(ts_h) == 2 and ts_h[0] != a) or ts_h[-1] != b:
            raise NotImplementedError("the variational kernel keeps the final state only (ts = [t_end] or [t_start, t_end])")
        y0 = field._flat(w0)
        ctrl = rt.make_ctrl(solver, rtol, atol, dtmin, dtmax, max_steps)
        M20 = rt.to_dev(y0[42:]).reshape(1, 6, 6, 6) if field.order == 2 else None
        wout, Mout, M2out, status, nsteps = rt.variational(field.pot, field.order, rt.to_dev(y0[:6]).reshape(1, 6), rt.to_dev(y0[6:42]).reshape(1, 6, 6),
                                                           M20, rt.to_dev([a]), b, ctrl)
        if int(status[0]) != 0:
            raise RuntimeError("integrate_field failed: " + ("max_steps reached" if int(status[0]) == 1 else "non-finite state"))
        first = len(ts_h) == 2
        stack = lambda y_start, y_end: np.concatenate([y_start[None], y_end.cpu().numpy()]) if first else y_end.cpu().numpy()
        ys = [stack(y0[:6], wout), stack(y0[6:42].reshape(6, 6), Mout)]
        if field.order == 2:
            ys.append(stack(y0[42:].reshape(6, 6, 6), M2out))
        return Solution(ts_h, ys, status[0].cpu().numpy(), nsteps[0])
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
