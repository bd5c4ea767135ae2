"""`Potential` base class - same names, arguments and error behaviour as the reference's
/root/reference/streamsculptor/main.py, dispatching into libssb200 (sm_100a CUDA) instead of JAX/diffrax.

Differences that are inherent to leaving JAX:
  * results are numpy arrays (or CUDA torch tensors when the inputs were CUDA tensors), not jax Arrays;
  * `solver` is any object whose class is named Dopri5 / Dopri8 (streamsculptor_b200.Dopri5/8 or diffrax's);
  * autodiff THROUGH a solve (adjoint=ForwardMode(), main.py:160) is not available;
  * `release_model`/`gen_stream_*` take an optional `normals=` array ([N,4] standard normals) that replaces the
    jax.random draws of main.py:263-266 (the threefry recipe itself is reproduced when it is omitted).
"""
import numpy as np

from . import _lib
from . import _runtime as rt
from .solvers import Dopri5, Dopri8, solver_id
from .units import dimensionless, resolve_G, usys  # noqa: F401  (usys re-exported like the reference)

DEFAULT_KVALS = (2.0, 0.3, 0.0, 0.0, 0.4, 0.4, 0.5, 0.5)   # main.py:214


class Solution:
    """The subset of diffrax.Solution the reference's callers touch: .ts, .ys, .stats, .result, .evaluate(t)."""

    def __init__(self, ts, ys, status, nsteps, dense=None):
        self.ts, self.ys = ts, ys
        self.result = status                       # 0 successful, 1 max_steps_reached, 2 non-finite
        ns = np.asarray(nsteps.cpu() if hasattr(nsteps, "cpu") else nsteps)
        self.stats = {"num_steps": ns[..., 0], "num_accepted_steps": ns[..., 1], "num_rejected_steps": ns[..., 2]}
        self._dense = dense

    def evaluate(self, t):
        if self._dense is None:
            raise ValueError("Dense solution has not been saved; pass dense=True.")
        return self._dense(t)


class Potential:
    def __init__(self, units, params):                # main.py:22-35
        self.units = dimensionless if units is None else units
        self._G = resolve_G(units)
        for name, param in params.items():
            setattr(self, name, param)

    # ---- lowering hook: subclasses append their components to the potential program ----
    def _lower(self, prog, track):
        raise NotImplementedError(
            f"{type(self).__name__} has no closed-form device implementation (arbitrary Python potentials cannot run in "
            "the CUDA kernels; there is no CPU fallback)")

    # ---- field evaluation (main.py:37-112) ----
    def potential(self, xyz, t):
        phi, = rt.potential_eval(self, xyz, t, ("phi",))
        return phi[0] if np.ndim(xyz) == 1 else phi

    def gradient(self, xyz, t):
        g, = rt.potential_eval(self, xyz, t, ("grad",))
        return g[0] if np.ndim(xyz) == 1 else g

    def acceleration(self, xyz, t):
        return -self.gradient(xyz, t)

    def jacobian_force(self, xyz, t):                 # main.py:59-65: jacfwd(gradient) = Hessian of Phi
        h, = rt.potential_eval(self, xyz, t, ("hess",))
        return h[0] if np.ndim(xyz) == 1 else h

    def density(self, xyz, t):                        # main.py:42-45
        h = self.jacobian_force(xyz, t)
        return (h[..., 0, 0] + h[..., 1, 1] + h[..., 2, 2]) / (4 * np.pi * self._G)

    def local_circular_velocity(self, xyz, t):        # main.py:51-57
        xyz = np.asarray(xyz, dtype=np.float64)
        r = np.sqrt(np.sum(xyz ** 2))
        return np.sqrt(r * np.sum(self.gradient(xyz, t) * xyz / r))

    def dphidr(self, x, t):                           # main.py:67-74
        x = np.asarray(x, dtype=np.float64)
        return np.sum(self.gradient(x, t) * x / np.linalg.norm(x))

    def d2phidr2(self, x, t):                         # main.py:76-85: rhat^T Hess rhat (rhat held fixed)
        x = np.asarray(x, dtype=np.float64)
        rhat = x / np.linalg.norm(x)
        return rhat @ self.jacobian_force(x, t) @ rhat

    def omega(self, x, v):                            # main.py:88-96
        x, v = np.asarray(x, dtype=np.float64), np.asarray(v, dtype=np.float64)
        return np.linalg.norm(np.cross(x, v) / (x[0] ** 2 + x[1] ** 2 + x[2] ** 2))

    def tidalr(self, x, v, Msat, t):                  # main.py:98-104
        return (self._G * Msat / (self.omega(x, v) ** 2 - self.d2phidr2(x, t))) ** (1.0 / 3.0)

    def lagrange_pts(self, x, v, Msat, t):            # main.py:106-112
        x = np.asarray(x, dtype=np.float64)
        r_tidal = self.tidalr(x, v, Msat, t)
        r_hat = x / np.linalg.norm(x)
        return x - r_hat * r_tidal, x + r_hat * r_tidal

    def velocity_acceleration(self, t, xv, args=None):   # main.py:116-120
        xv = np.asarray(xv, dtype=np.float64)
        return np.hstack([xv[3:], -self.gradient(xv[:3], t)])

    # ---- orbit integration (main.py:125-202) ----
    def integrate_orbit(self, w0=None, ts=None, dense=False, solver=Dopri8(scan_kind='bounded'), rtol=1e-7, atol=1e-7, dtmin=0.3,
                        dtmax=None, max_steps=10_000, t0=None, t1=None, steps=False, jump_ts=None, throw=False):
        """Integrate orbit associated with potential function (docstring of main.py:127-137 applies)."""
        if steps:
            raise NotImplementedError("SaveAt(steps=True) is not on the B200 hot path")
        if jump_ts is not None:
            raise NotImplementedError("jump_ts is None on every reference path (perturbative.py:54); not implemented")
        dev_in = rt.is_dev(w0)
        ts_d = rt.to_dev(ts).reshape(-1)
        t0v = float(ts_d.min()) if t0 is None else float(t0)          # main.py:152-153
        t1v = float(ts_d.max()) if t1 is None else float(t1)
        ctrl = rt.make_ctrl(solver, rtol, atol, dtmin, dtmax, max_steps)
        w0_d = rt.to_dev(w0).reshape(6)
        if dense or ts_d.shape[0] > 32:
            ys, status, nsteps, scratch = rt.orbit_dense(self, w0_d, t0v, t1v, ts_d, ctrl)
            dense_fn = None
            if dense:
                dense_fn = _DenseEval(self, scratch, ctrl, w0_d, t0v, t1v)
            sol = Solution(None if dense else rt.out(ts_d, dev_in), None if dense else rt.out(ys, dev_in), rt.out(status[0], dev_in), nsteps,
                           dense=dense_fn)
        else:
            tt = rt.torch()
            t0a = tt.full((1,), t0v, dtype=tt.float64, device=w0_d.device)
            t1a = tt.full((1,), t1v, dtype=tt.float64, device=w0_d.device)
            ys, status, nsteps = rt.orbit_integrate(self, w0_d.reshape(1, 6), t0a, t1a, ts_d, ctrl, ts_per_orbit=0)
            sol = Solution(rt.out(ts_d, dev_in), rt.out(ys[0], dev_in), rt.out(status[0], dev_in), nsteps[0])
        if throw and int(np.asarray(sol.result if not dev_in else sol.result.cpu())) != 0:
            raise RuntimeError("integrate_orbit failed: " + ("max_steps reached" if int(status[0]) == 1 else "non-finite state"))
        return sol

    def integrate_orbit_batch_vmapped(self, w0=None, ts=None, dense=False, solver=Dopri8(scan_kind='bounded'), rtol=1e-7, atol=1e-7,
                                      dtmin=0.3, dtmax=None, max_steps=10_000, t0=None, t1=None, steps=False, jump_ts=None):
        """Batch of orbits, ts either [M] (shared) or [N,M] (main.py:186-202).  t0/t1 may also be [N] arrays."""
        if dense or steps or jump_ts is not None:
            raise NotImplementedError("batched dense / steps / jump_ts output is not on the B200 hot path")
        dev_in = rt.is_dev(w0)
        tt = rt.torch()
        w0_d = rt.to_dev(w0).reshape(-1, 6)
        N = w0_d.shape[0]
        ts_d = rt.to_dev(ts)
        per_orbit = ts_d.dim() == 2
        if per_orbit and ts_d.shape[0] != N:
            raise ValueError("ts must be [M] or [N, M]")

        def ends(val, red):
            if val is not None:
                return rt.to_dev(np.broadcast_to(np.asarray(val, dtype=np.float64), (N,)).copy()) if not rt.is_dev(val) else val.reshape(-1).expand(N).contiguous()
            r = red(ts_d, dim=-1).values if per_orbit else red(ts_d).reshape(1).expand(N)
            return r.contiguous()
        t0a, t1a = ends(t0, tt.min), ends(t1, tt.max)
        ctrl = rt.make_ctrl(solver, rtol, atol, dtmin, dtmax, max_steps)
        ys, status, nsteps = rt.orbit_integrate(self, w0_d, t0a, t1a, ts_d, ctrl, ts_per_orbit=int(per_orbit))
        ts_out = ts_d if per_orbit else ts_d.reshape(1, -1).expand(N, -1)
        return Solution(rt.out(ts_out, dev_in), rt.out(ys, dev_in), rt.out(status, dev_in), nsteps)

    # the reference's scan variant is the same computation scheduled sequentially (main.py:166-183)
    integrate_orbit_batch_scan = integrate_orbit_batch_vmapped

    # ---- stream model (main.py:209-368) ----
    def release_model(self, x=None, v=None, Msat=None, i=None, t=None, seed_num=None, kval_arr=1.0, normals=None):
        kv = DEFAULT_KVALS if np.isscalar(kval_arr) else tuple(np.asarray(kval_arr, dtype=np.float64).reshape(8))   # main.py:212-218
        tt = rt.torch()
        prog = rt.to_dev(np.hstack([np.asarray(x, dtype=np.float64), np.asarray(v, dtype=np.float64)]).reshape(1, 6))
        outs = rt.release_spray(self, self._G, prog, rt.to_dev([float(Msat)]), rt.to_dev(np.array([int(i)]), tt.int64), rt.to_dev([float(t)]),
                                0 if seed_num is None else int(seed_num), kv, None if normals is None else rt.to_dev(normals).reshape(1, 4))
        return tuple(o[0].cpu().numpy() for o in outs)           # pos_lead, pos_trail, v_lead, v_trail

    def release_jacobian(self, prog, Msat, idx, ts, seed_num, kval_arr=1.0, normals=None):
        """jacfwd(release_model) over a batch (perturbative.py:281-296): [N,2,6,6], rows (pos, vel) of lead/trail, columns d/d(x, v)."""
        tt = rt.torch()
        prog_d = rt.to_dev(prog).reshape(-1, 6)
        n = prog_d.shape[0]
        kv = DEFAULT_KVALS if np.isscalar(kval_arr) else tuple(np.asarray(kval_arr, dtype=np.float64).reshape(8))
        Ms = rt.to_dev(np.broadcast_to(np.asarray(Msat, dtype=np.float64), (n,)).copy())
        jac = rt.release_jacobian(self, self._G, prog_d, Ms, rt.to_dev(np.asarray(idx), tt.int64), rt.to_dev(ts).reshape(-1),
                                  0 if seed_num is None else int(seed_num), kv, None if normals is None else rt.to_dev(normals).reshape(n, 4))
        return jac.cpu().numpy()

    def third_derivative(self, xyz, t):
        """d^3 Phi / dx_i dx_j dx_k (the reference obtains it by nested jacfwd, fields.py:278-283)."""
        out = rt.potential_third(self, xyz, t).cpu().numpy()
        return out[0] if np.ndim(xyz) == 1 else out

    def _stream_inputs(self, ts, prog_w0, Msat, kval_arr, normals):
        ts_d = rt.to_dev(ts).reshape(-1)
        n = ts_d.shape[0]
        Ms = rt.to_dev(np.broadcast_to(np.asarray(Msat, dtype=np.float64), (n,)).copy()) if not rt.is_dev(Msat) else (Msat * rt.torch().ones_like(ts_d)).contiguous()
        kv = DEFAULT_KVALS if np.isscalar(kval_arr) else tuple(np.asarray(kval_arr, dtype=np.float64).reshape(8))
        nr = None if normals is None else rt.to_dev(normals).reshape(n, 4)
        return ts_d, rt.to_dev(prog_w0).reshape(6), Ms, kv, nr

    def gen_stream_ics(self, ts=None, prog_w0=None, Msat=None, seed_num=None, solver=Dopri5(scan_kind='bounded'), kval_arr=1.0, rtol=1e-7,
                       atol=1e-7, dtmin=0.3, dtmax=None, max_steps=10_000, normals=None):
        dev_in = rt.is_dev(ts)
        tt = rt.torch()
        ts_d, w0_d, Ms, kv, nr = self._stream_inputs(ts, prog_w0, Msat, kval_arr, normals)
        ctrl = rt.make_ctrl(solver, rtol, atol, dtmin, dtmax, max_steps)
        ws, _, _, _ = rt.orbit_dense(self, w0_d, float(ts_d.min()), float(ts_d.max()), ts_d, ctrl)           # main.py:289
        idx = tt.arange(ts_d.shape[0], dtype=tt.int64, device=ts_d.device)
        outs = rt.release_spray(self, self._G, ws, Ms, idx, ts_d, 0 if seed_num is None else int(seed_num), kv, nr)
        return tuple(rt.out(o, dev_in) for o in outs)

    def gen_stream_vmapped(self, ts=None, prog_w0=None, Msat=None, seed_num=None, solver=Dopri5(scan_kind='bounded'), kval_arr=1.0,
                           rtol=1e-7, atol=1e-7, dtmin=0.3, dtmax=None, max_steps=10_000, throw=False, normals=None, _return_stats=False):
        """Generate stellar stream (main.py:343-368): lead[N-1,6], trail[N-1,6]."""
        dev_in = rt.is_dev(ts)
        ts_d, w0_d, Ms, kv, nr = self._stream_inputs(ts, prog_w0, Msat, kval_arr, normals)
        ctrl = rt.make_ctrl(solver, rtol, atol, dtmin, dtmax, max_steps)
        lead, trail, status, nsteps = rt.gen_stream(self, self, self._G, ts_d, w0_d, Ms, 0 if seed_num is None else int(seed_num), kv, nr, ctrl)
        if throw and bool((status != 0).any()):
            raise RuntimeError("gen_stream_vmapped: an orbit failed (max_steps reached or non-finite state)")
        if _return_stats:
            return rt.out(lead, dev_in), rt.out(trail, dev_in), rt.out(status, dev_in), rt.out(nsteps, dev_in)
        return rt.out(lead, dev_in), rt.out(trail, dev_in)

    def gen_stream_scan(self, ts=None, prog_w0=None, Msat=None, seed_num=None, solver=Dopri5(scan_kind='bounded'), kval_arr=1.0, rtol=1e-7,
                        atol=1e-7, dtmin=0.3, dtmax=None, max_steps=10_000, normals=None):
        """The sequential schedule of gen_stream_vmapped (main.py:312-340).  One quirk of the reference is kept: gen_stream_scan calls
        gen_stream_ics WITHOUT its `solver` (main.py:318), so the progenitor orbit and the release use the default Dopri5 whatever solver
        integrates the particles.  With solver = Dopri5 this is gen_stream_vmapped exactly (one fused enqueue); otherwise the release ICs
        come from a Dopri5 progenitor and the particles are integrated with `solver`."""
        if solver_id(solver) == 5:
            return self.gen_stream_vmapped(ts=ts, prog_w0=prog_w0, Msat=Msat, seed_num=seed_num, solver=solver, kval_arr=kval_arr, rtol=rtol,
                                           atol=atol, dtmin=dtmin, dtmax=dtmax, max_steps=max_steps, normals=normals)
        dev_in = rt.is_dev(ts)
        tt = rt.torch()
        ts_d = rt.to_dev(ts).reshape(-1)
        pl, pt, vl, vt = self.gen_stream_ics(ts=ts_d, prog_w0=prog_w0, Msat=Msat, seed_num=seed_num, solver=Dopri5(scan_kind='bounded'), kval_arr=kval_arr,
                                             rtol=rtol, atol=atol, dtmin=dtmin, dtmax=dtmax, max_steps=max_steps, normals=normals)
        n = ts_d.shape[0] - 1
        w0 = tt.cat([tt.cat([pl, vl], 1)[:n], tt.cat([pt, vt], 1)[:n]]).contiguous()            # lead block, trail block
        t0 = tt.cat([ts_d[:n], ts_d[:n]]).contiguous()
        t1 = ts_d[-1].expand(2 * n).contiguous()
        ctrl = rt.make_ctrl(solver, rtol, atol, dtmin, dtmax, max_steps)
        ys, _, _ = rt.orbit_integrate(self, w0, t0, t1, t1.reshape(-1, 1), ctrl, ts_per_orbit=1)
        return rt.out(ys[:n, 0].contiguous(), dev_in), rt.out(ys[n:, 0].contiguous(), dev_in)

    def gen_stream_vmapped_dense(self, ts=None, prog_w0=None, Msat=None, seed_num=None, solver=Dopri5(scan_kind='bounded'), kval_arr=1.0, rtol=1e-7,
                                 atol=1e-7, dtmin=0.3, dtmax=None, max_steps=10_000, normals=None, rec_cap=None):
        """Dense stream model (main.py:376-430): every particle's orbit from its release time to ts[-1] as a dense interpolant.
        Returns a DenseStream; `streamhelpers.eval_dense_stream(t, dense_stream)` gives (lead, trail) at any time (+inf for particles
        not yet released at t).  rec_cap = record slots per orbit (default: sized from a step-count pre-pass so that no orbit overflows)."""
        tt = rt.torch()
        ts_d, w0_d, Ms, kv, nr = self._stream_inputs(ts, prog_w0, Msat, kval_arr, normals)
        pl, pt, vl, vt = self.gen_stream_ics(ts=ts_d, prog_w0=w0_d, Msat=Ms, seed_num=seed_num, solver=solver, kval_arr=kval_arr, rtol=rtol, atol=atol,
                                             dtmin=dtmin, dtmax=dtmax, max_steps=max_steps, normals=normals)
        n = ts_d.shape[0] - 1
        w0 = tt.cat([tt.cat([pl, vl], 1)[:n], tt.cat([pt, vt], 1)[:n]]).contiguous()            # lead block, trail block
        t0 = tt.cat([ts_d[:n], ts_d[:n]]).contiguous()
        t1 = ts_d[-1].expand(2 * n).contiguous()
        ctrl = rt.make_ctrl(solver, rtol, atol, dtmin, dtmax, max_steps)
        return DenseStream(rt.DenseOrbits(self, w0, t0, t1, ctrl, rec_cap), n)

    gen_stream_scan_dense = gen_stream_vmapped_dense


class DenseStream:
    """What gen_stream_*_dense returns: 2(N-1) dense orbits (lead block, trail block) on the device."""

    def __init__(self, orbits, n):
        self.orbits, self.n = orbits, n

    @property
    def status(self):
        """Per-orbit solver status [2(N-1)] (0 ok, 1 max_steps / record slots exhausted, 2 non-finite): lead block, trail block."""
        return self.orbits.status

    def evaluate(self, t):
        ys = self.orbits.evaluate(t)
        return ys[: self.n], ys[self.n:]

    def evaluate_id(self, t, idx, lead=True):
        """One particle's trajectory at the times t (streamhelpers.eval_dense_stream_id)."""
        tq = np.atleast_1d(np.asarray(t, dtype=np.float64))
        row = int(idx) + (0 if lead else self.n)
        out = np.stack([self.orbits.evaluate(float(v))[row].cpu().numpy() for v in tq])
        return out[0] if np.ndim(t) == 0 else out


class _DenseEval:
    """Solution.evaluate for dense=True: interpolates inside the recorded steps on the device."""

    def __init__(self, pot, scratch, ctrl, w0, t0, t1):
        self.pot, self.scratch, self.ctrl, self.w0, self.t0, self.t1 = pot, scratch, ctrl, w0, t0, t1

    def __call__(self, t):
        tq = rt.to_dev(np.atleast_1d(np.asarray(t, dtype=np.float64)))
        ys = rt.empty((tq.shape[0], 6))
        _lib.check(_lib.lib().ssb_orbit_dense_eval_f64(self.ctrl.solver, rt.ptr(self.scratch), rt.ptr(tq), tq.shape[0], rt.ptr(ys), rt.stream_ptr()))
        res = ys.cpu().numpy()
        return res[0] if np.ndim(t) == 0 else res
