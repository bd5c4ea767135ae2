"""RestrictedNbody.py of the reference (/root/reference/streamsculptor/RestrictedNbody.py): a stream modelled as tracers in an
external potential plus the progenitor's fitted monopole on its interpolated orbit.

Hot path: `RestrictedNbody_generator.term` (RestrictedNbody.py:93-106) and its integration by
`integrate_field` with ALL N tracers as one ODE state (RestrictedNbody.py:131) -> K5 shared-step kernels (csrc/ssb_shared.cu).
The monopole re-fit at the interrupt times (RestrictedNbody.py:51-89) is a two-scalar host-side optimisation in the reference
(jaxopt.OptaxSolver(optax.adam)); here it is the same Adam recipe on two scalars with the densities evaluated on the device
and the gradient of the cost taken by central differences instead of autodiff.
"""
import numpy as np

from . import _runtime as rt
from . import potential as _pot
from .fields import integrate_field
from .solvers import Dopri8


_as_prog_track = _pot._track_of_interp      # diffrax dense Solution -> tabulated cubic track (potential.py:152, RestrictedNbody.py:99)


class RestrictedNbody_generator:
    def __init__(self, potential=None, progenitor_potential=None, interp_prog=None, init_mass=None, init_rs=None, r_esc=1.0, maxiter=250, lr=1e-3):
        self.potential = potential
        self.progenitor_potential = progenitor_potential
        self.interp_prog = interp_prog
        self.r_esc = r_esc
        self.lr = lr
        self.maxiter = maxiter
        self.init_mass = init_mass
        self.init_rs = init_rs
        self._track = _as_prog_track(interp_prog)
        self.pot_prog_curr = self.progenitor_potential(m=float(init_mass), r_s=float(init_rs), units=self.potential.units)   # RestrictedNbody.py:49
        # external + progenitor monopole about its interpolated centre (RestrictedNbody.py:98-104) as ONE potential program
        self.potential_total = _pot.Potential_Combine(
            [self.potential, _pot.TimeDepTranslatingPotential(self.pot_prog_curr, self._track, units=self.potential.units)], units=self.potential.units)

    # ---- monopole fit (RestrictedNbody.py:51-89) ----
    def cost_func(self, params, locs, t, inside_bool):
        mass_param, r_s_param = 10.0 ** np.asarray(params, dtype=np.float64)
        pot_prog_curr = self.progenitor_potential(m=float(mass_param), r_s=float(r_s_param), units=self.potential.units)
        density_at_locs = np.asarray(pot_prog_curr.density(locs, t))
        log_density = np.where(inside_bool, np.log(np.where(inside_bool, density_at_locs, 1.0)), 0.0)
        log_like = -mass_param + np.sum(log_density)
        return -log_like

    def fit_monopole(self, x, t, inside_bool, tol=1e-3, fd_step=1e-6):
        """optax.adam(lr) driven by jaxopt.OptaxSolver.run(maxiter): stop when ||grad||_2 <= tol or after maxiter updates."""
        inside_bool = np.asarray(inside_bool, dtype=bool)
        if inside_bool.sum() == 0:                      # dissolved (RestrictedNbody.py:74-77)
            return np.array([0.0, 1.0])
        x = np.asarray(x, dtype=np.float64)
        params = np.array([np.log10(self.init_mass), np.log10(self.init_rs)], dtype=np.float64)
        m1, m2 = np.zeros(2), np.zeros(2)
        b1, b2, eps = 0.9, 0.999, 1e-8
        for it in range(1, int(self.maxiter) + 1):
            g = np.empty(2)
            for k in range(2):
                dp = np.zeros(2); dp[k] = fd_step
                g[k] = (self.cost_func(params + dp, x, t, inside_bool) - self.cost_func(params - dp, x, t, inside_bool)) / (2 * fd_step)
            m1 = b1 * m1 + (1 - b1) * g
            m2 = b2 * m2 + (1 - b2) * g * g
            params = params - self.lr * (m1 / (1 - b1 ** it)) / (np.sqrt(m2 / (1 - b2 ** it)) + eps)
            if np.linalg.norm(g) <= tol:
                break
        return 10.0 ** params

    def get_params(self, t, coords, args=None):
        coords = np.asarray(coords.cpu() if hasattr(coords, "cpu") else coords, dtype=np.float64)
        x = coords[:, :3]
        prog_center = self._track(float(t))
        x_rel = x - prog_center
        r_rel = np.sqrt(np.sum(x_rel ** 2, axis=1))
        inside_bool = r_rel < self.r_esc
        mass_fit, r_s_fit = self.fit_monopole(x_rel, t, inside_bool)
        return mass_fit, r_s_fit

    # ---- the field (RestrictedNbody.py:93-106) ----
    def term(self, t, coords, args=None):
        dev_in = rt.is_dev(coords)
        w = rt.to_dev(coords).reshape(-1, 6)
        n = w.shape[0]
        g, = rt.potential_eval(self.potential_total, w[:, :3].contiguous(), rt.to_dev(np.full(n, float(t))), ("grad",))
        return rt.out(rt.torch().cat([w[:, 3:], -g], dim=1), dev_in)


def initialize_prog_params(w0=None, t0=None, field=None, maxiter=5_000):
    init_state = RestrictedNbody_generator(potential=field.potential, progenitor_potential=field.progenitor_potential, interp_prog=field._track,
                                           r_esc=field.r_esc, init_mass=field.init_mass, init_rs=field.init_rs, maxiter=maxiter)
    return init_state.get_params(t=t0, coords=w0, args=None)


def integrate_restricted_Nbody(w0=None, ts=None, interrupt_ts=None, solver=Dopri8(scan_kind='bounded'), field=None, args=None, rtol=1e-7, atol=1e-7,
                               dtmin=0.05, dtmax=None, maxiter=5, max_steps=1_000, mass_init=None, r_s_init=None):
    """RestrictedNbody.py:120-147: between consecutive interrupt times re-fit the progenitor monopole, then integrate all tracers
    as one ODE.  Returns [tstop[K], mass[K], r_s[K], w_at_tstop[K,N,6]] like the reference's lax.scan outputs."""
    ts = np.asarray(ts, dtype=np.float64).reshape(-1)
    interrupt = np.hstack([np.asarray(interrupt_ts, dtype=np.float64).reshape(-1), ts.max()])
    K = len(interrupt)
    dev_in = rt.is_dev(w0)
    wcurr = rt.to_dev(w0).reshape(-1, 6)
    tcurr, tstop, pm, prs = float(ts[0]), float(interrupt[0]), float(mass_init), float(r_s_init)
    out_t, out_m, out_rs, out_w = [], [], [], []
    for idx in range(K):
        tend = min(tstop, float(ts[-1]))
        curr_state = RestrictedNbody_generator(potential=field.potential, progenitor_potential=field.progenitor_potential, interp_prog=field._track,
                                               r_esc=field.r_esc, init_mass=pm, init_rs=prs, maxiter=maxiter)
        mass_curr, r_s_curr = curr_state.get_params(t=tcurr, coords=wcurr, args=None)
        new_field = RestrictedNbody_generator(potential=field.potential, progenitor_potential=field.progenitor_potential, interp_prog=field._track,
                                              r_esc=field.r_esc, init_mass=mass_curr, init_rs=r_s_curr)
        if tend != tcurr:
            wcurr = rt.to_dev(integrate_field(w0=wcurr, ts=np.array([tcurr, tend]), solver=solver, field=new_field, args=args, rtol=rtol, atol=atol,
                                              dtmin=dtmin, dtmax=dtmax, max_steps=max_steps).ys[-1])
        out_t.append(tstop); out_m.append(mass_curr); out_rs.append(r_s_curr); out_w.append(wcurr)
        tcurr, tstop = tstop, float(interrupt[min(idx + 1, K - 1)])     # jax clamps the out-of-range gather of RestrictedNbody.py:133
        pm, prs = float(mass_curr), float(r_s_curr)
    W = rt.torch().stack(out_w)
    return [np.asarray(out_t), np.asarray(out_m), np.asarray(out_rs), rt.out(W, dev_in)]
