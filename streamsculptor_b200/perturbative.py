"""perturbative.py of the reference: first-order stream response to subhalo impacts.

Implemented on the device: the production path GenerateMassRadiusPerturbation_Chen25.compute_perturbation_OTF
(perturbative.py:661-755) and the custom-base variant (perturbative.py:367-454), both of which integrate
fields.MassRadiusPerturbation_OTF per particle from ts[i] to ts[-1] and keep the final state.
"""
import numpy as np

from . import _runtime as rt
from .main import Potential
from .potential import Potential_Combine, SubhaloLinePotentialCustom_dRadius_fromFunc
from .solvers import Dopri5, Dopri8
from .units import usys


class BaseStreamModel(Potential):
    """Stream model without linear perturbations (perturbative.py:243-296): release ICs, forward progenitor track and the
    Jacobian of the release function w.r.t. the progenitor phase-space position."""

    def __init__(self, potential_base, prog_w0, ts, Msat, seednum, solver, units=None, dense=False, cpu=True, normals=None, **kwargs):
        super().__init__(units, {'potential_base': potential_base, 'prog_w0': prog_w0, 'ts': ts, 'Msat': Msat, 'seednum': seednum,
                                 'solver': solver, 'dense': dense, 'cpu': cpu})
        if dense:
            raise NotImplementedError("dense base streams (perturbative.py:271-276) are not on the B200 hot path")
        self.ts = np.asarray(ts, dtype=np.float64)
        self.solver = Dopri5(scan_kind='bounded') if solver is None else solver
        self._normals = normals
        self.streamICs = potential_base.gen_stream_ics(ts=self.ts, prog_w0=prog_w0, Msat=Msat, seed_num=seednum, solver=self.solver, normals=normals,
                                                       **kwargs)                                                    # perturbative.py:265
        self.IDs = np.arange(len(self.ts))
        self.prog_loc_fwd = np.asarray(potential_base.integrate_orbit(w0=prog_w0, ts=self.ts, t0=self.ts.min(), t1=self.ts.max(), solver=self.solver,
                                                                      **kwargs).ys)                                 # perturbative.py:268
        self.dRel_dIC = self.release_func_jacobian()
        self.stream_interp = None

    def release_func_jacobian(self):
        """[len(ts), 2, 6, 6] (perturbative.py:281-296)."""
        return self.potential_base.release_jacobian(self.prog_loc_fwd, self.Msat, self.IDs, self.ts, self.seednum, normals=self._normals)


class GenerateMassRadiusPerturbation(Potential):
    """Classic lead/trail perturbation generator (perturbative.py:24-221): the progenitor's own response is integrated backwards
    and mapped through the release Jacobian into the particles' perturbation ICs."""

    def __init__(self, potential_base, potential_perturbation, potential_structural=None, BaseStreamModel=None, units=None, **kwargs):
        super().__init__(units, {'potential_base': potential_base, 'potential_perturbation': potential_perturbation,
                                 'potential_structural': potential_structural, 'BaseStreamModel': BaseStreamModel})
        from . import fields
        self.gradient = None
        self.potential_base_total = potential_base
        self.subhalo_arrays = potential_perturbation._arrays
        self.jump_ts = None
        if BaseStreamModel is not None:
            self.base_stream = BaseStreamModel
            self.num_pert = potential_perturbation._arrays.n
            self.field_wobs = [BaseStreamModel.prog_loc_fwd[-1], np.zeros((self.num_pert, 12))]                       # perturbative.py:53
            flipped_times = np.flip(BaseStreamModel.ts)
            prog_fieldICs = fields.integrate_field(w0=self.field_wobs, ts=flipped_times, field=fields.MassRadiusPerturbation_OTF(self),
                                                   backwards_int=True, **kwargs)                                     # perturbative.py:57
            self.prog_base = prog_fieldICs
            self.prog_fieldICs = np.flipud(prog_fieldICs.ys[1])
            self.perturbation_ICs_lead, self.perturbation_ICs_trail = self.compute_perturbation_ICs()
            s = self.base_stream.streamICs
            self.base_realspace_ICs_lead = np.hstack([s[0], s[2]])
            self.base_realspace_ICs_trail = np.hstack([s[1], s[3]])

    def compute_base_stream(self, cpu=True):          # perturbative.py:68-81
        b = self.base_stream
        return self.potential_base.gen_stream_vmapped(ts=b.ts, prog_w0=b.prog_w0, Msat=b.Msat, seed_num=b.seednum, solver=b.solver, normals=b._normals)

    def compute_perturbation_ICs(self):               # perturbative.py:84-98
        J, F = self.base_stream.dRel_dIC, self.prog_fieldICs
        lead = np.dstack([np.einsum('ijk,ilk->ilj', J[:, 0], F[:, :, :6]), np.einsum('ijk,ilk->ilj', J[:, 0], F[:, :, 6:])])
        trail = np.dstack([np.einsum('ijk,ilk->ilj', J[:, 1], F[:, :, :6]), np.einsum('ijk,ilk->ilj', J[:, 1], F[:, :, 6:])])
        return lead, trail

    def compute_perturbation_OTF(self, cpu=True, solver=Dopri8(scan_kind='bounded'), rtol=1e-6, atol=1e-6, dtmin=0.05, dtmax=None, max_steps=10_000):
        """(lead_and_derivs, trail_and_derivs), each [w (N-1,6), D (N-1,N_sh,12)] (perturbative.py:101-135)."""
        ts = self.base_stream.ts
        n = len(ts) - 1
        ctrl = rt.make_ctrl(solver, rtol, atol, dtmin, dtmax, max_steps)
        outs = []
        for w0, D0 in ((self.base_realspace_ICs_lead, self.perturbation_ICs_lead), (self.base_realspace_ICs_trail, self.perturbation_ICs_trail)):
            wout, Dout, status, nsteps = rt.linear_response(self.potential_base_total, self.subhalo_arrays, rt.to_dev(w0[:n]), rt.to_dev(D0[:n]),
                                                            rt.to_dev(ts[:n]), float(ts[-1]), ctrl)
            if bool((status != 0).any()):
                raise RuntimeError("compute_perturbation_OTF: a particle failed (max_steps reached or non-finite state)")
            outs.append([wout.cpu().numpy(), Dout.cpu().numpy()])
        return outs[0], outs[1]


class CustomBaseStreamModel(Potential):
    """Stream model with user-supplied release offsets (perturbative.py:299-364)."""

    def __init__(self, potential_base=None, prog_w0=None, ts=None, pos_rel=None, vel_rel=None, solver=Dopri5(), units=None, dense=False,
                 cpu=True, **kwargs):
        super().__init__(units, {'potential_base': potential_base, 'prog_w0': prog_w0, 'ts': ts, 'pos_rel': pos_rel, 'vel_rel': vel_rel,
                                 'solver': solver, 'dense': dense, 'cpu': cpu})
        if dense:
            raise NotImplementedError("dense base streams (perturbative.py:331-333) are not on the B200 hot path")
        self.ts = np.asarray(ts, dtype=np.float64)
        self.solver = Dopri5(scan_kind='bounded') if solver is None else solver
        self.prog_at_ts = np.asarray(potential_base.integrate_orbit(w0=prog_w0, ts=self.ts, t0=self.ts.min(), t1=self.ts.max(),
                                                                    solver=self.solver, **kwargs).ys)          # perturbative.py:317
        self.prog_loc_fwd = self.prog_at_ts
        self.streamICs = np.hstack([self.prog_at_ts[:, :3] + np.asarray(pos_rel), self.prog_at_ts[:, 3:] + np.asarray(vel_rel)])
        self.IDs = np.arange(len(self.ts))
        # custom_release_model is additive, so its Jacobian w.r.t. the progenitor state is the identity (perturbative.py:349-364)
        self.dRel_dIC = np.broadcast_to(np.eye(6), (len(self.ts), 6, 6)).copy()
        self.stream_interp = None

    def gen_stream(self):                             # perturbative.py:340-345
        sol = self.potential_base.integrate_orbit_batch_vmapped(w0=self.streamICs[:-1], ts=np.full((len(self.ts) - 1, 1), self.ts[-1]),
                                                                t0=self.ts[:-1], t1=self.ts[-1], solver=self.solver)
        return sol


class _ResponseGenerator(Potential):
    """Shared implementation of compute_perturbation_OTF for the custom-base / Chen25 generators."""

    def _init_common(self, potential_base_total, potential_perturbation):
        self.gradient = None
        self.potential_base_total = potential_base_total
        self.num_pert = potential_perturbation._arrays.n
        self.subhalo_arrays = potential_perturbation._arrays
        self.jump_ts = None

    def compute_perturbation_OTF(self, cpu=True, solver=Dopri8(scan_kind='bounded'), rtol=1e-6, atol=1e-6, dtmin=0.05, max_steps=10_000,
                                 dtmax=None):
        """Returns [w (N-1,6), D (N-1,N_sh,12)] like the reference (perturbative.py:425-454, 726-755).  `cpu` only chose
        scan vs vmap in the reference; both schedules give the same result and there is one device schedule here."""
        ts = np.asarray(self.base_stream.ts, dtype=np.float64)
        n = len(ts) - 1
        w0 = rt.to_dev(np.asarray(self.base_realspace_ICs)[:n])
        pics = np.asarray(self.perturbation_ICs)[:n]
        D0 = None if not np.any(pics) else rt.to_dev(pics)
        ctrl = rt.make_ctrl(solver, rtol, atol, dtmin, dtmax, max_steps)
        wout, Dout, status, nsteps = rt.linear_response(self.potential_base_total, self.subhalo_arrays, w0, D0, rt.to_dev(ts[:n]), float(ts[-1]), ctrl)
        self.last_status, self.last_nsteps = status.cpu().numpy(), nsteps.cpu().numpy()
        if (self.last_status != 0).any():     # integrate_field leaves diffrax's throw=True (fields.py:85-98)
            raise RuntimeError("compute_perturbation_OTF: a particle failed (max_steps reached or non-finite state)")
        return [wout.cpu().numpy(), Dout.cpu().numpy()]


class GenerateMassRadiusPerturbation_CustomBase(_ResponseGenerator):      # perturbative.py:367-454
    def __init__(self, potential_base, potential_perturbation, potential_structural=None, BaseStreamModel=None, units=None,
                 perturbation_ICs=None, **kwargs):
        super().__init__(units, {'potential_base': potential_base, 'potential_perturbation': potential_perturbation,
                                 'potential_structural': potential_structural, 'BaseStreamModel': BaseStreamModel})
        self._init_common(potential_base, potential_perturbation)
        self.base_stream = BaseStreamModel
        self.base_realspace_ICs = self.base_stream.streamICs
        if perturbation_ICs is None:
            # perturbative.py:388-399: the progenitor's own response, integrated backwards from the observed position and saved at every
            # stripping time (ssb_linear_response_saveat_f64), mapped through the release Jacobian
            from . import fields
            self.field_wobs = [np.asarray(BaseStreamModel.prog_loc_fwd)[-1], np.zeros((self.num_pert, 12))]
            flipped_times = np.flip(np.asarray(BaseStreamModel.ts, dtype=np.float64))
            prog_fieldICs = fields.integrate_field(w0=self.field_wobs, ts=flipped_times, field=fields.MassRadiusPerturbation_OTF(self),
                                                   backwards_int=True, **kwargs)
            self.prog_base = prog_fieldICs
            self.prog_fieldICs = np.flipud(prog_fieldICs.ys[1])
            perturbation_ICs = self.compute_perturbation_ICs()
        self.perturbation_ICs = perturbation_ICs

    def compute_base_stream(self, cpu=True):          # perturbative.py:405-410
        return self.base_stream.gen_stream()

    def compute_perturbation_ICs(self):               # perturbative.py:413-422: [N_star, N_sh, 12]
        J, F = np.asarray(self.base_stream.dRel_dIC), self.prog_fieldICs
        return np.dstack([np.einsum('ijk,ilk->ilj', J, F[:, :, :6]), np.einsum('ijk,ilk->ilj', J, F[:, :, 6:])])


class GenerateMassRadiusPerturbation_CustomBase_SecondOrder(_ResponseGenerator):      # perturbative.py:491-585
    """Custom-base perturbation generator up to SECOND order in the subhalo mass (+ first-order radius response): state per particle
    [w(6), D(N_sh,12), E(N_sh,6)], field MassRadiusPerturbation_OTF_SecondOrder (fields.py:260-320).

    Perturbation ICs (perturbative.py:519-552): the progenitor's own first- and second-order response, integrated BACKWARDS from the observed
    position, evaluated at every stripping time and mapped through the release Jacobian.  The reference obtains all stripping times from ONE
    backward solve with SaveAt(ts); here every stripping time is its own backward solve ending at ts[i] (N copies of the progenitor in one
    launch of the second-order kernel, per-particle end times) - the same quantities to within the solver tolerance, without the dense
    output of a 6 + 18 N_sh dimensional state."""

    def __init__(self, potential_base, potential_perturbation, potential_structural=None, BaseStreamModel=None, units=None, solver=Dopri8(scan_kind='bounded'),
                 rtol=1e-7, atol=1e-7, dtmin=0.05, dtmax=None, max_steps=1_000, **kwargs):
        super().__init__(units, {'potential_base': potential_base, 'potential_perturbation': potential_perturbation,
                                 'potential_structural': potential_structural, 'BaseStreamModel': BaseStreamModel})
        self._init_common(potential_base, potential_perturbation)
        self.base_stream = BaseStreamModel
        ts = np.asarray(BaseStreamModel.ts, dtype=np.float64)
        n, nsh = len(ts), self.num_pert
        w_obs = np.asarray(BaseStreamModel.prog_loc_fwd)[-1]
        self.field_wobs = [w_obs, np.zeros((nsh, 12)), np.zeros((nsh, 6))]                              # perturbative.py:515
        ctrl = rt.make_ctrl(solver, rtol, atol, dtmin, dtmax, max_steps)
        w0 = rt.to_dev(np.broadcast_to(w_obs, (n - 1, 6)).copy())
        t_start = rt.to_dev(np.full(n - 1, ts[-1]))
        _, D, E, status, _ = rt.second_order_response(self.potential_base_total, self.subhalo_arrays, w0, None, None, t_start, rt.to_dev(ts[:-1].copy()), ctrl)
        if bool((status != 0).any()):
            raise RuntimeError("backward integration of the progenitor's second-order field failed (max_steps reached or non-finite state)")
        # the last stripping time is the start of the backward solve: its row is the initial (zero) field
        self.prog_fieldICs_first_order_mass = np.concatenate([D.cpu().numpy(), np.zeros((1, nsh, 12))])   # [N, N_sh, 12], time order of ts
        self.prog_fieldICs_second_order_mass = np.concatenate([E.cpu().numpy(), np.zeros((1, nsh, 6))])   # [N, N_sh, 6]
        self.perturbation_ICs = self.compute_perturbation_ICs()
        self.base_realspace_ICs = self.base_stream.streamICs

    def compute_base_stream(self, cpu=False):         # perturbative.py:530-535
        return self.base_stream.gen_stream()

    def compute_perturbation_ICs(self):               # perturbative.py:539-552: [N x N_sh x 12, N x N_sh x 6]
        J = np.asarray(self.base_stream.dRel_dIC)
        F1, F2 = self.prog_fieldICs_first_order_mass, self.prog_fieldICs_second_order_mass
        first = np.dstack([np.einsum('ijk,ilk->ilj', J, F1[:, :, :6]), np.einsum('ijk,ilk->ilj', J, F1[:, :, 6:])])
        return [first, np.einsum('ijk,ilk->ilj', J, F2)]

    def compute_perturbation_OTF(self, cpu=False, solver=Dopri8(scan_kind='bounded'), rtol=1e-8, atol=1e-8, dtmin=0.05, max_steps=10_000, dtmax=None):
        """[w (N-1,6), D (N-1,N_sh,12), E (N-1,N_sh,6)] (perturbative.py:556-585)."""
        ts = np.asarray(self.base_stream.ts, dtype=np.float64)
        n = len(ts) - 1
        ctrl = rt.make_ctrl(solver, rtol, atol, dtmin, dtmax, max_steps)
        D0, E0 = np.asarray(self.perturbation_ICs[0])[:n], np.asarray(self.perturbation_ICs[1])[:n]
        wout, Dout, Eout, status, nsteps = rt.second_order_response(self.potential_base_total, self.subhalo_arrays,
                                                                    rt.to_dev(np.asarray(self.base_realspace_ICs)[:n]), rt.to_dev(D0) if np.any(D0) else None,
                                                                    rt.to_dev(E0) if np.any(E0) else None, rt.to_dev(ts[:n]), float(ts[-1]), ctrl)
        self.last_status, self.last_nsteps = status.cpu().numpy(), nsteps.cpu().numpy()
        if (self.last_status != 0).any():
            raise RuntimeError("compute_perturbation_OTF: a particle failed (max_steps reached or non-finite state)")
        return [wout.cpu().numpy(), Dout.cpu().numpy(), Eout.cpu().numpy()]


class BaseStreamModelChen25(Potential):               # perturbative.py:588-658
    """Chen+25 base model (perturbative.py:588-658).  `key`: int seed (= jax.random.PRNGKey(seed)) or the two key words.
    Optionally the release can be supplied instead of drawn: stream_ics = (pos_lead, pos_trail, vel_lead, vel_trail) each [N,3]
    and prog_fwd [N,6] (the progenitor at ts)."""

    def __init__(self, pot_base, ts, prog_w0, Msat=None, key=None, solver=Dopri5(scan_kind='bounded'), rtol=1e-7, atol=1e-7, dtmin=0.3,
                 dtmax=None, max_steps=10_000, throw=False, prog_pot=None, units=usys, stream_ics=None, prog_fwd=None):
        super().__init__(units, {'pot_base': pot_base, 'ts': ts, 'prog_w0': prog_w0, 'Msat': Msat, 'key': key, 'solver': solver, 'rtol': rtol,
                                 'atol': atol, 'dtmin': dtmin, 'dtmax': dtmax, 'max_steps': max_steps, 'throw': throw, 'prog_pot': prog_pot})
        from .potential import CubicTrack, TimeDepTranslatingPotential
        from .streamhelpers import gen_stream_ics_Chen25
        ts = np.asarray(ts, dtype=np.float64)
        if stream_ics is None:                                                             # perturbative.py:615-625
            stream_ics, orb_fwd = gen_stream_ics_Chen25(pot_base=pot_base, ts=ts, prog_w0=prog_w0, Msat=Msat, key=key, solver=solver, rtol=rtol,
                                                        atol=atol, dtmin=dtmin, dtmax=dtmax, max_steps=max_steps)
            prog_fwd = np.asarray(orb_fwd.ys)
        if prog_fwd is None:
            prog_fwd = np.asarray(pot_base.integrate_orbit(w0=prog_w0, ts=ts, solver=solver, rtol=rtol, atol=atol, dtmin=dtmin, dtmax=dtmax,
                                                           max_steps=max_steps).ys)
        pl, pt, vl, vt = [np.asarray(a, dtype=np.float64) for a in stream_ics]
        pos_rel = np.vstack([pl - prog_fwd[:, :3], pt - prog_fwd[:, :3]])                  # perturbative.py:627-633
        vel_rel = np.vstack([vl - prog_fwd[:, 3:], vt - prog_fwd[:, 3:]])
        ts_stack = np.clip(np.hstack([ts, ts + 1e-12]), ts.min(), ts.max())                # perturbative.py:635-639
        order = np.argsort(ts_stack, kind="stable")
        pot_tot = pot_base
        if prog_pot is not None and getattr(prog_pot, "m", 0.0) != 0.0:
            pot_tot = Potential_Combine([pot_base, TimeDepTranslatingPotential(pot=prog_pot, center_spl=CubicTrack(ts, prog_fwd[:, :3]), units=usys)],
                                        units=usys)                                        # perturbative.py:642-644
        self.pot_tot = pot_tot
        self.BaseModel = CustomBaseStreamModel(potential_base=pot_tot, prog_w0=prog_w0, ts=ts_stack[order], pos_rel=pos_rel[order],
                                               vel_rel=vel_rel[order], solver=solver, units=usys, dense=False, cpu=False)


class GenerateMassRadiusPerturbation_Chen25(_ResponseGenerator):          # perturbative.py:661-755
    def __init__(self, potential_base, potential_perturbation, BaseStreamModel, units=None, **kwargs):
        super().__init__(units, {'potential_base': potential_base, 'potential_perturbation': potential_perturbation,
                                 'BaseStreamModel': BaseStreamModel})
        self._init_common(BaseStreamModel.pot_tot, potential_perturbation)               # perturbative.py:677
        self.potential_structural = SubhaloLinePotentialCustom_dRadius_fromFunc(
            func=potential_perturbation.func, m=potential_perturbation.m, r_s=potential_perturbation.r_s,
            subhalo_x0=potential_perturbation.subhalo_x0, subhalo_v=potential_perturbation.subhalo_v,
            subhalo_t0=potential_perturbation.subhalo_t0, t_window=potential_perturbation.t_window, units=potential_perturbation.units)
        self.base_stream = BaseStreamModel.BaseModel
        self.perturbation_ICs = np.zeros((len(self.base_stream.ts), self.num_pert, 12))   # perturbative.py:712
        self.base_realspace_ICs = self.base_stream.streamICs

    def gradientPotentialPerturbation_per_SH(self, xyz, t):
        return self.potential_perturbation.gradient_per_SH(xyz, t)

    def gradientPotentialStructural_per_SH(self, xyz, t):
        return self.potential_structural.gradient_per_SH(xyz, t)

    def compute_base_stream(self, cpu=True):
        return self.base_stream.gen_stream()

    def run_nonlinear_sim(self, pot_pert=None, solver=Dopri8(scan_kind='bounded'), rtol=1e-6, atol=1e-6, dtmin=0.05, max_steps=10_000):
        """The full non-linear simulation of the stream in base + perturbation, without perturbation theory (perturbative.py:775-813):
        gen_stream_vmapped_with_pert_Chen25_fixed_prog with this model's stream parameters; returns vstack([lead, trail]) [2(N-1), 6]."""
        from .streamhelpers import gen_stream_vmapped_with_pert_Chen25_fixed_prog
        if pot_pert is None:
            pot_pert = self.potential_perturbation
        bm = self.BaseStreamModel
        lead, trail = gen_stream_vmapped_with_pert_Chen25_fixed_prog(pot_base=self.potential_base, pot_pert=pot_pert, prog_pot=bm.prog_pot,
                                                                     prog_w0=self.base_stream.prog_w0, ts=bm.ts, key=bm.key, Msat=bm.Msat,
                                                                     max_steps=max_steps, atol=atol, rtol=rtol, solver=solver, dtmin=dtmin)
        return np.vstack([np.asarray(lead), np.asarray(trail)])

    def compute_perturbation_second_order_OTF(self, cpu=True, solver=Dopri8(scan_kind='bounded'), rtol=1e-6, atol=1e-6, dtmin=0.05, max_steps=10_000,
                                              dtmax=None):
        """[w (N-1,6), D (N-1,N_sh,12), E (N-1,N_sh,6)] (perturbative.py:757-772); zero perturbation ICs (perturbative.py:715)."""
        ts = np.asarray(self.base_stream.ts, dtype=np.float64)
        n = len(ts) - 1
        ctrl = rt.make_ctrl(solver, rtol, atol, dtmin, dtmax, max_steps)
        wout, Dout, Eout, status, nsteps = rt.second_order_response(self.potential_base_total, self.subhalo_arrays,
                                                                    rt.to_dev(np.asarray(self.base_realspace_ICs)[:n]), None, None, rt.to_dev(ts[:n]),
                                                                    float(ts[-1]), ctrl)
        self.last_status, self.last_nsteps = status.cpu().numpy(), nsteps.cpu().numpy()
        if (self.last_status != 0).any():
            raise RuntimeError("compute_perturbation_second_order_OTF: a particle failed (max_steps reached or non-finite state)")
        return [wout.cpu().numpy(), Dout.cpu().numpy(), Eout.cpu().numpy()]
