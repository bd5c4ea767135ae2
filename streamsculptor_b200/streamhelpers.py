"""streamhelpers.py of the reference: stream generation with perturbers (streamhelpers.py:56-198)."""
import numpy as np

from . import _runtime as rt
from .main import DEFAULT_KVALS
from .potential import Potential_Combine
from .solvers import Dopri5, Dopri8
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


# ---- dense streams and streaklines (streamhelpers.py:23-53, 656-732) -------------------------------------------------
def eval_dense_stream(t_eval=None, dense_stream=None):
    """(lead, trail) of a dense stream model at time t_eval (streamhelpers.py:23-31); device tensors [N-1, 6]."""
    return dense_stream.evaluate(t_eval)


def eval_dense_stream_id(time=None, interp_func=None, idx=None, lead=True):
    """Trajectory of particle `idx` of a dense stream (streamhelpers.py:33-53)."""
    return interp_func.evaluate_id(time, idx, lead=lead)


def get_Streakline_ICs(pot, prog_w0, Msat, t0, t1, Nstrip, solver=Dopri8(), rtol=1e-6, atol=1e-6):
    """Streakline initial conditions (streamhelpers.py:656-699): particles leave L1 / L2 with the cluster's angular velocity.
    The progenitor orbit and the Hessians come from the device; the O(Nstrip) vector algebra is numpy."""
    ts = np.linspace(t0, t1, Nstrip)
    prog = np.asarray(pot.integrate_orbit(w0=prog_w0, ts=ts, solver=solver, rtol=rtol, atol=atol).ys)
    x, v = prog[:, :3], prog[:, 3:]
    r = np.sqrt(np.sum(x ** 2, axis=1))
    r_hat = x / r[:, None]
    H = np.asarray(pot.jacobian_force(x, ts))                                   # Hessian of Phi at the progenitor (main.py:59-65)
    d2phidr2 = np.einsum("ni,nij,nj->n", r_hat, H, r_hat)                       # main.py:76-85
    omega = np.linalg.norm(np.cross(x, v), axis=1) / r ** 2                     # main.py:88-96
    r_tidal = (pot._G * Msat / (omega ** 2 - d2phidr2)) ** (1.0 / 3.0)          # main.py:98-104
    L_close, L_far = x - r_hat * r_tidal[:, None], x + r_hat * r_tidal[:, None]  # main.py:106-112
    v_hat = v / np.linalg.norm(v, axis=1)[:, None]
    sintheta = np.linalg.norm(np.cross(r_hat, v_hat), axis=1)

    def release_velocity(q_rel):
        return (omega * np.sqrt(np.sum(q_rel ** 2, axis=1)) / sintheta)[:, None] * v_hat

    return L_close, release_velocity(L_close), L_far, release_velocity(L_far), ts


def gen_streakline(pot, prog_w0, Msat, t0, t1, Nstrip, solver=Dopri8(), rtol=1e-6, atol=1e-6):
    """Streakline stream (streamhelpers.py:701-732): every particle integrated from its stripping time to t1."""
    pos_lead, vel_lead, pos_trail, vel_trail, tstrip = get_Streakline_ICs(pot=pot, prog_w0=prog_w0, Msat=Msat, t0=t0, t1=t1, Nstrip=Nstrip,
                                                                           solver=solver, rtol=rtol, atol=atol)
    w0 = np.vstack([np.hstack([pos_lead, vel_lead]), np.hstack([pos_trail, vel_trail])])
    tt = np.concatenate([tstrip, tstrip])
    sol = pot.integrate_orbit_batch_vmapped(w0=w0, ts=np.full((len(tt), 1), float(t1)), solver=solver, rtol=rtol, atol=atol, t0=tt,
                                            t1=np.full(len(tt), float(t1)))
    # a particle released AT t1 is a zero-length solve: diffrax's loop never runs and its SaveAt(ts=[t1]) row stays +inf - the same rows
    # the reference returns (and the kernels produce: tests assert inf for t0 == t1)
    ys = np.asarray(sol.ys)[:, 0]
    return ys[:Nstrip], ys[Nstrip:], tstrip


# ---- track summaries of a finished stream (streamhelpers.py:201-304): O(N) post-processing of the kernels' output ---------------
def _host(a):
    return np.asarray(a.cpu() if hasattr(a, "cpu") else a, dtype=np.float64)


def computed_binned_track(stream, phi1, bins=20):
    """Mean phase-space position in phi1 bins (streamhelpers.py:201-228), [bins-1, 6] with NaN for empty bins.  The reference's
    conventions are kept: edges = linspace(min, max, bins), `digitize` indices start at 1, so row 0 is always NaN, rows 1..bins-2 hold
    the particles with edges[r-1] <= phi1 < edges[r], and the particles of the last interval (and the maximum itself) are not used."""
    stream, phi1 = _host(stream), _host(phi1)
    edges = np.linspace(phi1.min(), phi1.max(), bins)
    dig = np.digitize(phi1, edges)
    out = np.full((len(edges) - 1, 6), np.nan)
    for b in range(len(edges) - 1):
        mask = dig == b
        if mask.any():
            out[b] = stream[mask].mean(axis=0)
    out[out == 0] = np.nan                               # streamhelpers.py:227
    return out


def compute_stream_length(stream, phi1, bins=20):
    """Length of the binned track: sum of the distances between consecutive bin means (streamhelpers.py:230-242)."""
    pos = computed_binned_track(stream, phi1, bins)[:, :3]
    return float(np.nansum(np.linalg.norm(pos[1:] - pos[:-1], axis=1)))


def compute_length_oscillations(pot, prog_today, first_stripped_lead, first_stripped_trail, t_age, length_today):
    """Stream length versus time from the separations of the progenitor and the first stripped stars integrated backwards
    (streamhelpers.py:245-304): dict(ts[2000] ascending from -t_age to 0, length_func[2000])."""
    ts = np.linspace(0.0, -float(t_age), 2_000)
    w0 = np.stack([_host(first_stripped_lead), _host(first_stripped_trail), _host(prog_today)])
    ys = np.asarray(pot.integrate_orbit_batch_vmapped(w0=w0, ts=ts, t0=0.0, t1=-float(t_age)).ys)          # three orbits, one launch
    diff = np.sqrt(np.sum((ys[0, :, :3] - ys[2, :, :3]) ** 2 + (ys[1, :, :3] - ys[2, :, :3]) ** 2, axis=1))
    diff_flip = diff[::-1]
    return dict(ts=ts[::-1].copy(), length_func=float(length_today) * diff_flip / diff_flip[-1])


# ---- sample_from_1D_pdf (streamhelpers.py:306-349): inverse-CDF sampling with jax.random.uniform's draws -----------------------
def _threefry2x32(k0, k1, c0, c1):
    """Threefry-2x32, 20 rounds (the generator behind jax.random), vectorised over the counter arrays c0, c1 (uint32)."""
    rot = ((13, 15, 26, 6), (17, 29, 16, 24))
    m = np.uint64(0xFFFFFFFF)
    ks = [np.uint64(k0) & m, np.uint64(k1) & m, (np.uint64(k0) ^ np.uint64(k1) ^ np.uint64(0x1BD11BDA)) & m]
    x0 = (np.asarray(c0, dtype=np.uint64) + ks[0]) & m
    x1 = (np.asarray(c1, dtype=np.uint64) + ks[1]) & m
    for g in range(5):
        for r in rot[g & 1]:
            x0 = (x0 + x1) & m
            x1 = ((x1 << np.uint64(r)) | (x1 >> np.uint64(32 - r))) & m
            x1 = x1 ^ x0
        x0 = (x0 + ks[(g + 1) % 3]) & m
        x1 = (x1 + ks[(g + 2) % 3] + np.uint64(g + 1)) & m
    return x0, x1


def jax_uniform(key, n):
    """jax.random.uniform(key, (n,)) in float64: 64 random bits per sample (counter pairs (j, n + j)), 52 of them as the mantissa of a
    number in [1, 2), minus 1."""
    k0, k1 = key_words(key)
    j = np.arange(n, dtype=np.uint64)
    hi, lo = _threefry2x32(k0, k1, j, j + np.uint64(n))
    bits = ((hi << np.uint64(32)) | lo) >> np.uint64(12) | np.uint64(0x3FF0000000000000)
    return bits.view(np.float64) - 1.0


def sample_from_1D_pdf(x, y, key, num_samples):
    """Random values from the 1-D density y(x) by the inverse-CDF method (streamhelpers.py:306-349): cumulative sum of the normalised
    density, linear interpolation of its inverse, jax.random.uniform draws for `key` (int seed or two key words)."""
    x, y = _host(x), _host(y)
    pdf = y / np.trapezoid(y, x)
    cdf = np.cumsum(pdf)
    cdf = cdf / cdf[-1]
    return np.interp(jax_uniform(key, int(num_samples)), cdf, x)
