"""TEST INFRASTRUCTURE ONLY - ctypes binding of the CPU oracle (oracle/ssb_oracle.cpp).

The oracle restates, on the CPU, the algorithm the reference (jnibauer/streamsculptor)
executes on its hot path; see the header of oracle/ssb_oracle.cpp for the file:line map and
for its PARITY STATUS (pinned by the reference's notebook goldens to the 9 digits they print; no runnable reference beyond that).

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline leg / --impl reference) may
import this package.  The product package `streamsculptor_b200` never does.

The builder API below is deliberately independent of the product's potential classes and
lowering code so that a parameter-order mistake in the product cannot cancel out.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

G_KPC_MYR_MSUN = 4.498502151469553e-12  # astropy G in kpc^3 Msun^-1 Myr^-2 (main.py:18,30; pinned by golden G1)

NFW, HERNQUIST, MIYAMOTO, PLUMMER, ISOCHRONE, TRIAXNFW, UNIFORM_ACC, SUBHALOS = range(8)
BAR, DEHNEN_BAR = 9, 10
LINEAR, CUBIC = 0, 1
PR_PLUMMER, PR_HERNQUIST, PR_NFW = 0, 1, 2

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_int64)


_VARIANTS = {None: ("all", "libssb_oracle.so"), "fma": ("fma", "libssb_oracle_fma.so"), "bench": ("bench", "libssb_oracle_bench.so")}


def build(force=False, variant=None):
    """Compile oracle/_build/libssb_oracle[_fma|_bench].so with the committed Makefile."""
    import fcntl
    target, name = _VARIANTS[variant]
    args = ["make", "-C", _HERE, target] + (["-B"] if force else [])
    os.makedirs(os.path.join(_HERE, "_build"), exist_ok=True)
    with open(os.path.join(_HERE, "_build", ".lock"), "w") as lock:      # several test ranks may import at once
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            subprocess.run(args, check=True, stdout=subprocess.DEVNULL)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return os.path.join(_HERE, "_build", name)


def _load(path):
    L = C.CDLL(path)
    L.orc_program_new.restype = C.c_void_p
    L.orc_normal1.restype = C.c_double
    L.orc_erfinv.restype = C.c_double
    L.orc_erfinv.argtypes = [C.c_double]
    L.orc_normal1.argtypes = [C.c_int64]
    return L


def lib():
    global _LIB
    if _LIB is None:
        _LIB = _load(build())
    return _LIB


_VARIANT_LIBS = {}


class variant:
    """Context manager: inside it every oracle call runs in another BUILD of the same sources (oracle/Makefile): "fma" = FMA contraction
    and AVX2 code generation (another legal rounding of the same algorithm), "bench" = the -O3 build bench.py times as the CPU baseline.
    Programs must be built and used inside the same context."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        global _LIB
        lib()
        if self.name not in _VARIANT_LIBS:
            _VARIANT_LIBS[self.name] = _load(build(variant=self.name))
        self._prev, _LIB = _LIB, _VARIANT_LIBS[self.name]
        return self

    def __exit__(self, *exc):
        global _LIB
        _LIB = self._prev
        return False


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(_dp)


class Program:
    """A sum of potential components (Potential_Combine, potential.py:1279-1296)."""

    def __init__(self, G=G_KPC_MYR_MSUN):
        self.G = float(G)
        self._h = C.c_void_p(lib().orc_program_new())
        self.n_sh = []

    def __del__(self):
        try:
            lib().orc_program_free(self._h)
        except Exception:
            pass

    # ---- builders -------------------------------------------------------------------------
    def track(self, kind, t, y, slopes=None):
        """slopes: knot derivatives of a cubic track given by the caller (Hermite data of any C1 piecewise cubic, e.g. the not-a-knot spline the
        reference builds with InterpolatedUnivariateSpline, potential.py:47-49) instead of interpax's 'cubic' slopes."""
        t, y = _d(t), _d(y)
        y = y.reshape(len(t), -1)
        idx = lib().orc_add_track(self._h, int(kind), len(t), _p(t), _p(y), y.shape[1])
        if slopes is not None:
            lib().orc_set_track_slopes(self._h, idx, _p(_d(slopes).reshape(len(t), -1)))
        return idx

    def _comp(self, typ, params, track=-1):
        p = np.zeros(8)
        p[: len(params)] = params
        self._last_comp = lib().orc_add_comp(self._h, typ, _p(p), int(track))
        return self

    def growing(self, t, factor, kind=LINEAR):
        """GrowingPotential (potential.py:464-477) around the component added LAST: Phi * growth_func(t), growth_func tabulated on (t, factor)."""
        lib().orc_set_growth(self._h, int(self._last_comp), int(self.track(kind, t, np.asarray(factor, dtype=np.float64).reshape(-1, 1))))
        return self

    def nfw(self, m, r_s, track=-1, soft=0.0):
        """soft: r -> sqrt(r^2 + soft).  0 = the reference today; 1e-3 = the revision that printed the notebook goldens D8 / S1."""
        return self._comp(NFW, [self.G * m, r_s, soft], track)

    def hernquist(self, m, r_s, soft=0.0, track=-1):
        return self._comp(HERNQUIST, [self.G * m, r_s, soft], track)

    def miyamoto(self, m, a, b, track=-1):
        return self._comp(MIYAMOTO, [self.G * m, a, b], track)

    def plummer(self, m, r_s, track=-1):
        return self._comp(PLUMMER, [self.G * m, r_s], track)

    def isochrone(self, m, a, track=-1):
        return self._comp(ISOCHRONE, [self.G * m, a], track)

    def triaxnfw(self, m, r_s, q1, q2, q3, track=-1):
        return self._comp(TRIAXNFW, [self.G * m, r_s, q1, q2, q3], track)

    def bar(self, m, a, b, c, Omega, track=-1):
        """BarPotential (potential.py:178-198)."""
        return self._comp(BAR, [self.G * m, a, b, c, Omega], track)

    def dehnen_bar(self, alpha, v0, R0, Rb, phib, Omega, track=-1):
        """DehnenBarPotential (potential.py:200-222)."""
        return self._comp(DEHNEN_BAR, [alpha, v0, R0, Rb, phib, Omega], track)

    def uniform_acc(self, t, vel):
        return self._comp(UNIFORM_ACC, [], self.track(LINEAR, t, vel))

    def subhalos(self, profile, m, r_s, x0, v, t0, t_window, dradius=False, track=-1):
        m, r_s, x0, v, t0 = _d(m), _d(r_s), _d(x0), _d(v), _d(t0)
        n = len(m)
        tw = _d(np.broadcast_to(_d(t_window), (n,)))
        idx = lib().orc_add_subhalos(self._h, int(profile), int(bool(dradius)), C.c_double(self.G), n, _p(m), _p(r_s),
                                     _p(x0), _p(v), _p(t0), _p(tw), int(track))
        self.n_sh.append(n)
        self.last_sh = idx
        return self

    # ---- field evaluation -------------------------------------------------------------------
    def _xt(self, xyz, t):
        xyz = _d(xyz).reshape(-1, 3)
        t = _d(np.broadcast_to(_d(t), (len(xyz),)))
        return xyz, t

    def potential(self, xyz, t=0.0):
        xyz, t = self._xt(xyz, t)
        out = np.empty(len(xyz))
        lib().orc_potential(self._h, len(xyz), _p(xyz), _p(t), _p(out))
        return out

    def gradient(self, xyz, t=0.0):
        xyz, t = self._xt(xyz, t)
        out = np.empty((len(xyz), 3))
        lib().orc_gradient(self._h, len(xyz), _p(xyz), _p(t), _p(out))
        return out

    def hessian(self, xyz, t=0.0):
        xyz, t = self._xt(xyz, t)
        out = np.empty((len(xyz), 3, 3))
        lib().orc_hessian(self._h, len(xyz), _p(xyz), _p(t), _p(out))
        return out

    def third(self, xyz, t=0.0):
        xyz, t = self._xt(xyz, t)
        out = np.empty((len(xyz), 3, 3, 3))
        lib().orc_third(self._h, len(xyz), _p(xyz), _p(t), _p(out))
        return out

    def per_sh(self, xyz, t, sh=0):
        xyz = _d(xyz).reshape(3)
        n = self.n_sh[sh]
        phi, grad = np.empty(n), np.empty((n, 3))
        lib().orc_per_sh(self._h, sh, _p(xyz), C.c_double(t), _p(phi), _p(grad))
        return phi, grad

    def track_eval(self, track, t):
        t = _d(t).reshape(-1)
        out, dout = np.empty((len(t), 3)), np.empty((len(t), 3))
        lib().orc_track_eval(self._h, int(track), len(t), _p(t), _p(out), _p(dout))
        return out, dout

    # ---- integration ------------------------------------------------------------------------
    def integrate_orbits(self, w0, t0, t1, ts=None, solver=8, rtol=1e-7, atol=1e-7, dtmin=0.3, dtmax=None,
                         max_steps=10_000, threads=1):
        """Batch of integrate_orbit calls (main.py:125-163).  ts: [M] or [N,M]; None -> [t1]."""
        w0 = _d(w0).reshape(-1, 6)
        N = len(w0)
        t0 = _d(np.broadcast_to(_d(t0), (N,)))
        t1 = _d(np.broadcast_to(_d(t1), (N,)))
        if ts is None:
            ts = t1[:, None]
        ts = _d(ts)
        if ts.ndim == 1:
            ts = np.broadcast_to(ts, (N, len(ts)))
        ts = _d(ts)
        M = ts.shape[1]
        ys = np.empty((N, M, 6))
        status = np.empty(N, dtype=np.int32)
        nsteps = np.empty((N, 3), dtype=np.int32)
        lib().orc_integrate_orbits(self._h, N, _p(w0), _p(t0), _p(t1), _p(ts), M, int(solver), C.c_double(rtol),
                                   C.c_double(atol), C.c_double(dtmin), C.c_double(np.inf if dtmax is None else dtmax),
                                   int(max_steps), _p(ys), status.ctypes.data_as(_ip), nsteps.ctypes.data_as(_ip), int(threads))
        return ys, status, nsteps

    def orbit_steps(self, w0, t0, t1, solver=8, rtol=1e-7, atol=1e-7, dtmin=0.3, dtmax=None, max_steps=10_000):
        cap = max_steps + 1
        tg, yg = np.empty(cap + 1), np.empty((cap + 1, 6))
        w0 = _d(w0)
        n = lib().orc_orbit_steps(self._h, _p(w0), C.c_double(t0), C.c_double(t1), int(solver), C.c_double(rtol),
                                  C.c_double(atol), C.c_double(dtmin), C.c_double(np.inf if dtmax is None else dtmax),
                                  int(max_steps), cap, _p(tg), _p(yg))
        return tg[: n + 1], yg[: n + 1]

    def orbit_trace(self, w0, t0, t1, solver=8, rtol=1e-7, atol=1e-7, dtmin=0.3, dtmax=None, max_steps=10_000):
        """Every step ATTEMPT of one adaptive solve: array [n_attempts, 4] = (tprev, dt, err, keep) in mirrored time, and the final state."""
        cap = max_steps + 1
        out, yfin = np.empty((cap, 4)), np.empty(6)
        w0 = _d(w0)
        n = lib().orc_orbit_trace(self._h, _p(w0), C.c_double(t0), C.c_double(t1), int(solver), C.c_double(rtol), C.c_double(atol),
                                  C.c_double(dtmin), C.c_double(np.inf if dtmax is None else dtmax), int(max_steps), cap, _p(out), _p(yfin))
        return out[:min(n, cap)].copy(), yfin

    # ---- release model ----------------------------------------------------------------------
    def release(self, xv, Msat, idx, t, seed, kvals=None, normals=None, jacobian=False):
        """release_model (main.py:209-280) for a batch; returns pos_lead, pos_trail, v_lead, v_trail [n,3]
        (or dRel_dIC [n,2,6,6] when jacobian=True, perturbative.py:281-296)."""
        xv = _d(xv).reshape(-1, 6)
        n = len(xv)
        Msat = _d(np.broadcast_to(_d(Msat), (n,)))
        t = _d(np.broadcast_to(_d(t), (n,)))
        idx = np.ascontiguousarray(idx, dtype=np.int64).reshape(n)
        kv = _d([2.0, 0.3, 0.0, 0.0, 0.4, 0.4, 0.5, 0.5] if kvals is None else kvals)   # main.py:214
        nr = None if normals is None else _d(normals).reshape(n, 4)
        nrp = _p(nr) if nr is not None else None
        if jacobian:
            out = np.empty((n, 2, 6, 6))
            lib().orc_release_jacobian(self._h, C.c_double(self.G), n, _p(xv), _p(Msat), idx.ctypes.data_as(_lp), _p(t),
                                       C.c_int64(int(seed)), _p(kv), nrp, _p(out))
            return out
        out = np.empty((n, 12))
        lib().orc_release(self._h, C.c_double(self.G), n, _p(xv), _p(Msat), idx.ctypes.data_as(_lp), _p(t),
                          C.c_int64(int(seed)), _p(kv), nrp, _p(out))
        return out[:, 0:3].copy(), out[:, 3:6].copy(), out[:, 6:9].copy(), out[:, 9:12].copy()

    def release_chen25(self, xv, Msat, t, key, mean, factor, normals=None):
        """release_model_Chen25 for a batch (streamhelpers.py:352-459): pos_lead, pos_trail, v_lead, v_trail [n,3]."""
        xv = _d(xv).reshape(-1, 6)
        n = len(xv)
        Msat, t = _d(np.broadcast_to(_d(Msat), (n,))), _d(np.broadcast_to(_d(t), (n,)))
        mean, factor = _d(mean).reshape(6), _d(factor).reshape(36)
        nr = None if normals is None else _d(normals).reshape(n, 6)
        out = np.empty((n, 12))
        k0, k1 = (0, 0) if key is None else key
        lib().orc_release_chen25(self._h, C.c_double(self.G), n, _p(xv), _p(Msat), _p(t), C.c_uint32(k0), C.c_uint32(k1), _p(mean), _p(factor),
                                 _p(nr) if nr is not None else None, _p(out))
        return out[:, 0:3].copy(), out[:, 3:6].copy(), out[:, 6:9].copy(), out[:, 9:12].copy()

    # ---- stream generation (main.py:287-368) ---------------------------------------------------
    def gen_stream_ics(self, ts, prog_w0, Msat, seed, solver=5, kvals=None, normals=None, **ctl):
        ts = _d(ts)
        prog, _, _ = self.integrate_orbits(prog_w0, ts.min(), ts.max(), ts=ts, solver=solver, **ctl)   # main.py:289
        return self.release(prog[0], Msat, np.arange(len(ts)), ts, seed, kvals, normals) + (prog[0],)

    def gen_stream(self, ts, prog_w0, Msat, seed, solver=5, kvals=None, normals=None, threads=1, **ctl):
        """gen_stream_vmapped / gen_stream_scan (main.py:312-368): returns lead[N-1,6], trail[N-1,6], nsteps."""
        ts = _d(ts)
        pl, pt, vl, vt, _ = self.gen_stream_ics(ts, prog_w0, Msat, seed, solver, kvals, normals, **ctl)
        w0 = np.concatenate([np.hstack([pl, vl])[:-1], np.hstack([pt, vt])[:-1]])
        t0 = np.concatenate([ts[:-1], ts[:-1]])
        ys, status, nsteps = self.integrate_orbits(w0, t0, ts[-1], solver=solver, threads=threads, **ctl)
        n = len(ts) - 1
        return ys[:n, -1], ys[n:, -1], status, nsteps


def linear_response(base, shprog, w0, t0, t1, D0=None, sh=0, solver=8, rtol=1e-6, atol=1e-6, dtmin=0.05, dtmax=None,
                    max_steps=10_000, threads=1):
    """compute_perturbation_OTF (perturbative.py:101-135, 726-755): returns w[N,6], D[N,nsh,12], status, nsteps."""
    w0 = _d(w0).reshape(-1, 6)
    N = len(w0)
    nsh = shprog.n_sh[sh]
    t0 = _d(np.broadcast_to(_d(t0), (N,)))
    D0p = None
    if D0 is not None:
        D0 = _d(D0).reshape(N, nsh, 12)
        D0p = _p(D0)
    wout, Dout = np.empty((N, 6)), np.empty((N, nsh, 12))
    status, nsteps = np.empty(N, dtype=np.int32), np.empty((N, 3), dtype=np.int32)
    lib().orc_linear_response(base._h, shprog._h, sh, N, _p(w0), D0p, _p(t0), C.c_double(t1), int(solver), C.c_double(rtol),
                              C.c_double(atol), C.c_double(dtmin), C.c_double(np.inf if dtmax is None else dtmax),
                              int(max_steps), _p(wout), _p(Dout), status.ctypes.data_as(_ip), nsteps.ctypes.data_as(_ip),
                              int(threads))
    return wout, Dout, status, nsteps


def linear_response_saveat(base, shprog, w0, t0, t1, ts, D0=None, sh=0, solver=8, rtol=1e-7, atol=1e-7, dtmin=0.05, dtmax=None, max_steps=1_000):
    """integrate_field(w0=[w, D], ts, field=MassRadiusPerturbation_OTF) for one trajectory (perturbative.py:53-60): ws[M,6], Ds[M,nsh,12]."""
    w0, ts = _d(w0).reshape(6), _d(ts).reshape(-1)
    nsh, M = shprog.n_sh[sh], len(_d(ts).reshape(-1))
    D0p = None
    if D0 is not None:
        D0 = _d(D0).reshape(nsh, 12)
        D0p = _p(D0)
    ws, Ds = np.empty((M, 6)), np.empty((M, nsh, 12))
    status, nsteps = np.zeros(1, dtype=np.int32), np.zeros(3, dtype=np.int32)
    lib().orc_linear_response_saveat(base._h, shprog._h, sh, _p(w0), D0p, C.c_double(t0), C.c_double(t1), _p(ts), M, int(solver), C.c_double(rtol),
                                     C.c_double(atol), C.c_double(dtmin), C.c_double(np.inf if dtmax is None else dtmax), int(max_steps), _p(ws), _p(Ds),
                                     status.ctypes.data_as(_ip), nsteps.ctypes.data_as(_ip))
    return ws, Ds, status, nsteps


def second_order_response(base, shprog, w0, t0, t1, D0=None, E0=None, sh=0, solver=8, rtol=1e-6, atol=1e-6, dtmin=0.05, dtmax=None,
                          max_steps=10_000, threads=1):
    """compute_perturbation_second_order_OTF (perturbative.py:757-772): w[N,6], D[N,nsh,12], E[N,nsh,6], status, nsteps."""
    w0 = _d(w0).reshape(-1, 6)
    N, nsh = len(w0), shprog.n_sh[sh]
    t0 = _d(np.broadcast_to(_d(t0), (N,)))
    D0 = None if D0 is None else _d(D0).reshape(N, nsh, 12)
    E0 = None if E0 is None else _d(E0).reshape(N, nsh, 6)
    wout, Dout, Eout = np.empty((N, 6)), np.empty((N, nsh, 12)), np.empty((N, nsh, 6))
    status, nsteps = np.empty(N, dtype=np.int32), np.empty((N, 3), dtype=np.int32)
    lib().orc_second_order_response(base._h, shprog._h, sh, N, _p(w0), None if D0 is None else _p(D0), None if E0 is None else _p(E0), _p(t0),
                                    C.c_double(t1), int(solver), C.c_double(rtol), C.c_double(atol), C.c_double(dtmin),
                                    C.c_double(np.inf if dtmax is None else dtmax), int(max_steps), _p(wout), _p(Dout), _p(Eout),
                                    status.ctypes.data_as(_ip), nsteps.ctypes.data_as(_ip), int(threads))
    return wout, Dout, Eout, status, nsteps


def second_order_term(base, shprog, t, y, sh=0):
    y = _d(y)
    dy = np.empty_like(y)
    lib().orc_second_order_term(base._h, shprog._h, sh, C.c_double(t), _p(y), _p(dy))
    return dy


def response_term(base, shprog, t, y, sh=0):
    y = _d(y)
    dy = np.empty_like(y)
    lib().orc_response_term(base._h, shprog._h, sh, C.c_double(t), _p(y), _p(dy))
    return dy



def shared_step_orbits(prog, w0, t0, t1, ts=None, solver=8, rtol=1e-7, atol=1e-7, dtmin=0.05, dtmax=None, max_steps=1_000):
    """integrate_field(w0[N,6], field=RestrictedNbody_generator) (RestrictedNbody.py:93-106,131): all N tracers are ONE ODE
    state with a shared controller.  Returns ys[M,N,6] (ts None -> [t1]), status, nsteps[3]."""
    w0 = _d(w0).reshape(-1, 6)
    N = len(w0)
    ts = _d([t1] if ts is None else ts).reshape(-1)
    ys = np.empty((len(ts), N, 6))
    status, nsteps = np.zeros(1, dtype=np.int32), np.zeros(3, dtype=np.int32)
    lib().orc_shared_step_orbits(prog._h, N, _p(w0), C.c_double(t0), C.c_double(t1), _p(ts), len(ts), int(solver), C.c_double(rtol), C.c_double(atol),
                                 C.c_double(dtmin), C.c_double(np.inf if dtmax is None else dtmax), int(max_steps), _p(ys),
                                 status.ctypes.data_as(_ip), nsteps.ctypes.data_as(_ip))
    return ys, int(status[0]), nsteps


def nbody(ext, masses, w0, t0, t1, ts=None, eps=1e-3, G=G_KPC_MYR_MSUN, solver=8, rtol=1e-7, atol=1e-7, dtmin=0.05, dtmax=None, max_steps=1_000):
    """integrate_field(w0[N,6], field=Nbody_field(ext_pot, masses, eps)) (fields.py:115-155).  ext: Program or None."""
    w0, masses = _d(w0).reshape(-1, 6), _d(masses).reshape(-1)
    N = len(w0)
    ts = _d([t1] if ts is None else ts).reshape(-1)
    ys = np.empty((len(ts), N, 6))
    status, nsteps = np.zeros(1, dtype=np.int32), np.zeros(3, dtype=np.int32)
    lib().orc_nbody(None if ext is None else ext._h, N, _p(masses), C.c_double(G), C.c_double(eps), _p(w0), C.c_double(t0), C.c_double(t1), _p(ts),
                    len(ts), int(solver), C.c_double(rtol), C.c_double(atol), C.c_double(dtmin), C.c_double(np.inf if dtmax is None else dtmax),
                    int(max_steps), _p(ys), status.ctypes.data_as(_ip), nsteps.ctypes.data_as(_ip))
    return ys, int(status[0]), nsteps


def nbody_term(ext, masses, t, y, eps=1e-3, G=G_KPC_MYR_MSUN):
    y, masses = _d(y).reshape(-1, 6), _d(masses).reshape(-1)
    dy = np.empty_like(y)
    lib().orc_nbody_term(None if ext is None else ext._h, len(y), _p(masses), C.c_double(G), C.c_double(eps), C.c_double(t), _p(y), _p(dy))
    return dy


def variational(prog, w0, t0, t1, order=1, M0=None, M20=None, solver=8, rtol=1e-7, atol=1e-7, dtmin=0.05, dtmax=None, max_steps=10_000, threads=1):
    """Variational equations along each orbit (higher_order_variationalEqn.ipynb cell 3): returns w[N,6], M[N,6,6], M2[N,6,6,6] | None."""
    w0 = _d(w0).reshape(-1, 6)
    N = len(w0)
    t0 = _d(np.broadcast_to(_d(t0), (N,)))
    M0p = None if M0 is None else _p(_d(M0).reshape(N, 6, 6))
    M20p = None if M20 is None else _p(_d(M20).reshape(N, 6, 6, 6))
    wout, Mout = np.empty((N, 6)), np.empty((N, 6, 6))
    M2out = np.empty((N, 6, 6, 6)) if order == 2 else np.empty(1)
    status, nsteps = np.empty(N, dtype=np.int32), np.empty((N, 3), dtype=np.int32)
    lib().orc_variational(prog._h, int(order), N, _p(w0), M0p, M20p, _p(t0), C.c_double(t1), int(solver), C.c_double(rtol), C.c_double(atol),
                          C.c_double(dtmin), C.c_double(np.inf if dtmax is None else dtmax), int(max_steps), _p(wout), _p(Mout), _p(M2out),
                          status.ctypes.data_as(_ip), nsteps.ctypes.data_as(_ip), int(threads))
    return wout, Mout, (M2out if order == 2 else None), status, nsteps


def variational_term(prog, t, y, order=1):
    y = _d(y).reshape(-1)
    dy = np.empty_like(y)
    lib().orc_variational_term(prog._h, int(order), C.c_double(t), _p(y), _p(dy))
    return dy

def threefry2x32(k0, k1, c0, c1):
    out = (C.c_uint32 * 2)()
    lib().orc_threefry(C.c_uint32(k0), C.c_uint32(k1), C.c_uint32(c0), C.c_uint32(c1), out)
    return int(out[0]), int(out[1])


def randint5(seed, lo=0, hi=1000):
    out = np.empty(5, dtype=np.int64)
    lib().orc_randint5(C.c_int64(seed), C.c_int64(lo), C.c_int64(hi), out.ctypes.data_as(_lp))
    return out


def normal1(seed):
    return lib().orc_normal1(C.c_int64(seed))


def erfinv(x):
    return lib().orc_erfinv(C.c_double(x))


def release_normals(seed, idx):
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    out = np.empty((len(idx), 4))
    lib().orc_release_normals(C.c_int64(seed), len(idx), idx.ctypes.data_as(_lp), _p(out))
    return out


def num_threads():
    return lib().orc_num_threads()
