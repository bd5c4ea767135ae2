"""Host-side plumbing between the Python API and libssb200 (device-pointer entry points).

torch is used for what the task allows it for: device memory, streams and (in parallel.py) torch.distributed.
All arithmetic happens in the CUDA library; there is no CPU fallback anywhere in this package.
"""
import ctypes as C

import numpy as np

from . import _lib
from .solvers import solver_id

_F64 = None


def torch():
    return _lib.require_cuda()


def device():
    t = torch()
    return t.device("cuda", t.cuda.current_device())


def to_dev(a, dtype=None):
    """numpy / list / torch (any device) -> contiguous CUDA tensor."""
    t = torch()
    dtype = dtype or t.float64
    if isinstance(a, t.Tensor):
        return a.to(device=device(), dtype=dtype).contiguous()
    arr = np.ascontiguousarray(np.asarray(a, dtype=np.float64 if dtype == t.float64 else None))
    return t.from_numpy(arr).to(device=device(), dtype=dtype).contiguous()


def is_dev(a):
    t = torch()
    return isinstance(a, t.Tensor) and a.is_cuda


def out(x, like_device):
    """Return type policy: CUDA tensors in -> CUDA tensors out; anything else -> numpy arrays."""
    return x if like_device else x.cpu().numpy()


def ptr(t):
    return C.c_void_p(0) if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    return C.c_void_p(torch().cuda.current_stream().cuda_stream)


def empty(shape, dtype=None):
    t = torch()
    return t.empty(shape, dtype=dtype or t.float64, device=device())


def make_ctrl(solver, rtol, atol, dtmin, dtmax, max_steps):
    return _lib.Ctrl(solver=solver_id(solver), max_steps=int(max_steps), rtol=float(rtol), atol=float(atol),
                     dtmin=float(dtmin), dtmax=float(np.inf if dtmax is None else dtmax))


# ------------------------------------------------------------------------------------------------
# lowering: Potential object tree -> flat potential program (walk potential_list
# recursively and dispatch on component type; never call .potential of force-only components)
# ------------------------------------------------------------------------------------------------
class Track:
    """A tabulated 3-vector function of time living on the device."""

    def __init__(self, kind, t, y, slopes=None):
        """slopes (cubic tracks only): knot derivatives [n, 3] of ANY C1 piecewise-cubic through the knots (Hermite data), e.g. a not-a-knot
        spline; None: interpax's 'cubic' slopes, computed on the device (ssb_track_slopes_f64)."""
        self.kind = int(kind)
        self.t_host = np.ascontiguousarray(np.asarray(t, dtype=np.float64)).reshape(-1)
        self.y_host = np.ascontiguousarray(np.asarray(y, dtype=np.float64)).reshape(len(self.t_host), 3)
        if len(self.t_host) < 2:
            raise ValueError("a track needs at least 2 knots")
        self.s_host = None
        if slopes is not None:
            if self.kind != _lib.TRACK_CUBIC:
                raise ValueError("knot slopes belong to cubic tracks")
            self.s_host = np.ascontiguousarray(np.asarray(slopes, dtype=np.float64)).reshape(len(self.t_host), 3)
        self._dev = None

    def dev(self):
        if self._dev is None:
            t, y = to_dev(self.t_host), to_dev(self.y_host)
            s = None
            if self.kind == _lib.TRACK_CUBIC and self.s_host is not None:
                s = to_dev(self.s_host)
            elif self.kind == _lib.TRACK_CUBIC:
                s = empty((len(self.t_host), 3))
                _lib.check(_lib.lib().ssb_track_slopes_f64(len(self.t_host), ptr(t), ptr(y), ptr(s), stream_ptr()))
            self._dev = (t, y, s)
        return self._dev

    def struct(self):
        t, y, s = self.dev()
        return _lib.Track(kind=self.kind, n=len(self.t_host), t=t.data_ptr(), y=y.data_ptr(), s=0 if s is None else s.data_ptr())

    def __call__(self, tq, derivative=False):
        tq_dev = to_dev(np.atleast_1d(np.asarray(tq, dtype=np.float64)))
        o, d = empty((len(tq_dev), 3)), empty((len(tq_dev), 3))
        st = self.struct()
        _lib.check(_lib.lib().ssb_track_eval_f64(C.byref(st), len(tq_dev), ptr(tq_dev), ptr(o), ptr(d), stream_ptr()))
        res = (d if derivative else o).cpu().numpy()
        return res[0] if np.ndim(tq) == 0 else res


def as_track(obj, default_kind=None):
    """Accept our Track objects or interpolator-like objects exposing knots (interpax.Interpolator1D: .x, .f, .method)."""
    if isinstance(obj, Track):
        return obj
    if hasattr(obj, "x") and hasattr(obj, "f"):
        method = getattr(obj, "method", "cubic")
        if method == "linear":
            kind = _lib.TRACK_LINEAR
        elif method == "cubic":
            kind = _lib.TRACK_CUBIC
        else:
            raise NotImplementedError(f"interpolation method {method!r} is not implemented on the device")
        return Track(kind, np.asarray(obj.x), np.asarray(obj.f))
    raise NotImplementedError(
        "time-dependent centres must be tabulated tracks (streamsculptor_b200.LinearTrack / CubicTrack or an "
        "interpax-like object with .x/.f); arbitrary Python callables cannot run inside the CUDA kernels")


class SubhaloArrays:
    def __init__(self, profile, G, m, rs, x0, v, t0, tw):
        m = np.atleast_1d(np.asarray(m, dtype=np.float64))
        n = len(m)
        self.n, self.profile, self.G = n, int(profile), float(G)
        self.host = dict(m=m, rs=np.broadcast_to(np.asarray(rs, dtype=np.float64), (n,)).copy(),
                         x0=np.asarray(x0, dtype=np.float64).reshape(n, 3).copy(), v=np.asarray(v, dtype=np.float64).reshape(n, 3).copy(),
                         t0=np.broadcast_to(np.asarray(t0, dtype=np.float64), (n,)).copy(),
                         tw=np.broadcast_to(np.asarray(tw, dtype=np.float64), (n,)).copy())
        self._dev = None

    def struct(self):
        if self._dev is None:
            self._dev = {k: to_dev(v) for k, v in self.host.items()}
        d = self._dev
        return _lib.Subhalos(n=self.n, profile=self.profile, G=self.G, m=d["m"].data_ptr(), rs=d["rs"].data_ptr(), x0=d["x0"].data_ptr(),
                             v=d["v"].data_ptr(), t0=d["t0"].data_ptr(), tw=d["tw"].data_ptr())


class PerturberArrays:
    """n moving spheres of one profile on a shared time grid (ssb_perturbers): GM[n], rs[n], t[nk], centres y[nk, n, 3]."""

    def __init__(self, profile, GM, rs, t, y):
        self.profile = int(profile)
        self.host = dict(GM=np.ascontiguousarray(np.asarray(GM, dtype=np.float64).reshape(-1)), rs=None, t=np.ascontiguousarray(np.asarray(t, dtype=np.float64).reshape(-1)), y=None)
        self.n, self.n_knots = len(self.host["GM"]), len(self.host["t"])
        self.host["rs"] = np.ascontiguousarray(np.broadcast_to(np.asarray(rs, dtype=np.float64), (self.n,)).copy())
        self.host["y"] = np.ascontiguousarray(np.asarray(y, dtype=np.float64).reshape(self.n_knots, self.n, 3))
        if self.n_knots < 2:
            raise ValueError("a perturber set needs at least 2 knots")
        self._dev = None

    def struct(self):
        if self._dev is None:
            self._dev = {k: to_dev(v) for k, v in self.host.items()}
        d = self._dev
        return _lib.Perturbers(n=self.n, n_knots=self.n_knots, profile=self.profile, t=d["t"].data_ptr(), y=d["y"].data_ptr(), GM=d["GM"].data_ptr(),
                               rs=d["rs"].data_ptr())


_PROFILE_OF_COMP = {}          # filled below: component type -> subhalo / perturber profile id


class Program:
    """Flat potential program under construction.  Component / track limits are enforced when the struct is built, after moving
    spheres that share a time grid have been packed into a perturber set (pack_perturbers) if the program would not fit otherwise."""

    def __init__(self):
        self.comps, self.tracks, self.shs, self.psets = [], [], [], []
        self.growth = 0          # > 0 while the components of a GrowingPotential are being added: index + 1 of its growth-factor track

    def add_track(self, track):
        for i, t in enumerate(self.tracks):
            if t is track:
                return i
        self.tracks.append(track)
        return len(self.tracks) - 1

    def add_perturbers(self, arrays):
        if len(self.psets) >= _lib.MAX_PSET:
            raise NotImplementedError(f"more than {_lib.MAX_PSET} perturber set(s) in one potential")
        self.psets.append(arrays)
        self.add(_lib.PERTURBERS, [], sh=len(self.psets) - 1)

    def pack_perturbers(self):
        """Move the largest group of translating spherical components (same type, linear tracks on one time grid, no growth factor) into
        a perturber set: what a Potential_Combine of many TimeDepTranslatingPotential objects (the reference's way to write N moving
        perturbers, potential.py:448-462) lowers to when it exceeds the component / track limits."""
        if len(self.psets) >= _lib.MAX_PSET:
            return False
        groups = {}
        for i, (typ, params, track, sh, growth) in enumerate(self.comps):
            if typ not in _PROFILE_OF_COMP or track < 0 or growth or self.tracks[track].kind != _lib.TRACK_LINEAR:
                continue
            if typ == _lib.HERNQUIST and params[2] != 0.0:
                continue
            t = self.tracks[track].t_host
            groups.setdefault((typ, len(t), float(t[0]), float(t[-1])), []).append(i)
        groups = {k: [i for i in v if np.array_equal(self.tracks[self.comps[i][2]].t_host, self.tracks[self.comps[v[0]][2]].t_host)] for k, v in groups.items()}
        if not groups:
            return False
        key, members = max(groups.items(), key=lambda kv: len(kv[1]))
        if len(members) < 2:
            return False
        tr = [self.tracks[self.comps[i][2]] for i in members]
        arrays = PerturberArrays(_PROFILE_OF_COMP[key[0]], [self.comps[i][1][0] for i in members], [self.comps[i][1][1] for i in members], tr[0].t_host,
                                 np.stack([t.y_host for t in tr], axis=1))
        keep = [c for i, c in enumerate(self.comps) if i not in set(members)]
        used = sorted({c[2] for c in keep if c[2] >= 0} | {c[4] - 1 for c in keep if c[4] > 0})
        remap = {old: new for new, old in enumerate(used)}
        self.tracks = [self.tracks[o] for o in used]
        self.comps = [(typ, params, remap.get(track, -1), sh, (remap[growth - 1] + 1) if growth > 0 else 0) for typ, params, track, sh, growth in keep]
        self.add_perturbers(arrays)
        return True

    def add(self, typ, params, track=-1, sh=-1):
        if self.growth and int(typ) in (_lib.UNIFORM_ACC, _lib.SUBHALOS):
            raise NotImplementedError("GrowingPotential around a force-only (UniformAcceleration) or subhalo-ensemble component is not supported")
        self.comps.append((int(typ), [float(p) for p in params], int(track), int(sh), int(self.growth)))

    def add_subhalos(self, arrays, track=-1):
        if len(self.shs) >= _lib.MAX_SH:
            raise NotImplementedError(f"more than {_lib.MAX_SH} subhalo sets in one potential")
        self.shs.append(arrays)
        self.add(_lib.SUBHALOS, [], track=track, sh=len(self.shs) - 1)

    def struct(self):
        if len(self.comps) > _lib.MAX_COMP or len(self.tracks) > _lib.MAX_TRACK:
            self.pack_perturbers()
        if len(self.comps) > _lib.MAX_COMP:
            raise NotImplementedError(f"more than {_lib.MAX_COMP} components in one potential (after packing moving spheres into a perturber set)")
        if len(self.tracks) > _lib.MAX_TRACK:
            raise NotImplementedError(f"more than {_lib.MAX_TRACK} tabulated tracks in one potential (after packing moving spheres into a perturber set)")
        P = _lib.Potential()
        P.n_comp, P.n_track, P.n_sh, P.n_pset = len(self.comps), len(self.tracks), len(self.shs), len(self.psets)
        for i, s in enumerate(self.psets):
            P.pset[i] = s.struct()
        for i, (typ, params, track, sh, growth) in enumerate(self.comps):
            P.comp[i].type, P.comp[i].track, P.comp[i].sh, P.comp[i].growth = typ, track, sh, growth
            for k, v in enumerate(params):
                P.comp[i].p[k] = v
        for i, t in enumerate(self.tracks):
            P.track[i] = t.struct()
        for i, s in enumerate(self.shs):
            P.sh[i] = s.struct()
        return P


_PROFILE_OF_COMP.update({_lib.PLUMMER: _lib.PROFILE_PLUMMER, _lib.HERNQUIST: _lib.PROFILE_HERNQUIST, _lib.NFW: _lib.PROFILE_NFW})


def lower(pot):
    """Potential object -> (ctypes struct, keepalive).  Cached on the object (parameters are immutable by convention)."""
    cached = getattr(pot, "_ssb_lowered", None)
    if cached is not None:
        return cached
    prog = Program()
    pot._lower(prog, -1)
    res = (prog.struct(), prog)
    try:
        pot._ssb_lowered = res
    except Exception:
        pass
    return res


# ------------------------------------------------------------------------------------------------
# calls
# ------------------------------------------------------------------------------------------------
def potential_eval(pot, xyz, t, want):
    dev_in = is_dev(xyz)
    x = to_dev(xyz).reshape(-1, 3)
    n = x.shape[0]
    tt = to_dev(np.broadcast_to(np.asarray(t, dtype=np.float64), (n,)).copy()) if not is_dev(t) else t.reshape(-1).expand(n).contiguous()
    P, _keep = lower(pot)
    phi = empty((n,)) if "phi" in want else None
    grad = empty((n, 3)) if "grad" in want else None
    hess = empty((n, 3, 3)) if "hess" in want else None
    _lib.check(_lib.lib().ssb_potential_eval_f64(C.byref(P), n, ptr(x), ptr(tt), ptr(phi), ptr(grad), ptr(hess), stream_ptr()))
    return tuple(out(v, dev_in) for v in (phi, grad, hess) if v is not None)


def orbit_integrate(pot, w0, t0, t1, ts, ctrl, ts_per_orbit):
    """Device tensors in, device tensors out: ys[N,M,6], status[N], nsteps[N,3]."""
    t = torch()
    P, _keep = lower(pot)
    N, M = w0.shape[0], ts.shape[-1]
    ys = empty((N, M, 6))
    status = empty((N,), t.int32)
    nsteps = empty((N, 3), t.int32)
    _lib.check(_lib.lib().ssb_orbit_integrate_f64(C.byref(P), N, ptr(w0), ptr(t0), ptr(t1), ptr(ts), M, int(ts_per_orbit), ctrl,
                                                  ptr(ys), ptr(status), ptr(nsteps), stream_ptr()))
    return ys, status, nsteps


def orbit_trace(pot, w0, t0, t1, ctrl, trace_cap=None):
    """Every step attempt of N adaptive solves (diagnostics): trace[N,cap,4] = (tprev, dt, err, keep) in mirrored time (NaN-filled beyond
    an orbit's attempts), yfin[N,6], status[N], nsteps[N,3]."""
    t = torch()
    P, _keep = lower(pot)
    N = w0.shape[0]
    cap = int(trace_cap if trace_cap is not None else ctrl.max_steps)
    trace = t.full((N, cap, 4), float("nan"), dtype=t.float64, device=device())
    yfin, status, nsteps = empty((N, 6)), empty((N,), t.int32), empty((N, 3), t.int32)
    _lib.check(_lib.lib().ssb_orbit_trace_f64(C.byref(P), N, ptr(w0), ptr(t0), ptr(t1), ctrl, cap, ptr(trace), ptr(yfin), ptr(status), ptr(nsteps),
                                              stream_ptr()))
    return trace, yfin, status, nsteps


def orbit_dense(pot, w0, t0, t1, ts, ctrl):
    """One orbit: returns ys[M,6], status[1], nsteps[3], scratch (kept for later dense evaluation)."""
    t = torch()
    P, _keep = lower(pot)
    M = ts.shape[0]
    nbytes = _lib.lib().ssb_scratch_bytes(ctrl.max_steps)
    scratch = empty(((nbytes + 7) // 8,))
    ys = empty((M, 6))
    status = empty((1,), t.int32)
    nsteps = empty((3,), t.int32)
    _lib.check(_lib.lib().ssb_orbit_dense_f64(C.byref(P), ptr(w0), float(t0), float(t1), ptr(ts), M, ctrl, ptr(ys), ptr(status), ptr(nsteps),
                                              ptr(scratch), nbytes, stream_ptr()))
    return ys, status, nsteps, scratch


def release_spray(pot, G, prog, Msat, idx, t, seed, kvals, normals):
    tt = torch()
    P, _keep = lower(pot)
    N = prog.shape[0]
    outs = [empty((N, 3)) for _ in range(4)]
    kv = (C.c_double * 8)(*[float(k) for k in kvals])
    _lib.check(_lib.lib().ssb_release_spray_f64(C.byref(P), float(G), N, ptr(prog), ptr(Msat), ptr(idx), ptr(t), int(seed), kv, ptr(normals),
                                                ptr(outs[0]), ptr(outs[1]), ptr(outs[2]), ptr(outs[3]), stream_ptr()))
    return outs


def shard_count(n_particles, rank, world):
    """Number of particles i = rank + k*world below n_particles (interleaved sharding)."""
    return max(0, (n_particles - rank + world - 1) // world)


def release_chen25(pot, G, prog, Msat, t, key, mean, factor, normals):
    P, _keep = lower(pot)
    N = prog.shape[0]
    outs = [empty((N, 3)) for _ in range(4)]
    kc = None if key is None else (C.c_uint32 * 2)(int(key[0]), int(key[1]))
    mc = (C.c_double * 6)(*[float(v) for v in mean])
    fc = (C.c_double * 36)(*[float(v) for v in np.asarray(factor, dtype=np.float64).reshape(36)])
    _lib.check(_lib.lib().ssb_release_chen25_f64(C.byref(P), float(G), N, ptr(prog), ptr(Msat), ptr(t), kc, mc, fc, ptr(normals), ptr(outs[0]),
                                                 ptr(outs[1]), ptr(outs[2]), ptr(outs[3]), stream_ptr()))
    return outs


def release_jacobian(pot, G, prog, Msat, idx, t, seed, kvals, normals):
    P, _keep = lower(pot)
    N = prog.shape[0]
    jac = empty((N, 2, 6, 6))
    kv = (C.c_double * 8)(*[float(k) for k in kvals])
    _lib.check(_lib.lib().ssb_release_jacobian_f64(C.byref(P), float(G), N, ptr(prog), ptr(Msat), ptr(idx), ptr(t), int(seed), kv, ptr(normals),
                                                   ptr(jac), stream_ptr()))
    return jac


def potential_third(pot, xyz, t):
    x = to_dev(xyz).reshape(-1, 3)
    n = x.shape[0]
    tt = to_dev(np.broadcast_to(np.asarray(t, dtype=np.float64), (n,)).copy())
    P, _keep = lower(pot)
    out3 = empty((n, 3, 3, 3))
    _lib.check(_lib.lib().ssb_potential_third_f64(C.byref(P), n, ptr(x), ptr(tt), ptr(out3), stream_ptr()))
    return out3


_STREAM_SCRATCH = {}


def _stream_scratch(nbytes):
    """Scratch of gen_stream, kept per (device, CUDA stream) between calls: the buffer is only touched by kernels of the call that owns it, and
    calls on one stream are ordered, so re-using it needs no synchronisation; a 1e6-particle stream needs 200 MB, and asking the allocator
    for it at every step sits on the critical path of a sharded step (2 ms at 8 GPUs)."""
    t = torch()
    key = (t.cuda.current_device(), t.cuda.current_stream().cuda_stream)
    buf = _STREAM_SCRATCH.get(key)
    if buf is None or buf.numel() * 8 < nbytes:
        if len(_STREAM_SCRATCH) > 16:
            _STREAM_SCRATCH.clear()
        buf = _STREAM_SCRATCH[key] = empty(((nbytes + 7) // 8,))
    return buf


def gen_stream(pot, pot_release, G, ts, prog_w0, Msat, seed, kvals, normals, ctrl, i_begin=0, i_stride=1, n_local=None):
    tt = torch()
    P, _k1 = lower(pot)
    PR, _k2 = lower(pot_release)
    Nts = ts.shape[0]
    i_begin, i_stride = int(i_begin), int(i_stride)
    n = shard_count(Nts - 1, i_begin, i_stride) if n_local is None else int(n_local)
    both = empty((2, n, 6))                       # lead and trail share one buffer: a sharded run gathers them in one collective
    lead, trail = both[0], both[1]
    lead._ssb_packed = both
    status, nsteps = empty((2, n), tt.int32), empty((2, n, 3), tt.int32)
    nbytes = _lib.lib().ssb_stream_scratch_bytes(Nts, ctrl.max_steps)
    scratch = _stream_scratch(nbytes)
    kv = (C.c_double * 8)(*[float(k) for k in kvals])
    _lib.check(_lib.lib().ssb_gen_stream_f64(C.byref(P), C.byref(PR), float(G), Nts, ptr(ts), ptr(prog_w0), ptr(Msat), int(seed), kv, ptr(normals),
                                             ctrl, i_begin, i_stride, n, ptr(lead), ptr(trail), ptr(status), ptr(nsteps), ptr(scratch), nbytes,
                                             stream_ptr()))
    return lead, trail, status, nsteps


def linear_response(pot_base, sharrays, w0, D0, t0, t1, ctrl):
    tt = torch()
    P, _keep = lower(pot_base)
    S = sharrays.struct()
    N = w0.shape[0]
    wout, Dout = empty((N, 6)), empty((N, sharrays.n, 12))
    status, nsteps = empty((N,), tt.int32), empty((N, 3), tt.int32)
    nbytes = _lib.lib().ssb_response_scratch_bytes(sharrays.n)
    scratch = empty(((nbytes + 7) // 8,))
    _lib.check(_lib.lib().ssb_linear_response_f64(C.byref(P), C.byref(S), N, ptr(w0), ptr(D0), ptr(t0), float(t1), ctrl, ptr(wout), ptr(Dout),
                                                  ptr(status), ptr(nsteps), ptr(scratch), nbytes, stream_ptr()))
    return wout, Dout, status, nsteps


def linear_response_saveat(pot_base, sharrays, w0, D0, t0, t1, ts, ctrl):
    """One trajectory of the coupled field saved at ts[M]: returns ws[M,6], Ds[M,n_sh,12], status[1], nsteps[3]."""
    tt = torch()
    P, _keep = lower(pot_base)
    S = sharrays.struct()
    M = ts.shape[0]
    ws, Ds = empty((M, 6)), empty((M, sharrays.n, 12))
    status, nsteps = empty((1,), tt.int32), empty((1, 3), tt.int32)
    nbytes = _lib.lib().ssb_response_saveat_scratch_bytes(sharrays.n)
    scratch = empty(((nbytes + 7) // 8,))
    _lib.check(_lib.lib().ssb_linear_response_saveat_f64(C.byref(P), C.byref(S), ptr(w0), ptr(D0), ptr(t0), float(t1), ptr(ts), M, ctrl, ptr(ws), ptr(Ds),
                                                         ptr(status), ptr(nsteps), ptr(scratch), nbytes, stream_ptr()))
    return ws, Ds, status, nsteps


def second_order_response(pot_base, sharrays, w0, D0, E0, t0, t1, ctrl):
    tt = torch()
    P, _keep = lower(pot_base)
    S = sharrays.struct()
    N = w0.shape[0]
    wout, Dout, Eout = empty((N, 6)), empty((N, sharrays.n, 12)), empty((N, sharrays.n, 6))
    status, nsteps = empty((N,), tt.int32), empty((N, 3), tt.int32)
    nbytes = _lib.lib().ssb_second_order_scratch_bytes(sharrays.n)
    scratch = empty(((nbytes + 7) // 8,))
    if is_dev(t1):            # per-particle end times
        _lib.check(_lib.lib().ssb_second_order_response_ends_f64(C.byref(P), C.byref(S), N, ptr(w0), ptr(D0), ptr(E0), ptr(t0), ptr(t1), ctrl, ptr(wout),
                                                                 ptr(Dout), ptr(Eout), ptr(status), ptr(nsteps), ptr(scratch), nbytes, stream_ptr()))
    else:
        _lib.check(_lib.lib().ssb_second_order_response_f64(C.byref(P), C.byref(S), N, ptr(w0), ptr(D0), ptr(E0), ptr(t0), float(t1), ctrl, ptr(wout),
                                                            ptr(Dout), ptr(Eout), ptr(status), ptr(nsteps), ptr(scratch), nbytes, stream_ptr()))
    return wout, Dout, Eout, status, nsteps


def second_order_term(pot_base, sharrays, t, y):
    P, _keep = lower(pot_base)
    S = sharrays.struct()
    yd = to_dev(y).reshape(-1)
    dy = empty(yd.shape)
    _lib.check(_lib.lib().ssb_second_order_term_f64(C.byref(P), C.byref(S), float(t), ptr(yd), ptr(dy), stream_ptr()))
    return dy


def response_term(pot_base, sharrays, t, y):
    P, _keep = lower(pot_base)
    S = sharrays.struct()
    yd = to_dev(y).reshape(-1)
    dy = empty(yd.shape)
    _lib.check(_lib.lib().ssb_response_term_f64(C.byref(P), C.byref(S), float(t), ptr(yd), ptr(dy), stream_ptr()))
    return dy


def subhalo_eval(sharrays, dradius, xyz, t):
    S = sharrays.struct()
    x = (C.c_double * 3)(*[float(v) for v in np.asarray(xyz, dtype=np.float64).reshape(3)])
    phi, grad = empty((sharrays.n,)), empty((sharrays.n, 3))
    _lib.check(_lib.lib().ssb_subhalo_eval_f64(C.byref(S), int(bool(dradius)), x, float(t), ptr(phi), ptr(grad), stream_ptr()))
    return phi, grad


def shared_step_orbits(pot, w0, t0, t1, ctrl):
    """N tracers as ONE ODE state with a shared controller (RestrictedNbody.py:93-106,131): wout[N,6], status[1], nsteps[3]."""
    tt = torch()
    P, _keep = lower(pot)
    N = w0.shape[0]
    wout = empty((N, 6))
    status, nsteps = empty((1,), tt.int32), empty((3,), tt.int32)
    nbytes = _lib.lib().ssb_shared_scratch_bytes(N)
    scratch = empty(((nbytes + 7) // 8,))
    _lib.check(_lib.lib().ssb_shared_step_orbits_f64(C.byref(P), N, ptr(w0), float(t0), float(t1), ctrl, ptr(wout), ptr(status), ptr(nsteps),
                                                     ptr(scratch), nbytes, stream_ptr()))
    return wout, status, nsteps


def _ext_struct(ext_pot):
    if ext_pot is None:
        return _lib.Potential(), None
    return lower(ext_pot)


def nbody_integrate(ext_pot, masses, G, eps, w0, t0, t1, ts, ctrl):
    """Nbody_field through integrate_field (fields.py:115-155): ys[M,N,6], status[1], nsteps[3]."""
    tt = torch()
    P, _keep = _ext_struct(ext_pot)
    N, M = w0.shape[0], ts.shape[0]
    ys = empty((M, N, 6))
    status, nsteps = empty((1,), tt.int32), empty((3,), tt.int32)
    nbytes = _lib.lib().ssb_nbody_scratch_bytes(N)
    scratch = empty(((nbytes + 7) // 8,))
    _lib.check(_lib.lib().ssb_nbody_integrate_f64(C.byref(P), N, ptr(masses), float(G), float(eps), ptr(w0), float(t0), float(t1), ptr(ts), M, ctrl,
                                                  ptr(ys), ptr(status), ptr(nsteps), ptr(scratch), nbytes, stream_ptr()))
    return ys, status, nsteps


def nbody_term(ext_pot, masses, G, eps, t, y):
    P, _keep = _ext_struct(ext_pot)
    yd = to_dev(y).reshape(-1, 6)
    N = yd.shape[0]
    dy, scratch = empty((N, 6)), empty((3 * N,))
    _lib.check(_lib.lib().ssb_nbody_term_f64(C.byref(P), N, ptr(masses), float(G), float(eps), float(t), ptr(yd), ptr(dy), ptr(scratch), 24 * N,
                                             stream_ptr()))
    return dy


def variational(pot, order, w0, M0, M20, t0, t1, ctrl):
    """Variational equations along N orbits: wout[N,6], Mout[N,6,6], M2out[N,6,6,6] | None, status[N], nsteps[N,3]."""
    tt = torch()
    P, _keep = lower(pot)
    N = w0.shape[0]
    wout, Mout = empty((N, 6)), empty((N, 6, 6))
    M2out = empty((N, 6, 6, 6)) if order == 2 else None
    status, nsteps = empty((N,), tt.int32), empty((N, 3), tt.int32)
    _lib.check(_lib.lib().ssb_variational_f64(C.byref(P), int(order), N, ptr(w0), ptr(M0), ptr(M20), ptr(t0), float(t1), ctrl, ptr(wout), ptr(Mout),
                                              ptr(M2out), ptr(status), ptr(nsteps), stream_ptr()))
    return wout, Mout, M2out, status, nsteps


def variational_term(pot, order, t, y):
    P, _keep = lower(pot)
    yd = to_dev(y).reshape(-1)
    dy = empty(yd.shape)
    _lib.check(_lib.lib().ssb_variational_term_f64(C.byref(P), int(order), float(t), ptr(yd), ptr(dy), stream_ptr()))
    return dy


class DenseOrbits:
    """N dense solutions living on the device (the vmapped `diffrax.Solution`s of gen_stream_*_dense, main.py:376-430):
    every accepted step of every orbit recorded by ssb_orbit_record_f64; evaluate(t) interpolates all of them."""

    MAX_BYTES = 64 << 30

    def __init__(self, pot, w0, t0, t1, ctrl, rec_cap=None):
        tt = torch()
        P, _keep = lower(pot)
        self.N, self.solver = int(w0.shape[0]), int(ctrl.solver)
        if rec_cap is None:
            # size the record from a step-count pre-pass (the final-state kernel: same stepper, same step sequences), so that no orbit
            # can run out of slots - the reference's dense Solution holds up to max_steps steps
            _, _, ns = orbit_integrate(pot, w0, t0, t1, t1.reshape(-1, 1), ctrl, ts_per_orbit=1)
            rec_cap = max(1, int(ns[:, 1].max().item())) if self.N > 0 else 1
        self.rec_cap = int(rec_cap)
        nbytes = _lib.lib().ssb_record_bytes(self.N, self.rec_cap)
        if nbytes > self.MAX_BYTES:
            raise MemoryError(f"dense solutions of {self.N} orbits x {self.rec_cap} steps need {nbytes / 2**30:.0f} GiB; lower rec_cap or save "
                              "snapshots (integrate_orbit_batch_vmapped with ts[N,M]) instead")
        self.recs = empty(((nbytes + 7) // 8,))
        self.status, self.nsteps = empty((self.N,), tt.int32), empty((self.N, 3), tt.int32)
        _lib.check(_lib.lib().ssb_orbit_record_f64(C.byref(P), self.N, ptr(w0), ptr(t0), ptr(t1), ctrl, self.rec_cap, ptr(self.recs), nbytes,
                                                   ptr(self.status), ptr(self.nsteps), stream_ptr()))
        n_over = int((self.nsteps[:, 1] > self.rec_cap).sum().item()) if self.N > 0 else 0
        if n_over:            # only possible with an explicit rec_cap: those orbits carry status 1 and evaluate to +inf past their last slot
            import warnings
            warnings.warn(f"DenseOrbits: {n_over} of {self.N} orbits took more than rec_cap = {self.rec_cap} accepted steps; they have status 1 and "
                          "evaluate to +inf beyond the recorded part (raise rec_cap, or leave it None to size it automatically)")

    def evaluate(self, t):
        """States of all N orbits at time t (scalar) or at t[i] (array of length N): device tensor [N, 6]; +inf outside an orbit's interval."""
        tq = to_dev(np.atleast_1d(np.asarray(t.cpu() if hasattr(t, "cpu") else t, dtype=np.float64)))
        per = int(tq.shape[0] == self.N and self.N > 1)
        if not per and tq.shape[0] != 1:
            raise ValueError("evaluate(t): t must be a scalar or have one entry per orbit")
        ys = empty((self.N, 6))
        _lib.check(_lib.lib().ssb_orbit_record_eval_f64(self.solver, self.N, ptr(self.recs), self.rec_cap, ptr(tq), per, ptr(ys), stream_ptr()))
        return ys
