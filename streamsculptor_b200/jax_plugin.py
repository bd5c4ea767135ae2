"""jax side of the XLA FFI binding (csrc/ssb_xla_ffi.cc): registration of the custom-call targets and `jax.ffi.ffi_call` wrappers with the
reference's calling conventions, so that `Potential.integrate_orbit`, `gen_stream_vmapped` and `compute_perturbation_OTF` can stay inside
`jax.jit` (main.py:139-162, 343-368; perturbative.py:726-755), plus the `jax.custom_jvp` rule that restores forward-mode differentiation of
`integrate_orbit` with respect to `w0` (the reference solves with `adjoint=ForwardMode()`, main.py:160) from the state-transition matrix
kernel (`ssb_variational_f64`).

IMPORT-GUARDED: this image has no jax, so nothing here is imported by the package and nothing has run under XLA; `flatten_program` (pure
numpy / ctypes) is covered by tests/test_host_cpu.py, the C++ side is compile-checked against tools/xla_ffi_stub.  With jax installed:

    from streamsculptor_b200 import jax_plugin as jp
    jp.register("libssb200_ffi.so")                     # once per process
    ys, status, nsteps = jp.integrate_orbit(pot, w0, t0, t1, ts, solver=8)          # jittable; differentiable in w0 (forward mode)
"""
import ctypes as C

import numpy as np

from . import _lib

_REGISTERED = False
TARGETS = ("ssb_orbit_integrate", "ssb_variational", "ssb_gen_stream", "ssb_linear_response")


class FfiProgram(C.Structure):
    """Pointer-free part of a potential program: must match `struct ssb_ffi_program` in csrc/ssb_xla_ffi.cc."""
    _fields_ = [("n_comp", C.c_int32), ("n_track", C.c_int32), ("n_sh", C.c_int32), ("_pad", C.c_int32),
                ("comp", _lib.Component * _lib.MAX_COMP),
                ("track_kind", C.c_int32 * _lib.MAX_TRACK), ("track_n", C.c_int32 * _lib.MAX_TRACK),
                ("sh_n", C.c_int32 * _lib.MAX_SH), ("sh_profile", C.c_int32 * _lib.MAX_SH), ("sh_G", C.c_double * _lib.MAX_SH)]


def flatten_program(pot):
    """Potential object -> (attribute bytes as a uint8 array, list of table arrays in operand order).
    Per track: t[n], y[n,3], s[n,3] (knot slopes of cubic tracks by the interpax 'cubic' rule; zeros for linear tracks); per subhalo
    set: m, r_s, x0, v, t0, t_window.  Only host data is touched (no CUDA call), so this also runs where there is no GPU."""
    from . import _runtime as rt
    prog = rt.Program()
    pot._lower(prog, -1)
    if prog.psets or len(prog.comps) > _lib.MAX_COMP or len(prog.tracks) > _lib.MAX_TRACK:
        raise NotImplementedError("the FFI shim carries components, tracks and subhalo sets within the program limits; packed perturber sets are not carried yet")
    P = FfiProgram()
    P.n_comp, P.n_track, P.n_sh = len(prog.comps), len(prog.tracks), len(prog.shs)
    for i, (typ, params, track, sh, growth) in enumerate(prog.comps):
        P.comp[i].type, P.comp[i].track, P.comp[i].sh, P.comp[i].growth = typ, track, sh, growth
        for k, v in enumerate(params):
            P.comp[i].p[k] = v
    tables = []
    for i, t in enumerate(prog.tracks):
        P.track_kind[i], P.track_n[i] = t.kind, len(t.t_host)
        tables += [t.t_host, t.y_host, cubic_slopes(t.t_host, t.y_host) if t.kind == _lib.TRACK_CUBIC else np.zeros_like(t.y_host)]
    for i, s in enumerate(prog.shs):
        P.sh_n[i], P.sh_profile[i], P.sh_G[i] = s.n, s.profile, s.G
        tables += [s.host[k] for k in ("m", "rs", "x0", "v", "t0", "tw")]
    return np.frombuffer(bytes(P), dtype=np.uint8).copy(), tables


def cubic_slopes(t, y):
    """Knot slopes of interpax's method='cubic' (centred secant average, one-sided at the ends): what ssb_track_slopes_f64 computes."""
    t, y = np.asarray(t, dtype=np.float64), np.asarray(y, dtype=np.float64)
    dx = np.diff(t)
    sec = np.where(dx[:, None] == 0.0, 0.0, np.diff(y, axis=0) / np.where(dx == 0.0, 1.0, dx)[:, None])
    s = np.empty_like(y)
    s[0], s[-1] = sec[0], sec[-1]
    s[1:-1] = 0.5 * (sec[:-1] + sec[1:])
    return s


def _jax():
    try:
        import jax
        import jax.numpy as jnp
    except ImportError as e:          # pragma: no cover - this image has no jax
        raise ImportError("streamsculptor_b200.jax_plugin needs jax (with jax.ffi) and the FFI library built from csrc/ssb_xla_ffi.cc") from e
    return jax, jnp


def register(ffi_library="libssb200_ffi.so"):
    """jax.ffi.register_ffi_target for every handler of the shim (platform CUDA)."""
    global _REGISTERED
    jax, _ = _jax()
    lib = C.CDLL(ffi_library)
    for name in TARGETS:
        jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(getattr(lib, name + "_ffi")), platform="CUDA")
    _REGISTERED = True


def _ctrl_attrs(solver, rtol, atol, dtmin, dtmax, max_steps):
    return dict(solver=np.int32(solver), max_steps=np.int32(max_steps), rtol=float(rtol), atol=float(atol), dtmin=float(dtmin),
                dtmax=float(np.inf if dtmax is None else dtmax))


def integrate_orbit(pot, w0, t0, t1, ts, solver=8, rtol=1e-7, atol=1e-7, dtmin=0.3, dtmax=None, max_steps=10_000):
    """integrate_orbit_batch_vmapped (main.py:186-202) as an XLA custom call: ys[N,M,6], status[N], nsteps[N,3].  Differentiable with respect
    to w0 in forward mode (jax.jvp / jax.jacfwd) for final-state solves (ts[N,1] == t1) through `_final_state`."""
    jax, jnp = _jax()
    program, tables = flatten_program(pot)
    N, M = w0.shape[0], ts.shape[-1]
    out = (jax.ShapeDtypeStruct((N, M, 6), jnp.float64), jax.ShapeDtypeStruct((N,), jnp.int32), jax.ShapeDtypeStruct((N, 3), jnp.int32))
    return jax.ffi.ffi_call("ssb_orbit_integrate", out)(w0, t0, t1, ts, *[jnp.asarray(a) for a in tables], program=program,
                                                       **_ctrl_attrs(solver, rtol, atol, dtmin, dtmax, max_steps))


def final_state(pot, w0, t0, t1, **ctl):
    """w(t1) for N orbits started at (w0, t0): the particle solves of gen_stream_vmapped (main.py:349-368), with a custom JVP whose tangent
    is the state-transition matrix of the variational kernel: d w(t1) = M @ d w0, M = d w(t1) / d w0 (ssb_variational_f64, order 1)."""
    jax, jnp = _jax()
    program, tables = flatten_program(pot)
    tabs = [jnp.asarray(a) for a in tables]
    attrs = dict(program=program, t1=float(t1), **_ctrl_attrs(ctl.get("solver", 8), ctl.get("rtol", 1e-7), ctl.get("atol", 1e-7), ctl.get("dtmin", 0.3),
                                                              ctl.get("dtmax"), ctl.get("max_steps", 10_000)))

    def call(w):
        N = w.shape[0]
        out = (jax.ShapeDtypeStruct((N, 6), jnp.float64), jax.ShapeDtypeStruct((N, 6, 6), jnp.float64), jax.ShapeDtypeStruct((N,), jnp.int32),
               jax.ShapeDtypeStruct((N, 3), jnp.int32))
        return jax.ffi.ffi_call("ssb_variational", out)(w, t0, *tabs, **attrs)

    @jax.custom_jvp
    def f(w):
        return call(w)[0]

    @f.defjvp
    def f_jvp(primals, tangents):
        (w,), (dw,) = primals, tangents
        wout, M, _, _ = call(w)
        return wout, jnp.einsum("nak,nk->na", M, dw)

    return f(w0)


def gen_stream(pot, ts, prog_w0, Msat, seed_num, kvals, pot_release=None, solver=8, rtol=1e-7, atol=1e-7, dtmin=0.3, dtmax=None, max_steps=10_000):
    """gen_stream_vmapped (main.py:343-368) as ONE custom call: lead[Nts-1,6], trail[Nts-1,6]."""
    jax, jnp = _jax()
    program, tables = flatten_program(pot)
    program_r, tables_r = flatten_program(pot if pot_release is None else pot_release)
    Nts = ts.shape[0]
    n = Nts - 1
    scratch = int(_lib.lib().ssb_stream_scratch_bytes(Nts, int(max_steps)))
    out = (jax.ShapeDtypeStruct((2, n, 6), jnp.float64), jax.ShapeDtypeStruct((2, n), jnp.int32), jax.ShapeDtypeStruct((2, n, 3), jnp.int32),
           jax.ShapeDtypeStruct((scratch,), jnp.uint8))
    both, status, nsteps, _ = jax.ffi.ffi_call("ssb_gen_stream", out)(
        ts, prog_w0, Msat * jnp.ones(Nts), *[jnp.asarray(a) for a in tables + tables_r], program=program, program_release=program_r,
        kvals=np.asarray(kvals, dtype=np.float64), G=float(pot._G), seed=np.int64(0 if seed_num is None else seed_num),
        **_ctrl_attrs(solver, rtol, atol, dtmin, dtmax, max_steps))
    return both[0], both[1]


def linear_response(pot_base, subhalos, w0, t0, t1, solver=8, rtol=1e-6, atol=1e-6, dtmin=0.01, dtmax=None, max_steps=10_000):
    """compute_perturbation_OTF (perturbative.py:726-755): w[N,6], D[N,n_sh,12].  `subhalos`: a SubhaloLinePotential* object of this package."""
    jax, jnp = _jax()
    program, tables = flatten_program(pot_base)
    arr = subhalos._arrays
    sh_tabs = [arr.host[k] for k in ("m", "rs", "x0", "v", "t0", "tw")]
    N = w0.shape[0]
    scratch = int(_lib.lib().ssb_response_scratch_bytes(arr.n))
    out = (jax.ShapeDtypeStruct((N, 6), jnp.float64), jax.ShapeDtypeStruct((N, arr.n, 12), jnp.float64), jax.ShapeDtypeStruct((N,), jnp.int32),
           jax.ShapeDtypeStruct((N, 3), jnp.int32), jax.ShapeDtypeStruct((scratch,), jnp.uint8))
    w, D, status, nsteps, _ = jax.ffi.ffi_call("ssb_linear_response", out)(
        w0, t0, *[jnp.asarray(a) for a in tables + sh_tabs], program=program, n_sh=np.int32(arr.n), profile=np.int32(arr.profile), G=float(arr.G),
        t1=float(t1), **_ctrl_attrs(solver, rtol, atol, dtmin, dtmax, max_steps))
    return w, D
