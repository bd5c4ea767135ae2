"""ctypes loader for libssb200.so (C ABI: include/ssb200.h) and the build recipe.

The library is built IN-TREE (streamsculptor_b200/_lib/libssb200.so) with nvcc for sm_100a.
There is no CPU fallback: if the shared object is missing or no CUDA device is present, calls raise.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
_OUT = os.path.join(_HERE, "_lib", "libssb200.so")
_SOURCES = ["ssb_kernels.cu", "ssb_response.cu", "ssb_response2.cu", "ssb_shared.cu", "ssb_variational.cu", "ssb_host.cu"]
_HEADERS = ["ssb_common.cuh", "ssb_response_wa.cuh", "ssb_jet.cuh", "ssb_potential.cuh", "ssb_rk.cuh", "ssb_fastmath.cuh", "ssb_tableau.h", "ssb_logtab.h", "../../include/ssb200.h"]

MAX_COMP, MAX_TRACK, MAX_SH, MAX_PSET = 12, 4, 2, 1

NFW, HERNQUIST, MIYAMOTO, PLUMMER, ISOCHRONE, TRIAXNFW, UNIFORM_ACC, SUBHALOS, PERTURBERS, BAR, DEHNEN_BAR = range(11)
TRACK_LINEAR, TRACK_CUBIC = 0, 1
PROFILE_PLUMMER, PROFILE_HERNQUIST, PROFILE_NFW = 0, 1, 2

_dp = C.c_void_p


class Component(C.Structure):
    _fields_ = [("type", C.c_int32), ("track", C.c_int32), ("sh", C.c_int32), ("growth", C.c_int32), ("p", C.c_double * 8)]


class Track(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n", C.c_int32), ("t", _dp), ("y", _dp), ("s", _dp), ("t_first", C.c_double), ("inv_dt", C.c_double)]


class Subhalos(C.Structure):
    _fields_ = [("n", C.c_int32), ("profile", C.c_int32), ("G", C.c_double), ("m", _dp), ("rs", _dp), ("x0", _dp),
                ("v", _dp), ("t0", _dp), ("tw", _dp)]


class Perturbers(C.Structure):
    _fields_ = [("n", C.c_int32), ("n_knots", C.c_int32), ("profile", C.c_int32), ("_pad", C.c_int32), ("t", _dp), ("y", _dp), ("GM", _dp), ("rs", _dp)]


class Potential(C.Structure):
    _fields_ = [("n_comp", C.c_int32), ("n_track", C.c_int32), ("n_sh", C.c_int32), ("n_pset", C.c_int32),
                ("comp", Component * MAX_COMP), ("track", Track * MAX_TRACK), ("sh", Subhalos * MAX_SH), ("pset", Perturbers * MAX_PSET)]


class Ctrl(C.Structure):
    _fields_ = [("solver", C.c_int32), ("max_steps", C.c_int32), ("rtol", C.c_double), ("atol", C.c_double),
                ("dtmin", C.c_double), ("dtmax", C.c_double)]


class SSBError(RuntimeError):
    pass


def build(force=False, verbose=False, out=None):
    """nvcc -gencode arch=compute_100a,code=sm_100a -> streamsculptor_b200/_lib/libssb200.so"""
    global _OUT
    if out is not None:
        _OUT, force = out, True
    srcs = [os.path.join(_CSRC, s) for s in _SOURCES]
    deps = srcs + [os.path.join(_CSRC, h) for h in _HEADERS]
    if not force and os.path.exists(_OUT) and all(os.path.getmtime(_OUT) >= os.path.getmtime(d) for d in deps):
        return _OUT
    os.makedirs(os.path.dirname(_OUT), exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    # -split-compile 1 (no parallel split of a translation unit): 3 minutes instead of 1, but the split changes inlining / FMA-contraction
    # decisions inside the hot kernels - measured on one B200 box with the same source (profiles/r2_split_compile_ab.txt): stream kernel
    # 11.39 -> 10.84 ms, response 34.2 -> 31.7 ms (production batch 38.1 -> 33.9 ms), saving kernel 20.5 -> 19.5 ms.  SSB_NVCC_SPLIT=0 for quick builds.
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo", "--expt-relaxed-constexpr", "--threads", "0", "-split-compile", os.environ.get("SSB_NVCC_SPLIT", "1"),
           "-Xptxas", "-v" if verbose else "-O3", "-shared", "-Xcompiler", "-fPIC", "-o", _OUT] + os.environ.get("SSB_NVCC_FLAGS", "").split() + srcs
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise SSBError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return _OUT


_LIB = None
_i64, _i32, _dbl = C.c_int64, C.c_int32, C.c_double
_PP, _SP = C.POINTER(Potential), C.POINTER(Subhalos)

_SIGNATURES = {
    "ssb_abi_version": ([], C.c_int),
    "ssb_last_error": ([], C.c_char_p),
    "ssb_launch_count": ([], C.c_ulonglong),
    "ssb_potential_eval_f64": ([_PP, _i64, _dp, _dp, _dp, _dp, _dp, _dp], C.c_int),
    "ssb_subhalo_eval_f64": ([_SP, C.c_int, C.POINTER(C.c_double), _dbl, _dp, _dp, _dp], C.c_int),
    "ssb_track_slopes_f64": ([_i64, _dp, _dp, _dp, _dp], C.c_int),
    "ssb_track_eval_f64": ([C.POINTER(Track), _i64, _dp, _dp, _dp, _dp], C.c_int),
    "ssb_orbit_integrate_f64": ([_PP, _i64, _dp, _dp, _dp, _dp, _i32, _i32, Ctrl, _dp, _dp, _dp, _dp], C.c_int),
    "ssb_orbit_dense_f64": ([_PP, _dp, _dbl, _dbl, _dp, _i64, Ctrl, _dp, _dp, _dp, _dp, C.c_size_t, _dp], C.c_int),
    "ssb_scratch_bytes": ([_i32], C.c_size_t),
    "ssb_orbit_record_f64": ([_PP, _i64, _dp, _dp, _dp, Ctrl, _i32, _dp, C.c_size_t, _dp, _dp, _dp], C.c_int),
    "ssb_orbit_trace_f64": ([_PP, _i64, _dp, _dp, _dp, Ctrl, _i32, _dp, _dp, _dp, _dp, _dp], C.c_int),
    "ssb_orbit_record_eval_f64": ([_i32, _i64, _dp, _i32, _dp, _i32, _dp, _dp], C.c_int),
    "ssb_record_bytes": ([_i64, _i32], C.c_size_t),
    "ssb_orbit_dense_eval_f64": ([_i32, _dp, _dp, _i64, _dp, _dp], C.c_int),
    "ssb_release_spray_f64": ([_PP, _dbl, _i64, _dp, _dp, _dp, _dp, _i64, C.POINTER(C.c_double), _dp, _dp, _dp, _dp, _dp, _dp], C.c_int),
    "ssb_release_chen25_f64": ([_PP, _dbl, _i64, _dp, _dp, _dp, C.POINTER(C.c_uint32), C.POINTER(C.c_double), C.POINTER(C.c_double), _dp, _dp, _dp, _dp,
                                _dp, _dp], C.c_int),
    "ssb_release_jacobian_f64": ([_PP, _dbl, _i64, _dp, _dp, _dp, _dp, _i64, C.POINTER(C.c_double), _dp, _dp, _dp], C.c_int),
    "ssb_potential_third_f64": ([_PP, _i64, _dp, _dp, _dp, _dp], C.c_int),
    "ssb_gen_stream_f64": ([_PP, _PP, _dbl, _i64, _dp, _dp, _dp, _i64, C.POINTER(C.c_double), _dp, Ctrl, _i64, _i64, _i64, _dp, _dp,
                            _dp, _dp, _dp, C.c_size_t, _dp], C.c_int),
    "ssb_stream_scratch_bytes": ([_i64, _i32], C.c_size_t),
    "ssb_linear_response_f64": ([_PP, _SP, _i64, _dp, _dp, _dp, _dbl, Ctrl, _dp, _dp, _dp, _dp, _dp, C.c_size_t, _dp], C.c_int),
    "ssb_response_scratch_bytes": ([_i32], C.c_size_t),
    "ssb_linear_response_saveat_f64": ([_PP, _SP, _dp, _dp, _dp, _dbl, _dp, _i32, Ctrl, _dp, _dp, _dp, _dp, _dp, C.c_size_t, _dp], C.c_int),
    "ssb_response_saveat_scratch_bytes": ([_i32], C.c_size_t),
    "ssb_second_order_response_f64": ([_PP, _SP, _i64, _dp, _dp, _dp, _dp, _dbl, Ctrl, _dp, _dp, _dp, _dp, _dp, _dp, C.c_size_t, _dp], C.c_int),
    "ssb_second_order_response_ends_f64": ([_PP, _SP, _i64, _dp, _dp, _dp, _dp, _dp, Ctrl, _dp, _dp, _dp, _dp, _dp, _dp, C.c_size_t, _dp], C.c_int),
    "ssb_second_order_scratch_bytes": ([_i32], C.c_size_t),
    "ssb_second_order_term_f64": ([_PP, _SP, _dbl, _dp, _dp, _dp], C.c_int),
    "ssb_response_term_f64": ([_PP, _SP, _dbl, _dp, _dp, _dp], C.c_int),
    "ssb_shared_step_orbits_f64": ([_PP, _i64, _dp, _dbl, _dbl, Ctrl, _dp, _dp, _dp, _dp, C.c_size_t, _dp], C.c_int),
    "ssb_shared_scratch_bytes": ([_i64], C.c_size_t),
    "ssb_nbody_integrate_f64": ([_PP, _i32, _dp, _dbl, _dbl, _dp, _dbl, _dbl, _dp, _i32, Ctrl, _dp, _dp, _dp, _dp, C.c_size_t, _dp], C.c_int),
    "ssb_nbody_scratch_bytes": ([_i32], C.c_size_t),
    "ssb_nbody_term_f64": ([_PP, _i32, _dp, _dbl, _dbl, _dbl, _dp, _dp, _dp, C.c_size_t, _dp], C.c_int),
    "ssb_variational_f64": ([_PP, _i32, _i64, _dp, _dp, _dp, _dp, _dbl, Ctrl, _dp, _dp, _dp, _dp, _dp, _dp], C.c_int),
    "ssb_variational_term_f64": ([_PP, _i32, _dbl, _dp, _dp, _dp], C.c_int),
    "ssb_orbit_integrate_host": ([_PP, _i64, _dp, _dp, _dp, _dp, _i32, _i32, Ctrl, _dp, _dp, _dp], C.c_int),
    "ssb_gen_stream_host": ([_PP, _PP, _dbl, _i64, _dp, _dp, _dp, _i64, C.POINTER(C.c_double), _dp, Ctrl, _i64, _i64, _i64, _dp, _dp,
                             _dp, _dp], C.c_int),
    "ssb_linear_response_host": ([_PP, _SP, _i64, _dp, _dp, _dp, _dbl, Ctrl, _dp, _dp, _dp, _dp], C.c_int),
    "ssb_fp64_peak_probe": ([C.c_int, C.POINTER(C.c_double), _dp], C.c_int),
}
EXPORTED = tuple(_SIGNATURES)


def lib():
    """Load the shared object (building it first if the sources are newer).  Raises SSBError if impossible."""
    global _LIB
    if _LIB is None:
        path = os.environ.get("SSB_LIB_PATH", _OUT)      # A/B builds of the same sources (tools/bench_k1.py)
        if not os.path.exists(path) or os.environ.get("SSB_REBUILD"):
            path = build()
        L = C.CDLL(path)
        for name, (args, res) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes, fn.restype = args, res
        if L.ssb_abi_version() != 2:
            raise SSBError("libssb200 ABI mismatch")
        _LIB = L
    return _LIB


def check(code):
    if code != 0:
        msg = lib().ssb_last_error().decode(errors="replace")
        raise SSBError(f"libssb200 error {code}: {msg}")


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise SSBError("streamsculptor_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch
