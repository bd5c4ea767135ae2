"""Host-side logic that needs no GPU: the C-ABI library loads and exports every symbol of include/ssb200.h, struct layouts
match the header, potential lowering, argument validation through the C ABI (error paths return before any CUDA call),
and the multi-process sharding / gather logic on the gloo backend (world_size 2)."""
import ctypes as C
import os
import re
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from streamsculptor_b200 import _lib
    path = _lib.build()
    L = C.CDLL(path)
    header = open(os.path.join(ROOT, "include", "ssb200.h")).read()
    declared = set(re.findall(r"\b(ssb_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/ssb200.h but not exported"
    assert set(_lib.EXPORTED) == declared
    assert L.ssb_abi_version() == 2          # round 2: growth factors, perturber sets


def test_struct_layouts_match_header():
    from streamsculptor_b200 import _lib
    assert C.sizeof(_lib.Component) == 16 + 64
    assert C.sizeof(_lib.Track) == 8 + 3 * 8 + 2 * 8
    assert C.sizeof(_lib.Subhalos) == 16 + 6 * 8
    assert C.sizeof(_lib.Perturbers) == 16 + 4 * 8
    assert C.sizeof(_lib.Potential) == 16 + 12 * 80 + 4 * 48 + 2 * 64 + 48
    assert C.sizeof(_lib.Ctrl) == 8 + 4 * 8
    # the CUDA side agrees (scratch sizes are computed from the same constants)
    L = _lib.lib()
    assert L.ssb_scratch_bytes(100) == 8 * (8 + 100 * 64)
    assert L.ssb_response_scratch_bytes(1000) >= 2 * 148 * 2 * 6 * 2000 * 8


def test_argument_validation_without_gpu():
    from streamsculptor_b200 import _lib
    L = _lib.lib()
    P = _lib.Potential()
    P.n_comp = 1
    P.comp[0].type = 99
    ctrl = _lib.Ctrl(solver=8, max_steps=10, rtol=1e-7, atol=1e-7, dtmin=0.3, dtmax=np.inf)
    z = C.c_void_p(0)
    assert L.ssb_orbit_integrate_f64(C.byref(P), 4, z, z, z, z, 1, 1, ctrl, z, z, z, z) == -2          # unsupported component
    assert b"component type" in L.ssb_last_error()
    P.comp[0].type = _lib.NFW
    assert L.ssb_orbit_integrate_f64(C.byref(P), 4, z, z, z, z, 1, 1, ctrl, z, z, z, z) == -1 and b"missing track" in L.ssb_last_error()   # track 0 of 0
    P.comp[0].track = -1
    bad = _lib.Ctrl(solver=4, max_steps=10, rtol=1e-7, atol=1e-7, dtmin=0.3, dtmax=np.inf)
    assert L.ssb_orbit_integrate_f64(C.byref(P), 4, z, z, z, z, 1, 1, bad, z, z, z, z) == -2           # Tsit5 & co are not on the path
    assert L.ssb_orbit_integrate_f64(C.byref(P), 4, z, z, z, z, 1, 1, ctrl, z, z, z, z) == -1          # NULL arrays
    assert L.ssb_orbit_integrate_f64(C.byref(P), 0, z, z, z, z, 1, 1, ctrl, z, z, z, z) == 0           # empty batch is a no-op
    assert L.ssb_orbit_integrate_f64(C.byref(P), -1, z, z, z, z, 1, 1, ctrl, z, z, z, z) == -1
    P.comp[0].type = _lib.UNIFORM_ACC
    P.comp[0].track = -1
    assert L.ssb_potential_eval_f64(C.byref(P), 1, z, z, z, z, z, z) == -1                               # uniform acceleration without a track
    assert L.ssb_gen_stream_host(C.byref(P), C.byref(P), 1.0, 1, z, z, z, 0, None, z, ctrl, 0, 1, 0, z, z, z, z) == -1   # Nts < 2


def test_lowering_flattens_the_object_tree():
    import streamsculptor_b200 as ssc
    from streamsculptor_b200 import _lib, _runtime as rt
    P = ssc.potential
    gala = P.GalaMilkyWayPotential(units=ssc.usys)
    prog = rt.Program()
    gala._lower(prog, -1)
    assert [c[0] for c in prog.comps] == [_lib.MIYAMOTO, _lib.HERNQUIST, _lib.HERNQUIST, _lib.NFW]        # potential.py:413 order
    assert prog.comps[0][1][:3] == [ssc.G_KPC_MYR_MSUN * 6.80e10, 3.0, 0.28]
    assert prog.comps[3][1][:2] == [ssc.G_KPC_MYR_MSUN * 5.4e11, 15.62]
    mn3 = P.MN3ExponentialDiskPotential(m=5e10, h_R=3.0, h_z=0.3, units=ssc.usys)
    prog = rt.Program(); mn3._lower(prog, -1)
    assert len(prog.comps) == 3 and all(c[0] == _lib.MIYAMOTO for c in prog.comps)
    assert abs(sum(mn3._ms) / 5e10 - 1.0) < 0.5                                                         # the three disks share the mass
    assert P.NFWPotential(m=1.0, r_s=1.0)._G == 1.0                                                     # units=None -> dimensionless (main.py:23-28)
    with pytest.raises(NotImplementedError):
        P.TimeDepTranslatingPotential(P.PlummerPotential(m=1.0, r_s=1.0, units=ssc.usys), center_spl=lambda t: np.zeros(3), units=ssc.usys)
    with pytest.raises(NotImplementedError):
        P.SubhaloLinePotentialCustom_fromFunc(func=P.Isochrone, m=[1.0], r_s=[1.0], subhalo_x0=np.zeros((1, 3)), subhalo_v=np.zeros((1, 3)),
                                              subhalo_t0=[0.0], t_window=1.0, units=ssc.usys)
    for name in ("CustomPotential", "ZhaoPotential"):
        with pytest.raises(NotImplementedError):
            getattr(P, name)()
    with pytest.raises(NotImplementedError):            # GrowingPotential: the growth factor must be tabulated (potential.py:464-477 takes any callable)
        P.GrowingPotential(P.NFWPotential(m=1.0, r_s=1.0, units=ssc.usys), growth_func=lambda t: 1.0, units=ssc.usys)
    from streamsculptor_b200 import _lib, _runtime as rt
    grow = P.GrowingPotential(P.Potential_Combine([P.NFWPotential(m=1e12, r_s=20.0, units=ssc.usys), P.PlummerPotential(m=1e10, r_s=1.0, units=ssc.usys)],
                                                  units=ssc.usys), growth_func=(np.linspace(-3000.0, 0.0, 11), np.linspace(0.5, 1.0, 11)), units=ssc.usys)
    prog = rt.Program()
    P.Potential_Combine([P.HernquistPotential(m=5e9, r_s=1.0, units=ssc.usys), grow], units=ssc.usys)._lower(prog, -1)
    assert [c[0] for c in prog.comps] == [_lib.HERNQUIST, _lib.NFW, _lib.PLUMMER] and [c[4] for c in prog.comps] == [0, 1, 1] and len(prog.tracks) == 1
    with pytest.raises(NotImplementedError):            # no growth factor on force-only components
        P.GrowingPotential(P.UniformAcceleration(ssc.LinearTrack(np.array([0.0, 1.0]), np.zeros((2, 3))), units=ssc.usys),
                           growth_func=(np.array([0.0, 1.0]), np.ones(2)), units=ssc.usys)._lower(rt.Program(), -1)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import streamsculptor_b200 as ssc
    from streamsculptor_b200._lib import SSBError
    pot = ssc.potential.NFWPotential(m=1e12, r_s=20.0, units=ssc.usys)
    with pytest.raises(SSBError):
        pot.gradient(np.array([1.0, 2.0, 3.0]), 0.0)
    with pytest.raises(SSBError):
        pot.integrate_orbit(w0=np.ones(6), ts=np.array([0.0, 1.0]))
    # the product package never imports the oracle
    src = "".join(open(os.path.join(ROOT, "streamsculptor_b200", f)).read() for f in os.listdir(os.path.join(ROOT, "streamsculptor_b200")) if f.endswith(".py"))
    assert "import oracle" not in src and "from oracle" not in src


def test_shard_math():
    from streamsculptor_b200 import parallel as par
    for n in (0, 1, 7, 1000, 1001):
        for world in (1, 2, 3, 8):
            idx = [par.shard_indices(n, r, world) for r in range(world)]
            assert [len(i) for i in idx] == [par.shard_count(n, r, world) for r in range(world)]
            assert sorted(np.concatenate(idx).tolist()) == list(range(n))


_WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import numpy as np, torch
import torch.distributed as dist
import oracle as O
from common import mw3_oracle
from streamsculptor_b200 import parallel as par
rank, world = par.init_from_env("gloo")
ts = np.linspace(-600.0, 0.0, 38)
w0 = [12.0, 3.0, -6.0, -0.05, 0.15, 0.03]
nr = np.random.Generator(np.random.PCG64(3)).standard_normal((38, 4))
orc = mw3_oracle()
lead_all, trail_all, _, _ = orc.gen_stream(ts, w0, 1e4, 0, solver=5, normals=nr)
def compute(rank, world, n_local):          # the oracle stands in for the CUDA call: host logic only
    sel = par.shard_indices(len(ts) - 1, rank, world)
    assert len(sel) == n_local
    return torch.from_numpy(lead_all[sel].copy()), torch.from_numpy(trail_all[sel].copy())
lead, trail = par.gen_stream_sharded(None, ts, w0, 1e4, 0, None, rank, world, compute=compute)
assert lead.shape == (37, 6) and np.array_equal(lead.numpy(), lead_all) and np.array_equal(trail.numpy(), trail_all)
# C4 sharding: particles dealt out interleaved, only final states + response summaries are exchanged
sh = dict(m=np.ones(3), rs=np.array([0.2, 0.3, 0.4]), x0=np.array([[5., 5, 5], [-5, 5, 0], [0, -8, 3]]), v=np.full((3, 3), 0.1), t0=np.array([-500., -300, -100]))
osh = O.Program().subhalos(O.PR_HERNQUIST, sh["m"], sh["rs"], sh["x0"], sh["v"], sh["t0"], 150.0)
w0r = np.vstack([lead_all[:9] * 0 + np.asarray(w0)]) + np.arange(9)[:, None] * 1e-3
t0r = ts[:9].copy()
w_all, D_all, _, _ = O.linear_response(orc, osh, w0r, t0r, 0.0, solver=5, rtol=1e-7, atol=1e-7)
def rcompute(w0_l, t0_l):
    wl, Dl, _, _ = O.linear_response(orc, osh, w0_l.numpy(), t0_l.numpy(), 0.0, solver=5, rtol=1e-7, atol=1e-7)
    return torch.from_numpy(wl), torch.from_numpy(Dl)
Msh = np.array([1e6, 3e7, 2e8])
wg, dg, D_l, sel = par.linear_response_sharded(None, None, torch.from_numpy(w0r), torch.from_numpy(t0r), 0.0, None, rank, world, Msh, dr_s=[0.1, 0.0, -0.2], compute=rcompute)
ref = np.einsum("s,nsk->nk", Msh, D_all[:, :, :6]) + np.einsum("s,nsk->nk", Msh * np.array([0.1, 0.0, -0.2]), D_all[:, :, 6:])
assert np.array_equal(wg.numpy(), w_all) and np.allclose(dg.numpy(), ref, rtol=1e-13, atol=0) and np.array_equal(D_l.numpy(), D_all[sel.numpy()])
s = par.allreduce_sum(torch.tensor([float(rank + 1)]), world)
assert s.item() == world * (world + 1) / 2
dist.barrier(); dist.destroy_process_group()
open(os.path.join(os.path.dirname(os.path.abspath(__file__)), f"rank{rank}.ok"), "w").write("ok")      # stdout of the ranks interleaves
'''


def test_gloo_world_size_2_shard_and_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % {"root": ROOT})
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", str(port),
           str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert (tmp_path / "rank0.ok").exists() and (tmp_path / "rank1.ok").exists(), res.stdout[-2000:]


def test_bench_reference_arm_runs_on_cpu():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"], capture_output=True,
                         text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    import json
    out = res.stdout.strip().splitlines()
    assert len(out) == 1, "bench.py must print exactly ONE line on stdout (library chatter goes to stderr)"
    line = json.loads(out[0])
    assert line["impl"] == "reference" and line["unit"] == "particle-steps/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["config"]["workload"].startswith("C2: ONE 1000000-particle mock stream") and line["metric"] == "fp64 particle-steps/sec"
    assert line["scaling"] == "strong"


def test_argument_validation_of_the_round1_entry_points():
    """A17 / A16 entry points: validation happens before any CUDA call (no GPU needed)."""
    from streamsculptor_b200 import _lib
    L = _lib.lib()
    P = _lib.Potential()
    P.n_comp = 1
    P.comp[0].type = _lib.NFW
    P.comp[0].track = -1
    ctrl = _lib.Ctrl(solver=8, max_steps=10, rtol=1e-7, atol=1e-7, dtmin=0.3, dtmax=np.inf)
    bad = _lib.Ctrl(solver=3, max_steps=10, rtol=1e-7, atol=1e-7, dtmin=0.3, dtmax=np.inf)
    z = C.c_void_p(0)
    assert L.ssb_variational_f64(C.byref(P), 1, 0, z, z, z, z, 0.0, ctrl, z, z, z, z, z, z) == 0            # empty batch
    assert L.ssb_variational_f64(C.byref(P), 3, 4, z, z, z, z, 0.0, ctrl, z, z, z, z, z, z) == -2           # order 3 does not exist
    assert L.ssb_variational_f64(C.byref(P), 1, 4, z, z, z, z, 0.0, ctrl, z, z, z, z, z, z) == -1           # NULL arrays
    assert L.ssb_variational_f64(C.byref(P), 1, 4, z, z, z, z, 0.0, bad, z, z, z, z, z, z) == -2            # solver
    assert L.ssb_shared_step_orbits_f64(C.byref(P), 0, z, 0.0, 1.0, ctrl, z, z, z, z, 0, z) == 0
    assert L.ssb_shared_step_orbits_f64(C.byref(P), 5, z, 0.0, 1.0, ctrl, z, z, z, z, 0, z) == -1
    assert L.ssb_shared_scratch_bytes(1000) >= 18 * 1000 * 8
    assert L.ssb_nbody_integrate_f64(None, 2000, z, 1.0, 0.1, z, 0.0, 1.0, z, 1, ctrl, z, z, z, z, 0, z) == -1 and b"1024" in L.ssb_last_error()
    assert L.ssb_nbody_integrate_f64(None, 3, z, 1.0, 0.1, z, 0.0, 1.0, z, 1, ctrl, z, z, z, z, 0, z) == -1           # NULL masses
    assert L.ssb_nbody_scratch_bytes(100) == 8 * 3 * 100 * 18


def test_run_nonlinear_sim_forwards_the_model_parameters(monkeypatch):
    """GenerateMassRadiusPerturbation_Chen25.run_nonlinear_sim (perturbative.py:775-813) is glue: the stream parameters of the base model go to
    gen_stream_vmapped_with_pert_Chen25_fixed_prog and (lead, trail) come back stacked.  Checked without a GPU by intercepting the call."""
    import types
    import numpy as np
    from streamsculptor_b200 import perturbative as pt, streamhelpers as sh
    seen = {}

    def fake(**kw):
        seen.update(kw)
        return np.zeros((3, 6)), np.ones((3, 6))
    monkeypatch.setattr(sh, "gen_stream_vmapped_with_pert_Chen25_fixed_prog", fake)
    gen = object.__new__(pt.GenerateMassRadiusPerturbation_Chen25)
    gen.potential_base, gen.potential_perturbation = "BASE", "PERT"
    gen.BaseStreamModel = types.SimpleNamespace(prog_pot="PROG", ts=np.linspace(-10.0, 0.0, 4), key=7, Msat=1e4)
    gen.base_stream = types.SimpleNamespace(prog_w0=[1.0, 2, 3, 4, 5, 6])
    out = gen.run_nonlinear_sim(rtol=1e-8, atol=1e-9, dtmin=0.02, max_steps=123)
    assert out.shape == (6, 6) and (out[:3] == 0).all() and (out[3:] == 1).all()
    assert seen["pot_base"] == "BASE" and seen["pot_pert"] == "PERT" and seen["prog_pot"] == "PROG" and seen["key"] == 7 and seen["Msat"] == 1e4
    assert seen["prog_w0"] == [1.0, 2, 3, 4, 5, 6] and seen["rtol"] == 1e-8 and seen["atol"] == 1e-9 and seen["dtmin"] == 0.02 and seen["max_steps"] == 123
    assert type(seen["solver"]).__name__ == "Dopri8"
    gen.run_nonlinear_sim(pot_pert="OTHER")
    assert seen["pot_pert"] == "OTHER"


def test_track_summaries_follow_the_reference_conventions():
    """computed_binned_track / compute_stream_length / compute_length_oscillations (streamhelpers.py:201-304): host post-processing of the
    kernels' output, checked against a particle-by-particle restatement of the reference's digitize / scan / nanmean recipe."""
    import types
    import numpy as np
    import streamsculptor_b200 as ssc
    rng = np.random.default_rng(0)
    n, bins = 500, 12
    stream, phi1 = rng.normal(size=(n, 6)), rng.uniform(-40.0, 25.0, n)
    edges = np.linspace(phi1.min(), phi1.max(), bins)
    want = np.full((bins - 1, 6), np.nan)
    sums, cnt = np.zeros((bins + 1, 6)), np.zeros(bins + 1)
    for x, w in zip(phi1, stream):
        d = int(np.searchsorted(edges, x, side="right"))        # jnp.digitize: edges[d-1] <= x < edges[d]
        sums[d] += w; cnt[d] += 1
    for b in range(bins - 1):                                   # the scan visits bin indices 0 .. bins-2 only
        if cnt[b] > 0:
            want[b] = sums[b] / cnt[b]
    got = ssc.computed_binned_track(stream, phi1, bins=bins)
    assert got.shape == (bins - 1, 6) and np.isnan(got[0]).all() and np.allclose(got[1:], want[1:], rtol=1e-13, atol=0, equal_nan=True)
    seg = np.linalg.norm(want[1:, :3] - want[:-1, :3], axis=1)
    assert np.isclose(ssc.compute_stream_length(stream, phi1, bins=bins), np.nansum(seg), rtol=1e-13)
    # length oscillations: three orbits integrated backwards in one batch; a fake potential supplies known separations
    calls = {}

    def batch(w0=None, ts=None, t0=None, t1=None):
        calls.update(w0=np.asarray(w0), ts=np.asarray(ts), t0=t0, t1=t1)
        ys = np.zeros((3, len(ts), 6))
        ys[0, :, 0] = 3.0 * (1.0 - ts / 100.0)                  # lead - progenitor separation grows into the past
        ys[1, :, 1] = 4.0 * (1.0 - ts / 100.0)
        return types.SimpleNamespace(ys=ys)
    pot = types.SimpleNamespace(integrate_orbit_batch_vmapped=batch)
    res = ssc.compute_length_oscillations(pot, np.arange(6.0), np.ones(6), 2 * np.ones(6), t_age=500.0, length_today=10.0)
    assert calls["t0"] == 0.0 and calls["t1"] == -500.0 and calls["w0"].shape == (3, 6) and np.array_equal(calls["w0"][2], np.arange(6.0))
    assert res["ts"][0] == -500.0 and res["ts"][-1] == 0.0 and len(res["ts"]) == 2000
    assert np.isclose(res["length_func"][-1], 10.0) and np.isclose(res["length_func"][0], 10.0 * 6.0)     # sqrt(3^2 + 4^2) * (1 + 5) / 5


def test_sample_from_1d_pdf_uses_jax_uniform_draws():
    """sample_from_1D_pdf (streamhelpers.py:306-349).  The host-side threefry / uniform recipe is checked against the oracle's generator,
    which the reference's printed release Jacobian and stream pin (goldens D8, OC): same Threefry-2x32 words, and the uniform it builds
    maps to the oracle's jax.random.normal through sqrt(2) erfinv(u (1 - lo) + lo)."""
    import numpy as np
    from scipy.special import erfinv
    import oracle as O
    from streamsculptor_b200 import streamhelpers as sh
    for k0, k1, c0, c1 in ((0, 0, 0, 0), (0x13198a2e, 0x03707344, 0x243f6a88, 0x85a308d3), (0xffffffff,) * 4):
        a = sh._threefry2x32(k0, k1, np.array([c0]), np.array([c1]))
        assert [int(a[0][0]), int(a[1][0])] == [int(v) for v in O.threefry2x32(k0, k1, c0, c1)]
    lo = np.nextafter(-1.0, 0.0)
    for seed in (0, 493, 90 * 1499, 2**40 + 17):
        u = sh.jax_uniform(seed, 1)[0]
        assert 0.0 <= u < 1.0 and abs(np.sqrt(2.0) * erfinv(max(lo, u * (1.0 - lo) + lo)) - O.normal1(seed)) < 1e-14
    # inverse-CDF sampling: a flat density returns the draws themselves on the grid's scale; a one-sided density stays in its support
    x = np.linspace(2.0, 6.0, 4001)
    s = sh.sample_from_1D_pdf(x, np.ones_like(x), key=7, num_samples=1000)
    assert np.abs(s - (2.0 + 4.0 * sh.jax_uniform(7, 1000))).max() < 2e-3
    y = np.where(x > 4.0, (x - 4.0) ** 2, 0.0)
    s = sh.sample_from_1D_pdf(x, y, key=(0, 7), num_samples=2000)
    assert s.min() >= 4.0 and s.max() <= 6.0 and abs(s.mean() - 5.5) < 0.03       # mean of 3 (x-4)^2 / 8 on [4, 6]
    assert np.array_equal(s, sh.sample_from_1D_pdf(x, y, key=7, num_samples=2000))  # PRNGKey(7) == key words (0, 7)


def test_xla_ffi_shim_compiles_against_the_api_stub():
    """csrc/ssb_xla_ffi.cc cannot be built for real here (no jaxlib headers); it must at least parse and type-check against the stub of the
    public XLA FFI API in tools/xla_ffi_stub (Buffer / RemainingArgs / Span / Error / Bind() builder)."""
    src = os.path.join(ROOT, "streamsculptor_b200", "csrc", "ssb_xla_ffi.cc")
    res = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Wno-comment", "-Werror", "-I", os.path.join(ROOT, "tools", "xla_ffi_stub"),
                          "-I", os.path.join(ROOT, "include"), src], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]
    # and the guard really is what keeps it out of normal builds: without the include path the translation unit is empty
    res = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), src], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]


def test_jax_plugin_flattens_programs_without_jax_or_cuda():
    """jax_plugin.flatten_program: pointer-free attribute bytes + table operands in the order the FFI handlers consume them."""
    import ctypes as C
    import numpy as np
    import streamsculptor_b200 as ssc
    from streamsculptor_b200 import _lib, jax_plugin as jp
    P = ssc.potential
    t = np.linspace(-3000.0, 0.0, 50)
    y = np.stack([np.sin(t / 500.0), np.cos(t / 700.0), t / 3000.0], axis=1)
    nsh = 7
    pot = P.Potential_Combine([
        P.NFWPotential(m=5.4e11, r_s=15.62, units=ssc.usys),
        P.TimeDepTranslatingPotential(P.PlummerPotential(m=1.5e11, r_s=10.8, units=ssc.usys), ssc.CubicTrack(t, y), units=ssc.usys),
        P.SubhaloLinePotentialCustom_fromFunc(func=P.HernquistPotential, m=np.ones(nsh), r_s=np.full(nsh, 0.3), subhalo_x0=np.zeros((nsh, 3)),
                                              subhalo_v=np.ones((nsh, 3)), subhalo_t0=np.linspace(-2000, -100, nsh), t_window=150.0, units=ssc.usys)],
        units=ssc.usys)
    attr, tables = jp.flatten_program(pot)
    assert attr.dtype == np.uint8 and attr.size == C.sizeof(jp.FfiProgram) == 16 + C.sizeof(_lib.Component) * _lib.MAX_COMP + 8 * _lib.MAX_TRACK + 16 * _lib.MAX_SH
    prog = jp.FfiProgram.from_buffer_copy(attr.tobytes())
    assert (prog.n_comp, prog.n_track, prog.n_sh) == (3, 1, 1)
    assert prog.comp[0].type == _lib.NFW and prog.comp[1].type == _lib.PLUMMER and prog.comp[1].track == 0 and prog.comp[2].type == _lib.SUBHALOS
    assert prog.track_kind[0] == _lib.TRACK_CUBIC and prog.track_n[0] == 50 and prog.sh_n[0] == nsh and prog.sh_profile[0] == _lib.PROFILE_HERNQUIST
    assert len(tables) == 3 + 6 and tables[0].shape == (50,) and tables[1].shape == (50, 3) and tables[2].shape == (50, 3) and tables[5].shape == (nsh, 3)
    # knot slopes: the interpax 'cubic' rule (what ssb_track_slopes_f64 computes on the device)
    s = tables[2]
    assert np.allclose(s[0], (y[1] - y[0]) / (t[1] - t[0])) and np.allclose(s[10], 0.5 * ((y[10] - y[9]) / (t[10] - t[9]) + (y[11] - y[10]) / (t[11] - t[10])))
    with pytest.raises(ImportError):
        jp.register("libssb200_ffi.so")


def test_many_moving_perturbers_pack_into_a_perturber_set():
    """BASELINE config 5 needs 100 moving perturbers; a program holds 12 components / 4 tracks.  A Potential_Combine of many
    TimeDepTranslatingPotential objects on one time grid (the reference's way to write them, potential.py:448-462) is lowered to ONE
    perturber-set component; other tracks and growth factors keep working and are re-indexed."""
    import numpy as np
    import streamsculptor_b200 as ssc
    from streamsculptor_b200 import _lib, _runtime as rt
    P = ssc.potential
    t = np.linspace(-3000.0, 0.0, 64)
    rng = np.random.default_rng(0)
    n = 30
    cen = rng.normal(size=(64, n, 3)) * 20.0
    ms, rs = 10 ** rng.uniform(8, 10, n), rng.uniform(0.5, 3.0, n)
    other = ssc.LinearTrack(np.linspace(-3000.0, 0.0, 17), rng.normal(size=(17, 3)))
    comps = [P.HernquistPotential(m=5e9, r_s=1.0, units=ssc.usys), P.MiyamotoNagaiDisk(m=6.8e10, a=3.0, b=0.28, units=ssc.usys),
             P.GrowingPotential(P.NFWPotential(m=5.4e11, r_s=15.62, units=ssc.usys), (np.array([-3000.0, 0.0]), np.array([0.5, 1.0])), units=ssc.usys),
             P.TimeDepTranslatingPotential(P.HernquistPotential(m=1e11, r_s=8.0, units=ssc.usys), other, units=ssc.usys)]
    comps += [P.TimeDepTranslatingPotential(P.PlummerPotential(m=ms[i], r_s=rs[i], units=ssc.usys), ssc.LinearTrack(t, cen[:, i]), units=ssc.usys) for i in range(n)]
    prog = rt.Program()
    P.Potential_Combine(comps, units=ssc.usys)._lower(prog, -1)
    assert len(prog.comps) == 4 + n and len(prog.tracks) == 2 + n
    assert prog.pack_perturbers()
    assert [c[0] for c in prog.comps] == [_lib.HERNQUIST, _lib.MIYAMOTO, _lib.NFW, _lib.HERNQUIST, _lib.PERTURBERS] and len(prog.tracks) == 2 and len(prog.psets) == 1
    assert prog.comps[2][4] == 1 and prog.comps[3][2] == 1          # growth track first (index 0 -> growth = 1), the lone translating Hernquist on track 1
    ps = prog.psets[0]
    assert ps.n == n and ps.n_knots == 64 and ps.profile == _lib.PROFILE_PLUMMER and ps.host["y"].shape == (64, n, 3)
    G = comps[0]._G
    assert np.allclose(ps.host["GM"], G * ms) and np.allclose(ps.host["rs"], rs) and np.array_equal(ps.host["y"][:, 7], cen[:, 7])
    # the explicit class gives the same arrays
    direct = P.PerturberSetPotential(P.PlummerPotential, ms, rs, t, cen, units=ssc.usys)
    assert np.allclose(direct._arrays.host["GM"], ps.host["GM"]) and np.array_equal(direct._arrays.host["y"], ps.host["y"])


def test_impact_generator_random_recipes():
    """jax.random on the host for the production driver's sampler (GenerateImpactParams.py:44,87-147): split + randint against the oracle's
    recipe (the one golden D8 hangs on), normal against the oracle's, choice and uniform by their distributions."""
    import numpy as np
    import oracle as O
    from streamsculptor_b200 import GenerateImpactParams as G
    for seed in (0, 493, 583, 9302, 2**33 + 5):
        assert np.array_equal(G.jax_randint(seed, 5, 0, 1000), O.randint5(seed, 0, 1000))
        assert abs(G.jax_normal(seed, 1)[0] - O.normal1(seed)) < 1e-14
    k = G.jax_split(11, 7)
    assert k.shape == (7, 2) and k.dtype == np.uint32 and len({tuple(r) for r in k}) == 7
    assert np.array_equal(G.jax_split(11, 7), k) and not np.array_equal(G.jax_split(12, 7), k)
    z = G.jax_normal(k[0], 20000)
    assert abs(z.mean()) < 0.03 and abs(z.std() - 1.0) < 0.03
    u = G.jax_uniform_range(k[1], 20000, -2.0, np.full(20000, 3.0))
    assert u.min() >= -2.0 and u.max() < 3.0 and abs(u.mean() - 0.5) < 0.05
    a = np.arange(4.0)
    c = G.jax_choice(k[2], a, 40000, [0.1, 0.2, 0.3, 0.4])
    assert np.abs(np.bincount(c.astype(int), minlength=4) / 40000 - [0.1, 0.2, 0.3, 0.4]).max() < 0.01


def test_oracle_rotating_bars_against_the_formulas():
    """The oracle's BarPotential / DehnenBarPotential (potential.py:178-222) against a direct numpy evaluation of the reference's formulas
    (rotation matrix, arctan2 and all), and its autodiff gradient against central differences of that."""
    import numpy as np
    import oracle as O
    G = O.Program().G

    def bar(x, t, m=1e10, a=3.5, b=0.5, c=0.6, Om=0.04):
        ang = -Om * t
        R = np.array([[np.cos(ang), -np.sin(ang), 0.0], [np.sin(ang), np.cos(ang), 0.0], [0.0, 0.0, 1.0]])
        xc = R @ x
        zz = b + np.sqrt(c ** 2 + xc[2] ** 2)
        Tp = np.sqrt((a + xc[0]) ** 2 + xc[1] ** 2 + zz ** 2)
        Tm = np.sqrt((a - xc[0]) ** 2 + xc[1] ** 2 + zz ** 2)
        return G * m / (2.0 * a) * np.log((xc[0] - a + Tm) / (xc[0] + a + Tp))

    def dehnen(x, t, alpha=0.01, v0=0.22, R0=8.0, Rb=3.4, phib=0.4, Om=0.05):
        phi = np.arctan2(x[1], x[0]); R = np.hypot(x[0], x[1]); r = np.linalg.norm(x)
        U = -(r / Rb) ** (-3) if r >= Rb else (r / Rb) ** 3 - 2.0
        return alpha * (v0 ** 2 / 3) * (R0 / Rb) ** 3 * (R ** 2 / r ** 2) * U * np.cos(2 * (phi - phib - Om * t))

    rng = np.random.default_rng(4)
    xyz = rng.normal(size=(60, 3)) * np.array([6.0, 6.0, 2.0])
    t = rng.uniform(-3000.0, 0.0, 60)
    for f, prog in ((bar, O.Program().bar(1e10, 3.5, 0.5, 0.6, 0.04)), (dehnen, O.Program().dehnen_bar(0.01, 0.22, 8.0, 3.4, 0.4, 0.05))):
        want = np.array([f(x, tt) for x, tt in zip(xyz, t)])
        got = prog.potential(xyz, t)
        assert np.abs(got - want).max() <= 1e-13 * np.abs(want).max()
        g = prog.gradient(xyz, t)
        h = 1e-5
        fd = np.array([[(f(x + h * e, tt) - f(x - h * e, tt)) / (2 * h) for e in np.eye(3)] for x, tt in zip(xyz, t)])
        assert np.abs(g - fd).max() <= 1e-7 * np.abs(fd).max()


def test_device_jet_algebra_against_oracle_autodiff(tmp_path):
    """csrc/ssb_jet.cuh (the Taylor jets the kernels use for the rotating bars) compiles as plain C++ too: potential, gradient, Hessian and
    third derivatives of both bars from one order-3 jet, against the oracle's nested dual numbers - the device algebra checked without a GPU."""
    import itertools
    import subprocess
    from math import factorial
    import numpy as np
    import oracle as O
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "jet_host_check")
    subprocess.run(["g++", "-std=c++17", "-O2", "-I", os.path.join(root, "streamsculptor_b200", "csrc"), os.path.join(root, "tests", "jet_host_check.cpp"), "-o", exe],
                   check=True)
    G = O.Program().G
    ex = [(3, 0, 0), (2, 1, 0), (2, 0, 1), (1, 2, 0), (1, 1, 1), (1, 0, 2), (0, 3, 0), (0, 2, 1), (0, 1, 2), (0, 0, 3)]
    rng = np.random.default_rng(1)
    for kind, prog in ((0, O.Program().bar(1e10, 3.5, 0.5, 0.6, 0.04)), (1, O.Program().dehnen_bar(0.01, 0.22, 8.0, 3.4, 0.4, 0.05))):
        for _ in range(12):
            x = rng.normal(size=3) * np.array([6.0, 6.0, 2.0])
            t = rng.uniform(-3000.0, 0.0)
            out = np.array(subprocess.run([exe, str(kind), *[repr(float(v)) for v in x], repr(float(t))], capture_output=True, text=True, check=True).stdout.split(),
                           dtype=float)
            J3, J2, J1 = out[:20], out[20:30], out[30:]
            assert np.allclose(J3[:10], J2, rtol=1e-13, atol=0) and np.allclose(J3[:4], J1, rtol=1e-13, atol=0)     # orders 1 and 2 are prefixes of order 3
            if kind == 0:
                J3 = J3 * (G / 4.498502151469554e-12)          # the check program hard-codes G
            tq = np.array([t])
            H = np.array([[2 * J3[4], J3[5], J3[6]], [J3[5], 2 * J3[7], J3[8]], [J3[6], J3[8], 2 * J3[9]]])
            T = np.zeros((3, 3, 3))
            for c, e in zip(J3[10:], ex):
                for perm in set(itertools.permutations([0] * e[0] + [1] * e[1] + [2] * e[2])):
                    T[perm] = c * factorial(e[0]) * factorial(e[1]) * factorial(e[2])
            for got, want, tol in ((J3[0], prog.potential(x[None], tq)[0], 1e-13), (J3[1:4], prog.gradient(x[None], tq)[0], 1e-12),
                                   (H, prog.hessian(x[None], tq)[0], 1e-11), (T, np.asarray(prog.third(x[None], tq)[0]), 1e-10)):
                assert np.abs(got - want).max() <= tol * np.abs(want).max()


def test_not_a_knot_spline_as_hermite_segments():
    """LMCPotential / the restricted N-body progenitor track use a C2 not-a-knot cubic spline in the reference (jax_cosmo
    InterpolatedUnivariateSpline(k=3), potential.py:47-49).  A cubic spline IS its knot values + knot slopes: a cubic track fed with those slopes
    reproduces scipy's not-a-knot spline (value and derivative) to rounding - checked on the oracle's track, the device evaluates the same
    Hermite form (tests/test_gpu_parity.py)."""
    import numpy as np
    from scipy.interpolate import CubicSpline
    import oracle as O
    rng = np.random.default_rng(2)
    t = np.sort(rng.uniform(-3000.0, 0.0, 40)); t[0], t[-1] = -3000.0, 0.0
    y = np.cumsum(rng.normal(size=(40, 3)), axis=0) * 5.0
    cs = CubicSpline(t, y, axis=0, bc_type="not-a-knot")
    prog = O.Program()
    tr = prog.track(O.CUBIC, t, y, slopes=cs(t, 1))
    tq = np.concatenate([rng.uniform(-3000.0, 0.0, 500), t])
    val, der = prog.track_eval(tr, tq)
    assert np.abs(val - cs(tq)).max() <= 1e-12 * np.abs(y).max()
    assert np.abs(der - cs(tq, 1)).max() <= 1e-11 * np.abs(cs(tq, 1)).max()
    # and it is NOT what the default (interpax 'cubic') slopes give
    tr2 = prog.track(O.CUBIC, t, y)
    assert np.abs(prog.track_eval(tr2, tq)[0] - cs(tq)).max() > 1e-3
