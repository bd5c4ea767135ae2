#!/usr/bin/env python
"""Benchmark of the hot path: BASELINE.json metric "fp64 particle-steps/sec; time-to-stream for 1e6 particles" on config C2
(ONE 1e6-particle mock stream, static MW3 potential, adaptive Dopri8, sharded over the GPUs of one box).

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run, one rank per GPU)
  python bench.py --impl reference ...                     (the CPU restatement of the reference on the host cores)

A "step" = one full pass of gen_stream_vmapped (main.py:343-368) over one synthetic stream: serial progenitor orbit with
dense output at every stripping time, particle-spray release (jax threefry recipe), then 2*(Nts-1) independent adaptive
Dopri8 solves from ts[i] to ts[-1].  One particle-step = one RK step ATTEMPT (accepted or rejected) of one particle.

Scaling (--scaling, default "strong"): BASELINE config 2 is ONE 1e6-particle stream "sharded across 8xB200" and the metric names its
time-to-stream, so the headline arm keeps the stream fixed at --particles (1e6) and rank r of N integrates the particles i = r (mod N);
the shares are all-gathered over NCCL inside the timed step.  The weak arm (--particles PER GPU, the N-fold stream) is measured in the same
run and reported under config.weak.  At N = 1 the two coincide.

The same JSON line carries a `c4` block: BASELINE config 4 (first-order response of 1e4 stream particles to 1000 Hernquist subhalos,
compute_perturbation_OTF, perturbative.py:726-755 / generate_derivs.py:193), pair-steps/s with its own roofline, at N GPUs with the particles
dealt out over the ranks (parallel.linear_response_sharded).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_STEP_DOPRI8 = 1920.0     # algorithmic flops per particle-step, unit-cost convention (SURVEY.md 8d / BASELINE.md section 3)
FP64_NOMINAL_TFLOPS = 37.2        # 148 SM x 64 DFMA/clk x 2 x 1.965 GHz
PROG_TODAY = [20.0, 0.0, 20.0, 0.0, 0.15, 0.0]
T_AGE, MSAT, SEED = 3000.0, 1e4, 583
CTRL = dict(rtol=1e-7, atol=1e-7, dtmin=0.3, dtmax=None, max_steps=10_000)
C4 = dict(n_particles=10_000, n_sh=1000, tol=1e-6, dtmin=0.01, t_window=150.0, n_prod=2000, tol_prod=1e-11)


def workload(n_release):
    """Synthetic C2 inputs: ts = linspace(-3000, 0, Nts), progenitor integrated back 3 Gyr."""
    ts = np.linspace(-T_AGE, 0.0, n_release + 1)
    return ts, np.full(n_release + 1, MSAT)


_JSON_FD = None


def claim_stdout():
    """Rank 0 must print ONE JSON line: everything else written to fd 1 by libraries (NCCL prints its version banner there) is sent
    to stderr for the duration of the run; emit() writes the result line to the real stdout."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def workload_name(total_particles, world, scaling):
    """config.workload of both arms (the reference arm times a bounded sample of this workload on the host cores)."""
    return (f"C2: ONE {total_particles}-particle mock stream ({scaling} scaling: {total_particles // max(world, 1)} particles per GPU on {world} GPU(s)), "
            "static MW3 (Hernquist+MiyamotoNagai+NFW), 3 Gyr, adaptive Dopri8 rtol=atol=1e-7 dtmin=0.3, final state kept "
            "(gen_stream_vmapped semantics), jax-threefry release draws")


def prog_start():
    """Progenitor state 3 Gyr ago, computed ONCE on the device (not part of the timed step)."""
    import streamsculptor_b200 as ssc
    pot = mw3()
    return np.asarray(pot.integrate_orbit(w0=PROG_TODAY, ts=np.array([0.0, -T_AGE]), t0=0.0, t1=-T_AGE, solver=ssc.Dopri8()).ys[-1])


def mw3():
    import streamsculptor_b200 as ssc
    P = ssc.potential
    return P.Potential_Combine([P.HernquistPotential(m=5e9, r_s=1.0, units=ssc.usys), P.MiyamotoNagaiDisk(m=6.8e10, a=3.0, b=0.28, units=ssc.usys),
                                P.NFWPotential(m=5.4e11, r_s=15.62, units=ssc.usys)], units=ssc.usys)


def mw3_oracle():
    import oracle as O
    return O.Program().hernquist(5e9, 1.0).miyamoto(6.8e10, 3.0, 0.28).nfw(5.4e11, 15.62)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.idx, self.lines, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.idx)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.th.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (C++ restatement of the reference algorithm) on the host cores.  Timed in its -O3 build (oracle/Makefile target
# `bench`: -O3, AVX2 + FMA code generation), as BASELINE.md section 5 plans; parity tests use the strict -O2 / no-contraction build.
# ----------------------------------------------------------------------------------------------------------------
def cpu_stream_rate(n_release, threads):
    """One pass of the same pipeline on the CPU: returns (particle_steps, seconds)."""
    import oracle as O
    with O.variant("bench"):
        orc = mw3_oracle()
        back, _, _ = orc.integrate_orbits(PROG_TODAY, 0.0, -T_AGE)
        ts, ms = workload(n_release)
        t = time.perf_counter()
        _, _, _, ns = orc.gen_stream(ts, back[0, 0], ms, SEED, solver=8, threads=threads, **{k: v for k, v in CTRL.items()})
        dt = time.perf_counter() - t
    return int(ns[:, 0].sum()), dt


def cpu_sample_size(threads, target_s):
    steps, dt = cpu_stream_rate(500, threads)            # calibration pass (1000 particles)
    rate = steps / dt
    per_release = steps / 500.0
    return int(np.clip(rate * target_s / per_release, 500, 500_000))


CPU_NOTE = ("C++ restatement of the reference algorithm (oracle/, -O3 AVX2+FMA build, dual-number autodiff of the scalar potentials as the "
            "reference does with jax), NOT jax[cpu]: jax/diffrax are not installable offline")


def run_reference(args, rank, world):
    import oracle as O
    if rank != 0:
        return
    threads = O.num_threads()
    n_rel = cpu_sample_size(threads, 6.0)
    for _ in range(args.warmup):
        cpu_stream_rate(max(500, n_rel // 8), threads)
    tot_steps, tot_s = 0, 0.0
    for _ in range(args.steps):
        s, dt = cpu_stream_rate(n_rel, threads)
        tot_steps += s; tot_s += dt
    value = tot_steps / tot_s
    total = args.particles if args.scaling == "strong" else args.particles * world
    sample = f"{2 * n_rel} particles of the C2 stream per step (ts=linspace(-3000,0,{n_rel + 1}), Dopri8 rtol=atol=1e-7)"
    line = {"impl": "reference", "metric": "fp64 particle-steps/sec", "value": value, "unit": "particle-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / args.steps, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(total, world, args.scaling), "total_particles": total,
                       "sample": f"each step = a {2 * n_rel}-particle stream of the same kind on the host cores (a rate, so the sub-sample is "
                                 "representative; the full stream would take minutes per step)", "particles_per_step": 2 * n_rel},
            "cpu_baseline": {"value": value, "unit": "particle-steps/s", "cores": threads, "kind": "port", "sample": sample, "note": CPU_NOTE},
            "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def _ncu_summary(pattern):
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", pattern)))
    return (open(files[-1]).read(), os.path.relpath(files[-1], ROOT)) if files else (None, None)


def ncu_numbers(pattern):
    """(dram bytes per launch, counted fp64 fraction, file) from the newest committed `ncu --set full` summary matching `pattern`.  These two
    counters cannot be read without a profiler, and a number taken under the profiler is never timed here: they describe the capture the
    file names (its launch type is stated there), the timing in `achieved` is this run's."""
    import re
    txt, path = _ncu_summary(pattern)
    if txt is None:
        return None, None, None
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = 0.0
    for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        m = re.search(re.escape(name) + r" \[(\w+)\] = ([0-9.eE+-]+)", txt)
        if not m:
            tot = None
            break
        tot += float(m.group(2)) * unit.get(m.group(1), 1.0)
    m = re.search(r"counted fp64 FLOP rate .* = ([0-9.]+)", txt)
    return tot, (float(m.group(1)) if m else None), path


# ----------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------
def c4_workload(pot, n_p, n_sh):
    """stream particles released along the C1/C2 progenitor orbit, 1000 Hernquist subhalos hitting the stream at random
    times (generate_derivs.py:155; GenerateImpactParams.py:24), numpy PCG64(1234); perturbation ICs zero (perturbative.py:712)."""
    import streamsculptor_b200 as ssc
    P = ssc.potential
    back = pot.integrate_orbit(w0=PROG_TODAY, ts=np.array([0.0, -T_AGE]), t0=0.0, t1=-T_AGE).ys[-1]
    ts = np.linspace(-T_AGE, 0.0, n_p // 2 + 1)
    nr = np.random.Generator(np.random.PCG64(0)).standard_normal((len(ts), 4))
    pl, pt, vl, vt = pot.gen_stream_ics(ts=ts, prog_w0=back, Msat=MSAT, seed_num=SEED, solver=ssc.Dopri8(), normals=nr)
    w0 = np.vstack([np.hstack([pl, vl])[:-1], np.hstack([pt, vt])[:-1]])
    t0 = np.concatenate([ts[:-1], ts[:-1]])
    rng = np.random.Generator(np.random.PCG64(1234))
    M = 10 ** rng.uniform(5, 9, n_sh); rs = 1.05 * np.sqrt(M / 1e8)
    t_imp = rng.uniform(-T_AGE, 0.0, n_sh)
    prog_at = pot.integrate_orbit(w0=back, ts=np.sort(t_imp), t0=-T_AGE, t1=0.0).ys[np.argsort(np.argsort(t_imp))]
    b = rng.uniform(0, 10 * rs)
    d = rng.normal(size=(n_sh, 3)); d /= np.linalg.norm(d, axis=1)[:, None]
    x0 = prog_at[:, :3] + b[:, None] * d
    v = rng.normal(size=(n_sh, 3)) * 0.184
    pert = P.SubhaloLinePotentialCustom_fromFunc(func=P.HernquistPotential, m=np.ones(n_sh), r_s=rs, subhalo_x0=x0, subhalo_v=v, subhalo_t0=t_imp,
                                                 t_window=C4["t_window"], units=ssc.usys)
    return w0, t0, pert, M, dict(rs=rs, x0=x0, v=v, t0=t_imp)


def run_c4(args, rank, world, dev, flush, barrier):
    """BASELINE config 4 in the driver-run record: 1e4 particles x 1000 subhalos, Dopri8; strong scaling over the ranks (particles dealt
    out interleaved, only final states + response summaries gathered)."""
    import torch
    import streamsculptor_b200 as ssc
    from streamsculptor_b200 import _lib, _runtime as rt, parallel as par
    d = par.dist()
    pot = mw3()
    out = {}
    for tag, n_p, tol in (("tol1e-6", C4["n_particles"], C4["tol"]), ("production_tol1e-11", C4["n_prod"], C4["tol_prod"])):
        w0, t0, pert, M_sh, _ = c4_workload(pot, n_p, C4["n_sh"])
        ctrl = rt.make_ctrl(ssc.Dopri8(), tol, tol, C4["dtmin"], None, 10_000)
        w0_d, t0_d = rt.to_dev(w0), rt.to_dev(t0)
        sel = torch.arange(rank, len(w0), world, device=dev)
        w0_l, t0_l = w0_d[sel].contiguous(), t0_d[sel].contiguous()

        def step():
            if world == 1:
                return rt.linear_response(pot, pert._arrays, w0_d, None, t0_d, 0.0, ctrl)
            par.linear_response_sharded(pot, pert._arrays, w0_d, t0_d, 0.0, ctrl, rank, world, M_sh)
            return None
        for _ in range(2):
            step()
        barrier()
        reps = max(2, min(args.steps, 5))
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for a, b in ev:
            flush.zero_()
            a.record(); step(); b.record()
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in ev) / reps
        _, _, st, ns = rt.linear_response(pot, pert._arrays, w0_l, None, t0_l, 0.0, ctrl)         # this rank's share: step counts for the metric
        assert int((st != 0).sum().item()) == 0, "a particle failed in the C4 benchmark"
        tt = torch.tensor([ms], dtype=torch.float64, device=dev)
        ps = torch.tensor([float(ns[:, 0].sum().item())], dtype=torch.float64, device=dev)
        if world > 1:
            d.all_reduce(tt, op=d.ReduceOp.MAX); d.all_reduce(ps, op=d.ReduceOp.SUM)
        ms, psteps = float(tt.item()), float(ps.item())
        out[tag] = {"particles": int(len(w0)), "subhalos": C4["n_sh"], "rtol_atol": tol, "ms": ms, "particle_steps": psteps,
                    "pair_steps_per_s": psteps * C4["n_sh"] / (ms * 1e-3), "out_bytes": int(len(w0)) * C4["n_sh"] * 96}
    if rank != 0:
        return None
    main_ = out["tol1e-6"]
    traffic, counted, src = ncu_numbers("*response*kernel_ncu.txt")
    peak = C.c_double(0.0)
    _lib.check(_lib.lib().ssb_fp64_peak_probe(20000, C.byref(peak), rt.stream_ptr()))
    blk = {"workload": f"C4: first-order (mass, radius) response of {C4['n_particles']} stream particles to {C4['n_sh']} Hernquist subhalos "
                       f"(compute_perturbation_OTF, field MassRadiusPerturbation_OTF), Dopri8 rtol=atol={C4['tol']:g} dtmin={C4['dtmin']}, t_window=150 Myr, "
                       f"{'one GPU' if world == 1 else f'particles dealt out over {world} GPUs, final states + response summaries all-gathered'}",
           "metric": "pair-steps/s (pair = particle x subhalo, step = RK attempt of the particle's coupled ODE)", "value": main_["pair_steps_per_s"],
           "ms": main_["ms"], "runs": out,
           "roofline": {"bound": "fp64", "kernel": "response_kernel_mp<8>", "unit": "fraction of the DFMA peak", "frac": counted,
                        "frac_source": "ncu-counted (2 DFMA + DADD + DMUL) per cycle / peak of the committed capture " + str(src) +
                                       " (a counter, not measurable without the profiler; the reference's per-pair flop count would credit work the "
                                       "propagator / unborn-subhalo paths legitimately skip, so it is not used)",
                        "peak_tflops": peak.value / 1e12, "traffic": traffic, "algorithmic_bytes": main_["out_bytes"] + 56 * main_["particles"]}}
    if world == 1:          # CPU port on a bounded sample of the same workload
        import oracle as O
        threads = O.num_threads()
        n_cpu = 60 * max(threads, 1)                # ~10-15 s of host work at ~0.2 s per particle and core
        w0, t0, pert, M_sh, sh = c4_workload(pot, 2 * (C4["n_particles"] // 2), C4["n_sh"])
        pick = np.linspace(0, len(w0) - 1, n_cpu).astype(int)
        with O.variant("bench"):
            base = mw3_oracle()
            shp = O.Program().subhalos(O.PR_HERNQUIST, np.ones(C4["n_sh"]), sh["rs"], sh["x0"], sh["v"], sh["t0"], np.full(C4["n_sh"], C4["t_window"]))
            t = time.perf_counter()
            _, _, _, ns = O.linear_response(base, shp, w0[pick], t0[pick], 0.0, solver=8, rtol=C4["tol"], atol=C4["tol"], dtmin=C4["dtmin"], threads=threads)
            dt = time.perf_counter() - t
        blk["cpu_baseline"] = {"value": float(ns[:, 0].sum()) * C4["n_sh"] / dt, "unit": "pair-steps/s", "cores": threads, "kind": "port",
                               "sample": f"{n_cpu} of the {C4['n_particles']} particles x {C4['n_sh']} subhalos, {dt:.1f} s", "note": CPU_NOTE}
    return blk


def run_ours(args, rank, world):
    import torch
    import streamsculptor_b200 as ssc
    from streamsculptor_b200 import _lib, _runtime as rt, parallel as par
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback (use --impl reference for the CPU arm)")
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    d = par.dist()
    lib = _lib.lib()
    pot = mw3()
    w0 = prog_start()
    ctrl = rt.make_ctrl(ssc.Dopri8(), **CTRL)
    kv = ssc.main.DEFAULT_KVALS
    flush = torch.empty(512 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)      # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            d.barrier()
        torch.cuda.synchronize()

    def device_arm(n_rel):
        """Time `steps` passes over ONE stream of 2 n_rel particles split over the ranks: (ms_total max over ranks, particle-steps per pass, outputs)."""
        ts, ms = workload(n_rel)
        ts_d, w0_d, ms_d = rt.to_dev(ts), rt.to_dev(w0), rt.to_dev(ms)
        n_local = par.shard_count(n_rel, rank, world)

        def step_device():
            lead, trail, status, nsteps = rt.gen_stream(pot, pot, pot._G, ts_d, w0_d, ms_d, SEED, kv, None, ctrl, i_begin=rank, i_stride=world, n_local=n_local)
            packed = getattr(lead, "_ssb_packed", None)          # [2, n_local, 6] storage shared by lead and trail
            if world > 1:          # ONE NCCL all-gather of the packed (lead, trail) shares + one permuting copy into global particle order
                both = par.gather_interleaved(lead if packed is None else packed, n_rel, rank, world, axis=1 if packed is not None else 0)
                if packed is not None:
                    lead, trail = both[0], both[1]
                else:
                    lead, trail = both, par.gather_interleaved(trail, n_rel, rank, world)
            return lead, trail, status, nsteps
        for _ in range(args.warmup):
            out = step_device()
        barrier()
        psteps_local = int(out[3][..., 0].sum().item())
        assert int((out[2] != 0).sum().item()) == 0, "an orbit failed in the benchmark stream"
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        l0 = lib.ssb_launch_count()
        barrier()
        for a, b in ev:
            flush.zero_()                      # L2 flush between timed iterations (untimed)
            a.record()
            step_device()
            b.record()
        barrier()
        launches = lib.ssb_launch_count() - l0
        ms_total = sum(a.elapsed_time(b) for a, b in ev)
        tt = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        ps = torch.tensor([float(psteps_local)], dtype=torch.float64, device=dev)
        if world > 1:
            d.all_reduce(tt, op=d.ReduceOp.MAX)
            d.all_reduce(ps, op=d.ReduceOp.SUM)
        return float(tt.item()), float(ps.item()), out, (ts, ms, ts_d, w0_d, ms_d, n_local), launches

    n_rel_strong = args.particles // 2
    n_rel_weak = (args.particles // 2) * world
    n_rel = n_rel_strong if args.scaling == "strong" else n_rel_weak
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total, psteps, out, (ts, ms, ts_d, w0_d, ms_d, n_local), launches = device_arm(n_rel)
    clocks = sampler.stop() if rank == 0 else None
    value = psteps * args.steps / (ms_total * 1e-3)
    other = None
    if world > 1:            # the other scaling arm, same run
        o_ms, o_ps, _, _, _ = device_arm(n_rel_weak if args.scaling == "strong" else n_rel_strong)
        other = {"scaling": "weak" if args.scaling == "strong" else "strong", "total_particles": 2 * (n_rel_weak if args.scaling == "strong" else n_rel_strong),
                 "ms_per_step": o_ms / args.steps, "value": o_ps * args.steps / (o_ms * 1e-3), "unit": "particle-steps/s"}

    # ---- e2e: the C-ABI host call (HOST buffers, H2D + D2H inside the timed region), same shard per rank ----
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h_ts, h_ms, h_w0 = pin(ts), pin(ms), pin(w0)
    # lead and trail are the two halves of ONE pinned [2, n_local, 6] buffer: with pinned outputs laid out like this the library lets the
    # orbit kernel write its results straight into host memory (zero-copy, csrc/ssb_host.cu), overlapping the transfer with the integration
    h_out = torch.empty((2, n_local, 6), dtype=torch.float64).pin_memory()
    h_lead, h_trail = h_out[0], h_out[1]
    h_stat, h_ns = torch.empty((2, n_local), dtype=torch.int32).pin_memory(), torch.empty((2, n_local, 3), dtype=torch.int32).pin_memory()
    P, _keep = rt.lower(pot)
    kvc = (C.c_double * 8)(*kv)
    hp = lambda t: C.c_void_p(t.data_ptr())

    def step_host():
        _lib.check(lib.ssb_gen_stream_host(C.byref(P), C.byref(P), pot._G, n_rel + 1, hp(h_ts), hp(h_w0), hp(h_ms), SEED, kvc, None, ctrl, rank, world,
                                           n_local, hp(h_lead), hp(h_trail), hp(h_stat), hp(h_ns)))
    for _ in range(min(args.warmup, 3)):
        step_host()

    def time_host():
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_host()
        barrier()
        te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            d.all_reduce(te, op=d.ReduceOp.MAX)
        return float(te.item())
    e2e_s = time_host()
    e2e_value = psteps * args.steps / e2e_s
    os.environ["SSB_HOST_ZEROCOPY"] = "0"            # same call with the outputs staged in HBM and copied back afterwards (reported for comparison)
    step_host()
    e2e_staged_s = time_host()
    del os.environ["SSB_HOST_ZEROCOPY"]
    staged = (h_out.clone(), h_stat.clone(), h_ns.clone())
    h_out.fill_(-1.0); h_stat.fill_(-1); h_ns.fill_(-1)
    step_host()
    assert torch.equal(staged[0], h_out) and torch.equal(staged[1], h_stat) and torch.equal(staged[2], h_ns), "zero-copy and staged host paths disagree"
    assert np.allclose(h_lead.numpy(), out[0][rank::world].cpu().numpy() if world > 1 else out[0].cpu().numpy(), rtol=0, atol=0), "host and device paths disagree"
    # bytes this rank moves per step: a shard uploads only its own stripping times (+ the two interval ends), see csrc/ssb_host.cu
    h2d = (h_ts.numel() * 8 + h_ms.numel() * 8 + 48) if world == 1 else ((n_local + 2) * 8 + n_local * 8 + 48)
    d2h = h_lead.numel() * 8 * 2 + h_stat.numel() * 4 + h_ns.numel() * 4

    c4 = None if args.no_c4 else run_c4(args, rank, world, dev, flush, barrier)
    if rank != 0:
        return
    # ---- roofline of the dominant kernel (orbit_kernel<8>): CUDA events around that launch alone ----
    pl, pt, vl, vt = pot.gen_stream_ics(ts=ts_d, prog_w0=w0_d, Msat=ms_d, seed_num=SEED, solver=ssc.Dopri8(), **CTRL)
    sel = torch.arange(rank, n_rel, world, device=dev)
    w0_all = torch.cat([torch.cat([pl, vl], 1)[sel], torch.cat([pt, vt], 1)[sel]]).contiguous()
    t0_all = torch.cat([ts_d[sel], ts_d[sel]]).contiguous()
    t1_all = torch.zeros_like(t0_all)
    kt = []
    for i in range(args.steps + 2):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ys, st, ns = rt.orbit_integrate(pot, w0_all, t0_all, t1_all, t1_all.reshape(-1, 1), ctrl, ts_per_orbit=1)
        b.record()
        torch.cuda.synchronize()
        if i >= 2:
            kt.append(a.elapsed_time(b))
    k_ms = float(np.mean(kt))
    k_steps = float(ns[:, 0].sum().item())
    achieved = FLOP_PER_STEP_DOPRI8 * k_steps / (k_ms * 1e-3) / 1e12
    peak = C.c_double(0.0)
    _lib.check(lib.ssb_fp64_peak_probe(20000, C.byref(peak), rt.stream_ptr()))
    peak_tf = peak.value / 1e12
    traffic, counted, src = ncu_numbers("*orbit_kernel_ncu.txt")
    roofline = {"bound": "fp64", "kernel": "orbit_kernel<8>", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                "traffic": traffic, "traffic_source": src, "ncu_counted_fp64_frac": counted,
                "kernel_ms": k_ms, "particle_steps_per_launch": k_steps, "flop_per_particle_step": FLOP_PER_STEP_DOPRI8,
                "peak_source": "measured in this run by ssb_fp64_peak_probe (independent DFMA chains on every SM); MEASURED_PEAKS.json has no fp64 entry; "
                               f"nominal 148 SM x 64 DFMA/clk x 2 x 1.965 GHz = {FP64_NOMINAL_TFLOPS} TFLOP/s",
                "frac_of_nominal": achieved / FP64_NOMINAL_TFLOPS,
                "algorithmic_bytes": w0_all.numel() * 8 + ys.numel() * 8 + t0_all.numel() * 16 + ns.numel() * 4 + st.numel() * 4,
                "hbm_gbs_of_kernel": (w0_all.numel() * 8 + ys.numel() * 8 + t0_all.numel() * 16 + ns.numel() * 4) / (k_ms * 1e-3) / 1e9}
    # ---- CPU baseline on this box's host cores: bounded sample of the same workload (rank 0, N = 1 only) ----
    cpu = None
    if world == 1:
        import oracle as O
        threads = O.num_threads()
        n_cpu = cpu_sample_size(threads, 12.0)
        cs, cdt = cpu_stream_rate(n_cpu, threads)
        cpu = {"value": cs / cdt, "unit": "particle-steps/s", "cores": threads, "kind": "port",
               "sample": f"{2 * n_cpu} particles of the same stream (ts=linspace(-3000,0,{n_cpu + 1})), {cdt:.1f} s", "note": CPU_NOTE}
    total = 2 * n_rel
    cfg = {"workload": workload_name(total, world, args.scaling), "total_particles": total, "particles_per_gpu": total // world,
           "particle_steps_per_step": psteps, "parallelism": f"dp{world} (particles interleaved over ranks, one all-gather of the (N,6) arms inside the step)",
           "l2": "512 MB buffer zeroed between timed iterations", "time_to_stream_ms": ms_total / args.steps}
    if other is not None:
        cfg[other["scaling"]] = other
    line = {"metric": "fp64 particle-steps/sec", "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg,
            # counted by the library (ssb_launch_count): kernels of libssb200 launched by THIS rank inside the timed device-resident region
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": "particle-steps/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": 1e3 * e2e_s / args.steps, "call": "ssb_gen_stream_host (C ABI, pinned host buffers; results written by the orbit "
                    "kernel directly into the pinned output buffer; a shard uploads only its own stripping times)",
                    "ms_per_step_staged_d2h": 1e3 * e2e_staged_s / args.steps, "bytes_are": "per rank"},
            "roofline": roofline, "cpu_baseline": cpu, "c4": c4}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: ONE stream of --particles split over the GPUs (BASELINE config 2, time-to-stream); weak: --particles per GPU")
    ap.add_argument("--particles", type=int, default=1_000_000, help="stream particles (2 per stripping time): total (strong) or per GPU (weak)")
    ap.add_argument("--no-c4", action="store_true", help="skip the C4 (1000-subhalo response) block")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    claim_stdout()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        from streamsculptor_b200 import parallel as par
        par.init_from_env("nccl")
    run_ours(args, rank, world)
    if world > 1:
        from streamsculptor_b200 import parallel as par
        par.dist().destroy_process_group()


if __name__ == "__main__":
    main()
