"""Production driver of the reference (generate_derivs.py:24-216): perturbation derivatives of a stream for batches of sampled subhalo
impacts, saved to disk.

Per batch the reference runs (generate_derivs.py:153-204): sample masses -> ImpactGenerator -> Hernquist subhalo arrays ->
GenerateMassRadiusPerturbation_Chen25.compute_perturbation_OTF -> jnp.save.  Here the base stream, the impact geometry (window means,
backward patch orbits) and the response solve stay on the device from one batch to the next - the base model is built ONCE, its release
conditions and stripping times are uploaded once, every batch only uploads the O(N_batch) subhalo parameters - and, besides the raw
derivatives the reference saves, a batch can be reduced on the device to the summaries a fit needs (binned first-order track
displacements, `summaries=`), which is what crosses NVLink in a sharded run (parallel.linear_response_sharded).
"""
import os

import numpy as np

from . import _runtime as rt
from . import perturbative as pert
from . import potential
from .GenerateImpactParams import ImpactGenerator, jax_randint, jax_split
from .solvers import Dopri8
from .streamhelpers import gen_stream_vmapped_Chen25, sample_from_1D_pdf
from .units import usys


def binned_response(stream_phi1, D, M_sh, edges):
    """Device summary of one batch: mean first-order displacement sum_sh M_sh D[:, sh, :6] of the stream in phi1 bins -> [n_bins, 6] (torch).
    Particles without a result (+inf rows: zero-length integration, as diffrax's SaveAt leaves them) are left out."""
    tt = rt.torch()
    disp = tt.einsum("s,nsk->nk", tt.as_tensor(M_sh, dtype=D.dtype, device=D.device), D[:, :, :6])
    idx = tt.bucketize(stream_phi1, tt.as_tensor(edges, dtype=D.dtype, device=D.device)) - 1
    nb = len(edges) - 1
    ok = (idx >= 0) & (idx < nb) & tt.isfinite(disp).all(dim=1) & tt.isfinite(stream_phi1)    # a particle released at the final time has +inf rows (diffrax)
    out = tt.zeros((nb, 6), dtype=D.dtype, device=D.device)
    cnt = tt.zeros((nb,), dtype=D.dtype, device=D.device)
    out.index_add_(0, idx[ok], disp[ok])
    cnt.index_add_(0, idx[ok], tt.ones_like(idx[ok], dtype=D.dtype))
    return out / cnt.clamp(min=1.0)[:, None]


def get_derivs(prog_wtoday, t_age, t_dissolve, log10_min_mass, log10_max_mass, phi1_bounds, phi1_exclude, stream_seednum, key, Msat, r_s, target_num,
               phi1_function, pot, path, N_batch=500, atol=1e-11, rtol=1e-11, bmax_fac=10.0, phi1window=0.5, N_arm=5_000, save_iter_start=0,
               summaries=None, save=True, progress=False, pipeline=4):
    """Same arguments and files as the reference (one `<path>/<i>.npy` per batch holding dict(pert_out, r_s_root, ImpactFrameParams)).
    Extras: summaries = phi1 bin edges -> every batch also stores `binned` ([n_bins, 6], computed on the device); save=False keeps nothing on
    disk and returns the list of per-batch dicts (tests, benchmarks); pipeline = batches in flight (each on its own CUDA stream).  A batch is
    bound by its slowest particle - with a compact progenitor a few particles take thousands of steps while the GPU idles (DESIGN.md section
    3) - so the next batch's solve starts while the previous one drains; results are identical to pipeline=1, batch by batch."""
    tt = rt.torch()
    t_age, t_dissolve = float(t_age), float(t_dissolve)
    IC = np.asarray(pot.integrate_orbit(w0=prog_wtoday, t0=0.0, t1=-t_age, ts=np.array([-t_age])).ys[0])               # generate_derivs.py:110
    ts = np.hstack([np.linspace(-t_age, t_dissolve, int(N_arm)), [0.0]])
    prog_pot = potential.PlummerPotential(m=Msat, r_s=r_s, units=usys)
    l, t = gen_stream_vmapped_Chen25(pot_base=pot, prog_w0=IC, ts=ts, key=stream_seednum, Msat=Msat, atol=1e-7, rtol=1e-7, solver=Dopri8(),
                                     prog_pot=prog_pot)                                                                 # generate_derivs.py:113-124
    stream = np.vstack([np.asarray(l), np.asarray(t)])
    phi1_model = np.asarray(phi1_function(stream))
    mass_lin = 10 ** np.linspace(log10_min_mass, log10_max_mass, 50_000)
    prob = mass_lin ** (-0.3)
    N_iter = int(np.ceil(target_num / N_batch))
    assert ts[-1] == 0.0
    base = pert.BaseStreamModelChen25(pot_base=pot, ts=ts, prog_w0=IC, Msat=Msat, key=stream_seednum, units=usys, prog_pot=prog_pot, rtol=1e-7, atol=1e-7,
                                      solver=Dopri8())                                                                  # generate_derivs.py:139-149
    keys = jax_split(key, N_iter)
    stripping = np.hstack([ts[:-1], ts[:-1]])
    ctrl = rt.make_ctrl(Dopri8(), rtol, atol, 0.01, None, 10_000)                                                       # generate_derivs.py:193-198
    out = []
    shared = {}                                          # what every batch's ImpactGenerator has in common (same stream): computed once,
    ImpactGenerator(pot=pot, tobs=0.0, stream=stream, stream_phi1=phi1_model, phi1_bounds=phi1_bounds, tImpactBounds=[-t_age, 0.0],     # on the caller's stream
                    phi1window=phi1window, NumImpacts=1, bImpact_bounds=[0, 1.0], stripping_times=stripping, phi1_exclude=phi1_exclude,
                    prog_today=prog_wtoday, seednum=0, shared=shared)

    def launch(i):
        """Everything of batch i up to the enqueued response solve (generate_derivs.py:153-198); no wait for its results."""
        mass = sample_from_1D_pdf(x=mass_lin, y=prob, key=keys[i], num_samples=N_batch)                                 # generate_derivs.py:154
        rs = 1.05 * np.sqrt(mass / 1e8)
        rand_int = jax_randint(jax_split(keys[i], 1)[0], 1, 0, 10_000_000)[0]                                           # generate_derivs.py:158-160
        gen = ImpactGenerator(pot=pot, tobs=0.0, stream=stream, stream_phi1=phi1_model, phi1_bounds=phi1_bounds, tImpactBounds=[-t_age, 0.0],
                              phi1window=phi1window, NumImpacts=len(mass), bImpact_bounds=[0, rs * bmax_fac], stripping_times=stripping,
                              phi1_exclude=phi1_exclude, prog_today=prog_wtoday, seednum=int(rand_int), shared=shared)
        imp = gen.get_subhalo_ImpactParams()
        cart = imp["CartesianImpactParams"]
        assert not np.isnan(cart.sum())                                                                                 # generate_derivs.py:180
        sub_pot = potential.SubhaloLinePotentialCustom_fromFunc(func=potential.HernquistPotential, m=np.ones(len(mass)), r_s=rs, subhalo_x0=cart[:, :3],
                                                                subhalo_v=cart[:, 3:], subhalo_t0=imp["ImpactFrameParams"]["tImpact"], t_window=150.0,
                                                                units=usys)
        pertgen = pert.GenerateMassRadiusPerturbation_Chen25(potential_base=pot, potential_perturbation=sub_pot, BaseStreamModel=base, units=usys)
        bs = pertgen.base_stream                         # compute_perturbation_OTF (perturbative.py:726-755) without its host round trip
        n = len(bs.ts) - 1
        w, D, status, _ = rt.linear_response(pertgen.potential_base_total, pertgen.subhalo_arrays, rt.to_dev(np.asarray(bs.streamICs)[:n]), None,
                                             rt.to_dev(np.asarray(bs.ts, dtype=np.float64)[:n]), float(bs.ts[-1]), ctrl)
        return dict(i=i, w=w, D=D, status=status, mass=mass, r_s_root=sub_pot.r_s, frame=imp["ImpactFrameParams"], keep=(pertgen, sub_pot))

    def finish(job):
        if bool((job["status"] != 0).any()):             # integrate_field leaves diffrax's throw=True (fields.py:85-98)
            raise RuntimeError("get_derivs: a particle failed (max_steps reached or non-finite state)")
        # D (0.5 GB for a production batch) goes through a pinned staging buffer kept per pipeline slot: 25 GB/s instead of a pageable copy
        slot_ = job["i"] % depth
        if pinned[slot_] is None or pinned[slot_].shape != job["D"].shape:
            pinned[slot_] = tt.empty(job["D"].shape, dtype=job["D"].dtype, pin_memory=True)
        pinned[slot_].copy_(job["D"], non_blocking=True)
        w_h = job["w"].cpu().numpy()                     # synchronises this batch's stream: the copy above has landed too
        D_h = pinned[slot_].numpy()
        rec = dict(pert_out=[w_h, D_h if save else D_h.copy()], r_s_root=job["r_s_root"], ImpactFrameParams=job["frame"])
        if summaries is not None:                        # phi1_function is the user's host callable; the binning runs on the device
            phi1_b = rt.to_dev(np.asarray(phi1_function(w_h)))
            rec.update(binned=binned_response(phi1_b, job["D"], job["mass"], summaries).cpu().numpy(), masses=job["mass"])
        if save:
            os.makedirs(path, exist_ok=True)
            np.save(os.path.join(path, str(job["i"] + save_iter_start)), rec, allow_pickle=True)
        else:
            out.append(rec)

    depth = max(1, int(pipeline))
    pinned = [None] * depth
    main_stream = tt.cuda.current_stream()
    streams = [tt.cuda.Stream() for _ in range(depth)] if depth > 1 else [main_stream]
    for s_ in streams:
        s_.wait_stream(main_stream)                      # the base model's tables were uploaded on the caller's stream
    it = range(N_iter)
    if progress:
        import tqdm
        it = tqdm.tqdm(it)
    jobs = []
    for i in it:
        s_ = streams[i % depth]
        with tt.cuda.stream(s_):
            jobs.append((s_, launch(i)))
        if len(jobs) == depth:                           # the oldest batch: wait for it (only its stream), fetch, save
            s0, job = jobs.pop(0)
            with tt.cuda.stream(s0):
                finish(job)
    for s0, job in jobs:
        with tt.cuda.stream(s0):
            finish(job)
    for s_ in streams:
        main_stream.wait_stream(s_)
    tt.cuda.current_stream().synchronize()
    return None if save else out
