"""Multi-GPU data parallelism for the hot path: one process per GPU, particles sharded, no data-path collective.

The reference is single-device (no pmap / shard_map / NCCL call sites); particles are fully
independent (main.py:343-368), so the stream shards trivially.  Particle i goes to rank i % world (interleaved: every
rank sees the same mix of integration spans ts[-1] - ts[i], which a contiguous split would not give).  NCCL over NVLink is
used only for the final gather of the (N, 6) stream (48 MB at 1e6 particles) and for summing response summaries.
"""
import os

import numpy as np


def dist():
    import torch.distributed as d
    return d


def init_from_env(backend=None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* (torchrun).  Returns (rank, world)."""
    import torch
    d = dist()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not d.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
        d.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world


def shard_indices(n, rank, world):
    """Indices of the units owned by `rank` under interleaved sharding."""
    return np.arange(rank, n, world)


def shard_count(n, rank, world):
    return max(0, (n - rank + world - 1) // world)


_GATHER_BUF = {}


def _gather_buffer(shape, dtype, device):
    """Receive buffers are cached between calls (a 1e6-particle share per GPU is 96 MB; re-allocating it every step costs more
    than the NVLink transfer itself)."""
    key = (tuple(shape), dtype, str(device))
    buf = _GATHER_BUF.get(key)
    if buf is None:
        import torch
        if len(_GATHER_BUF) > 8:
            _GATHER_BUF.clear()
        buf = _GATHER_BUF[key] = torch.empty(shape, dtype=dtype, device=device)
    return buf


def gather_interleaved(local, n_total, rank, world, group=None, axis=0):
    """All-gather per-rank shards (row k of rank r <-> global index r + k*world along `axis`) into the global order.
    ONE collective (all_gather_into_tensor) and ONE permuting copy; `local` is a torch tensor on the process group's device.
    axis = 1 gathers a packed [2, n_local, 6] (lead, trail) pair in a single call."""
    import torch
    if world == 1:
        return local
    d = dist()
    n_max = shard_count(n_total, 0, world)
    n_loc = local.shape[axis]
    src = local.contiguous()
    if n_loc != n_max:                                    # ragged tail: pad this rank's share to the common length
        shape = list(local.shape); shape[axis] = n_max
        src = torch.zeros(shape, dtype=local.dtype, device=local.device)
        src.narrow(axis, 0, n_loc).copy_(local)
    buf = _gather_buffer((world,) + tuple(src.shape), src.dtype, src.device)
    try:
        d.all_gather_into_tensor(buf, src, group=group)
    except (RuntimeError, NotImplementedError):           # backends without the flat collective
        d.all_gather(list(buf.unbind(0)), src, group=group)
    # buf[r, ..., k, ...] -> out[..., k*world + r, ...]
    perm = list(range(1, buf.dim()))
    perm.insert(axis + 1, 0)                              # [..., n_max, world, ...]
    out = buf.permute(perm)
    shape = list(local.shape); shape[axis] = n_max * world
    out = out.reshape(shape)                              # the one permuting copy
    if out.data_ptr() == buf.data_ptr():                  # degenerate shapes reshape without copying: never hand out the cached buffer
        out = out.clone()
    return out if n_max * world == n_total else out.narrow(axis, 0, n_total).contiguous()


def gen_stream_sharded(pot, ts, prog_w0, Msat, seed_num, solver, rank, world, kval_arr=1.0, rtol=1e-7, atol=1e-7, dtmin=0.3, dtmax=None,
                       max_steps=10_000, normals=None, gather=True, compute=None, group=None):
    """gen_stream_vmapped (main.py:343-368) over `world` ranks.  Every rank recomputes the (cheap, serial) progenitor
    orbit and release, integrates its interleaved share of the particles, and the shares are all-gathered.

    compute(i_begin, i_stride, n_local) -> (lead[n_local,6], trail[n_local,6]) may be injected (CPU/gloo tests run the
    host logic with the oracle standing in for the CUDA call)."""
    n = len(ts) - 1
    n_local = shard_count(n, rank, world)
    if compute is None:
        from . import _runtime as rt
        ts_d, w0_d, Ms, kv, nr = pot._stream_inputs(ts, prog_w0, Msat, kval_arr, normals)
        ctrl = rt.make_ctrl(solver, rtol, atol, dtmin, dtmax, max_steps)
        lead, trail, status, nsteps = rt.gen_stream(pot, pot, pot._G, ts_d, w0_d, Ms, 0 if seed_num is None else int(seed_num), kv, nr, ctrl,
                                                    i_begin=rank, i_stride=world, n_local=n_local)
    else:
        lead, trail = compute(rank, world, n_local)
    if not gather:
        return lead, trail
    import torch
    if isinstance(lead, torch.Tensor) and lead.shape == trail.shape:
        packed = getattr(lead, "_ssb_packed", None)          # rt.gen_stream writes both arms into one [2, n_local, 6] buffer
        both = gather_interleaved(torch.stack([lead, trail]) if packed is None else packed, n, rank, world, group, axis=1)   # one collective for both arms
        return both[0], both[1]
    return gather_interleaved(lead, n, rank, world, group), gather_interleaved(trail, n, rank, world, group)


def allreduce_sum(t, world, group=None):
    """Sum of per-rank response summaries (e.g. sum_sh M_sh D[:, sh, :6] binned along the stream)."""
    if world > 1:
        dist().all_reduce(t, group=group)
    return t


def response_summary(D, M_sh, dr_s=None):
    """First-order perturbed-stream displacement from the response derivatives (examples/linear_perturbation_stream.ipynb cell 24):
    sum_sh M_sh D[:, sh, :6] (+ M_sh dr_s[sh] D[:, sh, 6:]).  D: torch [n, n_sh, 12]; returns [n, 6] - the (small) quantity a sharded
    run exchanges instead of the raw derivatives (96 KB per particle at n_sh = 1000)."""
    import torch
    M = torch.as_tensor(M_sh, dtype=D.dtype, device=D.device)
    out = torch.einsum("s,nsk->nk", M, D[:, :, :6])
    if dr_s is not None:
        out = out + torch.einsum("s,nsk->nk", M * torch.as_tensor(dr_s, dtype=D.dtype, device=D.device), D[:, :, 6:])
    return out


def linear_response_sharded(pot_base, sharrays, w0, t0, t1, ctrl, rank, world, M_sh, dr_s=None, compute=None, group=None):
    """compute_perturbation_OTF (perturbative.py:726-755) over `world` ranks: particles (NOT subhalos - the subhalos of one particle
    share its step controller) are dealt out interleaved; every rank keeps its own D[n_local, n_sh, 12] and the ranks all-gather only
    the unperturbed final states and the response summary.  Returns (w[N,6], dstream[N,6], D_local, local_indices).

    compute(w0_local, t0_local) -> (w_local, D_local) may be injected (CPU/gloo tests: the oracle stands in for the CUDA call)."""
    import torch
    N = w0.shape[0]
    sel = torch.as_tensor(shard_indices(N, rank, world), dtype=torch.long, device=w0.device if isinstance(w0, torch.Tensor) else None)
    w0_l, t0_l = w0[sel], t0[sel]
    if compute is None:
        from . import _runtime as rt
        w_l, D_l, status, _ = rt.linear_response(pot_base, sharrays, w0_l.contiguous(), None, t0_l.contiguous(), float(t1), ctrl)
        if bool((status != 0).any()):
            raise RuntimeError("linear_response_sharded: a particle failed (max_steps reached or non-finite state)")
    else:
        w_l, D_l = compute(w0_l, t0_l)
    packed = torch.cat([w_l, response_summary(D_l, M_sh, dr_s)], dim=1)            # [n_local, 12]
    full = gather_interleaved(packed, N, rank, world, group)
    return full[:, :6], full[:, 6:], D_l, sel
