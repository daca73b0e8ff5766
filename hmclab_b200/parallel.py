"""Chain sharding across GPUs: one process per GPU, chains are independent units.

The reference's only parallelism is one OS process per chain (``ParallelSampleSMP``,
hmclab/Samplers.py:1757-1998).  Here a process owns a contiguous range of chains on its
GPU; the model constants are replicated, there is no collective inside a block of
proposals, and the per-chain diagnostics (acceptance counters, misfits) are gathered once
per sample block -- over NCCL/NVLink on GPUs, over gloo in the CPU tests.
Random streams are keyed by the *global* chain id, so results do not depend on how many
GPUs the chains are spread over.
"""
from __future__ import annotations

from typing import Tuple


def shard_range(total_chains: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Global chain ids [lo, hi) owned by ``rank``: contiguous, sizes differ by at most 1."""
    if not (0 <= rank < world_size):
        raise ValueError("rank outside [0, world_size)")
    base, extra = divmod(int(total_chains), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def world() -> Tuple[int, int]:
    """(rank, world_size) of the initialised torch.distributed group, else (0, 1)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def all_gather_shards(local, total: int):
    """``local`` [C_local, ...] holds this rank's contiguous shard (``shard_range``) of ``total`` rows;
    returns all ``total`` rows in global order on every rank (one all-gather of equal-size, zero-padded
    pieces: NCCL for CUDA tensors, gloo for CPU tensors).  Without a process group: ``local`` itself."""
    import torch
    import torch.distributed as dist

    rank, size = world()
    if size == 1:
        return local
    sizes = [shard_range(total, size, r) for r in range(size)]
    longest = max(hi - lo for lo, hi in sizes)
    piece = torch.zeros((longest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    piece[: local.shape[0]] = local
    out = torch.empty((size * longest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, piece)
    return torch.cat([out[r * longest: r * longest + hi - lo] for r, (lo, hi) in enumerate(sizes)])


def exchange_round(pairs, total: int, q_local, x_local, misfit_at, draw_uniform):
    """One replica-exchange round of the reference (hmclab/Samplers.py:589-669) over chains that are
    sharded across ranks.

    ``pairs`` [P, 2]: global chain positions (a, b) of the scheduled pairs, b the pair's master;
    every position occurs at most once; identical on all ranks.  ``q_local`` [C_local, d] /
    ``x_local`` [C_local]: models and misfits of this rank's contiguous shard (``shard_range``),
    updated IN PLACE.  ``misfit_at(models [C_local, d]) -> [C_local]``: chain i's OWN posterior at the
    model in row i (the device evaluation; rows of unscheduled chains hold their own model).
    ``draw_uniform(p, a, b)``: the acceptance uniform of pair p, asked only of the rank that owns the
    master b (its sampler's generator), or of every rank when it returns the same number everywhere.

    The exchange step is the one place where chains interact, hence the one collective on this path:
    an all-gather of the models (+ misfits) and one of the partner misfits, 8 (d + 2) bytes per chain,
    over NVLink with NCCL.  The decisions are then taken redundantly on every rank from identical
    numbers.  Returns ``(accepted [P] bool ndarray, q_all [total, d] (the models before the round))``."""
    import numpy as np
    import torch
    import torch.distributed as dist

    rank, size = world()
    lo, hi = shard_range(total, size, rank)
    pairs = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    partner = np.arange(total)
    partner[pairs[:, 0]], partner[pairs[:, 1]] = pairs[:, 1], pairs[:, 0]
    packed = all_gather_shards(torch.cat([q_local, x_local[:, None]], dim=1), total)
    q_all, x_all = packed[:, :-1], packed[:, -1]
    theirs = q_all[torch.as_tensor(partner[lo:hi], device=q_all.device)].contiguous()
    xex_all = all_gather_shards(misfit_at(theirs).reshape(-1, 1), total)[:, 0]
    u = torch.zeros(len(pairs), dtype=torch.float64)
    owner_draws = torch.zeros(len(pairs), dtype=torch.float64)
    for p, (a, b) in enumerate(pairs):
        if lo <= b < hi:
            u[p] = float(draw_uniform(p, int(a), int(b)))
            owner_draws[p] = 1.0
    if size > 1:
        both = torch.stack([u, owner_draws]).to(q_all.device)
        dist.all_reduce(both)
        u = (both[0] / both[1]).cpu()       # every pair has exactly one master, hence one owner
    improvement = (x_all - xex_all).cpu().numpy()
    xex = xex_all.cpu().numpy()
    with np.errstate(all="ignore"):
        accepted = np.exp(improvement[pairs[:, 0]] + improvement[pairs[:, 1]]) > u.numpy()
    for (a, b), ok in zip(pairs, accepted):
        if not ok:
            continue
        for mine, other in ((a, b), (b, a)):
            if lo <= mine < hi:
                q_local[mine - lo] = q_all[other]
                x_local[mine - lo] = float(xex[mine])
    return accepted, q_all


def gather_diagnostics(accepted, misfit, total_chains: int):
    """All-gather of per-chain diagnostics of this rank's shard.

    ``accepted`` int32 [C_local], ``misfit`` float64 [C_local] (CPU tensors with gloo, CUDA
    tensors with NCCL).  Returns (accepted_all [total], misfit_all [total]) in global chain
    order on every rank.  One collective pair per sample block; nothing here sits on the
    per-proposal path."""
    import torch
    import torch.distributed as dist

    rank, size = world()
    if size == 1:
        return accepted, misfit
    sizes = [shard_range(total_chains, size, r) for r in range(size)]
    longest = max(hi - lo for lo, hi in sizes)

    def padded(t):
        out = torch.zeros(longest, dtype=t.dtype, device=t.device)
        out[: t.numel()] = t
        return out

    acc_buf = [torch.empty(longest, dtype=accepted.dtype, device=accepted.device) for _ in range(size)]
    mis_buf = [torch.empty(longest, dtype=misfit.dtype, device=misfit.device) for _ in range(size)]
    dist.all_gather(acc_buf, padded(accepted))
    dist.all_gather(mis_buf, padded(misfit))
    acc_all = torch.cat([b[: hi - lo] for b, (lo, hi) in zip(acc_buf, sizes)])
    mis_all = torch.cat([b[: hi - lo] for b, (lo, hi) in zip(mis_buf, sizes)])
    return acc_all, mis_all
