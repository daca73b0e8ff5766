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
