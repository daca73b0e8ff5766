"""Chain sharding and the per-block diagnostics gather, world_size 2 over gloo on CPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hmclab_b200.parallel import gather_diagnostics, shard_range


def test_shard_ranges_partition_the_chains():
    for total in (1, 7, 8, 65536, 65537):
        for world in (1, 2, 3, 8):
            ranges = [shard_range(total, world, r) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [hi - lo for lo, hi in ranges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _worker(rank, world, port, total, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(total, world, rank)
    accepted = torch.arange(lo, hi, dtype=torch.int32) * 3
    misfit = torch.arange(lo, hi, dtype=torch.float64) + 0.5
    acc_all, mis_all = gather_diagnostics(accepted, misfit, total)
    np.save(os.path.join(out_dir, f"acc{rank}.npy"), acc_all.numpy())
    np.save(os.path.join(out_dir, f"mis{rank}.npy"), mis_all.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7])
def test_gather_diagnostics_world_size_2(tmp_path, total):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, total, str(tmp_path)), nprocs=2, join=True)
    for rank in range(2):
        acc = np.load(tmp_path / f"acc{rank}.npy")
        mis = np.load(tmp_path / f"mis{rank}.npy")
        assert np.array_equal(acc, np.arange(total) * 3)
        assert np.array_equal(mis, np.arange(total) + 0.5)
