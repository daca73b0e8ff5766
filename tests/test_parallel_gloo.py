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


# ---- replica exchange across ranks (parallel.exchange_round) ----------------------------------------

def _exchange_problem(total, d=5):
    rng = np.random.default_rng(11)
    mu, var = rng.normal(size=(total, d)), rng.uniform(0.5, 2.0, size=(total, d))
    q = rng.normal(size=(total, d))
    x = 0.5 * np.sum((q - mu) ** 2 / var, axis=1)
    rounds = [rng.permutation(total)[: 2 * (total // 2)].reshape(-1, 2) for _ in range(4)]
    uniforms = rng.uniform(size=(4, total // 2))
    return mu, var, q, x, rounds, uniforms


def _run_exchange(total, lo, hi):
    """Four exchange rounds on the shard [lo, hi) of the chains: every chain has its own Normal posterior
    (a tempering ladder); between rounds the models drift a little so that the rounds differ."""
    from hmclab_b200.parallel import exchange_round

    mu, var, q, x, rounds, uniforms = _exchange_problem(total)
    mu_t, var_t = torch.as_tensor(mu[lo:hi]), torch.as_tensor(var[lo:hi])
    q_loc, x_loc = torch.as_tensor(q[lo:hi]).clone(), torch.as_tensor(x[lo:hi]).clone()
    misfit_at = lambda m: 0.5 * torch.sum((m - mu_t) ** 2 / var_t, dim=1)
    decisions, before = [], []
    for r, pairs in enumerate(rounds):
        accepted, q_all = exchange_round(pairs, total, q_loc, x_loc, misfit_at,
                                         lambda p, a, b, r=r: uniforms[r, p] * 0.6)
        decisions.append(accepted)
        before.append(q_all.numpy().copy())
        q_loc += 0.05 * (r + 1)
        x_loc.copy_(misfit_at(q_loc))
    return q_loc.numpy(), x_loc.numpy(), np.array(decisions), np.array(before)


def _exchange_worker(rank, world, port, total, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(total, world, rank)
    q, x, decisions, before = _run_exchange(total, lo, hi)
    np.savez(os.path.join(out_dir, f"ex{rank}.npz"), q=q, x=x, decisions=decisions, before=before)
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7])
def test_exchange_round_world_size_2_equals_one_process(tmp_path, total):
    q1, x1, dec1, before1 = _run_exchange(total, 0, total)           # no process group: one shard
    assert dec1.any() and not dec1.all()                              # both outcomes occur
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_exchange_worker, args=(2, port, total, str(tmp_path)), nprocs=2, join=True)
    parts = [np.load(tmp_path / f"ex{r}.npz") for r in range(2)]
    assert np.array_equal(np.concatenate([p["q"] for p in parts]), q1)
    assert np.array_equal(np.concatenate([p["x"] for p in parts]), x1)
    for p in parts:
        assert np.array_equal(p["decisions"], dec1)
        assert np.array_equal(p["before"], before1)
    # an accepted pair really swapped models: after round 0 chain a holds b's model (+ the drift)
    mu, var, q, x, rounds, _ = _exchange_problem(total)
    a, b = rounds[0][np.argmax(dec1[0])]
    assert np.array_equal(before1[1][a], q[b] + 0.05) and np.array_equal(before1[1][b], q[a] + 0.05)
