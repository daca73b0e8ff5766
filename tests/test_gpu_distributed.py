"""HMC.sample(distributed=True) under torchrun on two GPUs: chain sharding, per-rank files,
NCCL gather of the diagnostics; the union of the rank files equals a single-GPU run.
Skipped on boxes with one GPU (the gloo CPU test covers the host logic there)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, {root!r})
from hmclab_b200 import workloads
from hmclab_b200.Samplers import HMC
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
w = workloads.source_location(events=4, stations=6, chains=10)
s = HMC(seed=3).sample(os.path.join({out!r}, "dist.npy"), w.posterior, stepsize=w.stepsize, proposals=8,
                       mass_matrix=w.mass_matrix, initial_model=w.initial_models, distributed=True)
assert s.accepted_proposals_all_chains.shape == (10,)
np.save(os.path.join({out!r}, f"acc{{dist.get_rank()}}.npy"), s.accepted_proposals_all_chains)
dist.destroy_process_group()
'''


def test_two_gpu_run_equals_single_gpu_run(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from hmclab_b200 import workloads
    from hmclab_b200.Samplers import HMC

    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, out=str(tmp_path)))
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                           "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                           "--master-port", "29541", str(script)], cwd=ROOT)
    w = workloads.source_location(events=4, stations=6, chains=10)
    single = HMC(seed=3).sample(str(tmp_path / "single.npy"), w.posterior, stepsize=w.stepsize,
                                proposals=8, mass_matrix=w.mass_matrix, initial_model=w.initial_models)
    whole = np.load(tmp_path / "single.npy").reshape(10, 8, -1)
    r0 = np.load(tmp_path / "dist.rank0.npy").reshape(5, 8, -1)
    r1 = np.load(tmp_path / "dist.rank1.npy").reshape(5, 8, -1)
    assert np.array_equal(np.concatenate([r0, r1]), whole)
    acc = np.load(tmp_path / "acc0.npy")
    assert np.array_equal(acc, np.load(tmp_path / "acc1.npy"))
    assert np.array_equal(acc, single.accepted_proposals_per_chain)


EXCHANGE_WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, {root!r})
sys.path.insert(0, os.path.join({root!r}, "tests"))
from test_gpu_distributed import tempering_run
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
smp = tempering_run(os.path.join({out!r}, "dist"))
np.save(os.path.join({out!r}, f"exacc{{dist.get_rank()}}.npy"), np.array([smp.exchanges_accepted]))
dist.destroy_process_group()
'''


def tempering_run(prefix):
    """Seven chains on a ladder of four posteriors (cold ... hot; the posteriors repeat, so the engines
    hold groups of one or two chains), replica exchange every third proposal, device random streams."""
    from hmclab_b200 import Distributions as D
    from hmclab_b200.Samplers import HMC, ParallelSampleSMP

    rng = np.random.default_rng(5)
    d, n = 12, 7
    mean, var = rng.normal(size=(d, 1)), rng.uniform(0.5, 1.5, size=(d, 1))
    ladder = [D.Normal(mean + 0.1 * t, var * (1.0 + 0.7 * t)) for t in range(4)]
    posts = [ladder[i % 4] for i in range(n)]
    os.makedirs(prefix, exist_ok=True)
    names = [os.path.join(prefix, f"chain{i}.npy") for i in range(n)]
    smp = ParallelSampleSMP(seed=21)
    smp.sample([HMC(seed=i) for i in range(n)], names, posts, overwrite_existing_files=True, proposals=30,
               exchange=True, exchange_interval=3, initial_model=[q[:, None] for q in rng.normal(size=(n, d))],
               kwargs=dict(stepsize=0.25, amount_of_steps=5, online_thinning=1, disable_progressbar=True))
    return smp


def test_replica_exchange_across_two_gpus_equals_single_gpu_run(tmp_path):
    """ParallelSampleSMP under torchrun: chains sharded over two ranks, exchange rounds over NCCL
    (parallel.exchange_round); every chain's file equals the one a single GPU writes, bit for bit."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "exchange_worker.py"
    script.write_text(EXCHANGE_WORKER.format(root=ROOT, out=str(tmp_path)))
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                           "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                           "--master-port", "29543", str(script)], cwd=ROOT)
    single = tempering_run(str(tmp_path / "single"))
    assert single.exchanges_accepted > 0
    assert int(np.load(tmp_path / "exacc0.npy")[0]) == single.exchanges_accepted
    assert int(np.load(tmp_path / "exacc1.npy")[0]) == single.exchanges_accepted
    for i in range(7):
        one = np.load(tmp_path / "single" / f"chain{i}.npy")
        two = np.load(tmp_path / "dist" / f"chain{i}.npy")
        assert one.shape == (30, 13) and np.array_equal(one, two), i


GOLDEN_EXCHANGE_WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, {root!r})
sys.path.insert(0, os.path.join({root!r}, "tests"))
from test_gpu_distributed import reference_exchange_run
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
reference_exchange_run({out!r})
dist.destroy_process_group()
'''


def reference_exchange_run(out_dir):
    """The reference's own ParallelSampleSMP run of tests/golden/exchange_runs.npz (two posteriors, four
    chains, host random streams in the reference's order) on whatever process group is initialised."""
    from hmclab_b200 import Distributions as D
    from hmclab_b200.Samplers import HMC, ParallelSampleSMP

    gold = np.load(os.path.join(ROOT, "tests", "golden", "exchange_runs.npz"))
    st = {k[len("setting_"):]: gold[k] for k in gold.files if k.startswith("setting_")}
    n = int(st["chains"])
    mean, var = gold["mean"][:, None], gold["var"][:, None]
    cold, hot = D.Normal(mean, var), D.Normal(mean + 0.3, 3.0 * var)
    names = [os.path.join(out_dir, f"gold_chain{i}.npy") for i in range(n)]
    smp = ParallelSampleSMP(seed=int(st["smp_seed"]))
    smp.sample([HMC(seed=int(s)) for s in st["sampler_seeds"]], names, [cold, hot, cold, hot],
               overwrite_existing_files=True, proposals=int(st["proposals"]), exchange=True,
               exchange_interval=int(st["exchange_interval"]), initial_model=[q[:, None] for q in gold["q0"]],
               kwargs=dict(stepsize=float(st["stepsize"]), amount_of_steps=int(st["amount_of_steps"]),
                           integrator=str(st["integrator"]), randomize_stepsize=bool(st["randomize_stepsize"]),
                           online_thinning=int(st["online_thinning"]), disable_progressbar=True, host_rng=True))
    return gold, names


def test_reference_exchange_run_is_reproduced_across_two_gpus(tmp_path):
    """The unmodified reference's multi-process replica-exchange run (golden) reproduced with the cold chains
    on one GPU and the hot chains on the other: every swap crosses the NVLink, the masters' acceptance
    uniforms come from the owning rank's sampler generators (all-reduced), and every file matches the
    reference's to 1e-10."""
    import torch

    from helpers import rel_err

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "gold_exchange_worker.py"
    script.write_text(GOLDEN_EXCHANGE_WORKER.format(root=ROOT, out=str(tmp_path)))
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                           "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                           "--master-port", "29545", str(script)], cwd=ROOT)
    gold = np.load(os.path.join(ROOT, "tests", "golden", "exchange_runs.npz"))
    for i in range(int(gold["setting_chains"])):
        got, ref = np.load(tmp_path / f"gold_chain{i}.npy"), gold[f"samples{i}"]
        assert got.shape == ref.shape and rel_err(got, ref) < 1e-10, i
