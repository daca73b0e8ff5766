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
