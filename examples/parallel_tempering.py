"""Replica exchange over a temperature ladder with the reference's ``ParallelSampleSMP`` API
(hmclab/Samplers.py:1807-1998), on one GPU or sharded over several.

    python examples/parallel_tempering.py [--chains 64] [--proposals 2000]
    torchrun --nproc-per-node 2 examples/parallel_tempering.py      # same files, chains split over 2 GPUs

A bimodal-looking target is imitated by a ladder of tempered Gaussian likelihoods: chain i samples
``prior x likelihood^(1/T_i)``; every ``--interval`` proposals scheduled pairs of chains try to swap their
models with the reference's acceptance rule.  Chains that share a temperature share one engine (one batch);
an exchange round is an all-gather of the models and of the partner misfits (NCCL between GPUs).  Every
chain writes its own reference-format samples file; the files do not depend on the number of GPUs.
"""
import argparse
import os
import tempfile

import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # run from a checkout
import hmclab_b200 as hmclab  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--chains", type=int, default=64)
ap.add_argument("--temperatures", type=int, default=4)
ap.add_argument("--proposals", type=int, default=2000)
ap.add_argument("--interval", type=int, default=10)
ap.add_argument("--out", default=None)
args = ap.parse_args()

distributed = "LOCAL_RANK" in os.environ
if distributed:
    import torch
    import torch.distributed as dist

    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

rng = np.random.default_rng(7)
dims, data = 40, 120
G = rng.normal(size=(data, dims)) / np.sqrt(data)
m_true = rng.normal(size=(dims, 1))
d_obs = G @ m_true + 0.05 * rng.normal(size=(data, 1))
prior = hmclab.Distributions.Normal(np.zeros((dims, 1)), 4.0)
ladder = [hmclab.Distributions.BayesRule([prior, hmclab.Distributions.LinearMatrix(G, d_obs, 0.05**2 * 4.0**t)])
          for t in range(args.temperatures)]                       # variance x T: the tempered likelihood
posteriors = [ladder[i % args.temperatures] for i in range(args.chains)]

out = args.out or tempfile.mkdtemp(prefix="tempering_")
os.makedirs(out, exist_ok=True)
names = [os.path.join(out, f"chain{i:03d}_T{i % args.temperatures}.npy") for i in range(args.chains)]
front = hmclab.Samplers.ParallelSampleSMP(seed=11)                 # the same seed on every rank
front.sample([hmclab.Samplers.HMC(seed=i) for i in range(args.chains)], names, posteriors,
             overwrite_existing_files=True, proposals=args.proposals, exchange=True,
             exchange_interval=args.interval,
             initial_model=[rng.normal(size=(dims, 1)) for _ in range(args.chains)],
             kwargs=dict(stepsize=0.02, amount_of_steps=10, online_thinning=5, disable_progressbar=True))

rank = dist.get_rank() if distributed else 0
if rank == 0:
    rounds = front.exchange_schedule.shape[0]
    print(f"{rounds} exchange rounds, {front.exchanges_accepted} accepted swaps "
          f"({front.exchanges_accepted / max(1, rounds * (args.chains // 2)):.2f} of the scheduled pairs)")
    cold = [n for n in names if n.endswith("_T0.npy") and os.path.exists(n)]
    s = np.concatenate([np.load(n)[len(np.load(n)) // 2:, :dims] for n in cold])
    print(f"cold chains on this rank: {len(cold)}; posterior mean error {np.abs(s.mean(0) - m_true[:, 0]).max():.3f}, "
          f"files in {out}")
if distributed:
    dist.destroy_process_group()
