"""Straight-ray travel-time tomography: sparse forward operator, Laplace (L1) prior.

    python examples/linear_tomography.py [--grid 40] [--rays 4000] [--chains 1024]

The reference expresses this problem as ``LinearMatrix`` with a user-built scipy sparse ``G``
(row = ray, entries = path length per cell); the same objects drive the CUDA engine here.
"""
import argparse
import os
import tempfile

import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # run from a checkout
import hmclab_b200 as hmclab  # noqa: E402
from hmclab_b200.workloads import straight_ray_matrix

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=40)
ap.add_argument("--rays", type=int, default=4000)
ap.add_argument("--chains", type=int, default=1024)
ap.add_argument("--proposals", type=int, default=400)
args = ap.parse_args()

n = args.grid
rng = np.random.default_rng(3)
G = straight_ray_matrix(n, n, args.rays, seed=3)                     # scipy CSR [rays x cells]
yy, xx = np.mgrid[0:n, 0:n]
s0 = np.full((n * n, 1), 0.5)                                        # background slowness
s_true = s0 + (0.1 * np.exp(-((xx - 0.6 * n) ** 2 + (yy - 0.4 * n) ** 2) / (0.02 * n * n))).reshape(-1, 1)
sigma2 = 0.05
d = G @ s_true + np.sqrt(sigma2) * rng.normal(size=(args.rays, 1))

posterior = hmclab.Distributions.BayesRule([
    hmclab.Distributions.Laplace(s0, np.full((n * n, 1), 0.1)),
    hmclab.Distributions.LinearMatrix(G, d, sigma2, premultiplication=False)])
start = s_true[:, 0][None, :] + 0.002 * rng.normal(size=(args.chains, n * n))
out = os.path.join(tempfile.mkdtemp(), "tomography.npy")
sampler = hmclab.Samplers.HMC(seed=2).sample(
    out, posterior, stepsize=0.002, amount_of_steps=10, proposals=args.proposals, online_thinning=20,
    chains=args.chains, initial_model=start, overwrite_existing_file=True)

with hmclab.Samples(out) as samples:
    mean = np.asarray(samples.samples).mean(axis=1)
err = np.abs(mean - s_true[:, 0])
print(f"engine path: {sampler.engine.path}; acceptance rate "
      f"{sampler.accepted_proposals / (args.chains * args.proposals):.2f}")
print(f"posterior-mean slowness error: median {np.median(err):.4f}, max {err.max():.4f} "
      f"(anomaly amplitude 0.1)")
assert np.median(err) < 0.02
