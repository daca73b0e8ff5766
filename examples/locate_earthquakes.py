"""Earthquake source location with thousands of chains (the reference's "locating quakes"
example, hmclab notebooks/examples, on the batched engine).

    python examples/locate_earthquakes.py [--chains 4096] [--proposals 2000]

Sixteen events, thirty surface stations, fixed medium velocity, a uniform box prior placed
directly in BayesRule so that trajectories reflect on its walls.  Every chain adapts its own
step size; the samples of all chains end up in one reference-format ``.npy`` file.
"""
import argparse
import os
import tempfile

import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # run from a checkout
import hmclab_b200 as hmclab  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--chains", type=int, default=4096)
ap.add_argument("--proposals", type=int, default=2000)
ap.add_argument("--out", default=None)
args = ap.parse_args()

rng = np.random.default_rng(4)
events, stations, v = 16, 30, 3.0
sx, sy, sz = rng.uniform(-10, 30, (1, stations)), rng.uniform(-10, 30, (1, stations)), np.zeros((1, stations))
ex, ey = rng.uniform(0, 20, (events, 1)), rng.uniform(0, 20, (events, 1))
ez, eT = rng.uniform(0, 10, (events, 1)), rng.uniform(0, 10, (events, 1))
tt = hmclab.Distributions.SourceLocation3D.forward(ex, ey, ez, eT, v, sx, sy, sz)
std = 0.1 * np.ones_like(tt)
tobs = tt + std * rng.normal(size=tt.shape)
tobs[3, 7] = np.nan                                   # a missing pick is simply skipped

likelihood = hmclab.Distributions.SourceLocation3D(sx, sy, sz, tobs, std, infer_velocity=False,
                                                   medium_velocity=v)
lo = np.tile(np.array([[-10.0], [-10.0], [0.0], [-5.0]]), (events, 1))
hi = np.tile(np.array([[30.0], [30.0], [20.0], [15.0]]), (events, 1))
posterior = hmclab.Distributions.BayesRule([hmclab.Distributions.Uniform(lo, hi), likelihood])

truth = np.hstack([ex, ey, ez, eT]).reshape(-1)
start = np.clip(truth[None, :] + 0.5 * rng.normal(size=(args.chains, 4 * events)), lo[:, 0] + 1e-3, hi[:, 0] - 1e-3)
out = args.out or os.path.join(tempfile.mkdtemp(), "quakes.npy")
sampler = hmclab.Samplers.HMC(seed=1).sample(
    out, posterior, stepsize=0.004, amount_of_steps=10, proposals=args.proposals, online_thinning=10,
    chains=args.chains, initial_model=start, autotuning=True, overwrite_existing_file=True)

with hmclab.Samples(out, burn_in=0) as samples:
    per = int(samples.read_attribute("samples_per_chain"))
    kept = np.stack([samples.chain(c)[:-1, per // 2:] for c in range(args.chains)])   # [C, d, n]
mean, sd = kept.mean(axis=(0, 2)), kept.std(axis=(0, 2))
worst = np.max(np.abs(mean - truth) / sd)
print(f"{args.chains} chains x {args.proposals} proposals -> {out}")
print(f"acceptance rate {sampler.accepted_proposals / (args.chains * args.proposals):.2f}, "
      f"adapted step sizes {np.min(sampler.stepsize):.4f}..{np.max(sampler.stepsize):.4f}")
print(f"posterior mean within {worst:.2f} posterior standard deviations of the true hypocentres")
assert worst < 4.0
