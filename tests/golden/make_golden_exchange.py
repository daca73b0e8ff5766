"""Golden vectors for replica exchange between DIFFERENT posteriors: the unmodified reference's
``ParallelSampleSMP().sample(...)`` (hmclab/Samplers.py:1807-1977, exchange step :589-669) run in
the build container, one OS process per chain, each with its own seeded Generator.

    python tests/golden/make_golden_exchange.py      (needs /root/reference)

Stored in ``exchange_runs.npz``: the settings, every chain's samples file as the reference wrote it
(rows [model, misfit]; the misfit column of a row written right after an accepted exchange still
holds the misfit of the model that left the chain -- the reference refreshes ``current_x`` only at
the next proposal), the exchange schedule and the accepted-proposal counters.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from _reference_shim import import_reference  # noqa: E402

SETTINGS = dict(dims=6, chains=4, proposals=14, exchange_interval=2, stepsize=0.35, amount_of_steps=4,
                integrator="lf", randomize_stepsize=True, online_thinning=1, smp_seed=5,
                sampler_seeds=[11, 12, 13, 14])


def posteriors(D, s):
    """Chains 0, 2 sample a cold Normal, chains 1, 3 a hot one (3x the variances, shifted mean)."""
    rng = np.random.default_rng(77)
    d = s["dims"]
    mean, var = rng.normal(size=(d, 1)), rng.uniform(0.5, 1.5, size=(d, 1))
    cold = D.Normal(mean, var)
    hot = D.Normal(mean + 0.3, 3.0 * var)
    return [cold, hot, cold, hot], dict(mean=mean, var=var)


def main():
    hmclab = import_reference()
    s = SETTINGS
    posts, inputs = posteriors(hmclab.Distributions, s)
    rng = np.random.default_rng(3)
    q0 = [rng.normal(size=(s["dims"], 1)) for _ in range(s["chains"])]
    out = {}
    runs = []
    for attempt in range(2):          # twice: the multi-process run has to be reproducible
        with tempfile.TemporaryDirectory() as tmp:
            names = [os.path.join(tmp, f"chain{i}.npy") for i in range(s["chains"])]
            samplers = [hmclab.Samplers.HMC(seed=seed) for seed in s["sampler_seeds"]]
            smp = hmclab.Samplers.ParallelSampleSMP(seed=s["smp_seed"])
            smp.sample(samplers, names, posts, overwrite_existing_files=True, proposals=s["proposals"],
                       exchange=True, exchange_interval=s["exchange_interval"], initial_model=q0,
                       kwargs=dict(stepsize=s["stepsize"], amount_of_steps=s["amount_of_steps"],
                                   integrator=s["integrator"], randomize_stepsize=s["randomize_stepsize"],
                                   online_thinning=s["online_thinning"], disable_progressbar=True))
            runs.append(([np.load(n) for n in names], np.array(smp.exchange_schedule)))
    for a, b in zip(runs[0][0], runs[1][0]):
        assert np.array_equal(a, b), "the reference's parallel run is not reproducible"
    files, schedule = runs[0]
    for i, f in enumerate(files):
        out[f"samples{i}"] = f
    out["schedule"] = schedule
    out["q0"] = np.stack([m[:, 0] for m in q0])
    out["mean"], out["var"] = inputs["mean"][:, 0], inputs["var"][:, 0]
    for k, v in s.items():
        out[f"setting_{k}"] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, "exchange_runs.npz"), **out)
    swaps = sum(int(np.any(np.diff(f[:, :-1], axis=0) != 0)) for f in files)
    print("wrote exchange_runs.npz", [f.shape for f in files], "schedule", schedule.shape)


if __name__ == "__main__":
    main()
