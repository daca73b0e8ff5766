"""HMC().sample(...) end to end on the GPU: reference reproduction from the same seed,
file format, determinism, chain-sharding invariance, early termination."""
import os
import time

import numpy as np
import pytest

import cases
from helpers import GOLDEN_DIR, build_mirror, load_golden, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-10


@pytest.mark.parametrize("name", ["dense_premult_cfg1", "srcloc_fixed_v", "normal_bounded"])
def test_same_seed_reproduces_the_reference_samples_file(tmp_path, name):
    """hmclab.Samplers.HMC(seed).sample(...) (run in the build container, stored in
    tests/golden/seeded_public_runs.npz) vs hmclab_b200 with host_rng=True, same seed."""
    from hmclab_b200.Samplers import HMC

    gold = np.load(os.path.join(GOLDEN_DIR, "seeded_public_runs.npz"))
    ref = gold[f"{name}__samples"]
    seed, thin = int(gold[f"{name}__seed"]), int(gold[f"{name}__thinning"])
    inp, _ = load_golden(name)
    s = cases.SETTINGS[name]
    post, mass = build_mirror(name, inp)
    fn = str(tmp_path / "run.npy")
    sampler = HMC(seed=seed).sample(
        fn, post, stepsize=s["stepsize"], randomize_stepsize=s["randomize"],
        amount_of_steps=s["steps"], mass_matrix=mass, integrator=s["integrator"],
        initial_model=inp["q0"][0].copy(), proposals=24, online_thinning=thin,
        disable_progressbar=True, host_rng=True, block_proposals=10)
    got = np.load(fn)
    assert got.shape == ref.shape
    # identical accept/reject sequence <=> identical pattern of repeated rows
    assert np.array_equal(np.all(np.diff(got, axis=0) == 0, axis=1),
                          np.all(np.diff(ref, axis=0) == 0, axis=1))
    assert rel_err(got, ref) < TOL
    assert sampler.accepted_proposals == int(gold[f"{name}__accepted"])
    assert sampler.current_model.shape == (post.dimensions, 1)
    if thin == 1:  # with thinning the last stored proposal (k % thin == 0) precedes the last one
        assert rel_err(sampler.current_model[:, 0], ref[-1, :-1]) < TOL
    assert isinstance(sampler.current_x, float)


@pytest.mark.parametrize("ext", [".npy", ".h5"])
def test_file_attributes_and_reader(tmp_path, ext):
    """Both of the reference's formats: NumPy + pickle, and HDF5 (h5py when installed, else the
    native writer of hmclab_b200._hdf5)."""
    from hmclab_b200 import Distributions as D
    from hmclab_b200.Samplers import HMC
    from hmclab_b200.Samples import Samples

    d, C = 6, 5
    post = D.Normal(np.zeros((d, 1)), 1.0)
    fn = str(tmp_path / f"multi{ext}")
    sampler = HMC(seed=3).sample(fn, post, stepsize=0.3, proposals=40, online_thinning=4,
                                 chains=C, disable_progressbar=True, block_proposals=12)
    with Samples(fn) as s:
        assert s.numpy.shape == (d + 1, C * 10)
        for key in ("write_index", "last_written_sample", "proposals", "acceptance_rate",
                    "online_thinning", "start_time", "sampler", "stepsize", "amount_of_steps",
                    "mass_matrix", "integrator", "end_time", "runtime", "runtime_seconds",
                    "chains", "samples_per_chain"):
            s.read_attribute(key)
        assert s.read_attribute("write_index") == C * 10
        assert s.read_attribute("sampler") == "Hamiltonian Monte Carlo"
        assert s.read_attribute("integrator") == "leapfrog integrator"
        assert s.read_attribute("mass_matrix") == "unit mass matrix"
        last = s.chain(C - 1)[:, -1]
        # stored misfit is chi of the stored model: 0.5 * |m|^2
        assert rel_err(np.ravel(s.misfits), 0.5 * np.sum(s.samples ** 2, axis=0)) < 1e-13
    assert last.shape == (d + 1,) and sampler.current_model.shape == (d, C)
    assert sampler.current_x.shape == (C,)
    assert sampler.amount_of_writes == 10 and sampler.current_proposal == 39
    assert 0 < sampler.accepted_proposals <= 40 * C
    with pytest.raises(FileExistsError):
        HMC(seed=3).sample(fn, post, proposals=4, chains=C)
    assert sampler.load_results(burn_in=2).shape == (d + 1, C * (10 - 2))   # burn-in is per chain


def test_device_rng_is_reproducible_and_independent_of_sharding(tmp_path):
    """Same seed -> same file; chains [0,8) in one engine == chains [0,4) + [4,8) in two
    engines (random streams are keyed by the global chain id)."""
    import torch

    from hmclab_b200._engine import Engine
    from hmclab_b200._lowering import describe, describe_mass, flatten
    from hmclab_b200 import workloads
    from hmclab_b200.Samplers import HMC

    w = workloads.source_location(events=3, stations=6, chains=8)
    a = HMC(seed=11).sample(str(tmp_path / "a.npy"), w.posterior, stepsize=w.stepsize, proposals=12,
                            mass_matrix=w.mass_matrix, initial_model=w.initial_models)
    b = HMC(seed=11).sample(str(tmp_path / "b.npy"), w.posterior, stepsize=w.stepsize, proposals=12,
                            mass_matrix=w.mass_matrix, initial_model=w.initial_models,
                            block_proposals=5)
    assert np.array_equal(np.load(a.samples_filename), np.load(b.samples_filename))
    c = HMC(seed=12).sample(str(tmp_path / "c.npy"), w.posterior, stepsize=w.stepsize, proposals=12,
                            mass_matrix=w.mass_matrix, initial_model=w.initial_models)
    assert not np.array_equal(np.load(a.samples_filename), np.load(c.samples_filename))

    plan, mplan = flatten(describe(w.posterior)), describe_mass(w.mass_matrix)

    def run(lo, hi):
        eng = Engine(plan, mplan, hi - lo, integrator="3s", amount_of_steps=4)
        q = torch.as_tensor(w.initial_models[lo:hi]).cuda().contiguous()
        x = eng.misfit(q)
        out = torch.zeros(6, hi - lo, w.dims + 1, dtype=torch.float64, device="cuda")
        eng.run_block(q, x, 6, stepsize=w.stepsize, seed=99, chain_offset=lo, out_samples=out)
        return out.cpu().numpy()

    whole = run(0, 8)
    assert np.array_equal(whole[:, :4], run(0, 4)) and np.array_equal(whole[:, 4:], run(4, 8))


def test_device_momenta_are_standard_normal():
    """One proposal with a huge mass-less free flight: q1 - q0 = eps * p0 exposes the draws."""
    import torch

    from hmclab_b200._engine import Engine

    d, C = 512, 2048
    big = 1e30   # prior variance so large that the gradient vanishes
    plan = {"dims": d, "terms": [{"kind": "normal", "offset": 0, "len": d, "a": np.zeros(d),
                                  "b": np.full(d, 1.0 / big), "const": 0.0}],
            "checks": [], "likelihood": None, "reflect_lb": None, "reflect_ub": None}
    eng = Engine(plan, {"kind": "unit", "dims": d}, C, integrator="lf", amount_of_steps=1)
    q = torch.zeros(C, d, dtype=torch.float64, device="cuda")
    x = eng.misfit(q)
    qp = torch.zeros(1, C, d, dtype=torch.float64, device="cuda")
    pp = torch.zeros(1, C, d, dtype=torch.float64, device="cuda")
    eng.run_block(q, x, 1, stepsize=1.0, randomize_stepsize=False, seed=5, out_q_prop=qp, out_p_prop=pp)
    z = pp[0].cpu().numpy()
    n = z.size
    assert abs(z.mean()) < 5 / np.sqrt(n)
    assert abs(z.var() - 1) < 5 * np.sqrt(2 / n)
    assert abs((z ** 3).mean()) < 5 * np.sqrt(15 / n)
    assert abs((z ** 4).mean() - 3) < 5 * np.sqrt(96 / n)
    assert np.abs(z).max() < 7 and np.abs(z).max() > 4
    # chains and coordinates are uncorrelated
    assert abs(np.corrcoef(z[0], z[1])[0, 1]) < 0.25
    assert abs(np.mean(z[:, 0] * z[:, 1])) < 5 / np.sqrt(C)
    # Kolmogorov-Smirnov against the normal CDF
    from scipy import stats

    assert stats.kstest(z.ravel()[:200000], "norm").pvalue > 1e-4


def test_sampling_statistics_of_a_known_target(tmp_path):
    from hmclab_b200 import Distributions as D
    from hmclab_b200.Samplers import HMC
    from hmclab_b200.Samples import Samples

    d, C = 8, 512
    mu = np.linspace(-1, 1, d).reshape(d, 1)
    var = np.linspace(0.5, 2.0, d).reshape(d, 1)
    post = D.Normal(mu, var)
    fn = str(tmp_path / "stat.npy")
    HMC(seed=21).sample(fn, post, stepsize=0.5, amount_of_steps=8, proposals=300, online_thinning=3,
                        chains=C, initial_model=np.tile(mu.T, (C, 1)), integrator="4s")
    with Samples(fn, burn_in=0) as s:
        per = int(s.read_attribute("samples_per_chain"))
        x = np.stack([s.chain(c)[:-1, per // 4:] for c in range(C)])   # [C, d, n]
        rate = s.read_attribute("acceptance_rate")
    assert 0.6 < rate <= 1.0
    n_eff = C * 20
    assert np.all(np.abs(x.mean(axis=(0, 2)) - mu[:, 0]) < 5 * np.sqrt(var[:, 0] / n_eff))
    assert np.all(np.abs(x.var(axis=(0, 2)) / var[:, 0] - 1) < 0.1)


@pytest.mark.parametrize("ext", [".npy", ".h5"])
def test_max_time_stops_early_and_leaves_a_valid_file(tmp_path, ext):
    from hmclab_b200 import workloads
    from hmclab_b200.Samplers import HMC
    from hmclab_b200.Samples import Samples

    w = workloads.normal_iid(dims=200, chains=256)
    fn = str(tmp_path / f"cut{ext}")
    t0 = time.time()
    sampler = HMC(seed=1).sample(fn, w.posterior, stepsize=0.05, proposals=2_000_000,
                                 online_thinning=1000, chains=256, initial_model=w.initial_models,
                                 max_time=0.5, block_proposals=2000)
    assert time.time() - t0 < 30
    done = sampler.current_proposal + 1
    # blocks are sized from the measured rate once max_time is set, so the run stops close to
    # the limit (the reference checks after every proposal), wherever that falls inside a block
    assert 0 < done < 2_000_000
    assert sampler.end_time is not None and (sampler.end_time - sampler.start_time).total_seconds() < 0.5 * 1.5 + 1.0
    with Samples(fn) as s:
        per = s.read_attribute("samples_per_chain")
        assert per == -(-done // 1000) and s.numpy.shape == (201, 256 * per)
        assert np.all(np.isfinite(s.numpy))


def test_autotuning_through_the_sampler(tmp_path):
    """autotuning=True: every chain adapts its own step size towards the target acceptance
    rate; chain 0 with host_rng reproduces what the reference's update rule gives."""
    from hmclab_b200 import Distributions as D
    from hmclab_b200.Samplers import HMC
    from hmclab_b200.Samples import Samples

    d, C = 20, 64
    post = D.Normal(np.zeros((d, 1)), np.linspace(0.5, 2.0, d).reshape(d, 1))
    fn = str(tmp_path / "auto.npy")
    sampler = HMC(seed=5).sample(fn, post, stepsize=0.05, amount_of_steps=6, proposals=600, chains=C,
                                 autotuning=True, target_acceptance_rate=0.65, learning_rate=0.75,
                                 block_proposals=100)
    assert sampler.stepsize.shape == (C,) and np.all(sampler.stepsize > 0.05)
    assert sampler.stepsizes.shape == (600, C) and sampler.acceptance_rates.shape == (600, C)
    assert np.all(sampler.stepsizes[0] == 0.05)
    late = np.minimum(sampler.acceptance_rates[300:], 1.0).mean()
    assert abs(late - 0.65) < 0.1
    with Samples(fn) as s:
        assert s.read_attribute("final_stepsizes").shape == (C,)
        assert s.read_attribute("stepsizes").shape == (600, C)


def test_parallel_sample_smp_front_end(tmp_path):
    """The reference's multi-chain API: one file per chain, one batch on the GPU."""
    from hmclab_b200 import workloads
    from hmclab_b200.Samplers import HMC, ParallelSampleSMP
    from hmclab_b200.Samples import Samples

    w = workloads.dense_small(chains=6)
    n, d = 6, w.dims
    names = [str(tmp_path / f"chain_{i}.npy") for i in range(n)]
    kw = dict(stepsize=w.stepsize, amount_of_steps=5, online_thinning=2, disable_progressbar=True)
    models = [w.initial_models[i].reshape(d, 1) for i in range(n)]

    def run(exchange, interval=1, seed=4):
        samplers = [HMC(seed=9) for _ in range(n)]
        par = ParallelSampleSMP(seed=seed).sample(samplers, names, [w.posterior] * n,
                                                   overwrite_existing_files=True, proposals=20,
                                                   exchange=exchange, exchange_interval=interval,
                                                   initial_model=models, kwargs=kw)
        out = []
        for name in names:
            with Samples(name) as smp:
                assert smp.numpy.shape == (d + 1, 10)
                assert smp.read_attribute("write_index") == 10
                assert 0.0 <= smp.read_attribute("acceptance_rate") <= 1.0
                out.append(np.array(smp.numpy))
        return np.stack(out), par

    plain, _ = run(False)
    batched = HMC(seed=9).sample(str(tmp_path / "b.npy"), w.posterior, proposals=20, chains=n,
                                 initial_model=w.initial_models, **kw)
    with Samples(batched.samples_filename) as smp:
        for c in range(n):
            assert np.array_equal(plain[c], smp.chain(c))
    swapped, par = run(True, interval=4)
    assert par.exchange_schedule.shape == (5, 6)
    assert all(sorted(row) == list(range(6)) for row in par.exchange_schedule)
    # proposal 0 is stored after the first exchange round: the same rows, permuted over the chains
    perm = np.arange(n)
    for a, b in par.exchange_schedule[0].reshape(-1, 2):
        perm[a], perm[b] = b, a
    assert np.array_equal(swapped[:, :, 0], plain[perm][:, :, 0])
    assert not np.array_equal(swapped, plain)
    again, _ = run(True, interval=4)
    assert np.array_equal(again, swapped)                       # same seeds -> same run
    far, _ = run(True, interval=50)                            # no exchange round inside 20 proposals
    assert np.array_equal(far, plain)
    with pytest.raises(AssertionError, match="overwriting"):
        ParallelSampleSMP().sample([HMC()], names[:1], [w.posterior])
    with pytest.raises(AssertionError, match="same dimensions"):   # different posteriors are fine, different sizes not
        other = workloads.normal_iid(dims=w.dims + 3, chains=1).posterior
        ParallelSampleSMP().sample([HMC(), HMC()], names[:2], [w.posterior, other],
                                   overwrite_existing_files=True)


@pytest.mark.parametrize("builder,kw", [("normal_iid", dict(dims=130, chains=70)),
                                        ("dense_small", dict(chains=40)),
                                        ("tomography", dict(nx=12, ny=9, rays=200, chains=33))])
def test_sample_host_equals_device_blocks(builder, kw):
    """hmcb_sample_host (host buffers, private streams, double-buffered D2H) returns exactly
    what hmcb_run_block produces on device tensors for the same seed."""
    import torch

    from hmclab_b200 import workloads
    from hmclab_b200._engine import Engine
    from hmclab_b200._lowering import describe, describe_mass, flatten

    w = workloads.BUILDERS[builder](**kw)
    C, d = w.chains, w.dims
    eng = Engine(flatten(describe(w.posterior)), describe_mass(w.mass_matrix), C,
                 integrator=w.integrator, amount_of_steps=w.amount_of_steps)
    P, thin = 12, 3
    q0 = torch.as_tensor(w.initial_models).contiguous().pin_memory()
    samples = torch.empty(P // thin, C, d + 1, dtype=torch.float64).pin_memory()
    acc = torch.empty(C, dtype=torch.int32).pin_memory()
    fq = torch.empty(C, d, dtype=torch.float64).pin_memory()
    fx = torch.empty(C, dtype=torch.float64).pin_memory()
    eng.sample_host(q0, P, stepsize=w.stepsize, thinning=thin, block_proposals=6, seed=17,
                    chain_offset=5, samples=samples, accepted=acc, final_q=fq, final_x=fx)
    q = q0.cuda()
    x = eng.misfit(q)
    ref = torch.zeros(P // thin, C, d + 1, dtype=torch.float64, device="cuda")
    racc = torch.zeros(C, dtype=torch.int32, device="cuda")
    eng.run_block(q, x, P, stepsize=w.stepsize, thinning=thin, seed=17, chain_offset=5,
                  out_samples=ref, accepted_total=racc)
    assert np.array_equal(samples.numpy(), ref.cpu().numpy())
    assert np.array_equal(acc.numpy(), racc.cpu().numpy())
    assert np.array_equal(fq.numpy(), q.cpu().numpy()) and np.array_equal(fx.numpy(), x.cpu().numpy())


def test_rwmh_sampler_statistics_and_format(tmp_path):
    from hmclab_b200 import Distributions as D
    from hmclab_b200.Samplers import RWMH
    from hmclab_b200.Samples import Samples

    d, C = 4, 512
    var = np.array([[0.5], [1.0], [2.0], [4.0]])
    post = D.Normal(np.zeros((d, 1)), var)
    fn = str(tmp_path / "rw.npy")
    sampler = RWMH(seed=2).sample(fn, post, stepsize=np.sqrt(var) * 1.2, proposals=600,
                                  online_thinning=3, chains=C, block_proposals=90)
    with Samples(fn) as s:
        assert s.read_attribute("sampler") == "Random Walk Metropolis Hastings"
        assert s.read_attribute("stepsize") == "ndarray"
        per = int(s.read_attribute("samples_per_chain"))
        x = np.stack([s.chain(c)[:-1, per // 3:] for c in range(C)])
        rate = s.read_attribute("acceptance_rate")
    assert 0.15 < rate < 0.7
    assert np.all(np.abs(x.var(axis=(0, 2)) / var[:, 0] - 1) < 0.15)
    tuned = RWMH(seed=2).sample(str(tmp_path / "rw2.npy"), post, stepsize=0.01, proposals=400, chains=64,
                                autotuning=True)
    assert tuned.stepsize.shape == (64,) and np.median(tuned.stepsize) > 0.05
    with pytest.raises(AssertionError, match="wrong shape"):
        RWMH().sample(str(tmp_path / "rw3.npy"), post, stepsize=np.ones((3, 1)), proposals=4)


def test_diagnostic_mode_reports_block_shares(tmp_path, capsys):
    from hmclab_b200 import Distributions as D
    from hmclab_b200.Samplers import HMC

    HMC(seed=1).sample(str(tmp_path / "diag.npy"), D.Normal(np.zeros((3, 1)), 1.0), proposals=20,
                       chains=4, diagnostic_mode=True)
    out = capsys.readouterr().out
    assert "Detailed statistics" in out and "device blocks" in out and "fused_priors" in out


def test_replica_exchange_between_different_posteriors_matches_the_reference(tmp_path):
    """ParallelSampleSMP with two different posteriors (a cold and a hot Normal, two chains each):
    the reference's own multi-process run (tests/golden/exchange_runs.npz, written by
    tests/golden/make_golden_exchange.py from the unmodified hmclab) is reproduced file by file --
    HMC proposals, exchange decisions (Samplers.py:589-669), swapped models, and the misfit column
    that still shows the pre-exchange value in the row written right after a swap."""
    from hmclab_b200 import Distributions as D
    from hmclab_b200.Samplers import HMC, ParallelSampleSMP
    from hmclab_b200.Samples import Samples

    gold = np.load(os.path.join(GOLDEN_DIR, "exchange_runs.npz"))
    st = {k[len("setting_"):]: gold[k] for k in gold.files if k.startswith("setting_")}
    n, d = int(st["chains"]), int(st["dims"])
    mean, var = gold["mean"][:, None], gold["var"][:, None]
    cold, hot = D.Normal(mean, var), D.Normal(mean + 0.3, 3.0 * var)
    posts = [cold, hot, cold, hot]
    names = [str(tmp_path / f"chain{i}.npy") for i in range(n)]
    samplers = [HMC(seed=int(s)) for s in st["sampler_seeds"]]
    smp = ParallelSampleSMP(seed=int(st["smp_seed"]))
    smp.sample(samplers, names, posts, overwrite_existing_files=True, proposals=int(st["proposals"]),
               exchange=True, exchange_interval=int(st["exchange_interval"]),
               initial_model=[q[:, None] for q in gold["q0"]],
               kwargs=dict(stepsize=float(st["stepsize"]), amount_of_steps=int(st["amount_of_steps"]),
                           integrator=str(st["integrator"]), randomize_stepsize=bool(st["randomize_stepsize"]),
                           online_thinning=int(st["online_thinning"]), disable_progressbar=True, host_rng=True))
    assert np.array_equal(smp.exchange_schedule, gold["schedule"])
    assert smp.exchanges_accepted > 0
    moved = 0
    for i in range(n):
        ref = gold[f"samples{i}"]
        got = np.load(names[i])
        assert got.shape == ref.shape
        assert rel_err(got, ref) < TOL
        # rows where the model changed although the misfit column did not: accepted exchanges
        same_x = np.diff(ref[:, -1]) == 0
        changed = np.any(np.diff(ref[:, :-1], axis=0) != 0, axis=1)
        moved += int(np.sum(same_x & changed))
        with Samples(names[i]) as s:
            assert s.numpy.shape == (d + 1, ref.shape[0])
    assert moved > 0      # the golden run does contain accepted exchanges
    # device random streams: the same call runs without host draws, chains keep their own posterior
    smp2 = ParallelSampleSMP(seed=1)
    smp2.sample([HMC(seed=i) for i in range(n)], names, posts, overwrite_existing_files=True, proposals=40,
                exchange=True, exchange_interval=4, initial_model=[q[:, None] for q in gold["q0"]],
                kwargs=dict(stepsize=0.3, amount_of_steps=4, online_thinning=2))
    for i in range(n):
        assert np.load(names[i]).shape == (20, d + 1) and np.all(np.isfinite(np.load(names[i])))


def test_sparse_covariance_likelihood_through_the_public_dispatcher(tmp_path):
    """LinearMatrix(sparse G, d, N x N covariance): the reference's dispatcher hands this to the
    sparse-covariance class with the covariance converted to float32 and solves in single precision
    (LinearMatrix.py:458-459); the engine computes in float64 on the float32-rounded values, so misfit and
    gradient agree with the oracle's restatement of that solve to single-precision level, and a sampling run
    through HMC.sample works end to end."""
    import scipy.sparse as sp

    from hmclab_b200 import Distributions as D
    from hmclab_b200._lowering import describe
    from hmclab_b200.Samplers import HMC
    from oracle import hmc_oracle as oracle

    rng = np.random.default_rng(3)
    N, d = 50, 30
    G = sp.csr_matrix(np.where(rng.uniform(size=(N, d)) < 0.15, rng.uniform(0.2, 1.2, size=(N, d)), 0.0))
    band = rng.uniform(-0.1, 0.1, size=N - 1)
    cov = np.diag(rng.uniform(0.8, 1.2, size=N)) + np.diag(band, 1) + np.diag(band, -1)
    lik = D.LinearMatrix(G, rng.normal(size=(N, 1)) + 1.0, cov)
    assert type(lik.Distribution).__name__ == "_LinearMatrix_sparse_forward_sparse_covariance"
    post = D.BayesRule([D.Normal(np.zeros((d, 1)), 3.0), lik])
    tree = describe(post)
    for _ in range(3):
        m = rng.normal(size=(d, 1))
        g_ref, x_ref = oracle.gradient(tree, m), oracle.misfit(tree, m)
        assert rel_err(post.gradient(m), g_ref) < 2e-5 and abs(post.misfit(m) - x_ref) < 2e-5 * abs(x_ref)
    name = str(tmp_path / "spcov.npy")
    smp = HMC(seed=2).sample(name, post, stepsize=0.05, amount_of_steps=8, proposals=60, chains=16,
                             overwrite_existing_file=True, disable_progressbar=True)
    rows = np.load(name)
    assert rows.shape == (16 * 60, d + 1) and np.all(np.isfinite(rows))
    assert 0.3 < smp.accepted_proposals_per_chain.mean() / 60 <= 1.0
