"""Parity at larger sizes: (i) CUDA vs the CPU oracle on seeded inputs at sizes the oracle
finishes in seconds, for every workload family of BASELINE.json; (ii) CUDA vs the CPU oracle at
BASELINE.json's FULL sizes for configs 2-5 (every chain is advanced on the GPU, the oracle
follows the first, the middle and the last chain: the full-size grids -- 25 856 SpMM blocks,
202 row chunks, 128 chain slabs -- are where indexing bugs would hide); (iii) size-independent
properties at full size: the fused register-resident kernel and the independent staged
(transposed, multi-launch) pipeline produce the same chains from the same device random
streams, trajectories are time reversible, the strip SpMM equals the L2-gather kernel."""
import os

import numpy as np
import pytest

from helpers import rel_err
from hmclab_b200 import workloads
from hmclab_b200._lowering import describe, describe_mass, flatten
from oracle import hmc_oracle as oracle

pytestmark = pytest.mark.gpu

TOL = 1e-10


def _compare_with_oracle(w, K=2, chains_checked=3, randomize=True, seed=0):
    import torch

    from hmclab_b200._engine import Engine

    C, d = w.chains, w.dims
    rng = np.random.default_rng(seed)
    z = rng.normal(size=(K, C, d))
    us, ua = rng.uniform(0.5, 1.5, size=(K, C)), rng.uniform(size=(K, C))
    tree, mtree = describe(w.posterior), describe_mass(w.mass_matrix)
    eng = Engine(flatten(tree), mtree, C, integrator=w.integrator, amount_of_steps=w.amount_of_steps)
    q = torch.as_tensor(w.initial_models).cuda().contiguous()
    x = eng.misfit(q)
    out = dict(out_samples=torch.zeros(K, C, d + 1, dtype=torch.float64, device="cuda"),
               out_accept=torch.zeros(K, C, dtype=torch.uint8, device="cuda"),
               out_h0=torch.zeros(K, C, dtype=torch.float64, device="cuda"),
               out_h1=torch.zeros(K, C, dtype=torch.float64, device="cuda"),
               out_q_prop=torch.zeros(K, C, d, dtype=torch.float64, device="cuda"),
               out_p_prop=torch.zeros(K, C, d, dtype=torch.float64, device="cuda"))
    eng.run_block(q, x, K, stepsize=w.stepsize, randomize_stepsize=randomize,
                  z=torch.as_tensor(z).cuda(), u_step=torch.as_tensor(us).cuda(),
                  u_accept=torch.as_tensor(ua).cuda(), **out)
    got = {k: v.cpu().numpy() for k, v in out.items()}
    sel = np.linspace(0, C - 1, chains_checked).astype(int)   # first, middle, last chain
    with np.errstate(all="ignore"):
        ref = oracle.run_chains(tree, mtree, q0=w.initial_models[sel], z=z[:, sel], u_step=us[:, sel],
                                u_acc=ua[:, sel], integrator=w.integrator, steps=w.amount_of_steps,
                                stepsize=w.stepsize, randomize=randomize)
    assert np.array_equal(got["out_accept"][:, sel].astype(bool), ref["accept"])
    assert rel_err(got["out_q_prop"][:, sel], ref["q_prop"]) < TOL
    assert rel_err(got["out_p_prop"][:, sel], ref["p_prop"]) < TOL
    assert rel_err(got["out_h0"][:, sel], ref["H0"]) < TOL
    assert rel_err(got["out_h1"][:, sel], ref["H1"]) < TOL
    assert rel_err(got["out_samples"][:, sel], ref["samples"]) < TOL
    return eng, ref


def test_config2_shape_vs_oracle():
    eng, ref = _compare_with_oracle(workloads.normal_iid(dims=1000, chains=300), K=3)
    assert eng.path == "fused_priors" and 0 < ref["accept"].mean() <= 1


def test_config3_direct_vs_oracle():
    w = workloads.dense_large(dims=333, data=700, chains=200)
    eng, _ = _compare_with_oracle(w, K=2)
    assert eng.path == "staged" and w.extra["form"] == "direct"


def test_config3_premultiplied_vs_oracle():
    eng, _ = _compare_with_oracle(workloads.dense_large(dims=333, data=700, chains=130,
                                                        premultiplication=True), K=2)
    assert eng.path == "staged"


def test_config4_tomography_vs_oracle():
    w = workloads.tomography(nx=30, ny=25, rays=1500, chains=150)
    eng, _ = _compare_with_oracle(w, K=2)
    assert eng.path == "staged"


def test_config5_source_location_vs_oracle():
    w = workloads.source_location(events=16, stations=30, chains=600)
    eng, _ = _compare_with_oracle(w, K=3)
    assert eng.path == "fused_srcloc"


@pytest.mark.parametrize("name,K", [("normal_iid", 2), ("dense_large", 1), ("dense_large_premult", 1),
                                    ("tomography", 1), ("source_location", 2)])
def test_full_size_config_vs_oracle(name, K):
    """BASELINE.json configs[1..4] at their full sizes against the numpy restatement
    (Samplers.py:1463-1492): trajectories within 1e-10, accept/reject bits identical."""
    w = workloads.BUILDERS[name]()
    assert w.chains >= 4096
    eng, ref = _compare_with_oracle(w, K=K)
    assert eng.path == {"normal_iid": "fused_priors", "source_location": "fused_srcloc"}.get(name, "staged")
    assert np.isfinite(ref["H1"]).all()


def test_config1_shape_vs_oracle():
    eng, _ = _compare_with_oracle(workloads.dense_small(chains=9), K=4, randomize=False)
    assert eng.path == "fused_dense"
    eng, _ = _compare_with_oracle(workloads.dense_small(chains=1000), K=3, chains_checked=5)
    assert eng.path == "fused_dense"


@pytest.mark.parametrize("prior", ["laplace", "uniform_box", "bounded_normal"])
def test_fused_dense_kernel_equals_staged_pipeline(prior):
    """Small premultiplied dense models: the whole-proposal tensor-core kernel and the staged
    multi-launch pipeline are independent implementations; same device random streams, same
    chains (also with thinning, two blocks, a 4-stage integrator and a diagonal mass)."""
    import torch

    from hmclab_b200 import Distributions as D
    from hmclab_b200 import MassMatrices as M
    from hmclab_b200._engine import Engine

    rng = np.random.default_rng(8)
    dims, data, C = 77, 150, 333
    G = rng.normal(size=(data, dims)) / np.sqrt(data)
    dvec = G @ rng.normal(size=(dims, 1)) + 0.3 * rng.normal(size=(data, 1))
    var = rng.uniform(0.5, 1.5, size=(data, 1))
    if prior == "laplace":
        pr = D.Laplace(np.zeros((dims, 1)), np.full((dims, 1), 2.0))
    elif prior == "uniform_box":      # reflection on the box + bound checks visible to the gradient
        pr = D.Uniform(np.full((dims, 1), -1.5), np.full((dims, 1), 1.5))
    else:
        pr = D.Normal(np.zeros((dims, 1)), 1.0, lower_bounds=np.full((dims, 1), -2.0),
                      upper_bounds=np.full((dims, 1), 2.5))
    post = D.BayesRule([pr, D.LinearMatrix(G, dvec, var)])
    mass = M.Diagonal(rng.uniform(0.5, 2.0, size=(dims, 1)))
    plan, mplan = flatten(describe(post)), describe_mass(mass)
    q0 = np.clip(rng.normal(size=(C, dims)), -1.4, 1.4)
    results = {}
    for label, force in (("fused", None), ("staged", "1")):
        if force:
            os.environ["HMCB_FORCE_STAGED"] = force
        try:
            eng = Engine(plan, mplan, C, integrator="4s", amount_of_steps=3)
        finally:
            os.environ.pop("HMCB_FORCE_STAGED", None)
        assert eng.path == ("staged" if force else "fused_dense")
        q = torch.as_tensor(q0).cuda().contiguous()
        x = eng.misfit(q)
        acc = torch.zeros(C, dtype=torch.int32, device="cuda")
        rows = []
        for lo, hi in ((0, 5), (5, 12)):
            n = eng.stored_rows(hi - lo, 3, lo)
            buf = torch.zeros(n, C, dims + 1, dtype=torch.float64, device="cuda")
            eng.run_block(q, x, hi - lo, stepsize=0.25, seed=31, thinning=3, proposal_offset=lo,
                          out_samples=buf, accepted_total=acc)
            rows.append(buf.cpu().numpy())
        results[label] = (np.concatenate(rows), q.cpu().numpy(), x.cpu().numpy(), acc.cpu().numpy())
    f, s = results["fused"], results["staged"]
    assert np.array_equal(f[3], s[3]) and 0.05 < f[3].mean() / 12 < 0.999
    assert rel_err(f[0], s[0]) < 1e-11 and rel_err(f[1], s[1]) < 1e-11 and rel_err(f[2], s[2]) < 1e-11


def test_large_dims_priors_only_uses_staged_path_and_matches_oracle():
    w = workloads.normal_iid(dims=5000, chains=40)
    eng, _ = _compare_with_oracle(w, K=2)
    assert eng.path == "staged"


def test_full_size_config2_fused_equals_staged_pipeline():
    """4096 chains x 1000 dims, device RNG: two independent implementations, same chains (in exact
    arithmetic bit for bit; the fused kernel's default FMA arithmetic within a few ulp of them)."""
    import torch

    from hmclab_b200._engine import Engine

    w = workloads.normal_iid()   # BASELINE.json configs[1]
    assert (w.chains, w.dims) == (4096, 1000)
    plan, mplan = flatten(describe(w.posterior)), describe_mass(w.mass_matrix)
    results = {}
    for label, force, exact in (("fused", None, True), ("staged", "1", True), ("fused_fma", None, False)):
        if force:
            os.environ["HMCB_FORCE_STAGED"] = force
        try:
            eng = Engine(plan, mplan, w.chains, integrator="lf", amount_of_steps=10, exact=exact)
        finally:
            os.environ.pop("HMCB_FORCE_STAGED", None)
        assert eng.path == ("staged" if force else "fused_priors")
        q = torch.as_tensor(w.initial_models).cuda().contiguous()
        x = eng.misfit(q)
        acc = torch.zeros(3, w.chains, dtype=torch.uint8, device="cuda")
        h0 = torch.zeros(3, w.chains, dtype=torch.float64, device="cuda")
        h1 = torch.zeros(3, w.chains, dtype=torch.float64, device="cuda")
        eng.run_block(q, x, 3, stepsize=0.15, seed=77, out_accept=acc, out_h0=h0, out_h1=h1)
        results[label] = (q.cpu().numpy(), x.cpu().numpy(), acc.cpu().numpy(), h0.cpu().numpy(),
                          h1.cpu().numpy())
    f, s = results["fused"], results["staged"]
    assert np.array_equal(f[2], s[2])                 # identical accept/reject decisions
    assert np.array_equal(f[0], s[0])                 # separable target: bit-identical positions
    assert rel_err(f[1], s[1]) < 1e-13 and rel_err(f[3], s[3]) < 1e-13 and rel_err(f[4], s[4]) < 1e-13
    assert 0.05 < f[2].mean() < 0.999                 # both decisions occur at this step size
    # energy error of a leapfrog trajectory is O(eps^2): a coarse sanity bound on H1 - H0
    assert np.median(np.abs(f[4] - f[3])) < 5.0
    # the default arithmetic of the fused kernel (one FMA per update instead of multiply + add): the
    # same decisions, positions and energies equal to a few ulp
    m = results["fused_fma"]
    assert np.array_equal(m[2], f[2])
    assert rel_err(m[0], f[0]) < 1e-12 and rel_err(m[3], f[3]) < 1e-12 and rel_err(m[4], f[4]) < 1e-12
    assert not np.array_equal(m[0], f[0])


@pytest.mark.parametrize("name", ["normal_iid", "dense_large", "dense_large_premult", "tomography",
                                  "source_location"])
def test_full_size_trajectories_are_time_reversible(name):
    """At BASELINE.json's full sizes, for EVERY chain (the oracle test above follows three):
    integrating forward, flipping
    the momentum and integrating again must return every chain to its start -- a property of
    the symmetric lf/3s/4s schemes that any error in the gradient, the mass matrix, the
    stage coefficients or the chain indexing of a kernel would break."""
    import torch

    from hmclab_b200._engine import Engine

    w = workloads.BUILDERS[name]()
    C, d = w.chains, w.dims
    mtree = describe_mass(w.mass_matrix)
    eng = Engine(flatten(describe(w.posterior)), mtree, C, integrator=w.integrator,
                 amount_of_steps=w.amount_of_steps)
    sqrtm = (torch.as_tensor(np.sqrt(mtree["diagonal"])).cuda() if mtree["kind"] == "diagonal"
             else torch.ones(d, dtype=torch.float64, device="cuda"))
    gen = torch.Generator(device="cuda").manual_seed(5)
    q0 = torch.as_tensor(w.initial_models).cuda().contiguous()
    z = torch.randn(1, C, d, dtype=torch.float64, device="cuda", generator=gen)
    us = torch.ones(1, C, dtype=torch.float64, device="cuda")
    ua = torch.zeros(1, C, dtype=torch.float64, device="cuda")

    def trajectory(q_start, z_in):
        q = q_start.clone()
        x = eng.misfit(q)
        q1 = torch.empty(1, C, d, dtype=torch.float64, device="cuda")
        p1 = torch.empty(1, C, d, dtype=torch.float64, device="cuda")
        h0 = torch.empty(1, C, dtype=torch.float64, device="cuda")
        h1 = torch.empty(1, C, dtype=torch.float64, device="cuda")
        eng.run_block(q, x, 1, stepsize=w.stepsize, randomize_stepsize=False, z=z_in, u_step=us,
                      u_accept=ua, out_q_prop=q1, out_p_prop=p1, out_h0=h0, out_h1=h1)
        return q1[0], p1[0], h0[0], h1[0]

    q1, p1, h0, h1 = trajectory(q0, z)
    assert torch.isfinite(h1).all()
    moved = (q1 - q0).abs().max().item()
    assert moved > 0
    # energy is conserved up to the integrator's O(eps^2) error
    assert (h1 - h0).abs().median().item() < 10.0
    q2, p2, _, _ = trajectory(q1.contiguous(), (-(p1 / sqrtm)).unsqueeze(0).contiguous())
    scale = max(q0.abs().max().item(), 1e-300)
    assert (q2 - q0).abs().max().item() < 1e-9 * scale
    p0 = z[0] * sqrtm
    assert (p2 + p0).abs().max().item() < 1e-8 * p0.abs().max().item()


def test_full_size_config4_strip_spmm_equals_l2_gather_kernel(monkeypatch):
    """BASELINE config 4 at full size (10k cells, 50k rays, 8192 chains): the shared-memory staged
    SpMM (TMA ring, strip tables, compact nonzeros) and the independent L2-gather kernel give the
    same misfits and gradients up to summation order, and the staged one is bit-reproducible."""
    import torch

    from hmclab_b200._engine import Engine

    w = workloads.tomography()
    tree, mtree = describe(w.posterior), describe_mass(w.mass_matrix)
    plan = flatten(tree)
    q = torch.as_tensor(w.initial_models).cuda().contiguous()
    eng = Engine(plan, mtree, w.chains, integrator=w.integrator, amount_of_steps=w.amount_of_steps)
    g, x = eng.gradient(q), eng.misfit(q)
    assert torch.equal(eng.gradient(q), g) and torch.equal(eng.misfit(q), x)
    eng.close()
    # default = the row-blocked kernel (rays cluster); then the plain strip kernel, then L2 gathers
    for env in ({"HMCB_SPMM_BLOCKED": "0"}, {"HMCB_SPMM_SHAPE": "-1"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        ref = Engine(plan, mtree, w.chains, integrator=w.integrator, amount_of_steps=w.amount_of_steps)
        g_ref, x_ref = ref.gradient(q), ref.misfit(q)
        assert float((g - g_ref).abs().max() / g_ref.abs().max()) < 1e-12
        assert float(((x - x_ref).abs() / x_ref.abs()).max()) < 1e-12
        ref.close()


def test_single_chain_single_dimension():
    w = workloads.normal_iid(dims=1, chains=1)
    eng, ref = _compare_with_oracle(w, K=5, chains_checked=1)
    assert eng.path == "fused_priors"


def test_composite_of_many_per_parameter_priors_vs_oracle():
    """The reference's common pattern: one prior object per parameter (or small group) inside a
    CompositeDistribution -- 12 priors, 10 of them bounded (more objects than the 8 the first
    version of the engine accepted), top level, so the composite reflects on its children's bounds."""
    from hmclab_b200 import Distributions as D
    from hmclab_b200 import MassMatrices as M

    rng = np.random.default_rng(12)
    parts, dims = [], 0
    for k in range(12):
        n = 1 + k % 3
        if k % 3 == 0:
            parts.append(D.Normal(rng.normal(size=(n, 1)), rng.uniform(0.5, 2.0, size=(n, 1)),
                                  lower_bounds=np.full((n, 1), -3.0), upper_bounds=np.full((n, 1), 3.0)))
        elif k % 3 == 1:
            parts.append(D.Laplace(rng.normal(size=(n, 1)), rng.uniform(0.5, 2.0, size=(n, 1)),
                                   lower_bounds=None if k == 1 else np.full((n, 1), -4.0)))
        else:
            parts.append(D.Uniform(np.full((n, 1), -2.0), np.full((n, 1), 2.5)) if k != 2 else
                         D.Normal(np.zeros((n, 1)), 1.0))
        dims += n
    post = D.CompositeDistribution(parts)
    C = 64
    w = workloads.Workload("composite12", post, M.Diagonal(rng.uniform(0.5, 2.0, size=(dims, 1))), C, "3s", 4,
                           0.3, np.clip(rng.normal(size=(C, dims)), -1.5, 1.5), "12 per-parameter priors")
    eng, ref = _compare_with_oracle(w, K=6, chains_checked=8)
    assert eng.path == "fused_priors" and ref["accept"].mean() > 0
