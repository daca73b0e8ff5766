"""The shared-memory staged CSR SpMM (hmclab_b200/csrc/spmm_strip.cuh) on awkward matrices:
empty rows, a full row, a full column, a chain count that is no multiple of the slab width,
every thread mapping the library is built with, both nonzero encodings (8-byte when the values
are exact in fp32, else 16-byte), strip limits small enough to force the builder's fallback
(more, narrower strips), and raw C-ABI input with unsorted rows and split
duplicate entries.  Checked against the numpy oracle evaluated chain by chain
(misfit/gradient contract of LinearMatrix.py:389-426)."""
import copy

import numpy as np
import pytest
import scipy.sparse as sp

from helpers import rel_err
from hmclab_b200 import Distributions as D
from hmclab_b200 import MassMatrices as M
from hmclab_b200._lowering import describe, describe_mass, flatten
from oracle import hmc_oracle as oracle

pytestmark = pytest.mark.gpu

TOL = 1e-11


def _awkward(premult, seed=11):
    rng = np.random.default_rng(seed)
    N, d = (700, 333) if not premult else (150, 90)
    G = sp.random(N, d, density=0.05, random_state=np.random.RandomState(seed), format="lil")
    G[5, :] = 0.0                       # empty row
    G[100:121, :] = 0.0                 # a run of empty rows
    G[7, :] = rng.normal(size=d)        # full row
    G[:, 11] = rng.normal(size=(N, 1))  # full column
    G = sp.csr_matrix(G)
    G.eliminate_zeros()
    dvec = rng.normal(size=(N, 1))
    var = rng.uniform(0.5, 1.5, size=(N, 1))
    lik = D.LinearMatrix(G, dvec, var, premultiplication=premult)
    prior = D.Normal(np.zeros((d, 1)), rng.uniform(0.5, 2.0, size=(d, 1)))
    return D.BayesRule([prior, lik]), d


def _check(plan, tree, mtree, d, chains=130, seed=3):
    import torch

    from hmclab_b200._engine import Engine

    rng = np.random.default_rng(seed)
    q = rng.normal(size=(chains, d))
    eng = Engine(plan, mtree, chains, integrator="lf", amount_of_steps=3)
    assert eng.path == "staged"
    qd = torch.as_tensor(q).cuda().contiguous()
    g = eng.gradient(qd).cpu().numpy()
    x = eng.misfit(qd).cpu().numpy()
    sel = [0, 1, 63, 64, 65, chains - 1]
    g_ref = np.stack([oracle.gradient(tree, q[c][:, None])[:, 0] for c in sel])
    x_ref = np.array([oracle.misfit(tree, q[c][:, None]) for c in sel])
    assert rel_err(g[sel], g_ref) < TOL
    assert rel_err(x[sel], x_ref) < TOL
    eng.close()
    return g, x


@pytest.mark.parametrize("premult", [False, True])
@pytest.mark.parametrize("env", [{}, {"HMCB_SPMM_SHAPE": "1"}, {"HMCB_SPMM_SHAPE": "2"},
                                 {"HMCB_SPMM_SHAPE": "3"}, {"HMCB_SPMM_SHAPE": "4"},
                                 {"HMCB_SPMM_COMPACT": "0"},
                                 {"HMCB_SPMM_KB": "7", "HMCB_SPMM_EMAX": "1", "HMCB_SPMM_STAGES": "4"},
                                 {"HMCB_SPMM_SHAPE": "-1"},
                                 # the row-blocked kernel, forced (these matrices have no row clusters)
                                 {"HMCB_SPMM_BLOCKED": "1"},
                                 {"HMCB_SPMM_BLOCKED": "1", "HMCB_SPMM_COMPACT": "0"},
                                 {"HMCB_SPMM_BLOCKED": "1", "HMCB_SPMM_BLOCK_GW": "2", "HMCB_SPMM_BLOCK_NB": "2",
                                  "HMCB_SPMM_STAGES": "3"},
                                 {"HMCB_SPMM_BLOCKED": "1", "HMCB_SPMM_BLOCK_WARPS": "15", "HMCB_SPMM_BLOCK_GW": "8",
                                  "HMCB_SPMM_STAGES": "4"},
                                 {"HMCB_SPMM_BLOCKED": "1", "HMCB_SPMM_BLOCK_WARPS": "3", "HMCB_SPMM_BLOCK_GW": "2",
                                  "HMCB_SPMM_STAGES": "4"}])
def test_strip_spmm_on_awkward_matrices(monkeypatch, premult, env):
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    post, d = _awkward(premult)
    tree, mtree = describe(post), describe_mass(M.Unit(d))
    _check(flatten(tree), tree, mtree, d)


def test_strip_spmm_accepts_unsorted_rows_and_duplicates():
    """The C ABI takes any valid CSR: rows need not be sorted and an entry may be split in two."""
    post, d = _awkward(False)
    tree, mtree = describe(post), describe_mass(M.Unit(d))
    plan = flatten(tree)
    g0, x0 = _check(plan, tree, mtree, d)

    def scramble(indptr, indices, data, rng, n_split=50):   # the ABI takes one nnz for G and G^T
        ip, ix, dv = [0], [], []
        for i in range(len(indptr) - 1):
            cols, vals = list(indices[indptr[i]:indptr[i + 1]]), list(data[indptr[i]:indptr[i + 1]])
            if cols and n_split > 0:   # split the first entry of the row into two halves (exact in binary)
                cols.append(cols[0]); vals.append(vals[0] / 2); vals[0] = vals[0] / 2
                n_split -= 1
            perm = rng.permutation(len(cols))
            ix += [cols[p] for p in perm]; dv += [vals[p] for p in perm]
            ip.append(len(ix))
        return (np.asarray(ip, dtype=np.int32), np.asarray(ix, dtype=np.int32), np.asarray(dv, dtype=np.float64))

    rng = np.random.default_rng(5)
    plan2 = copy.copy(plan)
    lik = dict(plan["likelihood"])
    lik["indptr"], lik["indices"], lik["data"] = scramble(lik["indptr"], lik["indices"], lik["data"], rng)
    lik["t_indptr"], lik["t_indices"], lik["t_data"] = scramble(lik["t_indptr"], lik["t_indices"], lik["t_data"], rng)
    plan2["likelihood"] = lik
    g1, x1 = _check(plan2, tree, mtree, d)
    assert rel_err(g1, g0) < 1e-13 and rel_err(x1, x0) < 1e-13


def test_strip_spmm_is_deterministic_and_equals_the_gather_kernel(monkeypatch):
    """Many blocks per SM and several ring turns per block: repeated evaluations are bit-identical
    (the mbarrier ring neither reads a stage early nor overwrites one late), and the independent
    L2-gather kernel gives the same numbers up to summation order."""
    import torch

    from hmclab_b200 import workloads
    from hmclab_b200._engine import Engine

    w = workloads.tomography(nx=60, ny=50, rays=9000, chains=1500)
    tree, mtree = describe(w.posterior), describe_mass(w.mass_matrix)
    plan = flatten(tree)
    q = torch.as_tensor(w.initial_models).cuda().contiguous()
    eng = Engine(plan, mtree, w.chains, integrator="lf", amount_of_steps=2)
    g0, x0 = eng.gradient(q).clone(), eng.misfit(q).clone()
    for _ in range(3):
        assert torch.equal(eng.gradient(q), g0) and torch.equal(eng.misfit(q), x0)
    eng.close()
    # straight rays cluster: the default path is the row-blocked kernel; the plain strip kernel and
    # the independent L2-gather kernel give the same numbers up to summation order
    for env in ({"HMCB_SPMM_BLOCKED": "0"}, {"HMCB_SPMM_SHAPE": "-1"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        ref = Engine(plan, mtree, w.chains, integrator="lf", amount_of_steps=2)
        assert rel_err(g0.cpu().numpy(), ref.gradient(q).cpu().numpy()) < 1e-12
        assert rel_err(x0.cpu().numpy(), ref.misfit(q).cpu().numpy()) < 1e-12
        ref.close()
