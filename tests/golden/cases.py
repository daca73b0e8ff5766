"""Small HMC problems shared by the golden-vector generator and the parity tests.

``make_inputs(name)`` draws the raw constructor inputs (seeded); ``build(name, inputs,
ns)`` constructs the posterior and mass matrix from those raw inputs with the classes of
``ns`` -- either the genuine ``hmclab`` package (generator) or ``hmclab_b200`` (tests).
The raw inputs are stored inside every ``.npz`` so the tests never re-draw them.
"""
from __future__ import annotations

import numpy as np


def _settings(integrator, steps, stepsize, randomize, chains, proposals):
    return dict(integrator=integrator, steps=steps, stepsize=stepsize,
                randomize=randomize, chains=chains, proposals=proposals)


SETTINGS = {
    # config-1 shape of BASELINE.json: dense 200 data x 100 params, premultiplied GtG
    "dense_premult_cfg1": _settings("lf", 10, 0.004, False, 3, 5),
    # config-2 shape (scaled): separable Normal posterior, Unit mass
    "normal_unit_lf": _settings("lf", 10, 0.6, True, 6, 6),
    # direct (G, Gt) form, vector variance, Diagonal mass, 4-stage
    "dense_direct_4s": _settings("4s", 3, 2.6, True, 4, 4),
    # premultiplied with vector variance (float32 GtG quirk), 3-stage
    "dense_premult_vecvar_3s": _settings("3s", 2, 2.2, True, 4, 4),
    # inner class built with dtype=float64 (pure fp64 G), direct form
    "dense_direct_f64": _settings("lf", 4, 0.02, True, 3, 3),
    # CSR tomography-like operator, Laplace prior
    "sparse_laplace_lf": _settings("lf", 5, 0.01, True, 4, 5),
    # sparse operator through the premultiplied (sparse GtG) form
    "sparse_premult": _settings("lf", 3, 0.01, False, 3, 3),
    # source location, fixed velocity, Uniform box in BayesRule -> reflection active
    "srcloc_fixed_v": _settings("lf", 8, 0.025, True, 6, 8),
    # source location with the velocity as the last parameter, 3-stage
    "srcloc_infer_v": _settings("3s", 3, 0.045, True, 4, 5),
    # composite prior (no reflection inside BayesRule) + dense likelihood
    "composite_in_bayes": _settings("4s", 2, 0.6, True, 4, 5),
    # priors only: top-level composite, reflection through the children's bounds
    "composite_top": _settings("lf", 6, 0.4, True, 6, 8),
    # bounded Normal on its own: reflection and out-of-bounds rejection
    "normal_bounded": _settings("lf", 5, 0.9, True, 6, 8),
    # 2-D source location with inferred velocity, Uniform box in BayesRule, Diagonal mass
    "srcloc2d_infer_v": _settings("lf", 6, 0.03, True, 5, 6),
    # dense (N x N) data covariance, premultiplied form (LinearMatrix.py:226-305)
    "dense_fullcov_premult": _settings("3s", 2, 0.35, True, 4, 4),
    # dense (N x N) data covariance, direct form: float32 Gt @ invcov re-formed per call in the reference
    "dense_fullcov_direct": _settings("lf", 4, 0.3, True, 4, 4),
    # sparse G with a sparse (banded) data covariance: an LU solve per evaluation in the reference
    "sparse_fullcov_lf": _settings("lf", 4, 0.05, True, 4, 4),
    # Full (dense) mass matrix (MassMatrices.py:241-327), dense direct likelihood, 3-stage
    "dense_full_mass_3s": _settings("3s", 2, 0.9, True, 4, 4),
    # Full mass matrix on a bounded priors-only target: reflection flips momenta between sub-steps
    "bounded_full_mass_lf": _settings("lf", 5, 0.3, True, 6, 6),
}

CASES = tuple(SETTINGS)


def make_inputs(name: str) -> dict:
    rng = np.random.default_rng(sum(map(ord, name)))
    s = SETTINGS[name]
    inp = {}
    if name == "dense_premult_cfg1":
        inp.update(G=rng.normal(size=(200, 100)), d=rng.normal(size=(200, 1)))
        dims = 100
    elif name == "normal_unit_lf":
        dims = 257
    elif name == "dense_direct_4s":
        dims = 80
        inp.update(G=rng.normal(size=(50, dims)) / np.sqrt(50), d=rng.normal(size=(50, 1)),
                   var=rng.uniform(0.5, 1.5, size=(50, 1)),
                   mass=rng.uniform(0.5, 2.0, size=(dims, 1)))
    elif name == "dense_premult_vecvar_3s":
        dims = 40
        inp.update(G=rng.normal(size=(90, dims)) / np.sqrt(90), d=rng.normal(size=(90, 1)),
                   var=rng.uniform(0.5, 1.5, size=(90, 1)),
                   mass=rng.uniform(0.5, 2.0, size=(dims, 1)))
    elif name == "srcloc2d_infer_v":
        E, S = 3, 6
        dims = 3 * E + 1
        sx = rng.uniform(-10, 30, size=(1, S))
        sz = np.zeros((1, S))
        ex, ez, eT = rng.uniform(0, 20, size=(E, 1)), rng.uniform(1, 10, size=(E, 1)), rng.uniform(0, 10, size=(E, 1))
        tt = eT + ((ex - sx) ** 2 + (ez - sz) ** 2) ** 0.5 / 3.0
        lo = np.vstack([np.tile(np.array([[-5.0], [0.0], [-2.0]]), (E, 1)), [[1.5]]])
        hi = np.vstack([np.tile(np.array([[25.0], [12.0], [12.0]]), (E, 1)), [[5.0]]])
        inp.update(sx=sx, sz=sz, tobs=tt + 0.1 * rng.normal(size=tt.shape),
                   std=rng.uniform(0.08, 0.2, size=tt.shape), lo=lo, hi=hi,
                   mass=rng.uniform(0.5, 2.0, size=(dims, 1)),
                   truth=np.hstack([ex, ez, eT]).reshape(-1, 1))
    elif name == "dense_fullcov_direct":
        dims = 28
        N = 19
        A = rng.normal(size=(N, N)) / np.sqrt(N)
        inp.update(G=rng.normal(size=(N, dims)) / np.sqrt(N), d=rng.normal(size=(N, 1)),
                   cov=A @ A.T + 0.5 * np.eye(N))
    elif name == "dense_fullcov_premult":
        dims = 24
        N = 40
        A = rng.normal(size=(N, N)) / np.sqrt(N)
        inp.update(G=rng.normal(size=(N, dims)) / np.sqrt(N), d=rng.normal(size=(N, 1)),
                   cov=A @ A.T + 0.5 * np.eye(N), mass=rng.uniform(0.5, 2.0, size=(dims, 1)))
    elif name == "sparse_fullcov_lf":
        dims, N = 40, 60
        mask = rng.uniform(size=(N, dims)) < 0.12
        band = rng.uniform(-0.2, 0.2, size=N - 1)
        inp.update(G=np.where(mask, rng.uniform(0.1, 1.4, size=(N, dims)), 0.0),
                   d=rng.normal(size=(N, 1)) + 2.0,
                   cov=np.diag(rng.uniform(0.6, 1.4, size=N)) + np.diag(band, 1) + np.diag(band, -1))
    elif name == "dense_full_mass_3s":
        dims = 30
        A = rng.normal(size=(dims, dims)) / np.sqrt(dims)
        inp.update(G=rng.normal(size=(45, dims)) / np.sqrt(45), d=rng.normal(size=(45, 1)),
                   var=rng.uniform(0.5, 1.5, size=(45, 1)), mass_full=A @ A.T + np.eye(dims))
    elif name == "bounded_full_mass_lf":
        dims = 10
        A = rng.normal(size=(dims, dims)) / np.sqrt(dims)
        inp.update(mu=rng.normal(size=(dims, 1)) * 0.2, var=rng.uniform(0.5, 2, size=(dims, 1)),
                   lo=np.full((dims, 1), -1.0), hi=np.full((dims, 1), 1.0),
                   mass_full=0.5 * (A @ A.T) + np.eye(dims))
    elif name == "dense_direct_f64":
        dims = 33
        inp.update(G=rng.normal(size=(21, dims)), d=rng.normal(size=(21, 1)),
                   var=rng.uniform(0.5, 1.5, size=(21, 1)))
    elif name in ("sparse_laplace_lf", "sparse_premult"):
        dims = 60
        N = 120
        mask = rng.uniform(size=(N, dims)) < 0.1
        inp.update(G=np.where(mask, rng.uniform(0.1, 1.4, size=(N, dims)), 0.0),
                   d=rng.normal(size=(N, 1)) + 3.0,
                   mu=rng.normal(size=(dims, 1)),
                   b=rng.uniform(0.5, 2.0, size=(dims, 1)))
    elif name in ("srcloc_fixed_v", "srcloc_infer_v"):
        E, S = (4, 7) if name == "srcloc_fixed_v" else (3, 5)
        dims = 4 * E + (name == "srcloc_infer_v")
        sx = rng.uniform(-10, 30, size=(1, S))
        sy = rng.uniform(-10, 30, size=(1, S))
        sz = np.zeros((1, S))
        ex, ey = rng.uniform(0, 20, size=(E, 1)), rng.uniform(0, 20, size=(E, 1))
        ez, eT = rng.uniform(0, 10, size=(E, 1)), rng.uniform(0, 10, size=(E, 1))
        v = 3.0
        tt = eT + ((ex - sx) ** 2 + (ey - sy) ** 2 + (ez - sz) ** 2) ** 0.5 / v
        tobs = tt + 0.1 * rng.normal(size=tt.shape)
        tobs[1, 2] = np.nan  # a missing pick
        std = rng.uniform(0.08, 0.2, size=tt.shape)
        lo = np.tile(np.array([[-5.0], [-5.0], [0.0], [-2.0]]), (E, 1))
        hi = np.tile(np.array([[25.0], [25.0], [12.0], [12.0]]), (E, 1))
        if name == "srcloc_infer_v":
            lo, hi = np.vstack([lo, [[1.5]]]), np.vstack([hi, [[5.0]]])
        inp.update(sx=sx, sy=sy, sz=sz, tobs=tobs, std=std, lo=lo, hi=hi,
                   mass=rng.uniform(0.5, 2.0, size=(dims, 1)),
                   truth=np.hstack([ex, ey, ez, eT]).reshape(-1, 1))
    elif name == "composite_in_bayes":
        dims = 9
        inp.update(G=rng.normal(size=(14, dims)), d=rng.normal(size=(14, 1)),
                   mu_n=rng.normal(size=(3, 1)), var_n=rng.uniform(0.5, 2, size=(3, 1)),
                   mu_l=rng.normal(size=(4, 1)), b_l=rng.uniform(0.5, 2, size=(4, 1)),
                   lo_u=np.full((2, 1), -1.5), hi_u=np.full((2, 1), 1.5))
    elif name == "composite_top":
        dims = 12
        inp.update(mu_n=rng.normal(size=(5, 1)), var_n=rng.uniform(0.5, 2, size=(5, 1)),
                   lo_n=np.full((5, 1), -1.0), hi_n=np.full((5, 1), 1.2),
                   lo_u=np.full((3, 1), -0.7), hi_u=np.full((3, 1), 0.9),
                   mu_l=rng.normal(size=(4, 1)) * 0.3, b_l=rng.uniform(0.5, 2, size=(4, 1)))
    elif name == "normal_bounded":
        dims = 10
        inp.update(mu=rng.normal(size=(dims, 1)) * 0.2, var=rng.uniform(0.5, 2, size=(dims, 1)),
                   lo=np.full((dims, 1), -1.0), hi=np.full((dims, 1), 1.0))
    else:
        raise KeyError(name)

    C, K = s["chains"], s["proposals"]
    inp["dims"] = np.int64(dims)
    if name.startswith("srcloc"):
        base = inp["truth"][:, 0]
        if name in ("srcloc_infer_v", "srcloc2d_infer_v"):
            base = np.concatenate([base, [3.0]])
        q0 = base[None, :] + 0.3 * rng.normal(size=(C, dims))
        q0 = np.clip(q0, inp["lo"][:, 0] + 1e-3, inp["hi"][:, 0] - 1e-3)
    elif name in ("composite_top", "normal_bounded", "bounded_full_mass_lf"):
        q0 = rng.uniform(-0.5, 0.5, size=(C, dims))
    elif name == "composite_in_bayes":
        q0 = rng.uniform(-0.5, 0.5, size=(C, dims))
    else:
        q0 = rng.normal(size=(C, dims))
    inp.update(q0=q0, z=rng.normal(size=(K, C, dims)),
               u_step=rng.uniform(0.5, 1.5, size=(K, C)),
               u_acc=rng.uniform(0.0, 1.0, size=(K, C)),
               probe=q0[:2] + 0.1 * rng.normal(size=(2, dims)))
    return inp


def build(name: str, inp: dict, ns):
    """-> (posterior, mass_matrix) built with the classes of package ``ns``."""
    D, M = ns.Distributions, ns.MassMatrices
    dims = int(inp["dims"])
    cp = lambda k: np.array(inp[k], copy=True)  # constructors may reshape in place
    mass = M.Unit(dims)
    if name == "dense_premult_cfg1":
        post = D.BayesRule([D.Normal(np.zeros((dims, 1)), 1.0),
                            D.LinearMatrix(cp("G"), cp("d"), 2.0)])
    elif name == "normal_unit_lf":
        post = D.Normal(np.zeros((dims, 1)), 1.0)
    elif name == "dense_direct_4s":
        post = D.BayesRule([D.Normal(np.zeros((dims, 1)), 1.0),
                            D.LinearMatrix(cp("G"), cp("d"), cp("var"))])
        mass = M.Diagonal(cp("mass"))
    elif name == "dense_premult_vecvar_3s":
        post = D.BayesRule([D.Normal(np.zeros((dims, 1)), 4.0),
                            D.LinearMatrix(cp("G"), cp("d"), cp("var"))])
        mass = M.Diagonal(cp("mass"))
    elif name == "srcloc2d_infer_v":
        lik = D.SourceLocation2D(cp("sx"), cp("sz"), cp("tobs"), cp("std"), infer_velocity=True)
        post = D.BayesRule([D.Uniform(cp("lo"), cp("hi")), lik])
        mass = M.Diagonal(cp("mass"))
    elif name == "dense_fullcov_premult":
        post = D.BayesRule([D.Normal(np.zeros((dims, 1)), 2.0),
                            D.LinearMatrix(cp("G"), cp("d"), cp("cov"))])
        mass = M.Diagonal(cp("mass"))
    elif name == "dense_fullcov_direct":
        post = D.BayesRule([D.Normal(np.zeros((dims, 1)), 2.0),
                            D.LinearMatrix(cp("G"), cp("d"), cp("cov"))])      # N < dims: direct form
    elif name == "sparse_fullcov_lf":
        import importlib
        import scipy.sparse as sp
        mod = importlib.import_module(D.LinearMatrix.__module__)
        # inner class directly: the float64 sparse covariance is kept as given (through the public
        # dispatcher it would be converted to float32 and solved in single precision)
        lik = mod._LinearMatrix_sparse_forward_sparse_covariance(
            sp.csr_matrix(cp("G")), cp("d"), sp.csr_matrix(cp("cov")))
        post = D.BayesRule([D.Normal(np.zeros((dims, 1)), 2.0), lik])
    elif name == "dense_full_mass_3s":
        post = D.BayesRule([D.Normal(np.zeros((dims, 1)), 1.0),
                            D.LinearMatrix(cp("G"), cp("d"), cp("var"), premultiplication=False)])
        mass = M.Full(cp("mass_full"))
    elif name == "bounded_full_mass_lf":
        post = D.Normal(cp("mu"), cp("var"), lower_bounds=cp("lo"), upper_bounds=cp("hi"))
        mass = M.Full(cp("mass_full"))
    elif name == "dense_direct_f64":
        inner = D.LinearMatrix.__module__
        import importlib
        mod = importlib.import_module(inner)
        lik = mod._LinearMatrix_dense_forward_simple_covariance(
            cp("G"), cp("d"), cp("var"), dtype=np.float64)
        post = D.BayesRule([D.Normal(np.zeros((dims, 1)), 1.0), lik])
    elif name in ("sparse_laplace_lf", "sparse_premult"):
        import scipy.sparse as sp
        G = sp.csr_matrix(cp("G"))
        lik = D.LinearMatrix(G, cp("d"), 0.25,
                             premultiplication=(name == "sparse_premult"))
        post = D.BayesRule([D.Laplace(cp("mu"), cp("b")), lik])
    elif name in ("srcloc_fixed_v", "srcloc_infer_v"):
        infer = name == "srcloc_infer_v"
        lik = D.SourceLocation3D(cp("sx"), cp("sy"), cp("sz"), cp("tobs"), cp("std"),
                                 infer_velocity=infer,
                                 medium_velocity=None if infer else 3.0)
        post = D.BayesRule([D.Uniform(cp("lo"), cp("hi")), lik])
        mass = M.Diagonal(cp("mass"))
    elif name == "composite_in_bayes":
        prior = D.CompositeDistribution([
            D.Normal(cp("mu_n"), cp("var_n")),
            D.Laplace(cp("mu_l"), cp("b_l")),
            D.Uniform(cp("lo_u"), cp("hi_u")),
        ])
        post = D.BayesRule([prior, D.LinearMatrix(cp("G"), cp("d"), 1.5)])
    elif name == "composite_top":
        post = D.CompositeDistribution([
            D.Normal(cp("mu_n"), cp("var_n"), lower_bounds=cp("lo_n"), upper_bounds=cp("hi_n")),
            D.Uniform(cp("lo_u"), cp("hi_u")),
            D.Laplace(cp("mu_l"), cp("b_l")),
        ])
    elif name == "normal_bounded":
        post = D.Normal(cp("mu"), cp("var"), lower_bounds=cp("lo"), upper_bounds=cp("hi"))
    else:
        raise KeyError(name)
    return post, mass
