"""Host-side contracts the reference's own tests pin (no device needed):
bounds validation texts (reference tests/test_distributions.py:147-166), momentum shape
errors (tests/test_mass_matrices.py:74-77,114-117), LinearMatrix constructor quirks
(tests/test_linear_dense_simple.py), lowering rules for bounds."""
import numpy as np
import pytest
import scipy.sparse as sp

from hmclab_b200 import Distributions as D
from hmclab_b200 import MassMatrices as M
from hmclab_b200._lowering import describe, describe_mass, flatten


def test_swapped_bounds_raise_the_reference_message():
    dist = D.Normal(np.zeros((3, 1)), 1.0)
    lo, hi = np.full((3, 1), -1.0), np.full((3, 1), 1.0)
    dist.update_bounds(lo, hi)
    with pytest.raises(ValueError, match="Bounds vectors are incompatible."):
        dist.update_bounds(hi, lo)
    assert dist.lower_bounds is lo and dist.upper_bounds is hi     # restored on failure
    with pytest.raises(ValueError, match="incorrect size"):
        dist.update_bounds(np.zeros((2, 1)), None)
    with pytest.raises(ValueError, match="not understood"):
        dist.update_bounds("low", None)


@pytest.mark.parametrize("mass", [M.Unit(4), M.Diagonal(np.ones((4, 1)) * 2.0)])
def test_wrong_momentum_shape_raises_value_error(mass):
    with pytest.raises(ValueError):
        mass.kinetic_energy(np.ones((5, 1)))
    with pytest.raises(ValueError):
        mass.kinetic_energy_gradient(np.ones((4,)))
    assert mass.dimensions == 4 and mass.matrix.shape == (4, 4)


def test_diagonal_mass_uses_the_rounded_reciprocal():
    diag = np.array([[3.0], [7.0], [0.1]])
    plan = describe_mass(M.Diagonal(diag))
    assert np.array_equal(plan["inverse_diagonal"], (1.0 / diag)[:, 0])
    assert describe_mass(M.Unit(3)) == {"kind": "unit", "dims": 3}
    with pytest.raises(NotImplementedError):
        describe_mass(type("BFGS", (), {"dimensions": 3})())


def test_full_mass_lowers_to_the_cholesky_factor_and_the_inverse():
    """MassMatrices.Full (MassMatrices.py:241-327): lower factor as scipy's cho_factor gives it, the
    inverse through cho_solve; an object without a stored inverse (the reference's) gets one."""
    rng = np.random.default_rng(3)
    A = rng.normal(size=(5, 5))
    full = A @ A.T + 5 * np.eye(5)
    mass = M.Full(full)
    plan = describe_mass(mass)
    assert plan["kind"] == "full" and plan["dims"] == 5
    assert np.allclose(plan["cholesky"] @ plan["cholesky"].T, full, atol=1e-12)
    assert np.array_equal(plan["cholesky"], np.tril(plan["cholesky"]))
    assert np.allclose(plan["inverse"] @ full, np.eye(5), atol=1e-12)
    bare = type("Full", (), {"dimensions": 5, "cholesky": mass.cholesky})()
    assert np.allclose(describe_mass(bare)["inverse"], plan["inverse"], atol=1e-14)
    with pytest.raises(AssertionError):
        M.Full(full + np.triu(np.ones((5, 5)), 1))          # not symmetric
    assert M.Full.create_default(4).matrix.shape == (4, 4)


def test_linear_matrix_dtype_quirks_are_inherited():
    rng = np.random.default_rng(0)
    G, d = rng.normal(size=(7, 3)), rng.normal(size=(7, 1))
    lik = D.LinearMatrix(G, d, 2.0)                        # N > d -> premultiplied by default
    inner = lik.Distribution
    assert inner.premultiplication and inner.GtG.shape == (3, 3)
    G32 = G.astype(np.float32).astype(np.float64)
    assert np.allclose(inner.GtG, G32.T @ G32 / 2.0, rtol=1e-15)      # float32-rounded G in f64 math
    assert not np.allclose(inner.GtG, G.T @ G / 2.0, rtol=1e-12)
    lik = D.LinearMatrix(G, d, 2.0, premultiplication=False)
    assert lik.Distribution.G.dtype == np.float32 and lik.Distribution.Gt.dtype == np.float64
    with pytest.raises(ValueError, match="data vector"):
        D.LinearMatrix(G, d[:, 0], 2.0)
    with pytest.raises(ValueError, match="forward model matrix"):
        D.LinearMatrix(G[:5], d, 2.0)
    with pytest.raises(ValueError, match="covariance"):
        D.LinearMatrix(G, d, np.ones((3, 1)))
    full = D.LinearMatrix(G, d, np.eye(7) * 2.0)            # dense covariance, premultiplied
    assert np.allclose(full.Distribution.GtG, inner.GtG, rtol=1e-6)
    direct = D.LinearMatrix(G, d, np.eye(7) * 2.0, premultiplication=False)   # dense covariance, direct form
    node = describe(direct)
    assert node["kind"] == "linear_dense" and not node["premult"] and node["Gt"].shape == (3, 7)
    inv32 = np.linalg.inv((np.eye(7) * 2.0).astype(np.float32))
    assert np.array_equal(node["Gt"], (G.astype(np.float32).T @ inv32).astype(np.float64))   # float32 product
    assert np.allclose(node["misfit_G"].T @ node["misfit_G"], G32.T @ G32 / 2.0, rtol=1e-6)
    with pytest.raises(ValueError, match="data covariance"):     # like the reference's dispatcher (LinearMatrix.py:74-95)
        D.LinearMatrix(sp.csr_matrix(G), d, sp.eye(7).tocsr())
    # sparse G + (N x N) covariance: the sparse-covariance class; lowered to the dense direct form
    cov = np.eye(7) * 2.0 + np.diag(np.full(6, 0.3), 1) + np.diag(np.full(6, 0.3), -1)
    spcov = D.LinearMatrix(sp.csr_matrix(G), d, cov)
    assert type(spcov.Distribution).__name__ == "_LinearMatrix_sparse_forward_sparse_covariance"
    assert spcov.Distribution.data_covariance.dtype == np.float32        # the dispatcher path rounds it
    from hmclab_b200.Distributions.LinearMatrix import _LinearMatrix_sparse_forward_sparse_covariance as SpCov
    node = describe(SpCov(sp.csr_matrix(G), d, sp.csr_matrix(cov)))
    assert node["kind"] == "linear_dense" and not node["premult"] and node["cov_csc"][0].dtype == np.float64
    assert np.allclose(node["Gt"], G32.T @ np.linalg.inv(cov), rtol=1e-12, atol=1e-14)
    assert np.allclose(node["chol_upper"].T @ node["chol_upper"], np.linalg.inv(cov), rtol=1e-12, atol=1e-14)
    with pytest.raises(ValueError, match="symmetric"):
        describe(SpCov(sp.csr_matrix(G), d, sp.csr_matrix(cov + np.diag(np.full(6, 0.1), 1))))
    sparse = D.LinearMatrix(sp.csr_matrix(G), d, 0.5, premultiplication=False)
    node = describe(sparse)
    assert node["kind"] == "linear_csr" and node["indices"].dtype == np.int32
    assert node["t_indptr"].size == 4 and node["indptr"].size == 8


def test_bayesrule_reflects_on_direct_children_only():
    lo, hi = np.full((4, 1), -1.0), np.full((4, 1), 2.0)
    rng = np.random.default_rng(1)
    lik = D.LinearMatrix(rng.normal(size=(3, 4)), rng.normal(size=(3, 1)), 1.0)
    direct = flatten(describe(D.BayesRule([D.Uniform(lo, hi), lik])))
    assert np.array_equal(direct["reflect_lb"], lo[:, 0]) and np.array_equal(direct["reflect_ub"], hi[:, 0])
    nested = D.CompositeDistribution([D.Uniform(lo[:2], hi[:2]), D.Normal(np.zeros((2, 1)), 1.0)])
    inside = flatten(describe(D.BayesRule([nested, lik])))
    assert inside["reflect_lb"] is None and inside["reflect_ub"] is None      # rejection only
    assert any(c["in_gradient"] and c["len"] == 2 for c in inside["checks"])
    top = flatten(describe(nested))                         # bound-less composite: children's bounds
    assert np.array_equal(top["reflect_lb"], [-1.0, -1.0, -np.inf, -np.inf])
    two = D.BayesRule([D.Uniform(lo, hi), D.Normal(np.zeros((4, 1)), 1.0, lower_bounds=lo + 0.5)])
    collapsed = flatten(describe(two))
    assert np.array_equal(collapsed["reflect_lb"], (lo + 0.5)[:, 0])          # max of the lowers


def test_unsupported_objects_are_refused_by_name():
    with pytest.raises(NotImplementedError, match="full covariance|Full-covariance"):
        D.Normal(np.zeros((2, 1)), np.eye(2))
    with pytest.raises(AttributeError):
        D.Laplace(np.zeros((2, 1)), 1.0)                   # dispersions must be an ndarray
    lik = D.SourceLocation3D(np.zeros((1, 3)), np.ones((1, 3)), np.zeros((1, 3)), np.ones((2, 3)),
                             np.ones((2, 3)), infer_velocity=True)
    assert lik.dimensions == 9
    with pytest.raises(NotImplementedError, match="Only one coupled likelihood"):
        flatten(describe(D.BayesRule([lik, lik])))


def _trees_equal(a, b, path="root"):
    assert type(a) is type(b) or (np.isscalar(a) and np.isscalar(b)), path
    if isinstance(a, dict):
        assert sorted(a) == sorted(b), path
        for k in a:
            _trees_equal(a[k], b[k], f"{path}.{k}")
    elif isinstance(a, (list, tuple)):
        assert len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            _trees_equal(x, y, f"{path}[{i}]")
    elif isinstance(a, np.ndarray):
        assert a.dtype == b.dtype and np.array_equal(a, b, equal_nan=True), path
    elif isinstance(a, float) and np.isnan(a):
        assert np.isnan(b), path
    else:
        assert a == b, path


@pytest.mark.skipif(not __import__("os").path.isdir("/root/reference/hmclab"), reason="reference not mounted")
def test_genuine_reference_objects_lower_to_the_same_plan():
    """A posterior built with the reference's own classes is accepted as is: the lowering reads
    it by class and attribute name and yields exactly what the mirror classes yield."""
    import hmclab_b200
    import cases
    from _reference_shim import import_reference
    from hmclab_b200.Samplers import _is_distribution, _is_mass_matrix

    hmclab = import_reference()
    for name in cases.CASES:
        inp = cases.make_inputs(name)
        ref_post, ref_mass = cases.build(name, inp, hmclab)
        our_post, our_mass = cases.build(name, inp, hmclab_b200)
        assert _is_distribution(ref_post) and _is_mass_matrix(ref_mass)
        _trees_equal(describe(ref_post), describe(our_post), name)
        _trees_equal(describe_mass(ref_mass), describe_mass(our_mass), name)
