"""Lowering of a distribution object tree to the flat plan the CUDA engine consumes.

``describe`` walks either ``hmclab_b200`` objects or genuine ``hmclab`` objects (it
dispatches on the class *name* and reads the reference's attribute names, so a
reference user's posterior can be handed over unchanged) and returns a tree of
plain dicts holding float64 numpy arrays.  ``flatten`` turns that tree into

* elementwise prior terms on coordinate ranges (Normal-diagonal, Laplace),
* bound checks (every object with bounds contributes ``+inf`` to the misfit and,
  for the priors/containers, to the gradient on its range; base.py:361-374,
  564-570, 703-710, 786, 867-898, 1036-1054),
* the reflection bounds the trajectory uses (base.py:239-270, 913-978, 1111-1142),
* at most one coupled likelihood (dense / CSR LinearMatrix, SourceLocation3D).
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import numpy as np

LIKELIHOOD_KINDS = ("linear_dense", "linear_csr", "srcloc3d", "srcloc2d")


def _vec(x, n: int, what: str) -> np.ndarray:
    """Scalar or (n,1)/(n,) -> contiguous float64 (n,)."""
    a = np.asarray(x, dtype=np.float64)
    if a.ndim == 0:
        return np.full(n, float(a))
    if a.size != n:
        raise ValueError(f"{what}: expected {n} values, got shape {a.shape}")
    return np.ascontiguousarray(a.reshape(n))


def _bounds(obj, n: int):
    lb = getattr(obj, "lower_bounds", None)
    ub = getattr(obj, "upper_bounds", None)
    return (
        None if lb is None else _vec(lb, n, "lower_bounds"),
        None if ub is None else _vec(ub, n, "upper_bounds"),
    )


def _csr_arrays(mat):
    import scipy.sparse as sp

    m = sp.csr_matrix(mat)
    m.sum_duplicates()
    m.sort_indices()
    return (
        np.ascontiguousarray(m.indptr, dtype=np.int32),
        np.ascontiguousarray(m.indices, dtype=np.int32),
        np.ascontiguousarray(m.data, dtype=np.float64),
    )


def describe(dist) -> Dict[str, Any]:
    """Distribution object -> tree of plain dicts (see module docstring)."""
    cls = type(dist).__name__
    n = int(dist.dimensions)
    lb, ub = _bounds(dist, n)
    node: Dict[str, Any] = {"dims": n, "lb": lb, "ub": ub}

    if cls == "Normal":
        if not getattr(dist, "diagonal", True):
            raise NotImplementedError("Normal with full covariance is not lowered.")
        node.update(
            kind="normal",
            means=_vec(dist.means, n, "means"),
            inv_cov=_vec(dist.inverse_covariance, n, "inverse_covariance"),
            const=float(dist.normalization_constant),
        )
    elif cls == "Laplace":
        node.update(
            kind="laplace",
            means=_vec(dist.means, n, "means"),
            inv_disp=_vec(dist.inverse_dispersions, n, "inverse_dispersions"),
            const=float(dist.normalization_constant),
        )
    elif cls == "Uniform":
        node.update(kind="uniform")
    elif cls in ("AdditiveDistribution", "BayesRule"):
        node.update(
            kind="additive", children=[describe(c) for c in dist.separate_distributions]
        )
    elif cls == "CompositeDistribution":
        node.update(
            kind="composite", children=[describe(c) for c in dist.separate_distributions]
        )
    elif cls == "LinearMatrix":
        inner = describe(dist.Distribution)
        # wrapper adds its own misfit_bounds on top of the inner ones (LinearMatrix.py:114-116)
        inner["wrapper_lb"], inner["wrapper_ub"] = lb, ub
        return inner
    elif cls in ("_LinearMatrix_dense_forward_simple_covariance",
                 "_LinearMatrix_dense_forward_dense_covariance"):
        node.update(kind="linear_dense", premult=bool(dist.premultiplication))
        if cls.endswith("dense_covariance") and not dist.premultiplication:
            # LinearMatrix.py:257-288: gradient Gt @ invcov @ (G m - d), evaluated left to right, so
            # the reference forms W = Gt @ invcov (a product in the class's dtype, float32 by default)
            # at every call; it is formed ONCE here with the same numpy expression.  Misfit:
            # 0.5 |U (G m - d)|^2 with the upper Cholesky factor U of the inverse covariance; the
            # engine gets U G and U d (float64 products of the float32 operands).
            N = int(dist.G.shape[0])
            G = np.ascontiguousarray(dist.G, dtype=np.float64)
            W = np.ascontiguousarray(dist.Gt @ dist.invcov, dtype=np.float64)
            U = np.ascontiguousarray(dist.cholesky_upper_inv_covariance, dtype=np.float64)
            dvec = _vec(dist.d, N, "d")
            node.update(N=N, G=G, Gt=W, d=dvec, var=np.ones(N), sigma=np.ones(N), chol_upper=U,
                        misfit_G=np.ascontiguousarray(U @ G), misfit_d=np.ascontiguousarray(U @ dvec))
        elif dist.premultiplication:
            node.update(
                GtG=np.ascontiguousarray(dist.GtG, dtype=np.float64),
                Gtd0=_vec(dist.Gtd0, n, "Gtd0"),
                dtd=float(dist.dtd),
            )
        else:
            N = int(dist.G.shape[0])
            G = np.ascontiguousarray(dist.G, dtype=np.float64)
            Gt = np.ascontiguousarray(dist.Gt, dtype=np.float64)
            node.update(
                N=N,
                G=G,
                Gt=None if np.array_equal(Gt, G.T) else Gt,
                d=_vec(dist.d, N, "d"),
                var=_vec(dist.data_variance, N, "data_variance"),
                sigma=_vec(dist.data_sigma, N, "data_sigma"),
            )
    elif cls == "_LinearMatrix_sparse_forward_sparse_covariance":
        # LinearMatrix.py:444-519: gradient Gt @ solve(cov, G m - d) with a sparse LU factorisation
        # formed once by the constructor.  The batched engine has no sparse triangular solve; the
        # factorisation is applied to the identity ONCE here (host set-up, like the reference's
        # constructor) and the products run in the dense direct dense-covariance form:
        # W = Gt invcov, misfit 0.5 |U (G m - d)|^2 with U^T U = invcov.  Everything is float64 on
        # the stored values; a float32 covariance (what the public dispatcher produces) makes the
        # reference solve in single precision, which this path does not imitate.
        import scipy.sparse.linalg as spla
        N = int(dist.G.shape[0])
        if 8.0 * N * (N + 2 * n) > 6e9:
            raise NotImplementedError(
                f"sparse-covariance LinearMatrix with {N} data x {n} parameters: the dense lowering "
                "(N x N inverse covariance, N x d operators) would need more than 6 GB")
        cov = dist.data_covariance.tocsc()
        invcov = spla.splu(cov.astype(np.float64)).solve(np.eye(N))
        sym = 0.5 * (invcov + invcov.T)
        if not np.allclose(invcov, sym, rtol=1e-9, atol=1e-12 * np.max(np.abs(invcov))):
            raise ValueError("the data covariance of a LinearMatrix must be symmetric")
        try:
            U = np.ascontiguousarray(np.linalg.cholesky(sym).T)
        except np.linalg.LinAlgError as err:
            raise ValueError("the data covariance of a LinearMatrix must be positive definite") from err
        G = np.ascontiguousarray(dist.G.toarray(), dtype=np.float64)
        dvec = _vec(dist.d, N, "d")
        node.update(kind="linear_dense", premult=False, N=N, G=G,
                    Gt=np.ascontiguousarray(G.T @ invcov), d=dvec, var=np.ones(N), sigma=np.ones(N),
                    chol_upper=U, misfit_G=np.ascontiguousarray(U @ G),
                    misfit_d=np.ascontiguousarray(U @ dvec),
                    cov_csc=(cov.data.copy(), cov.indices.copy(), cov.indptr.copy()),
                    G_csr=_csr_arrays(dist.G), d_stored=np.asarray(dist.d).copy())
    elif cls == "_LinearMatrix_sparse_forward_simple_covariance":
        node.update(kind="linear_csr", premult=bool(dist.premultiplication))
        if dist.premultiplication:
            indptr, indices, data = _csr_arrays(dist.GtG)
            node.update(
                N=n,
                indptr=indptr,
                indices=indices,
                data=data,
                Gtd0=_vec(dist.Gtd0, n, "Gtd0"),
                dtd=float(dist.dtd),
            )
        else:
            N = int(dist.G.shape[0])
            indptr, indices, data = _csr_arrays(dist.G)
            t_indptr, t_indices, t_data = _csr_arrays(dist.Gt)
            node.update(
                N=N,
                indptr=indptr,
                indices=indices,
                data=data,
                t_indptr=t_indptr,
                t_indices=t_indices,
                t_data=t_data,
                d=_vec(dist.d, N, "d"),
                var=_vec(dist.data_variance, N, "data_variance"),
                sigma=_vec(dist.data_sigma, N, "data_sigma"),
            )
    elif cls == "SourceLocation3D":
        E, S = int(dist.number_of_events), int(dist.number_of_stations)
        node.update(
            kind="srcloc3d",
            events=E,
            stations=S,
            rx=_vec(dist.receiver_array_x, S, "receiver_array_x"),
            ry=_vec(dist.receiver_array_y, S, "receiver_array_y"),
            rz=_vec(dist.receiver_array_z, S, "receiver_array_z"),
            tobs=np.ascontiguousarray(dist.observed_data, dtype=np.float64).reshape(E, S),
            std=np.ascontiguousarray(dist.data_std, dtype=np.float64).reshape(E, S),
            infer_velocity=bool(dist.infer_velocity),
            velocity=(
                float("nan")
                if dist.infer_velocity
                else float(np.asarray(dist.medium_velocity).reshape(-1)[0])
            ),
        )
    elif cls == "SourceLocation2D":
        E, S = int(dist.number_of_events), int(dist.number_of_stations)
        node.update(
            kind="srcloc2d",
            events=E,
            stations=S,
            rx=_vec(dist.receiver_array_x, S, "receiver_array_x"),
            rz=_vec(dist.receiver_array_z, S, "receiver_array_z"),
            tobs=np.ascontiguousarray(dist.observed_data, dtype=np.float64).reshape(E, S),
            std=np.ascontiguousarray(dist.data_std, dtype=np.float64).reshape(E, S),
            infer_velocity=bool(dist.infer_velocity),
            velocity=(
                float("nan")
                if dist.infer_velocity
                else float(np.asarray(dist.medium_velocity).reshape(-1)[0])
            ),
        )
    else:
        raise NotImplementedError(
            f"Distribution type `{cls}` is not on the batched B200 path "
            "(supported: Normal (diagonal), Laplace, Uniform, CompositeDistribution, "
            "AdditiveDistribution/BayesRule, LinearMatrix with scalar/vector variance, "
            "SourceLocation3D, SourceLocation2D)."
        )
    return node


def describe_mass(mass) -> Dict[str, Any]:
    cls = type(mass).__name__
    n = int(mass.dimensions)
    if cls == "Unit":
        return {"kind": "unit", "dims": n}
    if cls == "Diagonal":
        return {
            "kind": "diagonal",
            "dims": n,
            "diagonal": _vec(mass.diagonal, n, "diagonal"),
            "inverse_diagonal": _vec(mass.inverse_diagonal, n, "inverse_diagonal"),
        }
    if cls == "Full":
        # MassMatrices.py:241-327; genuine reference objects carry the factor but no inverse
        from scipy.linalg import cho_solve

        chol = np.ascontiguousarray(np.tril(np.asarray(mass.cholesky, dtype=np.float64)))
        if chol.shape != (n, n):
            raise ValueError("Full mass matrix: the Cholesky factor has the wrong shape.")
        inverse = getattr(mass, "inverse", None)
        if inverse is None:
            inverse = cho_solve((chol, True), np.eye(n))
        return {"kind": "full", "dims": n, "cholesky": chol,
                "inverse": np.ascontiguousarray(inverse, dtype=np.float64)}
    raise NotImplementedError(
        f"Mass matrix `{cls}` is not on the batched B200 path (Unit, Diagonal, Full)."
    )


# --------------------------------------------------------------------------------------
def _reflection(tree) -> (Optional[np.ndarray], Optional[np.ndarray]):
    """Bounds the *top-level* object's ``corrector`` reflects on."""
    n = tree["dims"]
    lb, ub = tree["lb"], tree["ub"]
    if tree["kind"] in LIKELIHOOD_KINDS and "wrapper_lb" in tree:
        lb, ub = tree["wrapper_lb"], tree["wrapper_ub"]
    if tree["kind"] == "composite" and lb is None and ub is None:
        # base.py:946-978: falls through to the direct children's own bounds
        lo = np.full(n, -np.inf)
        hi = np.full(n, np.inf)
        any_lo = any_hi = False
        off = 0
        for child in tree["children"]:
            c_lb, c_ub = child["lb"], child["ub"]
            if child["kind"] in LIKELIHOOD_KINDS and "wrapper_lb" in child:
                c_lb, c_ub = child["wrapper_lb"], child["wrapper_ub"]
            if c_lb is not None:
                lo[off : off + child["dims"]] = c_lb
                any_lo = True
            if c_ub is not None:
                hi[off : off + child["dims"]] = c_ub
                any_hi = True
            off += child["dims"]
        return (lo if any_lo else None), (hi if any_hi else None)
    return lb, ub


def flatten(tree) -> Dict[str, Any]:
    """Tree from ``describe`` -> flat engine plan."""
    plan: Dict[str, Any] = {
        "dims": tree["dims"],
        "terms": [],
        "checks": [],
        "likelihood": None,
    }

    def add_check(offset, n, lb, ub, in_gradient):
        if lb is None and ub is None:
            return
        for chk in plan["checks"]:
            same = (
                chk["offset"] == offset
                and chk["len"] == n
                and (chk["lb"] is None) == (lb is None)
                and (chk["ub"] is None) == (ub is None)
                and (lb is None or np.array_equal(chk["lb"], lb))
                and (ub is None or np.array_equal(chk["ub"], ub))
            )
            if same:
                chk["in_gradient"] = chk["in_gradient"] or in_gradient
                return
        plan["checks"].append(
            {"offset": offset, "len": n, "lb": lb, "ub": ub, "in_gradient": in_gradient}
        )

    def visit(node, offset):
        kind, n = node["kind"], node["dims"]
        if kind in LIKELIHOOD_KINDS:
            if plan["likelihood"] is not None:
                raise NotImplementedError("Only one coupled likelihood per posterior.")
            if offset != 0 or n != tree["dims"]:
                raise NotImplementedError(
                    "A coupled likelihood must span all coordinates of the posterior."
                )
            plan["likelihood"] = node
            add_check(offset, n, node["lb"], node["ub"], False)
            add_check(offset, n, node.get("wrapper_lb"), node.get("wrapper_ub"), False)
            return
        add_check(offset, n, node["lb"], node["ub"], True)
        if kind == "normal":
            plan["terms"].append(
                {"kind": "normal", "offset": offset, "len": n, "a": node["means"],
                 "b": node["inv_cov"], "const": node["const"]}
            )
        elif kind == "laplace":
            plan["terms"].append(
                {"kind": "laplace", "offset": offset, "len": n, "a": node["means"],
                 "b": node["inv_disp"], "const": node["const"]}
            )
        elif kind == "uniform":
            pass
        elif kind == "additive":
            for child in node["children"]:
                if child["dims"] != n:
                    raise ValueError("Additive children must share dimensions.")
                visit(child, offset)
        elif kind == "composite":
            off = offset
            for child in node["children"]:
                visit(child, off)
                off += child["dims"]
        else:  # pragma: no cover
            raise NotImplementedError(kind)

    visit(tree, 0)
    plan["reflect_lb"], plan["reflect_ub"] = _reflection(tree)
    return plan
