"""Gaussian likelihood with a linear forward model ``G m = d`` (host-side mirror).

Mirrors the constructor semantics of the reference so that the arrays handed to
the CUDA engine are the ones the reference would multiply with:

* dispatcher ``LinearMatrix``  -- hmclab/Distributions/LinearMatrix.py:15-135
* dense G, scalar/vector variance  -- LinearMatrix.py:139-222
* sparse G, scalar/vector variance -- LinearMatrix.py:309-440
* dense G, dense data covariance (both forms) -- LinearMatrix.py:226-305
* sparse G, sparse data covariance -- LinearMatrix.py:444-519

Inherited quirks (they are part of the contract):

* the inner classes round ``G``, ``d`` and a vector variance to ``dtype``
  (default ``numpy.single``); the dispatcher's own ``dtype`` keyword is not
  forwarded, so through ``LinearMatrix(...)`` the matrices are always float32
  values that are then used in float64 arithmetic;
* dense + premultiplication with a *vector* variance forms ``GtG`` in float32;
* ``Gtd0`` and the dense ``Gt`` use the caller's un-rounded ``G``.

Only parameters are prepared here (once, at construction, exactly as the
reference constructor does); ``misfit``/``gradient`` run on the GPU.
"""
from __future__ import annotations

import numpy as _numpy
import scipy.sparse as _sparse

from hmclab_b200.Distributions.base import _AbstractDistribution


class LinearMatrix(_AbstractDistribution):
    def __init__(self, G, d, data_covariance, dtype=None, **kwargs):
        self.name = "linear forward model with Gaussian data errors"
        if dtype is not None and _numpy.dtype(dtype) != G.dtype:
            # The reference assigns ``G.dtype = dtype`` (LinearMatrix.py:25-26), which
            # re-interprets the raw buffer.  That is never what a caller wants.
            raise ValueError(
                "LinearMatrix(dtype=...) must equal G.dtype; the reference would "
                "re-interpret G's memory instead of converting it."
            )
        d = d.astype(G.dtype)
        if type(data_covariance) not in (float, _numpy.float32, _numpy.float64):
            data_covariance = data_covariance.astype(G.dtype)

        self.dimensions = int(G.shape[1])
        if not (type(d) is _numpy.ndarray and d.shape == (d.size, 1)):
            raise ValueError(
                "Didn't understand the data vector object. Should be a NumPy column "
                f"vector (ndarray: [datapoints, 1]). {type(d)}, {d.shape}"
            )
        if type(G) is _numpy.ndarray and G.shape == (d.size, self.dimensions):
            dense = True
        elif _sparse.issparse(G) and G.shape == (d.size, self.dimensions):
            dense = False
        else:
            raise ValueError("Didn't understand the forward model matrix object.")

        if type(data_covariance) is float or (
            type(data_covariance) == _numpy.ndarray
            and data_covariance.shape == (d.size, 1)
        ):
            simple = True
        elif type(data_covariance) is _numpy.ndarray and data_covariance.shape == (
            d.size,
            d.size,
        ):
            simple = False
        else:
            # like the reference (LinearMatrix.py:74-95) the dispatcher only recognises an ndarray
            # covariance; a scipy.sparse covariance goes to the inner class directly
            raise ValueError("Didn't understand the data covariance object.")
        if not simple and not dense:
            inner = _LinearMatrix_sparse_forward_sparse_covariance
        elif not simple:
            inner = _LinearMatrix_dense_forward_dense_covariance
        elif dense:
            inner = _LinearMatrix_dense_forward_simple_covariance
        else:
            inner = _LinearMatrix_sparse_forward_simple_covariance
        self.Distribution = inner(G, d, data_covariance, **kwargs)

    @staticmethod
    def create_default(dimensions: int, dtype=_numpy.dtype("float64")) -> "LinearMatrix":
        return LinearMatrix(
            _numpy.eye(dimensions, dtype=dtype), _numpy.ones((dimensions, 1)), 1.0
        )


def _weights(data_variance):
    """1/variance with the dtype numpy gives the reference (LinearMatrix.py:170-173)."""
    if type(data_variance) == float:
        return 1.0 / data_variance
    return 1.0 / data_variance[:, 0]


class _LinearMatrix_dense_forward_simple_covariance(_AbstractDistribution):
    def __init__(self, G, d, data_variance, dtype=_numpy.single, premultiplication=None):
        self.name = "dense linear forward model"
        self.dimensions = int(G.shape[1])
        self.G = G.astype(dtype)
        self.d = d.astype(dtype)
        if type(data_variance) == _numpy.ndarray:
            self.data_variance = data_variance.astype(dtype)
        else:
            self.data_variance = data_variance
        self.data_sigma = self.data_variance**0.5
        if premultiplication is not None:
            self.premultiplication = premultiplication
        else:
            self.premultiplication = self.G.shape[0] > self.G.shape[1]

        if self.premultiplication:
            # The reference multiplies with an explicit N x N diagonal matrix
            # (LinearMatrix.py:169-177).  A diagonal factor contributes exactly one
            # rounded product per entry, so scaling columns gives the same numbers
            # without the N^2 temporary; dtypes follow numpy's promotion so a float32
            # vector variance keeps GtG in float32 like the reference does.
            w = _weights(self.data_variance)
            if type(self.data_variance) == float:
                scaled_Gt = self.G.T.astype(_numpy.float64) * w
                scaled_G0t = G.T.astype(_numpy.float64) * w
                scaled_dt = self.d.T.astype(_numpy.float64) * w
            else:
                scaled_Gt = self.G.T * w[None, :]
                scaled_G0t = G.T * w[None, :]
                scaled_dt = self.d.T * w[None, :]
            # C-ordered like the reference's matmul results, so BLAS sees the same call
            self.GtG = _numpy.ascontiguousarray(scaled_Gt) @ self.G
            self.Gtd0 = _numpy.ascontiguousarray(scaled_G0t) @ self.d
            self.dtd = (_numpy.ascontiguousarray(scaled_dt) @ self.d).item()
            del self.G, self.d, self.data_variance, self.data_sigma
        else:
            self.Gt = G.T


class _LinearMatrix_dense_forward_dense_covariance(_AbstractDistribution):
    """Dense G with a dense (N x N) data covariance (LinearMatrix.py:226-305).  The premultiplied
    form (the default when N > dimensions) reduces to the GtG / Gtd0 / dtd arithmetic of the
    simple-covariance class.  The direct form keeps G, the inverse covariance and the upper
    Cholesky factor of the inverse covariance like the reference (all in ``dtype``, default
    float32): gradient ``Gt @ invcov @ (G m - d)``, misfit ``0.5 |U (G m - d)|^2``."""

    def __init__(self, G, d, data_covariance, dtype=_numpy.single, premultiplication=None):
        self.name = "dense linear forward model, dense data covariance"
        self.dimensions = int(G.shape[1])
        self.G = G.astype(dtype)
        self.d = d.astype(dtype)
        self.data_covariance = data_covariance.astype(dtype)
        if premultiplication is not None:
            self.premultiplication = premultiplication
        else:
            self.premultiplication = self.G.shape[0] > self.G.shape[1]
        self.invcov = _numpy.linalg.inv(self.data_covariance)
        if self.premultiplication:
            self.GtG = self.G.T @ self.invcov @ self.G
            self.Gtd0 = G.T @ self.invcov @ self.d
            self.dtd = (self.d.T @ self.invcov @ self.d).item()
            del self.G, self.d, self.data_covariance
        else:
            self.Gt = self.G.T
            self.cholesky_upper_inv_covariance = _numpy.linalg.cholesky(self.invcov).T


class _LinearMatrix_sparse_forward_simple_covariance(_AbstractDistribution):
    def __init__(
        self,
        G,
        d,
        data_variance,
        dtype=_numpy.single,
        premultiplication=None,
        use_mkl=False,
    ):
        self.name = "sparse linear forward model"
        self.dimensions = int(G.shape[1])
        self.G = _sparse.csr_matrix(G, dtype=dtype)
        self.d = d.astype(dtype)
        if type(data_variance) == _numpy.ndarray:
            self.data_variance = data_variance.astype(dtype)
        else:
            self.data_variance = data_variance
        self.data_sigma = self.data_variance**0.5
        # MKL was a CPU accelerator for exactly this product (LinearMatrix.py:362-387);
        # the CUDA SpMM replaces it, the flag is accepted and ignored.
        self.use_mkl = False
        self.dtype = dtype
        if premultiplication is not None:
            self.premultiplication = premultiplication
        else:
            self.premultiplication = self.G.shape[0] > self.G.shape[1]

        if self.premultiplication:
            w = _weights(self.data_variance)
            if type(self.data_variance) == float:
                invcov = _sparse.eye(self.d.size).tocsr() / self.data_variance
            else:
                invcov = _sparse.diags(w, offsets=0).tocsr()
            self.GtG = self.G.T @ invcov @ self.G
            self.Gtd0 = G.T @ invcov @ self.d
            self.dtd = (self.d.T @ invcov @ self.d).item()
            del self.G, self.d, self.data_variance, self.data_sigma
        else:
            self.Gt = self.G.T.astype(dtype)


class _LinearMatrix_sparse_forward_sparse_covariance(_AbstractDistribution):
    """Sparse G with a sparse (N x N) data covariance (LinearMatrix.py:444-519): gradient
    ``Gt @ solve(cov, G m - d)``, misfit ``0.5 (G m - d)^T solve(cov, G m - d)``.  The reference
    factorises the covariance once (SuperLU) and solves per evaluation; the batched engine applies
    the inverse as a dense operator instead (``_lowering.describe``), so only the constructor
    arithmetic is mirrored here: ``G`` and ``d`` rounded to ``dtype`` (default float32), a
    covariance that is not sparse yet converted to CSR in ``dtype``, a sparse one kept as given."""

    def __init__(self, G, d, data_covariance, dtype=_numpy.single):
        self.name = "sparse linear forward model, sparse data covariance"
        self.dimensions = int(G.shape[1])
        self.G = G.astype(dtype)
        self.d = d.astype(dtype)
        if not _sparse.issparse(data_covariance):
            data_covariance = _sparse.csr_matrix(data_covariance, dtype=dtype)
        self.data_covariance = data_covariance
        self.Gt = self.G.T.tocsr()
        self.dt = d.T.astype(dtype)

    @staticmethod
    def create_default(dimensions: int, dtype=_numpy.dtype("float64")):
        return _LinearMatrix_sparse_forward_sparse_covariance(
            _sparse.eye(dimensions, dtype=dtype).tocsr(), _numpy.ones((dimensions, 1)),
            _sparse.eye(dimensions).tocsr(), dtype=dtype)
