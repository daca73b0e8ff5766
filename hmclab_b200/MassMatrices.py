"""Mass matrices (HMC metrics) of the batched engine: ``Unit``, ``Diagonal`` and ``Full``.

Host-side mirror of hmclab/MassMatrices.py:27-327.  The objects hold the metric;
momentum generation, kinetic energy and its gradient are fused into the CUDA
integrator kernels.  The single-vector methods below run those kernels on a batch
of one chain (they exist so the reference's mass-matrix tests can be replayed).
"""
from __future__ import annotations

import numpy as _numpy


class _AbstractMassMatrix:
    name: str = "mass matrix abstract base class"
    dimensions: int = -1
    rng = None

    def accept(self):  # MassMatrices.py:75-79: no-ops for Unit / Diagonal
        pass

    def reject(self):
        pass

    # -- single-vector protocol, evaluated on the device -----------------------------
    def _check(self, momentum):
        if momentum.shape != (self.dimensions, 1):
            raise ValueError()

    def _evaluator(self):
        from hmclab_b200._evaluator import mass_evaluator_for

        return mass_evaluator_for(self)

    def kinetic_energy(self, momentum: _numpy.ndarray) -> float:
        self._check(momentum)
        return float(self._evaluator().kinetic_energy_batch(momentum.T)[0])

    def kinetic_energy_gradient(self, momentum, position=None, g=None):
        self._check(momentum)
        return self._evaluator().kinetic_gradient_batch(momentum.T).T.copy()

    def generate_momentum(self) -> _numpy.ndarray:
        """One momentum draw (d,1).  Standard normals come from ``self.rng`` (a numpy
        Generator, as in the reference) and are scaled on the device."""
        rng = self.rng if self.rng is not None else _numpy.random.default_rng()
        z = rng.normal(size=(self.dimensions, 1))
        return self._evaluator().scale_momentum_batch(z.T).T.copy()


class Unit(_AbstractMassMatrix):
    """Identity metric (MassMatrices.py:82-156)."""

    def __init__(self, dimensions: int = -1, rng=None):
        self.name = "unit mass matrix"
        self.dimensions = int(dimensions)
        if rng is not None:
            self.rng = rng

    @property
    def matrix(self):
        return _numpy.eye(self.dimensions)

    @staticmethod
    def create_default(dimensions: int, rng=None) -> "Unit":
        return Unit(dimensions, rng)


class Diagonal(_AbstractMassMatrix):
    """Diagonal metric (MassMatrices.py:159-238); ``inverse_diagonal = 1/diagonal`` is
    precomputed so the kinetic gradient is a multiplication by the rounded
    reciprocal, as in the reference."""

    def __init__(self, diagonal, rng=None):
        self.name = "diagonal mass matrix"
        diagonal = _numpy.asarray(diagonal, dtype=_numpy.float64)
        self.diagonal = diagonal.reshape(diagonal.size, 1)
        self.dimensions = int(diagonal.size)
        self.inverse_diagonal = 1.0 / self.diagonal
        if rng is not None:
            self.rng = rng

    @property
    def matrix(self):
        return _numpy.diagflat(self.diagonal)

    @staticmethod
    def create_default(dimensions: int, rng=None) -> "Diagonal":
        return Diagonal(_numpy.ones((dimensions, 1)), rng=rng)


class Full(_AbstractMassMatrix):
    """Dense symmetric positive definite metric (MassMatrices.py:241-327): momentum
    ``cholesky @ normal``, kinetic energy ``0.5 p . M^-1 p``, gradient ``M^-1 p``.  The reference
    solves with the Cholesky factor (scipy ``cho_solve``) per call; on the batch ``M^-1`` (formed
    once from that factor) and the factor itself are applied as fp64 tensor-core products."""

    def __init__(self, full, rng=None, do_hermitian_check=True):
        from scipy.linalg import cho_factor, cho_solve

        self.name = "diagonal mass matrix"      # sic: the reference's Full carries this name (:258)
        self.mass_matrix = _numpy.asarray(full, dtype=_numpy.float64)
        if do_hermitian_check:
            assert _numpy.allclose(self.mass_matrix, self.mass_matrix.T)
        self.cholesky, self.cholesky_lower = cho_factor(self.mass_matrix, lower=True)
        self.cholesky = _numpy.tril(self.cholesky)
        self.dimensions = int(self.mass_matrix.shape[0])
        self.inverse = cho_solve((self.cholesky, self.cholesky_lower), _numpy.eye(self.dimensions))
        if rng is not None:
            self.rng = rng

    @property
    def matrix(self):
        return self.mass_matrix

    @staticmethod
    def create_default(dimensions: int, rng=None) -> "Full":
        mass_matrix = (_numpy.eye(dimensions) + 0.1 * _numpy.eye(dimensions, k=-1)
                       + 0.1 * _numpy.eye(dimensions, k=1))
        return Full(mass_matrix, rng=rng)
