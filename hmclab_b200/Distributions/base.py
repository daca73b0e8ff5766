"""Host-side mirror of the hmclab distribution protocol for the batched B200 engine.

These classes carry *parameters only*.  All arithmetic of the hot path
(``misfit``/``gradient``/``corrector`` inside the HMC trajectory) happens in the
CUDA engine; ``misfit(m)`` and ``gradient(m)`` on these objects call the same
device kernels on a batch of one chain, so there is no CPU implementation of the
path anywhere in this package.

Mirrored interface (reference file:line, relative to the hmclab repository):

* ``_AbstractDistribution``  -- hmclab/Distributions/base.py:21-374
  (``dimensions``, ``lower_bounds``/``upper_bounds``, ``update_bounds``,
  ``misfit``, ``gradient``, ``corrector``)
* ``Normal`` (diagonal branch)  -- base.py:427-574
* ``Laplace``  -- base.py:646-710
* ``Uniform``  -- base.py:747-786
* ``CompositeDistribution``  -- base.py:803-1002
* ``AdditiveDistribution`` / ``BayesRule``  -- base.py:1005-1205

Attribute names are the reference's, because ``hmclab_b200._lowering.describe``
reads either these objects or genuine ``hmclab`` objects by attribute name.
"""
from __future__ import annotations

from typing import List, Optional, Union

import numpy as _numpy

_SCALAR_TYPES = (float, int, _numpy.float64, _numpy.float32)


class _AbstractDistribution:
    """Parameter holder + batched device evaluation (base.py:21-374)."""

    name: str = None
    dimensions: int = -1
    lower_bounds: Optional[_numpy.ndarray] = None
    upper_bounds: Optional[_numpy.ndarray] = None
    normalized: bool = False

    # -- bounds bookkeeping (base.py:272-359) --------------------------------------
    def update_bounds(self, lower=None, upper=None):
        previous = (self.lower_bounds, self.upper_bounds)
        if type(upper) == list:
            upper = _numpy.array(upper)[:, None]
        if type(lower) == list:
            lower = _numpy.array(lower)[:, None]
        self.lower_bounds, self.upper_bounds = lower, upper

        def _fail(msg):
            self.lower_bounds, self.upper_bounds = previous
            raise ValueError(msg)

        if lower is not None and type(lower) is not _numpy.ndarray:
            _fail("Lower bounds object not understood.")
        if upper is not None and type(upper) is not _numpy.ndarray:
            _fail("Upper bounds object not understood.")
        for b in (lower, upper):
            if b is not None and b.shape != (self.dimensions, 1):
                _fail("Bounds vectors are of incorrect size.")
        if lower is not None and upper is not None and _numpy.any(upper <= lower):
            _fail("Bounds vectors are incompatible.")

    # -- device evaluation ---------------------------------------------------------
    def _evaluator(self):
        from hmclab_b200._evaluator import evaluator_for

        return evaluator_for(self)

    def misfit(self, coordinates: _numpy.ndarray) -> float:
        """chi(m) for one column vector (d,1); evaluated by the CUDA engine."""
        coordinates = _numpy.asarray(coordinates, dtype=_numpy.float64)
        if coordinates.shape != (self.dimensions, 1):
            raise ValueError(
                f"Expected a column vector of shape {(self.dimensions, 1)}, "
                f"got {coordinates.shape}."
            )
        return float(self._evaluator().misfit_batch(coordinates.T)[0])

    def gradient(self, coordinates: _numpy.ndarray) -> _numpy.ndarray:
        """grad chi(m) for one column vector (d,1); evaluated by the CUDA engine."""
        coordinates = _numpy.asarray(coordinates, dtype=_numpy.float64)
        if coordinates.shape != (self.dimensions, 1):
            raise ValueError(
                f"Expected a column vector of shape {(self.dimensions, 1)}, "
                f"got {coordinates.shape}."
            )
        return self._evaluator().gradient_batch(coordinates.T).T.copy()

    def misfit_batch(self, coordinates: _numpy.ndarray) -> _numpy.ndarray:
        """chi for a batch [chains x dimensions] (host array in, host array out)."""
        return self._evaluator().misfit_batch(coordinates)

    def gradient_batch(self, coordinates: _numpy.ndarray) -> _numpy.ndarray:
        """grad chi for a batch [chains x dimensions]."""
        return self._evaluator().gradient_batch(coordinates)

    def corrector(self, coordinates: _numpy.ndarray, momentum: _numpy.ndarray):
        """One-shot mirror reflection on the bounds, in place (base.py:239-270).

        Runs the engine's reflection kernel on a batch of one chain."""
        q, p = self._evaluator().corrector_batch(
            _numpy.asarray(coordinates, dtype=_numpy.float64).T,
            _numpy.asarray(momentum, dtype=_numpy.float64).T,
        )
        coordinates[...] = q.T
        momentum[...] = p.T

    def generate(self, repeat=1, rng=None):
        raise NotImplementedError(
            "Drawing direct samples is not part of the batched HMC path."
        )


class Normal(_AbstractDistribution):
    """Uncorrelated Gaussian (the diagonal branch of base.py:427-574).

    ``covariance`` is a scalar or a (d,1)/list vector of variances.  A full
    (d,d) covariance matrix is outside the accelerated path and is refused."""

    def __init__(
        self,
        means,
        covariance,
        inverse_covariance=None,
        lower_bounds=None,
        upper_bounds=None,
    ):
        self.name = "Gaussian (normal) distribution"
        if type(means) == list:
            means = _numpy.array(means)[:, None]
        if type(covariance) == list:
            covariance = _numpy.array(covariance)[:, None]

        if type(means) in (float, int):
            self.dimensions = 1
            means = _numpy.ones((1, 1)) * means
        else:
            means = _numpy.asarray(means)
            self.dimensions = int(means.size)
            means = means.reshape(self.dimensions, 1)
        self.means = means

        self.normalization_constant = 0.0
        if type(covariance) in _SCALAR_TYPES:
            covariance = _numpy.float64(covariance)
        else:
            covariance = _numpy.asarray(covariance)
            if covariance.shape == (self.dimensions, self.dimensions) and self.dimensions > 1:
                raise NotImplementedError(
                    "Full-covariance Normal is outside the batched B200 path "
                    "(only scalar / diagonal variances are lowered)."
                )
            covariance = covariance.reshape(self.dimensions, 1)
        self.diagonal = True
        self.covariance = covariance
        if inverse_covariance is not None:
            self.inverse_covariance = inverse_covariance
        else:
            self.inverse_covariance = 1.0 / self.covariance
        self.update_bounds(lower_bounds, upper_bounds)

    def normalize(self):
        # base.py:576-598
        if _numpy.ndim(self.covariance) == 0:
            determinant = self.covariance**self.dimensions
        else:
            determinant = _numpy.prod(self.covariance)
        self.normalization_constant = 0.5 * (
            _numpy.log(_numpy.abs(determinant))
            + self.dimensions * _numpy.log(2 * _numpy.pi)
        )

    @staticmethod
    def create_default(dimensions: int, diagonal=True) -> "Normal":
        means = _numpy.random.rand(dimensions, 1)
        variances = (_numpy.random.rand(dimensions, 1) + 1.0) ** 2
        return Normal(means, variances)


class Laplace(_AbstractDistribution):
    """Uncorrelated Laplace / L1 distribution (base.py:646-710)."""

    def __init__(self, means, dispersions, lower_bounds=None, upper_bounds=None):
        self.name = "Laplace distribution"
        if type(means) == list:
            means = _numpy.array(means)[:, None]
        if type(dispersions) == list:
            dispersions = _numpy.array(dispersions)[:, None]
        means = _numpy.asarray(means)
        self.dimensions = int(means.size)
        self.means = means.reshape(self.dimensions, 1)
        dispersions = _numpy.asarray(dispersions)
        if dispersions.ndim == 0:
            raise AttributeError("dispersions must be an ndarray of shape (d,1).")
        self.dispersions = dispersions.reshape(self.dimensions, 1)
        self.inverse_dispersions = 1.0 / self.dispersions
        self.normalization_constant = 0.0
        self.update_bounds(lower_bounds, upper_bounds)

    def normalize(self):
        # base.py:712-727
        self.normalization_constant = _numpy.log(
            1.0 / (2.0 * (_numpy.prod(self.dispersions) ** (1.0 / self.dimensions)))
        )

    @staticmethod
    def create_default(dimensions: int) -> "Laplace":
        return Laplace(
            _numpy.random.rand(dimensions, 1), 10 ** _numpy.random.rand(dimensions, 1)
        )


class Uniform(_AbstractDistribution):
    """Box prior: 0 inside, +inf outside (base.py:747-800)."""

    def __init__(self, lower_bounds, upper_bounds):
        self.name = "uniform distribution"
        lower_bounds = _numpy.asarray(lower_bounds, dtype=_numpy.float64)
        upper_bounds = _numpy.asarray(upper_bounds, dtype=_numpy.float64)
        lower_bounds = lower_bounds.reshape(lower_bounds.size, 1).copy()
        upper_bounds = upper_bounds.reshape(upper_bounds.size, 1).copy()
        self.dimensions = int(lower_bounds.size)
        self.update_bounds(lower_bounds, upper_bounds)

    @staticmethod
    def create_default(dimensions: int) -> "Uniform":
        return Uniform(
            _numpy.random.rand(dimensions, 1) * 5 - 10,
            _numpy.random.rand(dimensions, 1) * 5 + 10,
        )


class CompositeDistribution(_AbstractDistribution):
    """Block concatenation of distributions on disjoint coordinate ranges
    (base.py:803-1002)."""

    def __init__(
        self,
        list_of_distributions: List[_AbstractDistribution] = None,
        lower_bounds=None,
        upper_bounds=None,
    ):
        self.name = "composite distribution"
        self.separate_distributions = list_of_distributions
        sizes = [int(dist.dimensions) for dist in list_of_distributions]
        self.enumerated_dimensions = _numpy.array(sizes, dtype=float)
        self.dimensions = int(sum(sizes))
        self.enumerated_dimensions_cumulative = _numpy.cumsum(sizes, dtype="int")[:-1]
        self.lower_bounds = lower_bounds
        self.upper_bounds = upper_bounds


class AdditiveDistribution(_AbstractDistribution):
    """Sum of misfits over the same coordinates (base.py:1005-1202).

    As in the reference, bounds of the *direct* children are collapsed into this
    object (max of lowers, min of uppers; base.py:1061-1101) and only those
    collapsed bounds drive the reflection during a trajectory (base.py:1111-1142)."""

    def __init__(
        self,
        list_of_distributions: List[_AbstractDistribution],
        lower_bounds=None,
        upper_bounds=None,
    ):
        self.name = "additive distribution"
        self.dimensions = int(list_of_distributions[0].dimensions)
        self.separate_distributions = list_of_distributions
        for dist in list_of_distributions:
            assert dist.dimensions == self.dimensions
        self.lower_bounds = lower_bounds
        self.upper_bounds = upper_bounds
        self.collapse_bounds()

    def collapse_bounds(self):
        for dist in self.separate_distributions:
            assert dist.dimensions == self.dimensions
            for attr, pick in (("lower_bounds", _numpy.maximum), ("upper_bounds", _numpy.minimum)):
                theirs = getattr(dist, attr)
                if theirs is None:
                    continue
                assert theirs.shape == (self.dimensions, 1)
                mine = getattr(self, attr)
                setattr(self, attr, theirs if mine is None else pick(mine, theirs))

    def add_distribution(self, distribution: _AbstractDistribution):
        assert distribution.dimensions == self.dimensions
        self.separate_distributions.append(distribution)
        self.collapse_bounds()
        self.__dict__.pop("_hmcb_evaluator", None)


class BayesRule(AdditiveDistribution):
    """Unnormalised Bayes' rule: prior(s) + likelihood (base.py:1205)."""

    pass
