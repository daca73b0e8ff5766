"""Distributions understood by the batched B200 HMC engine.

Same names and constructor arguments as ``hmclab.Distributions`` for the classes on
the accelerated path; genuine ``hmclab`` objects of these
classes are accepted by the sampler as well (they are read by attribute name).
"""
from hmclab_b200.Distributions.base import (
    AdditiveDistribution,
    BayesRule,
    CompositeDistribution,
    Laplace,
    Normal,
    Uniform,
    _AbstractDistribution,
)
from hmclab_b200.Distributions.LinearMatrix import LinearMatrix
from hmclab_b200.Distributions.SourceLocation import SourceLocation2D, SourceLocation3D

__all__ = [
    "_AbstractDistribution",
    "Normal",
    "Laplace",
    "Uniform",
    "CompositeDistribution",
    "AdditiveDistribution",
    "BayesRule",
    "LinearMatrix",
    "SourceLocation2D",
    "SourceLocation3D",
]
