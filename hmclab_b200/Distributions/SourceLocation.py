"""Earthquake source location in a homogeneous medium, 3-D (host-side mirror).

Mirror of ``hmclab.Distributions.SourceLocation3D``
(hmclab/Distributions/SourceLocation.py:369-713): parameters are interleaved
``x, y, z, T`` per event, optionally followed by one medium velocity; data are
travel times ``T + |x - r| / v`` at every station, with NaN marking a missing
pick (the reference sums with ``nansum``; SourceLocation.py:482-540).

The object only validates and stores the station geometry and observations; the
misfit and its analytic gradient are evaluated by the fused CUDA kernel.
"""
from __future__ import annotations

import math as _math

import numpy as _numpy

from hmclab_b200.Distributions.base import _AbstractDistribution


def _row(values: _numpy.ndarray) -> _numpy.ndarray:
    values = _numpy.array(values, dtype=_numpy.float64)
    return values.reshape(1, values.size)


def _events_by_stations(values: _numpy.ndarray, shape, what: str) -> _numpy.ndarray:
    if values.shape == shape:
        return values
    if values.T.shape == shape:
        return values.T
    raise AssertionError(f"Wrong shape for the {what}, not sure what to do.")


class SourceLocation3D(_AbstractDistribution):
    name = "Earthquake source location in 3D"

    def __init__(
        self,
        receiver_array_x,
        receiver_array_y,
        receiver_array_z,
        observed_data,
        data_std,
        infer_velocity: bool = True,
        medium_velocity=None,
    ):
        self.receiver_array_x = _row(receiver_array_x)
        self.receiver_array_y = _row(receiver_array_y)
        self.receiver_array_z = _row(receiver_array_z)
        self.number_of_stations = int(self.receiver_array_z.size)
        assert (
            self.receiver_array_x.size == self.number_of_stations
            and self.receiver_array_y.size == self.number_of_stations
        ), "Receiver coordinate arrays differ in length."

        self.infer_velocity = bool(infer_velocity)
        self.medium_velocity = None
        if not self.infer_velocity:
            assert medium_velocity is not None
            self.medium_velocity = medium_velocity

        observed_data = _numpy.array(observed_data, dtype=_numpy.float64)
        assert observed_data.size % self.number_of_stations == 0
        self.number_of_events = int(observed_data.size // self.number_of_stations)
        self.number_of_datums = self.number_of_events * self.number_of_stations
        shape = (self.number_of_events, self.number_of_stations)
        self.observed_data = _events_by_stations(observed_data, shape, "observed data")

        if type(data_std) is float:
            data_std = _numpy.ones(shape) * data_std
        data_std = _numpy.array(data_std, dtype=_numpy.float64)
        self.data_std = _events_by_stations(data_std, shape, "data uncertainty")

        self.dimensions = self.number_of_events * 4 + int(self.infer_velocity)

    @staticmethod
    def forward(x, y, z, T, v, receiver_array_x, receiver_array_y, receiver_array_z):
        """Synthetic travel times for building test problems (setup-time helper,
        SourceLocation.py:547-563); not used inside sampling."""
        return (
            T
            + (
                (x - receiver_array_x) ** 2.0
                + (y - receiver_array_y) ** 2.0
                + (z - receiver_array_z) ** 2.0
            )
            ** 0.5
            / v
        )

    @staticmethod
    def create_default(dimensions, seed=127, stations=3):
        events = _math.floor(dimensions / 4)
        if dimensions < 4 or dimensions % 4 not in (0, 1):
            raise ValueError("SourceLocation3D needs 4*events (+1) parameters.")
        infer_velocity = dimensions % 4 == 1
        rng = _numpy.random.RandomState(seed)
        sx = rng.rand(1, stations) * 40 - 10
        sy = rng.rand(1, stations) * 40 - 10
        sz = _numpy.zeros_like(sx)
        x = rng.rand(events, 1) * 20
        y = rng.rand(events, 1) * 20
        z = rng.rand(events, 1) * 10
        T = rng.rand(events, 1) * 10
        v = rng.rand(1, 1) * 3 + 1
        data = SourceLocation3D.forward(x, y, z, T, v, sx, sy, sz)
        std = 1.0 * rng.randn(*data.shape)
        data = data + std * rng.randn(*std.shape)
        return SourceLocation3D(
            sx, sy, sz, data, std,
            infer_velocity=infer_velocity,
            medium_velocity=None if infer_velocity else v,
        )


class SourceLocation2D(_AbstractDistribution):
    """Mirror of ``hmclab.Distributions.SourceLocation2D`` (SourceLocation.py:11-366):
    parameters interleaved ``x, z, T`` per event, optionally followed by one medium velocity.
    Evaluated by the same kernel as the 3-D problem (stations and sources in the plane y = 0)."""

    name = "Earthquake source location in 2D"

    def __init__(self, receiver_array_x, receiver_array_z, observed_data, data_std,
                 infer_velocity: bool = True, medium_velocity=None):
        self.receiver_array_x = _row(receiver_array_x)
        self.receiver_array_z = _row(receiver_array_z)
        self.number_of_stations = int(self.receiver_array_x.size)
        assert self.receiver_array_z.size == self.number_of_stations, (
            "Receiver coordinate arrays differ in length.")
        self.infer_velocity = bool(infer_velocity)
        self.medium_velocity = None
        if not self.infer_velocity:
            assert medium_velocity is not None
            self.medium_velocity = medium_velocity
        observed_data = _numpy.array(observed_data, dtype=_numpy.float64)
        assert observed_data.size % self.number_of_stations == 0
        self.number_of_events = int(observed_data.size // self.number_of_stations)
        shape = (self.number_of_events, self.number_of_stations)
        self.observed_data = _events_by_stations(observed_data, shape, "observed data")
        if type(data_std) is float:
            data_std = _numpy.ones(shape) * data_std
        data_std = _numpy.array(data_std, dtype=_numpy.float64)
        self.data_std = _events_by_stations(data_std, shape, "data uncertainty")
        self.dimensions = self.number_of_events * 3 + int(self.infer_velocity)

    @staticmethod
    def forward(x, z, T, v, receiver_array_x, receiver_array_z):
        return T + ((x - receiver_array_x) ** 2.0 + (z - receiver_array_z) ** 2.0) ** 0.5 / v
