"""Synthetic HMC problems of the shapes named in BASELINE.json.

Every builder is deterministic (seeded) and returns a ``Workload``: posterior and mass
matrix objects of this package (same constructors as the reference's), the sampler
settings and a seeded batch of initial models.  ``bench.py`` and the full-size property
tests use them; the ``scale`` arguments shrink a problem for fast tests.

The straight-ray tomography operator has no module in the reference (README lists it, the
code base only has LinearMatrix fed with a user-built sparse G); ``straight_ray_matrix``
is the deterministic generator of that G.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict

import numpy as np

from hmclab_b200 import Distributions as D
from hmclab_b200 import MassMatrices as M


@dataclass
class Workload:
    name: str
    posterior: Any
    mass_matrix: Any
    chains: int
    integrator: str
    amount_of_steps: int
    stepsize: float
    initial_models: np.ndarray  # [chains, dims]
    description: str
    extra: Dict[str, Any] = field(default_factory=dict)

    @property
    def dims(self) -> int:
        return int(self.posterior.dimensions)

    @property
    def grads_per_proposal(self) -> int:
        return self.amount_of_steps * {"lf": 1, "3s": 3, "4s": 4}[self.integrator]


# --------------------------------------------------------------------------- config 1/2 --

def dense_small(chains: int = 1) -> Workload:
    """configs[0]: 100 params x 200 data, Normal prior, leapfrog (test_linear_dense_simple shape)."""
    rng = np.random.default_rng(0)
    G = rng.normal(size=(200, 100))
    d = rng.normal(size=(200, 1))
    post = D.BayesRule([D.Normal(np.zeros((100, 1)), 1.0), D.LinearMatrix(G, d, 2.0)])
    q0 = np.zeros((chains, 100)) if chains == 1 else rng.normal(size=(chains, 100)) * 0.1
    return Workload("dense_small", post, M.Unit(100), chains, "lf", 10, 0.01, q0,
                    "LinearMatrix dense 100 params x 200 data (premultiplied GtG), Normal prior, lf L=10",
                    extra={"flops_per_grad": 2.0 * 100 * 100, "form": "premult"})


def normal_iid(dims: int = 1000, chains: int = 4096) -> Workload:
    """configs[1]: standard-normal posterior, Unit mass, leapfrog (integrator/accept path)."""
    rng = np.random.default_rng(1)
    post = D.Normal(np.zeros((dims, 1)), 1.0)
    q0 = rng.normal(size=(chains, dims))
    # fp64 instructions per gradient evaluation per chain (estimate): 4 per coordinate and leapfrog
    # step since the two updates are one FMA each (6 in exact mode), plus ~63 per coordinate and
    # proposal for Box-Muller, energies and reductions (executed-instruction counts of
    # profiles/ncu_fused_priors_r01.txt), spread over L=10
    return Workload("normal_iid", post, M.Unit(dims), chains, "lf", 10, 0.05, q0,
                    f"StandardNormal {dims}-dim posterior, {chains} chains, lf L=10, Unit mass",
                    extra={"fp64_ops_per_grad": dims * (4.0 + 63.0 / 10)})


# ----------------------------------------------------------------------------- config 3 --

def dense_large(dims: int = 2000, data: int = 10000, chains: int = 8192,
                premultiplication: bool = False) -> Workload:
    """configs[2]: dense LinearMatrix, vector variance, Normal prior, Diagonal mass, 4-stage L=5.

    Built through the public ``LinearMatrix`` class, so G and d carry the reference's
    float32 rounding and are then used in float64 arithmetic."""
    rng = np.random.default_rng(2)
    G = rng.normal(size=(data, dims)) / np.sqrt(data)
    m_true = rng.normal(size=(dims, 1))
    var = rng.uniform(0.5, 1.5, size=(data, 1))
    d = G @ m_true + np.sqrt(var) * rng.normal(size=(data, 1))
    lik = D.LinearMatrix(G, d, var, premultiplication=premultiplication)
    post = D.BayesRule([D.Normal(np.zeros((dims, 1)), 1.0), lik])
    mass = M.Diagonal(rng.uniform(0.5, 2.0, size=(dims, 1)))
    q0 = rng.normal(size=(chains, dims))
    form = "premultiplied GtG" if premultiplication else "direct G, G^T"
    return Workload("dense_large", post, mass, chains, "4s", 5, 0.05, q0,
                    f"LinearMatrix dense fp64 {dims} params x {data} data ({form}), {chains} chains, "
                    "4s L=5, Diagonal mass",
                    extra={"flops_per_grad": (2.0 * dims * dims) if premultiplication else 4.0 * data * dims,
                           "form": "premult" if premultiplication else "direct"})


# ----------------------------------------------------------------------------- config 4 --

def straight_ray_matrix(nx: int, ny: int, rays: int, seed: int = 3):
    """CSR matrix [rays x nx*ny] of path lengths of straight rays through a grid of unit
    cells; ray end points are seeded-uniform on the boundary of the [0,nx]x[0,ny] box."""
    import scipy.sparse as sp

    rng = np.random.default_rng(seed)

    def boundary_points(n):
        side = rng.integers(0, 4, size=n)
        t = rng.uniform(0.0, 1.0, size=n)
        x = np.where(side == 0, t * nx, np.where(side == 1, nx, np.where(side == 2, t * nx, 0.0)))
        y = np.where(side == 0, 0.0, np.where(side == 1, t * ny, np.where(side == 2, ny, t * ny)))
        return x, y, side

    x0, y0, s0 = boundary_points(rays)
    x1, y1, s1 = boundary_points(rays)
    same = s0 == s1  # both ends on one side: move the second end to the opposite side
    x1 = np.where(same & (s0 == 1), 0.0, np.where(same & (s0 == 3), float(nx), x1))
    y1 = np.where(same & (s0 == 0), float(ny), np.where(same & (s0 == 2), 0.0, y1))
    dx, dy = x1 - x0, y1 - y0
    length = np.hypot(dx, dy)
    with np.errstate(divide="ignore", invalid="ignore"):
        tx = (np.arange(nx + 1)[None, :] - x0[:, None]) / dx[:, None]
        ty = (np.arange(ny + 1)[None, :] - y0[:, None]) / dy[:, None]
    t = np.concatenate([tx, ty, np.zeros((rays, 1)), np.ones((rays, 1))], axis=1)
    t = np.where(np.isfinite(t), t, 0.0)
    t = np.sort(np.clip(t, 0.0, 1.0), axis=1)
    seg = np.diff(t, axis=1)
    mid = 0.5 * (t[:, 1:] + t[:, :-1])
    cx = np.clip(np.floor(x0[:, None] + mid * dx[:, None]).astype(np.int64), 0, nx - 1)
    cy = np.clip(np.floor(y0[:, None] + mid * dy[:, None]).astype(np.int64), 0, ny - 1)
    vals = seg * length[:, None]
    keep = vals > 1e-12
    rows = np.broadcast_to(np.arange(rays)[:, None], vals.shape)[keep]
    cols = (cy * nx + cx)[keep]
    G = sp.coo_matrix((vals[keep], (rows, cols)), shape=(rays, nx * ny)).tocsr()
    G.sum_duplicates()
    G.sort_indices()
    return G


def tomography(nx: int = 100, ny: int = 100, rays: int = 50000, chains: int = 8192) -> Workload:
    """configs[3]: straight-ray tomography, CSR G, Laplace prior, leapfrog, direct form."""
    rng = np.random.default_rng(3)
    G = straight_ray_matrix(nx, ny, rays, seed=3)
    dims = nx * ny
    s0 = np.full((dims, 1), 0.5)  # background slowness
    yy, xx = np.mgrid[0:ny, 0:nx]
    anomaly = 0.1 * np.exp(-((xx - 0.6 * nx) ** 2 + (yy - 0.4 * ny) ** 2) / (0.02 * nx * ny))
    s_true = s0 + anomaly.reshape(dims, 1)
    sigma2 = 0.25
    d = G @ s_true + np.sqrt(sigma2) * rng.normal(size=(rays, 1))
    lik = D.LinearMatrix(G, d, sigma2, premultiplication=False)
    prior = D.Laplace(s0, np.full((dims, 1), 0.1))
    post = D.BayesRule([prior, lik])
    # chains start near the true model (as after burn-in): far from it the Laplace kinks make
    # the leapfrog energy error of a 10 000-dimensional chain enormous and nothing is accepted
    q0 = s_true[:, 0][None, :] + 0.002 * rng.normal(size=(chains, dims))
    return Workload("tomography", post, M.Unit(dims), chains, "lf", 10, 0.001, q0,
                    f"Straight-ray tomography {nx}x{ny} grid, {rays} rays, CSR G (nnz={G.nnz}), "
                    f"Laplace prior, {chains} chains, lf L=10",
                    extra={"nnz": int(G.nnz), "rays": rays})


# ----------------------------------------------------------------------------- config 5 --

def source_location(events: int = 16, stations: int = 30, chains: int = 8192) -> Workload:
    """configs[4]: SourceLocation3D, fixed velocity, Uniform box prior directly in BayesRule
    (reflection active), Diagonal mass, leapfrog."""
    rng = np.random.default_rng(4)
    sx = rng.uniform(-10, 30, size=(1, stations))
    sy = rng.uniform(-10, 30, size=(1, stations))
    sz = np.zeros((1, stations))
    ex, ey = rng.uniform(0, 20, size=(events, 1)), rng.uniform(0, 20, size=(events, 1))
    ez, eT = rng.uniform(0, 10, size=(events, 1)), rng.uniform(0, 10, size=(events, 1))
    v = 3.0
    tt = D.SourceLocation3D.forward(ex, ey, ez, eT, v, sx, sy, sz)
    std = 0.1 * np.ones_like(tt)
    tobs = tt + std * rng.normal(size=tt.shape)
    lik = D.SourceLocation3D(sx, sy, sz, tobs, std, infer_velocity=False, medium_velocity=v)
    lo = np.tile(np.array([[-10.0], [-10.0], [0.0], [-5.0]]), (events, 1))
    hi = np.tile(np.array([[30.0], [30.0], [20.0], [15.0]]), (events, 1))
    post = D.BayesRule([D.Uniform(lo, hi), lik])
    dims = 4 * events
    mass = M.Diagonal(rng.uniform(0.5, 2.0, size=(dims, 1)))
    truth = np.hstack([ex, ey, ez, eT]).reshape(-1)
    q0 = np.clip(truth[None, :] + 0.05 * rng.normal(size=(chains, dims)), lo[:, 0] + 1e-3, hi[:, 0] - 1e-3)
    # fp64 instructions per gradient evaluation per chain: 24 per event-station pair in the
    # gradient loop (profiles/ncu_fused_srcloc_r01.txt) + 1/10 of the 16-op misfit loop
    return Workload("source_location", post, mass, chains, "lf", 10, 0.004, q0,
                    f"SourceLocation3D {events} events x {stations} stations ({dims} params), Uniform box "
                    f"prior, {chains} chains per GPU, lf L=10, Diagonal mass",
                    extra={"fp64_ops_per_grad": events * stations * (24.0 + 1.6) + 12.0 * dims})


BUILDERS = {
    "dense_small": dense_small,
    "normal_iid": normal_iid,
    "dense_large": dense_large,
    "dense_large_premult": lambda **kw: dense_large(premultiplication=True, **kw),
    "tomography": tomography,
    "source_location": source_location,
}
