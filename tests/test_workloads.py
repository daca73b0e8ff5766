"""Synthetic problems of the BASELINE.json shapes (scaled down) and the lowering they feed."""
import numpy as np

from hmclab_b200 import workloads
from hmclab_b200._lowering import describe, describe_mass, flatten


def test_straight_ray_matrix_rows_are_ray_lengths():
    G = workloads.straight_ray_matrix(20, 15, 400, seed=3)
    assert G.shape == (400, 300) and G.nnz > 0
    assert G.data.min() > 0 and G.data.max() <= np.sqrt(2) + 1e-12   # chord of a unit cell
    lengths = np.asarray(G.sum(axis=1)).ravel()
    assert lengths.min() > 0 and lengths.max() <= np.hypot(20, 15) + 1e-9
    G2 = workloads.straight_ray_matrix(20, 15, 400, seed=3)
    assert (G != G2).nnz == 0                                         # deterministic
    # a straight ray crosses at most nx + ny - 1 cells
    assert np.diff(G.indptr).max() <= 20 + 15


def test_workloads_lower_to_the_expected_plans():
    w = workloads.normal_iid(dims=50, chains=4)
    plan = flatten(describe(w.posterior))
    assert plan["likelihood"] is None and [t["kind"] for t in plan["terms"]] == ["normal"]
    w = workloads.dense_small(chains=2)
    plan = flatten(describe(w.posterior))
    assert plan["likelihood"]["kind"] == "linear_dense" and plan["likelihood"]["premult"]
    w = workloads.dense_large(dims=32, data=80, chains=4)
    plan = flatten(describe(w.posterior))
    assert not plan["likelihood"]["premult"] and plan["likelihood"]["G"].shape == (80, 32)
    assert describe_mass(w.mass_matrix)["kind"] == "diagonal"
    # the public LinearMatrix class rounds G to float32 (reference quirk)
    assert np.array_equal(plan["likelihood"]["G"], plan["likelihood"]["G"].astype(np.float32))
    w = workloads.tomography(nx=12, ny=10, rays=300, chains=4)
    plan = flatten(describe(w.posterior))
    assert plan["likelihood"]["kind"] == "linear_csr" and not plan["likelihood"]["premult"]
    assert plan["likelihood"]["indices"].dtype == np.int32
    assert [t["kind"] for t in plan["terms"]] == ["laplace"]
    w = workloads.source_location(events=3, stations=5, chains=4)
    plan = flatten(describe(w.posterior))
    assert plan["likelihood"]["kind"] == "srcloc3d" and plan["reflect_lb"] is not None
    assert w.grads_per_proposal == 10 and w.initial_models.shape == (4, 12)
