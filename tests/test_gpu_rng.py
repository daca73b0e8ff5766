"""The on-device random streams against their CPU replica (oracle/device_rng.py): stream
definition (Philox counters/keys), accuracy of the specialised Box-Muller transcendentals,
and the use of the uniforms for the step size and the Metropolis test."""
import numpy as np
import pytest

from oracle import device_rng

pytestmark = pytest.mark.gpu


def _free_flight_engine(d, C, steps=1):
    from hmclab_b200._engine import Engine

    tiny = 1e-300   # inverse variance: the gradient is exactly negligible next to the momentum
    plan = {"dims": d, "terms": [{"kind": "normal", "offset": 0, "len": d, "a": np.zeros(d),
                                  "b": np.full(d, tiny), "const": 0.0}],
            "checks": [], "likelihood": None, "reflect_lb": None, "reflect_ub": None}
    return Engine(plan, {"kind": "unit", "dims": d}, C, integrator="lf", amount_of_steps=steps)


@pytest.mark.parametrize("d,C,chain_offset", [(1000, 64, 0), (33, 257, 4096), (5000, 8, 7)])
def test_normals_match_the_replica(d, C, chain_offset):
    import torch

    seed = 0x1234_5678_9ABC_DEF1
    eng = _free_flight_engine(d, C)      # d = 5000 exercises the staged path's draw kernel
    q = torch.zeros(C, d, dtype=torch.float64, device="cuda")
    x = eng.misfit(q)
    K = 3
    pp = torch.zeros(K, C, d, dtype=torch.float64, device="cuda")
    qp = torch.zeros(K, C, d, dtype=torch.float64, device="cuda")
    eng.run_block(q, x, K, stepsize=1.0, randomize_stepsize=True, seed=seed, chain_offset=chain_offset,
                  proposal_offset=10, out_q_prop=qp, out_p_prop=pp)
    z_dev, q_dev = pp.cpu().numpy(), qp.cpu().numpy()
    for k in range(K):
        z_ref = device_rng.normals(seed, C, 10 + k, d, chain_offset)
        assert np.max(np.abs(z_dev[k] - z_ref)) < 5e-15        # a few ulp of a value <= 8.6
        # step-size factor: q1 - q0 = eps * z with eps = u_step * 1.0
        u_step, _ = device_rng.uniforms(seed, C, 10 + k, chain_offset)
        start = np.zeros((C, d)) if k == 0 else None
        if start is not None:
            ratio = q_dev[k] / z_dev[k]
            assert np.max(np.abs(ratio - u_step[:, None])) < 1e-12
    assert np.all((z_dev != 0).mean(axis=(1, 2)) > 0.999)


def test_metropolis_test_uses_the_replica_uniform():
    import torch

    from hmclab_b200 import workloads
    from hmclab_b200._engine import Engine
    from hmclab_b200._lowering import describe, describe_mass, flatten

    w = workloads.normal_iid(dims=200, chains=512)
    eng = Engine(flatten(describe(w.posterior)), describe_mass(w.mass_matrix), w.chains,
                 integrator="lf", amount_of_steps=10)
    q = torch.as_tensor(w.initial_models).cuda().contiguous()
    x = eng.misfit(q)
    K, seed = 6, 99
    acc = torch.zeros(K, w.chains, dtype=torch.uint8, device="cuda")
    h0 = torch.zeros(K, w.chains, dtype=torch.float64, device="cuda")
    h1 = torch.zeros(K, w.chains, dtype=torch.float64, device="cuda")
    eng.run_block(q, x, K, stepsize=0.45, seed=seed, out_accept=acc, out_h0=h0, out_h1=h1)
    acc, h0, h1 = acc.cpu().numpy().astype(bool), h0.cpu().numpy(), h1.cpu().numpy()
    assert 0.2 < acc.mean() < 0.95
    for k in range(K):
        _, u_acc = device_rng.uniforms(seed, w.chains, k)
        expect = np.exp(h0[k] - h1[k]) > u_acc
        margin = np.abs(np.exp(h0[k] - h1[k]) - u_acc)
        bad = (expect != acc[k]) & (margin > 1e-12)
        assert not bad.any()
