"""GPU parity: the CUDA engine (through the C ABI) against the reference's own outputs
stored in tests/golden/*.npz, on identical inputs with injected momenta, step-size factors
and acceptance uniforms.

Bar (BASELINE.json north_star): gradients and full trajectories within 1e-10 relative in
fp64; the accept/reject sequence bit-identical.
"""
import numpy as np
import pytest

import cases
from helpers import FLOAT32_BLAS_CASES, build_mirror, load_golden, rel_err, tol_for
from hmclab_b200._lowering import describe, describe_mass, flatten

pytestmark = pytest.mark.gpu

TOL = 1e-10


def _engine(name, inp, chains):
    import torch
    from hmclab_b200._engine import Engine

    s = cases.SETTINGS[name]
    post, mass = build_mirror(name, inp)
    plan, mplan = flatten(describe(post)), describe_mass(mass)
    eng = Engine(plan, mplan, chains, integrator=s["integrator"], amount_of_steps=s["steps"])
    return eng, torch


def _dev(torch, a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64).cuda()


@pytest.mark.parametrize("name", cases.CASES)
def test_trajectories_and_decisions_match_reference(name):
    inp, ref = load_golden(name)
    s = cases.SETTINGS[name]
    K, C, d = inp["z"].shape
    eng, torch = _engine(name, inp, C)
    G = eng.grads_per_proposal
    q = _dev(torch, inp["q0"])
    x = eng.misfit(q)
    out = dict(
        out_samples=torch.zeros(K, C, d + 1, dtype=torch.float64, device="cuda"),
        out_accept=torch.zeros(K, C, dtype=torch.uint8, device="cuda"),
        out_h0=torch.zeros(K, C, dtype=torch.float64, device="cuda"),
        out_h1=torch.zeros(K, C, dtype=torch.float64, device="cuda"),
        accepted_total=torch.zeros(C, dtype=torch.int32, device="cuda"),
        out_q_prop=torch.zeros(K, C, d, dtype=torch.float64, device="cuda"),
        out_p_prop=torch.zeros(K, C, d, dtype=torch.float64, device="cuda"),
        trace_q=torch.zeros(K, G, C, d, dtype=torch.float64, device="cuda"),
        trace_g=torch.zeros(K, G, C, d, dtype=torch.float64, device="cuda"),
    )
    eng.run_block(q, x, K, stepsize=s["stepsize"], randomize_stepsize=s["randomize"],
                  z=_dev(torch, inp["z"]), u_step=_dev(torch, inp["u_step"]),
                  u_accept=_dev(torch, inp["u_acc"]), **out)
    torch.cuda.synchronize()
    got = {k: v.cpu().numpy() for k, v in out.items()}
    mism = got["out_accept"].astype(bool) != ref["accept"]
    if mism.any():
        margin = np.abs(np.exp(ref["H0"] - ref["H1"]) - inp["u_acc"])[mism]
        pytest.fail(f"accept/reject differs at {np.argwhere(mism).tolist()} (decision margins {margin})")
    for key, rk in (("trace_q", "trace_q"), ("trace_g", "trace_g"), ("out_q_prop", "q_prop"),
                    ("out_p_prop", "p_prop"), ("out_h0", "H0"), ("out_h1", "H1"),
                    ("out_samples", "samples")):
        err = rel_err(got[key], ref[rk])
        assert err < tol_for(name, TOL), f"{key}: rel err {err:.3e}"
    assert np.array_equal(got["accepted_total"], ref["accept"].sum(axis=0))
    assert rel_err(q.cpu().numpy(), ref["samples"][-1, :, :d]) < tol_for(name, TOL)
    assert rel_err(x.cpu().numpy(), ref["samples"][-1, :, d]) < tol_for(name, TOL)
    if name in FLOAT32_BLAS_CASES:
        # the reference's float32 operator product is host dependent: 1e-10 against the same-host oracle
        from oracle import hmc_oracle as oracle

        post, mass = build_mirror(name, inp)
        with np.errstate(all="ignore"):
            here = oracle.run_chains(describe(post), describe_mass(mass), q0=inp["q0"], z=inp["z"],
                                     u_step=inp["u_step"], u_acc=inp["u_acc"], integrator=s["integrator"],
                                     steps=s["steps"], stepsize=s["stepsize"], randomize=s["randomize"])
        assert np.array_equal(got["out_accept"].astype(bool), here["accept"])
        for key, rk in (("out_q_prop", "q_prop"), ("out_p_prop", "p_prop"), ("out_h0", "H0"),
                        ("out_h1", "H1"), ("out_samples", "samples")):
            assert rel_err(got[key], here[rk]) < TOL, key


@pytest.mark.parametrize("name", cases.CASES)
def test_misfit_gradient_contract(name):
    inp, ref = load_golden(name)
    pts = ref["probe_points"]
    eng, torch = _engine(name, inp, pts.shape[0])
    q = _dev(torch, pts)
    x = eng.misfit(q).cpu().numpy()
    g = eng.gradient(q).cpu().numpy()
    assert rel_err(x, ref["probe_misfit"]) < tol_for(name, TOL)
    for i in range(pts.shape[0]):
        assert rel_err(g[i], ref["probe_gradient"][i]) < tol_for(name, TOL), i


@pytest.mark.parametrize("name", ["normal_unit_lf", "dense_direct_4s", "srcloc_fixed_v"])
def test_blocks_thinning_and_untraced_run_agree(name):
    """Two blocks with thinning and no debug outputs reproduce the one-block traced run."""
    inp, ref = load_golden(name)
    s = cases.SETTINGS[name]
    K, C, d = inp["z"].shape
    eng, torch = _engine(name, inp, C)
    q = _dev(torch, inp["q0"])
    x = eng.misfit(q)
    z, us, ua = _dev(torch, inp["z"]), _dev(torch, inp["u_step"]), _dev(torch, inp["u_acc"])
    thin, k1 = 2, 3
    acc = torch.zeros(C, dtype=torch.int32, device="cuda")
    rows = []
    for lo, hi in ((0, k1), (k1, K)):
        n = eng.stored_rows(hi - lo, thin, lo)
        buf = torch.zeros(n, C, d + 1, dtype=torch.float64, device="cuda")
        eng.run_block(q, x, hi - lo, stepsize=s["stepsize"], randomize_stepsize=s["randomize"],
                      thinning=thin, proposal_offset=lo, z=z[lo:hi].contiguous(),
                      u_step=us[lo:hi].contiguous(), u_accept=ua[lo:hi].contiguous(),
                      out_samples=buf, accepted_total=acc)
        rows.append(buf.cpu().numpy())
    got = np.concatenate(rows)
    assert rel_err(got, ref["samples"][::thin]) < TOL
    assert np.array_equal(acc.cpu().numpy(), ref["accept"].sum(axis=0))


def test_mass_matrix_and_reflection_entry_points():
    name = "srcloc_fixed_v"
    inp, ref = load_golden(name)
    K, C, d = inp["z"].shape
    eng, torch = _engine(name, inp, C)
    z = _dev(torch, inp["z"][0])
    m = inp["mass"][:, 0]
    p = eng.scale_momentum(z)
    assert np.array_equal(p.cpu().numpy(), np.sqrt(m)[None, :] * inp["z"][0])
    inv = 1.0 / m
    assert np.array_equal(eng.kinetic_gradient(p).cpu().numpy(), inv[None, :] * p.cpu().numpy())
    k_ref = 0.5 * np.sum(p.cpu().numpy() * (inv[None, :] * p.cpu().numpy()), axis=1)
    assert rel_err(eng.kinetic_energy(p).cpu().numpy(), k_ref) < 1e-14
    # one-shot reflection: below the lower bound, above the upper bound, far outside
    lo, hi = inp["lo"][:, 0], inp["hi"][:, 0]
    q = np.tile(0.5 * (lo + hi), (C, 1))
    q[0, 0] = lo[0] - 1.0
    q[1, 1] = hi[1] + 2.0
    q[2, 2] = lo[2] - 10 * (hi[2] - lo[2])
    pq = np.ones_like(q)
    qe, pe = q.copy(), pq.copy()
    low = qe < lo
    qe[low] += 2 * (np.broadcast_to(lo, q.shape)[low] - qe[low])
    pe[low] *= -1
    high = qe > hi
    qe[high] += 2 * (np.broadcast_to(hi, q.shape)[high] - qe[high])
    pe[high] *= -1
    qd, pd = _dev(torch, q), _dev(torch, pq)
    eng.reflect_(qd, pd)
    assert np.array_equal(qd.cpu().numpy(), qe) and np.array_equal(pd.cpu().numpy(), pe)


def test_engine_refuses_bad_configuration():
    import torch  # noqa: F401
    from hmclab_b200._engine import Engine, HmcbError

    inp, _ = load_golden("normal_bounded")
    post, mass = build_mirror("normal_bounded", inp)
    plan, mplan = flatten(describe(post)), describe_mass(mass)
    with pytest.raises(ValueError):
        Engine(plan, mplan, 4, integrator="rk4")
    with pytest.raises(HmcbError):
        Engine(plan, mplan, 4, amount_of_steps=0)
    bad = dict(mplan, dims=mplan["dims"] + 1)
    with pytest.raises(ValueError):
        Engine(plan, bad, 4)


@pytest.mark.parametrize("name", ["normal_unit_lf", "dense_direct_4s", "srcloc_fixed_v", "sparse_laplace_lf"])
def test_autotuned_chains_match_reference(name):
    """Per-chain step-size adaptation on the device vs reference chains run with
    autotuning=True and the same replayed draws (tests/golden/autotuned_runs.npz); in two
    blocks, so the adapted step sizes also have to survive the block boundary."""
    import os

    from helpers import GOLDEN_DIR

    gold = np.load(os.path.join(GOLDEN_DIR, "autotuned_runs.npz"))
    inp, _ = load_golden(name)
    s = cases.SETTINGS[name]
    K, C, d = inp["z"].shape
    eng, torch = _engine(name, inp, C)
    q = _dev(torch, inp["q0"])
    x = eng.misfit(q)
    z, us, ua = _dev(torch, inp["z"]), _dev(torch, inp["u_step"]), _dev(torch, inp["u_acc"])
    eps = torch.full((C,), s["stepsize"], dtype=torch.float64, device="cuda")
    samples, steps = [], []
    for lo, hi in ((0, 2), (2, K)):
        buf = torch.zeros(hi - lo, C, d + 1, dtype=torch.float64, device="cuda")
        hist = torch.zeros(hi - lo, C, dtype=torch.float64, device="cuda")
        eng.run_block(q, x, hi - lo, stepsize=s["stepsize"], randomize_stepsize=s["randomize"],
                      proposal_offset=lo, z=z[lo:hi].contiguous(), u_step=us[lo:hi].contiguous(),
                      u_accept=ua[lo:hi].contiguous(), out_samples=buf, stepsize_chain=eps,
                      autotune=True, target_acceptance_rate=0.65, learning_rate=0.75,
                      out_stepsize=hist)
        samples.append(buf.cpu().numpy())
        steps.append(hist.cpu().numpy())
    got, got_steps = np.concatenate(samples), np.concatenate(steps)
    ref = gold[f"{name}__samples"]
    assert np.array_equal(np.all(np.diff(got, axis=0) == 0, axis=2), np.all(np.diff(ref, axis=0) == 0, axis=2))
    assert rel_err(got, ref) < TOL
    assert rel_err(eps.cpu().numpy(), gold[f"{name}__final_stepsize"]) < TOL
    ref_steps = gold[f"{name}__stepsizes"]
    known = ~np.isnan(ref_steps)
    assert rel_err(got_steps[known], ref_steps[known]) < TOL


def test_clear_target_allows_a_new_configuration_on_the_same_engine():
    """hmcb_clear_target frees the device constants; the engine can be described again."""
    import ctypes as C

    import torch

    from hmclab_b200._engine import load_library

    lib = load_library()
    dp = C.POINTER(C.c_double)
    h = C.c_void_p()
    d, n = 5, 4
    assert lib.hmcb_create(0, n, d, C.byref(h)) == 0
    q = torch.linspace(-1, 1, n * d, dtype=torch.float64, device="cuda").reshape(n, d).contiguous()
    x = torch.empty(n, dtype=torch.float64, device="cuda")
    for scale in (1.0, 4.0):
        a = np.zeros(d)
        b = np.full(d, 1.0 / scale)
        assert lib.hmcb_add_prior(h, 0, 0, d, a.ctypes.data_as(dp), b.ctypes.data_as(dp), 0.0) == 0
        assert lib.hmcb_finalize(h) == 0
        assert lib.hmcb_finalize(h) != 0 and b"twice" in lib.hmcb_last_error()
        assert lib.hmcb_add_prior(h, 0, 0, d, a.ctypes.data_as(dp), b.ctypes.data_as(dp), 0.0) != 0
        assert lib.hmcb_misfit(h, q.data_ptr(), x.data_ptr(), None) == 0
        torch.cuda.synchronize()
        assert rel_err(x.cpu().numpy(), 0.5 * (q.cpu().numpy() ** 2).sum(axis=1) / scale) < 1e-14
        assert lib.hmcb_clear_target(h) == 0
        assert lib.hmcb_misfit(h, q.data_ptr(), x.data_ptr(), None) != 0    # not finalized any more
    # bad arguments are refused with a message, not dereferenced
    assert lib.hmcb_add_prior(h, 7, 0, d, a.ctypes.data_as(dp), b.ctypes.data_as(dp), 0.0) != 0
    assert lib.hmcb_add_prior(h, 0, 3, d, a.ctypes.data_as(dp), b.ctypes.data_as(dp), 0.0) != 0
    assert lib.hmcb_set_likelihood_srcloc3d(h, 2, 3, a.ctypes.data_as(dp), a.ctypes.data_as(dp),
                                            a.ctypes.data_as(dp), a.ctypes.data_as(dp),
                                            a.ctypes.data_as(dp), 0, 3.0) != 0   # dims != 4 * events
    assert b"4*events" in lib.hmcb_last_error()
    assert lib.hmcb_destroy(h) == 0


@pytest.mark.parametrize("name,kind,tune", [("normal_bounded", "scalar", False),
                                            ("dense_premult_cfg1", "vector", False),
                                            ("srcloc_fixed_v", "scalar", True),
                                            ("sparse_laplace_lf", "vector", True)])
def test_rwmh_chains_match_reference(name, kind, tune):
    """hmcb_run_block_rwmh vs reference RWMH chains with the same replayed draws."""
    import os

    from helpers import GOLDEN_DIR

    gold = np.load(os.path.join(GOLDEN_DIR, "rwmh_runs.npz"))
    inp, _ = load_golden(name)
    K, C, d = inp["z"].shape
    eng, torch = _engine(name, inp, C)
    q = _dev(torch, inp["q0"])
    x = eng.misfit(q)
    vector = _dev(torch, gold[f"{name}__step_vector"][:, 0]) if kind == "vector" else None
    scalar = 1.0 if kind == "vector" else float(gold[f"{name}__step_scalar"])
    samples = torch.zeros(K, C, d + 1, dtype=torch.float64, device="cuda")
    acc = torch.zeros(C, dtype=torch.int32, device="cuda")
    extra = {}
    if tune:
        extra = dict(stepsize_chain=torch.full((C,), scalar, dtype=torch.float64, device="cuda"),
                     autotune=True, target_acceptance_rate=0.65, learning_rate=0.75)
    eng.run_block_rwmh(q, x, K, stepsize=scalar, step_vector=vector, z=_dev(torch, inp["z"]),
                       u_accept=_dev(torch, inp["u_acc"]), out_samples=samples, accepted_total=acc,
                       **extra)
    got, ref = samples.cpu().numpy(), gold[f"{name}__samples"]
    assert np.array_equal(acc.cpu().numpy(), gold[f"{name}__accepted"])
    assert rel_err(got, ref) < TOL
    if tune:
        assert rel_err(extra["stepsize_chain"].cpu().numpy(), gold[f"{name}__final_stepsize"]) < TOL
