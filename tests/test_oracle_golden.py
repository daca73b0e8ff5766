"""Pins oracle/hmc_oracle.py (and the mirror constructors + lowering that feed it) to
outputs of the unmodified reference stored in tests/golden/*.npz."""
import numpy as np
import pytest

import cases
from helpers import build_mirror, load_golden, rel_err, tol_for
from hmclab_b200._lowering import describe, describe_mass
from oracle import hmc_oracle as oracle

TOL = 1e-13  # same numpy/BLAS calls in the same order: agreement is at rounding level


@pytest.mark.parametrize("name", cases.CASES)
def test_oracle_reproduces_reference(name):
    inp, ref = load_golden(name)
    s = cases.SETTINGS[name]
    post, mass = build_mirror(name, inp)
    tree, mtree = describe(post), describe_mass(mass)
    with np.errstate(all="ignore"):
        got = oracle.run_chains(
            tree, mtree, q0=inp["q0"], z=inp["z"], u_step=inp["u_step"],
            u_acc=inp["u_acc"], integrator=s["integrator"], steps=s["steps"],
            stepsize=s["stepsize"], randomize=s["randomize"], record_trace=True)
    assert np.array_equal(got["accept"], ref["accept"]), "accept/reject sequence differs"
    for key in ("H0", "H1", "q_prop", "p_prop", "samples", "trace_q", "trace_g"):
        assert rel_err(got[key], ref[key]) < tol_for(name, TOL), key


@pytest.mark.parametrize("name", cases.CASES)
def test_oracle_misfit_gradient_contract(name):
    inp, ref = load_golden(name)
    post, _ = build_mirror(name, inp)
    tree = describe(post)
    d = tree["dims"]
    with np.errstate(all="ignore"):
        for i, point in enumerate(ref["probe_points"]):
            m = point.reshape(d, 1).copy()
            assert rel_err(oracle.misfit(tree, m), ref["probe_misfit"][i]) < tol_for(name, TOL)
            assert rel_err(oracle.gradient(tree, m)[:, 0], ref["probe_gradient"][i]) < tol_for(name, TOL)


def test_golden_cases_exercise_both_decisions_and_bounds():
    mixed = 0
    nonfinite = 0
    for name in cases.CASES:
        _, ref = load_golden(name)
        mixed += 0 < ref["accept"].mean() < 1
        nonfinite += (~np.isfinite(ref["H1"])).any()
    assert mixed >= 10 and nonfinite >= 2


AUTOTUNE_CASES = ("normal_unit_lf", "dense_direct_4s", "srcloc_fixed_v", "sparse_laplace_lf")


@pytest.mark.parametrize("name", AUTOTUNE_CASES)
def test_oracle_autotuning_reproduces_reference(name):
    """Reference chains run with autotuning=True (tests/golden/autotuned_runs.npz)."""
    import os

    from helpers import GOLDEN_DIR

    gold = np.load(os.path.join(GOLDEN_DIR, "autotuned_runs.npz"))
    inp, _ = load_golden(name)
    s = cases.SETTINGS[name]
    post, mass = build_mirror(name, inp)
    tree, mtree = describe(post), describe_mass(mass)
    with np.errstate(all="ignore"):
        got = oracle.run_chains(
            tree, mtree, q0=inp["q0"], z=inp["z"], u_step=inp["u_step"], u_acc=inp["u_acc"],
            integrator=s["integrator"], steps=s["steps"], stepsize=s["stepsize"],
            randomize=s["randomize"], autotuning=True, target_acceptance_rate=0.65, learning_rate=0.75)
    assert rel_err(got["samples"], gold[f"{name}__samples"]) < TOL
    assert rel_err(got["final_stepsize"], gold[f"{name}__final_stepsize"]) < TOL
    ref_steps = gold[f"{name}__stepsizes"]
    known = ~np.isnan(ref_steps)       # the reference drops the last entry of its history
    assert rel_err(got["stepsizes"][known], ref_steps[known]) < TOL
    assert (gold[f"{name}__final_stepsize"] != s["stepsize"]).all()


RWMH_CASES = {"normal_bounded": ("scalar", False), "dense_premult_cfg1": ("vector", False),
              "srcloc_fixed_v": ("scalar", True), "sparse_laplace_lf": ("vector", True)}


@pytest.mark.parametrize("name", list(RWMH_CASES))
def test_oracle_rwmh_reproduces_reference(name):
    """Reference RWMH chains with replayed draws (tests/golden/rwmh_runs.npz)."""
    import os

    from helpers import GOLDEN_DIR

    gold = np.load(os.path.join(GOLDEN_DIR, "rwmh_runs.npz"))
    kind, tune = RWMH_CASES[name]
    inp, _ = load_golden(name)
    post, _ = build_mirror(name, inp)
    tree = describe(post)
    K, C, d = inp["z"].shape
    vector = gold[f"{name}__step_vector"][:, 0] if kind == "vector" else None
    scalar = 1.0 if kind == "vector" else float(gold[f"{name}__step_scalar"])
    for c in range(C):
        draws = oracle.ReplayDraws(inp["z"][:, c], inp["u_step"][:, c], inp["u_acc"][:, c])
        with np.errstate(all="ignore"):
            got = oracle.run_chain_rwmh(tree, stepsize=scalar, step_vector=vector, q0=inp["q0"][c],
                                        proposals=K, draws=draws, autotuning=tune)
        assert rel_err(got["samples"], gold[f"{name}__samples"][:, c]) < TOL
        assert got["accept"].sum() == gold[f"{name}__accepted"][c]
        if tune:
            assert rel_err(got["final_stepsize"], gold[f"{name}__final_stepsize"][c]) < TOL
