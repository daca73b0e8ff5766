"""Generate golden input/output vectors by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

For every case in ``cases.py`` the reference's own ``hmclab.Samplers.HMC`` object is
driven proposal by proposal (``_propose`` / ``_evaluate_acceptance``, Samplers.py:1463-
1492) with its random generator replaced by a replay object, so momenta, step-size
factors and acceptance uniforms are the stored arrays.  One ``.npz`` per case holds the
raw constructor inputs, the injected draws and what the reference produced.  One case is
additionally run through the public ``HMC().sample(...)`` call and must give the same
samples file.
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import cases  # noqa: E402
from _reference_shim import import_reference  # noqa: E402


class ReplayRNG:
    """Duck-typed numpy Generator handing out pre-drawn values in the reference's order."""

    def __init__(self, z, u_step, u_acc):
        self.z, self.u_step, self.u_acc = z, u_step, u_acc
        self.iz = self.istep = self.iacc = 0
        self.log = []

    def normal(self, size=None, **kw):
        out = np.array(self.z[self.iz], dtype=np.float64).reshape(size)
        self.iz += 1
        self.log.append("normal")
        return out

    def uniform(self, low=0.0, high=1.0, size=None):
        if (low, high) == (0.5, 1.5):
            out = float(self.u_step[self.istep])
            self.istep += 1
            self.log.append("step")
        elif (low, high) == (0, 1):
            out = float(self.u_acc[self.iacc])
            self.iacc += 1
            self.log.append("acc")
        else:
            raise AssertionError((low, high))
        return out


def drive_reference(hmclab, name, inp, tmpdir):
    s = cases.SETTINGS[name]
    C, K, d = s["chains"], s["proposals"], int(inp["dims"])
    n_grad = s["steps"] * {"lf": 1, "3s": 3, "4s": 4}[s["integrator"]]
    out = dict(
        accept=np.zeros((K, C), dtype=bool), H0=np.zeros((K, C)), H1=np.zeros((K, C)),
        q_prop=np.zeros((K, C, d)), p_prop=np.zeros((K, C, d)),
        samples=np.zeros((K, C, d + 1)),
        trace_q=np.zeros((K, n_grad, C, d)), trace_g=np.zeros((K, n_grad, C, d)),
    )
    post, mass = cases.build(name, inp, hmclab)
    real_gradient = post.gradient
    trace = []

    def logging_gradient(m):
        g = real_gradient(m)
        trace.append((m[:, 0].copy(), g[:, 0].copy()))
        return g

    post.gradient = logging_gradient
    for c in range(C):
        sampler = hmclab.Samplers.HMC(seed=0)
        rng = ReplayRNG(inp["z"][:, c], inp["u_step"][:, c], inp["u_acc"][:, c])
        sampler.rng = rng
        sampler._init_sampler(
            samples_filename=os.path.join(tmpdir, f"{name}_{c}.npy"), distribution=post,
            initial_model=inp["q0"][c].copy(), proposals=K, online_thinning=1,
            overwrite_existing_file=True, max_time=None, disable_progressbar=True,
            diagnostic_mode=False, stepsize=s["stepsize"],
            randomize_stepsize=s["randomize"], amount_of_steps=s["steps"],
            mass_matrix=mass, integrator=s["integrator"], autotuning=False,
            target_acceptance_rate=0.65, learning_rate=0.75,
        )
        with np.errstate(all="ignore"):
            for k in range(K):
                sampler.current_proposal = k
                before = sampler.accepted_proposals
                trace.clear()
                sampler._propose()
                assert len(trace) == n_grad
                out["trace_q"][k, :, c] = np.stack([t[0] for t in trace])
                out["trace_g"][k, :, c] = np.stack([t[1] for t in trace])
                out["q_prop"][k, c] = sampler.proposed_model[:, 0]
                out["p_prop"][k, c] = sampler.proposed_momentum[:, 0]
                sampler._evaluate_acceptance()
                out["accept"][k, c] = sampler.accepted_proposals > before
                out["H0"][k, c] = sampler.current_h
                out["H1"][k, c] = sampler.proposed_h
                out["samples"][k, c, :d] = sampler.current_model[:, 0]
                out["samples"][k, c, d] = sampler.current_x
        sampler.samples.close()
        expected = ["normal"] + (["step"] if s["randomize"] else []) + ["acc"]
        assert rng.log == expected * K, rng.log[:6]
    post.gradient = real_gradient

    # direct misfit / gradient contract at probe points (inside and far outside bounds)
    probes = np.vstack([inp["probe"], inp["probe"][:1] * 50.0 + 40.0])
    with np.errstate(all="ignore"):
        out["probe_points"] = probes
        out["probe_misfit"] = np.array([post.misfit(p.reshape(d, 1).copy()) for p in probes])
        out["probe_gradient"] = np.stack(
            [post.gradient(p.reshape(d, 1).copy())[:, 0] for p in probes])
    return out, post, mass


def check_public_sample_call(hmclab, name, inp, golden, tmpdir):
    """Same chain through HMC().sample(...) -> .npy must equal the driven run."""
    s = cases.SETTINGS[name]
    post, mass = cases.build(name, inp, hmclab)
    sampler = hmclab.Samplers.HMC(seed=0)
    sampler.rng = ReplayRNG(inp["z"][:, 0], inp["u_step"][:, 0], inp["u_acc"][:, 0])
    fn = os.path.join(tmpdir, "public.npy")
    with np.errstate(all="ignore"):
        sampler.sample(fn, post, stepsize=s["stepsize"], randomize_stepsize=s["randomize"],
                       amount_of_steps=s["steps"], mass_matrix=mass,
                       integrator=s["integrator"], initial_model=inp["q0"][0].copy(),
                       proposals=s["proposals"], overwrite_existing_file=True,
                       disable_progressbar=True)
    stored = np.load(fn)  # (n, d+1) on disk; the reader transposes (Samples.py:160-161)
    assert np.array_equal(stored, golden["samples"][:, 0, :]), name


def seeded_public_runs(hmclab, tmpdir):
    """Reference HMC(seed).sample(...) with its own numpy Generator (no replay): what a
    user gets for a given seed.  hmclab_b200's ``host_rng=True`` mode must reproduce these
    sample files from the same seed."""
    out = {}
    for name, seed, thin in (("dense_premult_cfg1", 42, 1), ("srcloc_fixed_v", 7, 2),
                             ("normal_bounded", 3, 1)):
        s = cases.SETTINGS[name]
        inp = cases.make_inputs(name)
        post, mass = cases.build(name, inp, hmclab)
        fn = os.path.join(tmpdir, f"seeded_{name}.npy")
        sampler = hmclab.Samplers.HMC(seed=seed)
        with np.errstate(all="ignore"):
            sampler.sample(fn, post, stepsize=s["stepsize"], randomize_stepsize=s["randomize"],
                           amount_of_steps=s["steps"], mass_matrix=mass, integrator=s["integrator"],
                           initial_model=inp["q0"][0].copy(), proposals=24, online_thinning=thin,
                           overwrite_existing_file=True, disable_progressbar=True)
        out[f"{name}__samples"] = np.load(fn)
        out[f"{name}__seed"] = np.int64(seed)
        out[f"{name}__thinning"] = np.int64(thin)
        out[f"{name}__accepted"] = np.int64(sampler.accepted_proposals)
    np.savez_compressed(os.path.join(HERE, "seeded_public_runs.npz"), **out)
    print("seeded_public_runs", {k: v.shape for k, v in out.items() if k.endswith("samples")})


AUTOTUNE_CASES = ("normal_unit_lf", "dense_direct_4s", "srcloc_fixed_v", "sparse_laplace_lf")


def autotuned_runs(hmclab, tmpdir):
    """Reference chains with autotuning=True (Samplers.py:1494-1522), replayed draws."""
    out = {}
    for name in AUTOTUNE_CASES:
        s = cases.SETTINGS[name]
        inp = cases.make_inputs(name)
        C, K, d = s["chains"], s["proposals"], int(inp["dims"])
        post, mass = cases.build(name, inp, hmclab)
        accept = np.zeros((K, C), dtype=bool)
        stepsizes = np.zeros((K, C))
        samples = np.zeros((K, C, d + 1))
        final = np.zeros(C)
        for c in range(C):
            sampler = hmclab.Samplers.HMC(seed=0)
            sampler.rng = ReplayRNG(inp["z"][:, c], inp["u_step"][:, c], inp["u_acc"][:, c])
            fn = os.path.join(tmpdir, f"auto_{name}_{c}.npy")
            with np.errstate(all="ignore"):
                sampler.sample(fn, post, stepsize=s["stepsize"], randomize_stepsize=s["randomize"],
                               amount_of_steps=s["steps"], mass_matrix=mass, integrator=s["integrator"],
                               initial_model=inp["q0"][c].copy(), proposals=K, autotuning=True,
                               target_acceptance_rate=0.65, learning_rate=0.75,
                               overwrite_existing_file=True, disable_progressbar=True)
            samples[:, c] = np.load(fn)
            accept[1:, c] = np.any(np.diff(samples[:, c], axis=0) != 0, axis=1)
            accept[0, c] = np.any(samples[0, c, :d] != inp["q0"][c])
            stepsizes[: len(sampler.stepsizes), c] = sampler.stepsizes[:, 0]
            # the reference truncates its histories to current_proposal entries (off by one)
            stepsizes[len(sampler.stepsizes):, c] = np.nan
            final[c] = sampler.stepsize
        out[f"{name}__samples"] = samples
        out[f"{name}__stepsizes"] = stepsizes
        out[f"{name}__final_stepsize"] = final
        print("autotuned", name, "final stepsizes", np.round(final, 4))
    np.savez_compressed(os.path.join(HERE, "autotuned_runs.npz"), **out)


RWMH_CASES = {   # name -> (stepsize kind, autotuning)
    "normal_bounded": ("scalar", False),
    "dense_premult_cfg1": ("vector", False),
    "srcloc_fixed_v": ("scalar", True),
    "sparse_laplace_lf": ("vector", True),
}


def rwmh_step(name, dims):
    rng = np.random.default_rng(len(name))
    base = {"normal_bounded": 0.4, "dense_premult_cfg1": 0.02, "srcloc_fixed_v": 0.05,
            "sparse_laplace_lf": 0.03}[name]
    return base, base * rng.uniform(0.5, 1.5, size=(dims, 1))


def rwmh_runs(hmclab, tmpdir):
    """Reference RWMH chains (Samplers.py:777-1102) with replayed draws."""
    out = {}
    for name, (kind, tune) in RWMH_CASES.items():
        s = cases.SETTINGS[name]
        inp = cases.make_inputs(name)
        C, K, d = s["chains"], s["proposals"], int(inp["dims"])
        post, _ = cases.build(name, inp, hmclab)
        scalar, vector = rwmh_step(name, d)
        samples = np.zeros((K, C, d + 1))
        accepted = np.zeros(C, dtype=np.int64)
        final = np.zeros(C)
        for c in range(C):
            sampler = hmclab.Samplers.RWMH(seed=0)
            sampler.rng = ReplayRNG(inp["z"][:, c], inp["u_step"][:, c], inp["u_acc"][:, c])
            fn = os.path.join(tmpdir, f"rwmh_{name}_{c}.npy")
            with np.errstate(all="ignore"):
                sampler.sample(fn, post, stepsize=scalar if kind == "scalar" else vector.copy(),
                               initial_model=inp["q0"][c].copy(), proposals=K, autotuning=tune,
                               overwrite_existing_file=True, disable_progressbar=True)
            samples[:, c] = np.load(fn)
            accepted[c] = sampler.accepted_proposals
            final[c] = sampler.stepsize if tune else np.nan
        out[f"{name}__samples"] = samples
        out[f"{name}__accepted"] = accepted
        out[f"{name}__final_stepsize"] = final
        out[f"{name}__step_scalar"] = np.float64(scalar)
        out[f"{name}__step_vector"] = vector
        print("rwmh", name, kind, "autotune" if tune else "", "accepted", accepted, np.round(final, 4))
    np.savez_compressed(os.path.join(HERE, "rwmh_runs.npz"), **out)


def main():
    hmclab = import_reference()
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from hmclab_b200._lowering import describe, describe_mass

    with tempfile.TemporaryDirectory() as tmpdir:
        if "--autotune-only" in sys.argv:
            autotuned_runs(hmclab, tmpdir)
            return
        if "--rwmh-only" in sys.argv:
            rwmh_runs(hmclab, tmpdir)
            return
        if not any(a.startswith("--only=") for a in sys.argv):
            seeded_public_runs(hmclab, tmpdir)
            if "--seeded-only" in sys.argv:
                return
            autotuned_runs(hmclab, tmpdir)
            rwmh_runs(hmclab, tmpdir)
        only = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--only=")]
        for name in cases.CASES:
            if only and name not in only:
                continue
            inp = cases.make_inputs(name)
            golden, post, mass = drive_reference(hmclab, name, inp, tmpdir)
            if name in ("dense_premult_cfg1", "srcloc_fixed_v"):
                check_public_sample_call(hmclab, name, inp, golden, tmpdir)
            describe(post), describe_mass(mass)  # the lowering must read reference objects
            payload = {f"in_{k}": v for k, v in inp.items()}
            payload.update({f"ref_{k}": v for k, v in golden.items()})
            np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **payload)
            acc = golden["accept"].mean()
            finite = np.isfinite(golden["H1"]).mean()
            print(f"{name:28s} accept={acc:.2f} finite_H1={finite:.2f} "
                  f"size={os.path.getsize(os.path.join(HERE, name + '.npz')) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
