"""How fast is the CPU restatement (oracle/hmc_oracle.py) next to the UNMODIFIED reference on the
bench workloads?  Run in the build container (needs /root/reference); one core, a few seconds per
workload.  Writes profiles/port_vs_reference_r02.json, which bench.py quotes in `cpu_baseline.sample`
(the reference itself cannot travel to the GPU box: its package is not part of the repository).

    python profiles/tools/port_vs_reference.py
"""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
os.environ.setdefault("OMP_NUM_THREADS", "1")

from _reference_shim import import_reference  # noqa: E402

from hmclab_b200 import workloads  # noqa: E402
from hmclab_b200._lowering import describe, describe_mass  # noqa: E402
from oracle import hmc_oracle as oracle  # noqa: E402


def reference_objects(hmclab, name, kw):
    """The same problem built by the same builder from the reference's own classes."""
    mine = (workloads.D, workloads.M)
    workloads.D, workloads.M = hmclab.Distributions, hmclab.MassMatrices
    try:
        w = workloads.BUILDERS[name](**kw)
    finally:
        workloads.D, workloads.M = mine
    return w.posterior, w.mass_matrix


def main():
    hmclab = import_reference()
    out = {"how": "one core of the build container, same workload objects, seconds-long runs; "
                  "rate = proposals x gradient evaluations per proposal / wall time"}
    for name in ("dense_small", "normal_iid", "dense_large", "tomography", "source_location"):
        kw = {"chains": 1}
        w = workloads.BUILDERS[name](**kw)
        try:
            ref_post, ref_mass = reference_objects(hmclab, name, kw)
        except Exception as exc:  # noqa: BLE001
            print(name, "reference objects:", repr(exc))
            continue
        q0 = w.initial_models[0][:, None].copy()
        # the reference, through its public call
        props = {"dense_small": 3000, "normal_iid": 1500, "dense_large": 3, "tomography": 6, "source_location": 400}[name]
        with tempfile.TemporaryDirectory() as tmp:
            s = hmclab.Samplers.HMC(seed=1)
            t0 = time.perf_counter()
            s.sample(os.path.join(tmp, "x.npy"), ref_post, stepsize=w.stepsize, amount_of_steps=w.amount_of_steps,
                     integrator=w.integrator, mass_matrix=ref_mass, initial_model=q0, proposals=props,
                     disable_progressbar=True, overwrite_existing_file=True)
            t_ref = time.perf_counter() - t0
        tree, mtree = describe(w.posterior), describe_mass(w.mass_matrix)
        with np.errstate(all="ignore"):
            oracle.run_chain(tree, mtree, integrator=w.integrator, steps=w.amount_of_steps, stepsize=w.stepsize,
                             randomize=True, q0=w.initial_models[0], proposals=1, draws=oracle.GeneratorDraws(0))
            t0 = time.perf_counter()
            oracle.run_chain(tree, mtree, integrator=w.integrator, steps=w.amount_of_steps, stepsize=w.stepsize,
                             randomize=True, q0=w.initial_models[0], proposals=props, draws=oracle.GeneratorDraws(1))
            t_port = time.perf_counter() - t0
        g = w.grads_per_proposal
        rec = {"reference": props * g / t_ref, "port": props * g / t_port,
               "port_over_reference": (props * g / t_port) / (props * g / t_ref), "proposals": props}
        out[name] = rec
        print(name, rec, flush=True)
    with open(os.path.join(ROOT, "profiles", "port_vs_reference_r02.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
