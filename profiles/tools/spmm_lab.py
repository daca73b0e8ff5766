"""Times the two SpMMs of one gradient evaluation of the tomography workload (config 4) for the
L2-gather kernel and the shared-memory staged kernel under several thread mappings / strip
limits, and checks that they agree.  Run on a B200:
    python profiles/tools/spmm_lab.py [--chains 8192] [--out gpurun_out/spmm_lab.json]
Knobs are the HMCB_SPMM_* environment variables read by hmcb_finalize."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hmclab_b200 import workloads  # noqa: E402
from hmclab_b200._engine import Engine  # noqa: E402
from hmclab_b200._lowering import describe, describe_mass, flatten  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", type=int, default=8192)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default="gpurun_out/spmm_lab.json")
    ap.add_argument("--configs", default="")
    args = ap.parse_args()
    t0 = time.time()
    w = workloads.tomography(chains=args.chains)
    print(f"workload built in {time.time() - t0:.1f} s", flush=True)
    tree, mtree = describe(w.posterior), describe_mass(w.mass_matrix)
    plan = flatten(tree)
    q = torch.as_tensor(w.initial_models, dtype=torch.float64).cuda().contiguous()
    q += 0.01 * torch.randn_like(q)
    configs = [dict(HMCB_SPMM_SHAPE="-1")]
    for shape in ("0", "1", "2", "3"):
        configs.append(dict(HMCB_SPMM_SHAPE=shape))
    configs += [dict(HMCB_SPMM_SHAPE="0", HMCB_SPMM_KB="176", HMCB_SPMM_EMAX="1024"),
                dict(HMCB_SPMM_SHAPE="0", HMCB_SPMM_KB="120", HMCB_SPMM_EMAX="640", HMCB_SPMM_STAGES="3"),
                dict(HMCB_SPMM_SHAPE="1", HMCB_SPMM_KB="176", HMCB_SPMM_EMAX="1024"),
                dict(HMCB_SPMM_SHAPE="1", HMCB_SPMM_KB="120", HMCB_SPMM_EMAX="640", HMCB_SPMM_STAGES="3"),
                dict(HMCB_SPMM_SHAPE="2", HMCB_SPMM_KB="320", HMCB_SPMM_EMAX="1536")]
    if args.configs:
        configs = json.loads(args.configs)
    ref = None
    results = []
    for cfg in configs:
        for k in list(os.environ):
            if k.startswith("HMCB_SPMM_"):
                del os.environ[k]
        os.environ.update(cfg)
        t0 = time.time()
        try:
            eng = Engine(plan, mtree, args.chains, integrator=w.integrator, amount_of_steps=w.amount_of_steps)
        except Exception as exc:  # noqa: BLE001
            print(cfg, "FAILED:", exc, flush=True)
            results.append(dict(cfg=cfg, error=str(exc)))
            continue
        setup = time.time() - t0
        g = eng.gradient(q)
        x = eng.misfit(q)
        torch.cuda.synchronize()
        if ref is None:
            ref = (g.clone(), x.clone())
        gerr = float((g - ref[0]).abs().max() / ref[0].abs().max())
        xerr = float(((x - ref[1]).abs() / ref[1].abs()).max())
        times = []
        eng.kernel_timing_begin()
        for _ in range(args.reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            eng.gradient(q)
            b.record()
            torch.cuda.synchronize()
            times.append(a.elapsed_time(b))
        (pass_ms, passes), _ = eng.kernel_timing_end()
        res = dict(cfg=cfg, setup_s=round(setup, 2), grad_ms=min(times), grad_ms_all=times, g_relerr=gerr,
                   x_relerr=xerr, spmm_pair_ms=pass_ms / max(passes, 1))
        print(json.dumps(res), flush=True)
        results.append(res)
        eng.close()
        del eng, g, x
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
