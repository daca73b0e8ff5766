#!/usr/bin/env python
"""Throughput of the batched HMC hot path: gradient evaluations per second
(chains x leapfrog steps), BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

A "step" is one block of ``--block`` HMC proposals for every chain of the workload
(momentum draw on the device, full trajectory, energies, Metropolis decision, sample
rows written).  Default workload: BASELINE.json configs[1] (1000-dim standard normal,
4096 chains, leapfrog L=10, Unit mass).  Under torchrun every rank owns its own
``chains`` chains (weak scaling, chain ids keyed by rank, no collective in the step; one
NCCL all-gather of the acceptance counters per block).

One JSON line on stdout (rank 0).  ``value`` is timed with CUDA events around every step
(inputs resident in HBM, L2 flushed between steps); ``e2e`` is the same metric through
``hmcb_sample_host`` with pinned HOST buffers, copies inside the timed region.
``--impl reference`` times the CPU restatement of the reference's path (oracle/, one
process per host core) on the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "gradient evals/s (chains x leapfrog steps)"
UNIT = "grad_evals/s"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": float(p["hbm_gbs"]), "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


# fp64 peak is not in MEASURED_PEAKS.json; profiles/fp64_peak_r01.json holds this pool's
# measured cuBLAS DGEMM / DFMA numbers once measured, else the datasheet value is used.
def fp64_peak_tflops():
    path = os.path.join(ROOT, "profiles", "fp64_peak_r01.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["dgemm_tflops"]), "measured cuBLAS DGEMM (profiles/fp64_peak_r01.json)"
    return 37.0, "datasheet (unmeasured)"


class ClockSampler:
    """SM clock and throttle reasons polled through NVML during the timed region (the
    same counters `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.*` prints)."""

    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
               "sw_power_cap": 0x4, "hw_power_brake_slowdown": 0x80}

    def __init__(self, index, period=0.004):
        self.index, self.period = index, period
        self.sm, self.power, self.mask, self.smax = [], [], 0, None
        self._stop = threading.Event()
        self.thread = None
        self.error = None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[self.index]) if visible and visible.replace(",", "").isdigit() else self.index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception as exc:  # pragma: no cover
            self.error = repr(exc)

    def _poll(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.handle) / 1000.0)
                self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
            except Exception as exc:  # pragma: no cover
                self.error = repr(exc)
                break
            time.sleep(self.period)

    def stop(self):
        self._stop.set()
        if self.thread is not None:
            self.thread.join(timeout=2)
        reasons = sorted(name for name, bit in self.REASONS.items() if self.mask & bit)
        out = {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.smax,
               "reasons": reasons, "samples": len(self.sm),
               "power_w_max": max(self.power) if self.power else None}
        if self.error:
            out["error"] = self.error
        return out


def algorithmic_bytes_per_step(w, block, thinning):
    """HBM bytes one step must move with the trajectory fused on chip:
    q read + q write + stored sample rows + per-chain scalars."""
    C, d = w.chains, w.dims
    rows = block // thinning
    return 8.0 * C * d * 2 + 8.0 * C * (d + 1) * rows + 8.0 * C * 2 + 4.0 * C


def cpu_oracle_rate(w, seconds, seed=0):
    """grad evals/s of the numpy restatement (one chain, one core) on a bounded sample."""
    from hmclab_b200._lowering import describe, describe_mass
    from oracle import hmc_oracle as oracle

    tree, mtree = describe(w.posterior), describe_mass(w.mass_matrix)
    draws = oracle.GeneratorDraws(seed)
    q0 = w.initial_models[0]
    done, t0 = 0, time.perf_counter()
    chunk = 1
    with np.errstate(all="ignore"):
        while True:
            res = oracle.run_chain(tree, mtree, integrator=w.integrator, steps=w.amount_of_steps,
                                   stepsize=w.stepsize, randomize=True, q0=q0, proposals=chunk,
                                   draws=draws)
            q0 = res["final_q"]
            done += chunk
            el = time.perf_counter() - t0
            if el >= seconds:
                break
            chunk = max(1, min(4 * chunk, int(chunk * (seconds - el) / max(el, 1e-3) * 0.5) or 1))
    return done * w.grads_per_proposal / el, done, el


def _reference_worker(args):
    name, kwargs, proposals, seed, chain = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from hmclab_b200 import workloads
    from hmclab_b200._lowering import describe, describe_mass
    from oracle import hmc_oracle as oracle

    w = _WORKLOAD_CACHE.get(name)
    if w is None:
        w = workloads.BUILDERS[name](**kwargs)
        _WORKLOAD_CACHE[name] = w
    tree, mtree = describe(w.posterior), describe_mass(w.mass_matrix)
    with np.errstate(all="ignore"):
        res = oracle.run_chain(tree, mtree, integrator=w.integrator, steps=w.amount_of_steps,
                               stepsize=w.stepsize, randomize=True,
                               q0=w.initial_models[chain % w.chains], proposals=proposals,
                               draws=oracle.GeneratorDraws(seed))
    return int(res["accept"].sum())


_WORKLOAD_CACHE = {}


def run_reference(args, w, name, kwargs):
    """CPU arm: the oracle port of the reference's path, one process per host core (the
    reference's ParallelSampleSMP layout: one OS process per chain)."""
    import multiprocessing as mp

    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = "1"
    cores = os.cpu_count() or 1
    # size a step: about (150 s / (steps + warmup)) of work per core, bounded
    rate1, _, _ = cpu_oracle_rate(w, 3.0)
    budget = max(1.0, min(20.0, 150.0 / (args.steps + args.warmup)))
    proposals = max(1, int(rate1 * budget / w.grads_per_proposal))
    _WORKLOAD_CACHE[name] = w
    ctx = mp.get_context("fork")
    times = []
    with ctx.Pool(cores) as pool:
        for it in range(args.warmup + args.steps):
            jobs = [(name, kwargs, proposals if it >= args.warmup else max(1, proposals // 10),
                     1000 * it + c, c) for c in range(cores)]
            t0 = time.perf_counter()
            pool.map(_reference_worker, jobs)
            if it >= args.warmup:
                times.append(time.perf_counter() - t0)
    total = sum(times)
    evals = cores * proposals * w.grads_per_proposal * args.steps
    value = evals / total
    sample = (f"{cores} processes x 1 chain x {proposals} proposals per step "
              f"(oracle/hmc_oracle.py, numpy {np.__version__}, 1 BLAS thread each)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": w.description, "chains_per_step": cores,
                                        "proposals_per_step": proposals},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="normal_iid")
    ap.add_argument("--block", type=int, default=0, help="proposals per step (default: per workload)")
    ap.add_argument("--thinning", type=int, default=1)
    ap.add_argument("--chains", type=int, default=0, help="override chains per GPU")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl != "reference" and args.warmup < 3:
        args.warmup = 3   # timing rules: at least 3 warm-up steps (the line reports what was run)
    args.steps = max(1, args.steps)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    from hmclab_b200 import workloads

    kwargs = {"chains": args.chains} if args.chains else {}
    if args.impl == "reference":
        if rank != 0:
            return
        w = workloads.BUILDERS[args.workload](**kwargs)
        run_reference(args, w, args.workload, kwargs)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; hmclab_b200 has no CPU path to time")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from hmclab_b200._engine import Engine
    from hmclab_b200._lowering import describe, describe_mass, flatten

    w = workloads.BUILDERS[args.workload](**kwargs)
    if not args.block:
        # long enough steps that host launch overhead is invisible, short enough for minutes
        args.block = {"normal_iid": 50, "source_location": 50, "dense_small": 50}.get(args.workload, 1)
    C, d, B, thin = w.chains, w.dims, args.block, args.thinning
    assert B % thin == 0
    eng = Engine(flatten(describe(w.posterior)), describe_mass(w.mass_matrix), C,
                 integrator=w.integrator, amount_of_steps=w.amount_of_steps, device=local_rank)
    dev = torch.device("cuda", local_rank)
    q = torch.as_tensor(w.initial_models, dtype=torch.float64).to(dev).contiguous()
    x = eng.misfit(q)
    rows = B // thin
    samples = torch.empty(rows, C, d + 1, dtype=torch.float64, device=dev)
    accepted = torch.zeros(C, dtype=torch.int32, device=dev)
    gathered = torch.zeros(world * C, dtype=torch.int32, device=dev) if world > 1 else None
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    chain_offset = rank * C

    def step(k):
        eng.run_block(q, x, B, stepsize=w.stepsize, randomize_stepsize=True, thinning=thin,
                      proposal_offset=k * B, chain_offset=chain_offset, seed=2026,
                      out_samples=samples, accepted_total=accepted)
        if world > 1:  # diagnostics gather, once per sample block
            dist.all_gather_into_tensor(gathered, accepted)

    for k in range(args.warmup):
        step(k)
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()      # NVML init takes milliseconds: keep it in front of the barrier
    launches0 = eng.launch_count
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)           # evict the previous step's lines from L2 (untimed)
        starts[i].record()
        step(args.warmup + i)
        ends[i].record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - wall0
    if world > 1:
        dist.barrier()
    launches = eng.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    total_ms = sum(step_ms)
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    per_rank_ms = [total_ms / args.steps]
    if world > 1:
        every = torch.zeros(world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(every, t / args.steps)
        per_rank_ms = [float(v) for v in every.cpu()]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    evals_per_step = world * C * B * w.grads_per_proposal
    value = evals_per_step * args.steps / (total_ms * 1e-3)
    acc_rate = float(accepted.double().mean().item()) / ((args.warmup + args.steps) * B)

    # ---- end to end: host buffers, copies inside the timed region -----------------------
    e2e = None
    if not args.no_e2e:
        q0_host = torch.as_tensor(w.initial_models, dtype=torch.float64).contiguous().pin_memory()
        out_host = torch.empty(rows, C, d + 1, dtype=torch.float64).pin_memory()
        acc_host = torch.empty(C, dtype=torch.int32).pin_memory()
        n_e2e = max(3, min(args.steps, 10))
        # sub-blocks so that the D2H of one sub-block overlaps the kernels of the next
        e2e_block = max(thin, (B // 5) // thin * thin) if B >= 5 * thin else B
        for _ in range(2):
            eng.sample_host(q0_host, B, stepsize=w.stepsize, thinning=thin, block_proposals=e2e_block,
                            seed=7, chain_offset=chain_offset, samples=out_host, accepted=acc_host)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for i in range(n_e2e):
            eng.sample_host(q0_host, B, stepsize=w.stepsize, thinning=thin, block_proposals=e2e_block,
                            seed=8 + i, chain_offset=chain_offset, samples=out_host, accepted=acc_host)
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        te = torch.tensor([el], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        el = float(te.item())
        e2e = {"value": evals_per_step * n_e2e / el, "unit": UNIT,
               "h2d_bytes_per_step": int(q0_host.numel() * 8),
               "d2h_bytes_per_step": int(out_host.numel() * 8 + acc_host.numel() * 4),
               "steps": n_e2e, "api": "hmcb_sample_host (C ABI, pinned host buffers)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ---------------------------------------------------
    peaks = measured_peaks()
    ms_per_step = total_ms / args.steps
    traffic = None   # dram bytes of one launch from the committed ncu capture, if it matches this run
    tpath = os.path.join(ROOT, "profiles", "traffic_r01.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            rec = json.load(f).get(args.workload)
        if rec and rec["proposals_per_launch"] == B and thin == 1 and not args.chains:
            traffic = rec["dram_bytes_read"] + rec["dram_bytes_write"]
    if eng.path in ("fused_priors", "fused_srcloc"):
        abytes = algorithmic_bytes_per_step(w, B, thin)
        achieved = abytes / (ms_per_step * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": achieved / peaks["hbm_gbs"], "traffic": traffic,
                    "kernel": "hmc_fused_priors_kernel" if eng.path == "fused_priors" else "hmc_fused_srcloc_kernel",
                    "algorithmic_bytes_per_launch": abytes, "peak_source": peaks["source"],
                    "note": "one launch per step; fp64 SIMT pipe, not HBM, limits this kernel (see DESIGN.md)"}
        ops = w.extra.get("fp64_ops_per_grad")
        if ops:
            # the pipe that actually binds: fp64 instructions (an FMA counts once) against the
            # measured DFMA issue rate of this pool's B200 (profiles/fp64_peak_r01.json)
            peak_path = os.path.join(ROOT, "profiles", "fp64_peak_r01.json")
            ginst = 16940.0
            if os.path.exists(peak_path):
                with open(peak_path) as f:
                    ginst = float(json.load(f).get("dfma_ginst_per_s", ginst))
            rate = ops * value / world / 1e9
            roofline["fp64_pipe"] = {"achieved_ginst_per_s": rate, "peak_ginst_per_s": ginst,
                                     "frac": rate / ginst, "fp64_ops_per_grad_eval": ops}
    else:
        peak, src = fp64_peak_tflops()
        flops = w.extra.get("flops_per_grad")
        if flops is None and "nnz" in w.extra:
            flops = 4.0 * w.extra["nnz"]
        achieved = flops * value / world / 1e12
        roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "frac": achieved / peak, "traffic": None,
                    "kernel": {"fused_dense": "hmc_fused_dense_kernel (one launch per step)"}.get(
                        eng.path, "dmma_gemm_kernel / csr_spmm_kernel (whole step time attributed)"),
                    "algorithmic_flops_per_grad_eval": flops, "peak_source": src}
        if "nnz" in w.extra:
            # SpMM path: also the HBM view of the same step, from the
            # algorithmic bytes 16 (d + N) + 24 nnz / C per gradient evaluation and chain
            n_data = w.extra.get("data", w.extra.get("rays", 0))
            abytes = 16.0 * (w.dims + n_data) + 24.0 * w.extra["nnz"] / w.chains
            gbs = abytes * value / world / 1e9
            roofline["kernel"] = "csr_spmm_strip_kernel (two launches per gradient evaluation; whole step time attributed)"
            roofline["hbm"] = {"achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                               "frac": gbs / peaks["hbm_gbs"], "algorithmic_bytes_per_grad_eval": abytes}
            roofline["note"] = ("fp64 FMA work (4 nnz flop per gradient evaluation) against the measured DGEMM "
                                "peak; the kernel itself is bound by the shared-memory data pipe "
                                "(one 8-byte gather per FMA), see DESIGN.md")

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rate, n_prop, el = cpu_oracle_rate(w, args.cpu_seconds)
        cpu = {"value": rate, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"1 chain x {n_prop} proposals of the same workload in {el:.1f} s "
                         f"(oracle/hmc_oracle.py numpy restatement, host cores available: {os.cpu_count()})"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w.description, "chains_per_gpu": C, "dims": d,
                   "proposals_per_step": B, "online_thinning": thin, "integrator": w.integrator,
                   "amount_of_steps": w.amount_of_steps, "path": eng.path, "rng": "on-device Philox4x32-10",
                   "l2": "256 MiB flush write between timed steps"},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
        "cpu_baseline": cpu, "acceptance_rate": acc_rate, "wall_s_timed_region": wall,
        "ms_per_step_by_rank": per_rank_ms,
        "step_ms_rank0": {"min": min(step_ms), "median": float(np.median(step_ms)), "max": max(step_ms)},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
