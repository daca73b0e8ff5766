#!/usr/bin/env python
"""Throughput of the batched HMC hot path: gradient evaluations per second
(chains x leapfrog steps), BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME|all] [--impl reference]

A "step" is one block of HMC proposals for every chain of a workload (momentum draw on the
device, full trajectory, energies, Metropolis decision, sample rows written).

Default (``--workload all``): the headline is BASELINE.json configs[2] -- the largest
single-GPU configuration, dense LinearMatrix 2000 x 10000, 8192 chains, 4-stage integrator,
Diagonal mass (the fp64 tensor path) -- timed for ``--steps`` steps; the same JSON line carries
``per_config`` with full sub-records (value, ms_per_step, e2e, roofline, cpu_baseline, clocks)
for configs[0] (its shape on 4096 chains, plus one chain), configs[1], configs[3] and configs[4].
Under torchrun every rank owns its own ``chains_per_gpu`` chains of every workload (weak
scaling, chain ids keyed by rank, no collective in the step; one NCCL all-gather of the
acceptance counters per block), so configs[4] runs 8192 chains per GPU = 65 536 chains on 8.

One JSON line on stdout (rank 0).  ``value`` is timed with CUDA events around every step
(inputs resident in HBM, L2 flushed between steps); ``e2e`` is the same metric through
``hmcb_sample_host`` with pinned HOST buffers, copies inside the timed region; ``roofline``
is the dominant kernel's algorithmic work over its own CUDA-event duration
(hmcb_kernel_timing_*), against a peak measured in this run (fp64: cuBLAS DGEMM / DFMA / DMMA
bursts with their own clock record) or by the driver (HBM: MEASURED_PEAKS.json).
``--impl reference`` times the CPU restatement of the reference's path (oracle/, one process
per host core) on the same workloads.
"""
from __future__ import annotations

import argparse
import ctypes
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
HEADLINE = "dense_large"
SUB_WORKLOADS = ["dense_small", "normal_iid", "tomography", "source_location"]
BASELINE_CONFIG = {"dense_small": 0, "normal_iid": 1, "dense_large": 2, "dense_large_premult": 2,
                   "tomography": 3, "source_location": 4}
# proposals per step: long enough that host launch overhead is invisible, short enough for minutes
DEFAULT_BLOCK = {"normal_iid": 50, "source_location": 50, "dense_small": 50}
# steps of the sub-records (the headline uses --steps)
SUB_STEPS = {"dense_small": 10, "normal_iid": 20, "tomography": 4, "source_location": 20}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": float(p["hbm_gbs"]), "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


def committed_json(name):
    path = os.path.join(ROOT, "profiles", name)
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return {}


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
        return self

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


# ------------------------------------------------------------------- CPU side (oracle) ----

_WORKLOAD_CACHE = {}


def get_workload(name, chains=0):
    from hmclab_b200 import workloads

    key = (name, chains)
    if key not in _WORKLOAD_CACHE:
        kw = {"chains": chains} if chains else {}
        if name == "dense_small" and not chains:
            kw = {"chains": 4096}   # configs[0]'s shape on a batch (its single chain is timed too)
        _WORKLOAD_CACHE[key] = workloads.BUILDERS[name](**kw)
    return _WORKLOAD_CACHE[key]


_TREE_CACHE = {}


def get_tree(name, chains=0):
    """Plain-dict form of the workload's posterior / mass matrix for the oracle (cached: the
    oracle keys its matrix cache on the tree's identity)."""
    from hmclab_b200._lowering import describe, describe_mass

    key = (name, chains)
    if key not in _TREE_CACHE:
        w = get_workload(name, chains)
        _TREE_CACHE[key] = (describe(w.posterior), describe_mass(w.mass_matrix))
    return _TREE_CACHE[key]


def _cpu_worker(job):
    """One process = one chain of the numpy restatement (the reference's ParallelSampleSMP
    layout: one OS process per chain, one BLAS thread each)."""
    name, chains, proposals, seed, chain = job
    from oracle import hmc_oracle as oracle

    w = get_workload(name, chains)
    tree, mtree = get_tree(name, chains)
    t0 = time.perf_counter()
    with np.errstate(all="ignore"):
        res = oracle.run_chain(tree, mtree, integrator=w.integrator, steps=w.amount_of_steps,
                               stepsize=w.stepsize, randomize=True,
                               q0=w.initial_models[chain % w.chains], proposals=proposals,
                               draws=oracle.GeneratorDraws(seed))
    return int(res["accept"].sum()), time.perf_counter() - t0


class CpuArm:
    """Pool of one process per host core running the oracle port; created BEFORE CUDA is
    initialised (fork) and reused for every workload."""

    def __init__(self, names, chains=0):
        import multiprocessing as mp

        for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
            os.environ[var] = "1"
        try:
            from threadpoolctl import threadpool_limits

            threadpool_limits(1)
        except Exception:
            pass
        self.cores = os.cpu_count() or 1
        self.chains = chains
        for n in names:   # build once in the parent; the forked workers share the pages
            get_tree(n, chains)
        self.pool = mp.get_context("fork").Pool(self.cores)
        self.factors = committed_json("port_vs_reference_r02.json")

    def close(self):
        self.pool.close()
        self.pool.join()

    def step(self, name, proposals, seed0):
        jobs = [(name, self.chains, proposals, seed0 + c, c) for c in range(self.cores)]
        t0 = time.perf_counter()
        self.pool.map(_cpu_worker, jobs, chunksize=1)
        return time.perf_counter() - t0

    def calibrate(self, name, seconds):
        """proposals per chain so that one step (all cores busy) lasts about `seconds`."""
        t = self.step(name, 1, 10_000)          # also builds the workers' matrix caches
        t = self.step(name, 1, 20_000)
        return max(1, int(seconds / max(t, 1e-4)))

    def describe(self, name, proposals, steps):
        f = self.factors.get(name)
        note = (f"; in the build container the unmodified reference runs this workload at {1.0 / f['port_over_reference']:.2f}x "
                f"the port's rate ({f['reference']:.3g} vs {f['port']:.3g} grad evals/s/core)") if f else ""
        return (f"{self.cores} processes x 1 chain x {proposals} proposals x {steps} step(s) "
                f"(oracle/hmc_oracle.py, numpy {np.__version__}, 1 BLAS thread each{note})")

    def baseline(self, name, seconds):
        """Bounded sample of the workload on all host cores -> cpu_baseline record."""
        w = get_workload(name, self.chains)
        proposals = self.calibrate(name, seconds)
        el = self.step(name, proposals, 30_000)
        value = self.cores * proposals * w.grads_per_proposal / el
        return {"value": value, "unit": UNIT, "cores": self.cores, "kind": "port",
                "seconds": el, "sample": self.describe(name, proposals, 1)}


def run_reference(args, names):
    """CPU arm: the oracle port of the reference's path, one process per host core."""
    arm = CpuArm(names, args.chains)
    head = names[0]
    w = get_workload(head, args.chains)
    budget = max(1.0, min(20.0, 150.0 / (args.steps + args.warmup)))
    proposals = arm.calibrate(head, budget)
    times = []
    for it in range(args.warmup + args.steps):
        warm = it < args.warmup
        t = arm.step(head, max(1, proposals // 10) if warm else proposals, 1000 * it)
        if not warm:
            times.append(t)
    total = sum(times)
    value = arm.cores * proposals * w.grads_per_proposal * args.steps / total
    per_config = {}
    for n in names[1:]:
        rec = arm.baseline(n, 8.0)
        per_config[n] = {"baseline_config": BASELINE_CONFIG.get(n), "value": rec["value"], "unit": UNIT,
                         "workload": get_workload(n, args.chains).description, "cpu_baseline": rec}
    arm.close()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": w.description, "baseline_config": BASELINE_CONFIG.get(head),
                   "chains_per_step": arm.cores, "proposals_per_step": proposals},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.cores, "kind": "port",
                         "sample": arm.describe(head, proposals, args.steps)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if per_config:
        line["per_config"] = per_config
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------- GPU side --------

def bind_to_gpu_numa_node(local_rank):
    """Run this rank (and first-touch its pinned buffers) on the NUMA node of its GPU."""
    try:
        import torch

        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        with open(path) as f:
            node = int(f.read().strip())
        if node < 0:
            return {"numa_node": node, "bound": False}
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return {"numa_node": node, "bound": False}
        os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "bound": True, "cpus": len(cpus)}
    except Exception as exc:  # pragma: no cover
        return {"bound": False, "error": repr(exc)}


def measure_fp64_peak(torch, dev, local_rank, seconds=2.0):
    """fp64 roofline denominators, measured here and now: cuBLAS DGEMM 8192^3 (burst = best
    launch, sustained = median of a `seconds` long back-to-back loop) with the clocks seen, plus
    the DFMA / DMMA register loops of the library (hmcb_debug_fp64_peak)."""
    from hmclab_b200._engine import load_library

    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    c = torch.empty(n, n, dtype=torch.float64, device=dev)
    for _ in range(2):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank).start()
    times, t0 = [], time.perf_counter()
    while time.perf_counter() - t0 < seconds or len(times) < 5:
        batch = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(8)]
        for s, e in batch:
            s.record()
            torch.matmul(a, b, out=c)
            e.record()
        torch.cuda.synchronize()
        times += [s.elapsed_time(e) for s, e in batch]
    clocks = sampler.stop()
    flop = 2.0 * n ** 3
    out = {"dgemm_tflops_burst": flop / min(times) / 1e9,
           "dgemm_tflops_sustained": flop / float(np.median(times)) / 1e9,
           "dgemm_launches": len(times), "dgemm_n": n, "clocks": clocks,
           "how": f"torch.matmul float64 {n}^3 (cuBLAS DGEMM) back to back for {seconds:.0f} s inside this run; "
                  "DFMA / DMMA register loops of libhmcb.so (hmcb_debug_fp64_peak)"}
    del a, b, c
    lib = load_library()
    for kind, key in ((0, "dfma"), (1, "dmma")):
        ms, fl = ctypes.c_double(), ctypes.c_double()
        if lib.hmcb_debug_fp64_peak(local_rank, kind, 20000, 3, ctypes.byref(ms), ctypes.byref(fl)) == 0:
            out[key + "_tflops"] = fl.value / ms.value / 1e9
    if "dfma_tflops" in out:
        out["dfma_ginst_per_s"] = out["dfma_tflops"] * 1e3 / 2.0
    return out


def measure_d2h_ceiling(torch, dist, dev, world, gib=1.0, reps=4):
    """Bare device->pinned-host copy rate with every rank copying at once: the ceiling of the
    end-to-end sample stream (and, in the other direction, of the initial-model upload)."""
    n = int(gib * (1 << 30)) // 8
    src = torch.empty(n, dtype=torch.float64, device=dev).normal_()
    dst = torch.empty(n, dtype=torch.float64).pin_memory()
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    e.record()
    torch.cuda.synchronize()
    ms = torch.tensor([s.elapsed_time(e)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    per_rank = n * 8.0 * reps / (float(ms.item()) * 1e-3) / 1e9
    del src, dst
    return {"d2h_ceiling_gbs_per_gpu": per_rank, "d2h_ceiling_gbs_total": per_rank * world,
            "how": f"{world} rank(s) x {reps} x {gib:g} GiB cudaMemcpyAsync device->pinned host at once, max over ranks"}


def algorithmic_bytes_per_step(w, block, thinning):
    """HBM bytes one step must move with the trajectory fused on chip:
    q read + q write + stored sample rows + per-chain scalars."""
    C, d = w.chains, w.dims
    rows = block // thinning
    return 8.0 * C * d * 2 + 8.0 * C * (d + 1) * rows + 8.0 * C * 2 + 4.0 * C


class Ctx:
    pass


def bench_workload(ctx, name, steps, warmup, block=0, chains=0):
    """Time one workload on this rank's GPU (all ranks call this together) -> sub-record."""
    torch, dist = ctx.torch, ctx.dist
    from hmclab_b200._engine import Engine
    from hmclab_b200._lowering import describe, describe_mass, flatten

    rank, world, dev = ctx.rank, ctx.world, ctx.dev
    w = get_workload(name, chains)
    B = block or DEFAULT_BLOCK.get(name, 1)
    thin = ctx.thinning
    C, d = w.chains, w.dims
    assert B % thin == 0
    eng = Engine(flatten(describe(w.posterior)), describe_mass(w.mass_matrix), C,
                 integrator=w.integrator, amount_of_steps=w.amount_of_steps, device=ctx.local_rank)
    q = torch.as_tensor(w.initial_models, dtype=torch.float64).to(dev).contiguous()
    x = eng.misfit(q)
    rows = B // thin
    samples = torch.empty(rows, C, d + 1, dtype=torch.float64, device=dev)
    accepted = torch.zeros(C, dtype=torch.int32, device=dev)
    gathered = torch.zeros(world * C, dtype=torch.int32, device=dev) if world > 1 else None
    chain_offset = rank * C

    def step(k):
        eng.run_block(q, x, B, stepsize=w.stepsize, randomize_stepsize=True, thinning=thin,
                      proposal_offset=k * B, chain_offset=chain_offset, seed=2026,
                      out_samples=samples, accepted_total=accepted)
        if world > 1:  # diagnostics gather, once per sample block
            dist.all_gather_into_tensor(gathered, accepted)

    for k in range(warmup):
        step(k)
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    sampler = ClockSampler(ctx.local_rank)
    if rank == 0:
        sampler.start()      # NVML init takes milliseconds: keep it in front of the barrier
    launches0 = eng.launch_count
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    eng.kernel_timing_begin()
    wall0 = time.perf_counter()
    for i in range(steps):
        ctx.flush.fill_(i & 0xFF)       # evict the previous step's lines from L2 (untimed)
        starts[i].record()
        step(warmup + i)
        ends[i].record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - wall0
    ktimes = eng.kernel_timing_end()
    if world > 1:
        dist.barrier()
    launches = eng.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    total_ms = sum(step_ms)
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    per_rank_ms = [total_ms / steps]
    if world > 1:
        every = torch.zeros(world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(every, t / steps)
        per_rank_ms = [float(v) for v in every.cpu()]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    evals_per_step = world * C * B * w.grads_per_proposal
    value = evals_per_step * steps / (total_ms * 1e-3)
    acc_rate = float(accepted.double().mean().item()) / ((warmup + steps) * B)

    # ---- end to end: host buffers, copies inside the timed region -----------------------
    e2e = None
    if ctx.e2e:
        q0_host = torch.as_tensor(w.initial_models, dtype=torch.float64).contiguous().pin_memory()
        out_host = torch.empty(rows, C, d + 1, dtype=torch.float64).pin_memory()
        acc_host = torch.empty(C, dtype=torch.int32).pin_memory()
        n_e2e = max(3, min(steps, 10))
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
        d2h = int(out_host.numel() * 8 + acc_host.numel() * 4)
        e2e = {"value": evals_per_step * n_e2e / el, "unit": UNIT,
               "h2d_bytes_per_step": int(q0_host.numel() * 8), "d2h_bytes_per_step": d2h,
               "steps": n_e2e, "api": "hmcb_sample_host (C ABI, pinned host buffers)",
               "d2h_gbs_per_gpu": d2h * n_e2e / el / 1e9}
        if ctx.d2h_ceiling:
            e2e["d2h_ceiling_gbs_per_gpu"] = ctx.d2h_ceiling["d2h_ceiling_gbs_per_gpu"]
            e2e["d2h_frac_of_ceiling"] = e2e["d2h_gbs_per_gpu"] / ctx.d2h_ceiling["d2h_ceiling_gbs_per_gpu"]
        del q0_host, out_host, acc_host
    path = eng.path
    oz_pairs = eng.tcgen05_slice_pairs
    eng.close()
    del q, x, samples, accepted
    torch.cuda.empty_cache()
    if rank != 0:
        return None

    # ---- roofline of the dominant kernel ---------------------------------------------------
    ms_per_step = total_ms / steps
    (grad_ms, grad_passes), (mis_ms, mis_passes) = ktimes
    kernel_ms = grad_ms / max(grad_passes, 1)       # one gradient pass / one fused block launch
    share = (grad_ms + mis_ms) / sum(step_ms) if step_ms else None
    traffic_rec = ctx.traffic.get(name)
    traffic = None
    if traffic_rec and traffic_rec.get("proposals_per_launch", B) == B and thin == 1 and not chains:
        traffic = traffic_rec["dram_bytes_read"] + traffic_rec["dram_bytes_write"]
    peaks, fp64 = ctx.peaks, ctx.fp64
    timing = {"kernel_ms": kernel_ms, "launch_groups_timed": grad_passes,
              "share_of_step": share, "how": "CUDA event pairs on the launching stream around every launch "
              "group of the dominant kernel inside the timed region (hmcb_kernel_timing_*)"}
    if path in ("fused_priors", "fused_srcloc"):
        abytes = algorithmic_bytes_per_step(w, B, thin)
        achieved = abytes / (kernel_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": achieved / peaks["hbm_gbs"], "traffic": traffic,
                    "kernel": "hmc_fused_priors_kernel" if path == "fused_priors" else "hmc_fused_srcloc_kernel",
                    "algorithmic_bytes_per_launch": abytes, "peak_source": peaks["source"],
                    "note": "one launch per step; the fp64 SIMT pipe, not HBM, limits this kernel (see fp64_pipe and DESIGN.md)"}
        ops = w.extra.get("fp64_ops_per_grad")
        if ops and fp64.get("dfma_ginst_per_s"):
            # the pipe that actually binds: fp64 instructions (an FMA counts once) against the
            # DFMA issue rate measured in this run
            rate = ops * (C * B * w.grads_per_proposal) / (kernel_ms * 1e-3) / 1e9
            roofline["fp64_pipe"] = {"achieved_ginst_per_s": rate, "peak_ginst_per_s": fp64["dfma_ginst_per_s"],
                                     "frac": rate / fp64["dfma_ginst_per_s"], "fp64_ops_per_grad_eval": ops,
                                     "ops_source": "instruction-count estimate (workloads.py); the measured pipe "
                                                   "activity is ncu_counters.fp64_pipe_active_pct"}
    else:
        peak = fp64["dgemm_tflops_sustained"] if ms_per_step > 50 else fp64["dgemm_tflops_burst"]
        flops = w.extra.get("flops_per_grad")
        if flops is None and "nnz" in w.extra:
            flops = 4.0 * w.extra["nnz"]
        per_pass = flops * C if path != "fused_dense" else flops * C * B * w.grads_per_proposal
        achieved = per_pass / (kernel_ms * 1e-3) / 1e12
        kernel = {"fused_dense": "hmc_fused_dense_kernel (one launch per step)"}.get(
            path, "dmma_gemm_kernel (launch group = the GEMMs of one gradient evaluation of the batch)")
        if oz_pairs:
            kernel = ("i8_gemm_groups_kernel (tcgen05 int8 slice products) + slicing / recombination kernels "
                      "(launch group = everything of one gradient evaluation of the batch)")
        roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "frac": achieved / peak, "traffic": traffic, "kernel": kernel,
                    "algorithmic_flops_per_grad_eval": flops,
                    "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (%s), see fp64_peak" %
                                   ("sustained: median of a 2 s loop" if ms_per_step > 50 else "burst: best launch"),
                    "whole_step_tflops": flops * value / world / 1e12}
        if oz_pairs:
            # Ozaki scheme: the fp64 products run as exact int8 slice products on tcgen05.  The roofline of
            # the kernel is therefore the int8 tensor roof: `achieved` = int8 operations the algorithm
            # needs (slice pairs x 2 N d per chain; the pairs follow from the 1e-10 parity bar, DESIGN.md
            # section 4) over the duration of the WHOLE launch group (slicing + products + recombination);
            # the fp64 view of the same launches (4 N d flop per gradient evaluation against the DGEMM peak
            # measured in this run) is kept next to it -- its ratio may exceed 1, which is the point.
            pairs = oz_pairs                      # slice products of G q and G^T r together (premultiplied: of GtG q)
            # int8 operations per slice product and chain: 2 N d (half the 4 N d flops of the direct form's two
            # products), or 2 d^2 = all the flops of the premultiplied form's single product
            per_product = flops if w.extra.get("form") == "premult" else flops / 2.0
            i8_ops = pairs * per_product * C
            measured = committed_json("../MEASURED_PEAKS.json")
            sustained = measured.get("bf16_tflops_sustained") if ms_per_step > 50 else measured.get("bf16_tflops")
            i8_peak = 2.0 * sustained if sustained else 4500.0
            fp64_view = {"achieved": achieved, "peak": peak, "unit": "TFLOP/s", "ratio": achieved / peak,
                         "algorithmic_flops_per_grad_eval": flops, "peak_source": roofline["peak_source"],
                         "whole_step_tflops": roofline["whole_step_tflops"],
                         "note": "fp64-equivalent rate of the int8-sliced products against the native fp64 tensor "
                                 "path's measured peak: > 1 means faster than any DMMA/DGEMM kernel could be"}
            roofline = {"bound": "tensor", "achieved": i8_ops / (kernel_ms * 1e-3) / 1e12, "peak": i8_peak,
                        "unit": "TFLOP/s", "traffic": traffic, "kernel": kernel, "dtype_of_peak": "int8 (dense, TOP/s)",
                        "slice_pairs": pairs, "algorithmic_int8_ops_per_grad_eval": pairs * per_product,
                        "peak_source": ("2 x the driver-measured %s bf16 rate (MEASURED_PEAKS.json): tcgen05 kind::i8 "
                                        "runs at twice the bf16 rate" % ("sustained" if ms_per_step > 50 else "burst")
                                        if sustained else "nominal dense int8"),
                        "fp64_equivalent": fp64_view}
            roofline["frac"] = roofline["achieved"] / roofline["peak"]
        if "nnz" in w.extra:
            # SpMM path: also the HBM view of the same launches, from the
            # algorithmic bytes 16 (d + N) + 24 nnz / C per gradient evaluation and chain
            n_data = w.extra.get("data", w.extra.get("rays", 0))
            abytes = 16.0 * (w.dims + n_data) + 24.0 * w.extra["nnz"] / w.chains
            gbs = abytes * C / (kernel_ms * 1e-3) / 1e9
            roofline["kernel"] = "csr_spmm kernels (launch group = G q and G^T r of one gradient evaluation of the batch)"
            roofline["hbm"] = {"achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                               "frac": gbs / peaks["hbm_gbs"], "algorithmic_bytes_per_grad_eval": abytes}
            roofline["note"] = ("fp64 FMA work (4 nnz flop per gradient evaluation and chain) against the DGEMM peak "
                                "measured in this run: at nnz/row = 117 the fp64 pipe, not HBM, is the binding roof "
                                "(SURVEY 8d); the HBM view of the same launches is in `hbm`")
    roofline["timing"] = timing
    if traffic_rec:
        roofline["traffic_source"] = traffic_rec.get("source", "committed ncu capture (profiles/)")
        if traffic_rec.get("ncu"):     # pipe activity counters of the same committed capture
            roofline["ncu_counters"] = traffic_rec["ncu"]

    return {
        "baseline_config": BASELINE_CONFIG.get(name), "workload": w.description,
        "value": value, "unit": UNIT, "ms_per_step": ms_per_step, "steps": steps, "warmup": warmup,
        "chains_per_gpu": C, "chains_total": C * world, "dims": d, "proposals_per_step": B,
        "online_thinning": thin, "integrator": w.integrator, "amount_of_steps": w.amount_of_steps,
        "path": path, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
        "acceptance_rate": acc_rate, "wall_s_timed_region": wall, "ms_per_step_by_rank": per_rank_ms,
        "step_ms_rank0": {"min": min(step_ms), "median": float(np.median(step_ms)), "max": max(step_ms)},
    }


def sharding_check(ctx, per_rank=256, proposals=2):
    """Chains are keyed by their GLOBAL id: every rank advances its shard with the on-device
    random streams, rank 0 recomputes the 16 chains around every shard boundary on its own GPU
    through `chain_offset` and compares positions, misfits and decisions bit for bit."""
    torch, dist = ctx.torch, ctx.dist
    from hmclab_b200._engine import Engine
    from hmclab_b200 import workloads
    from hmclab_b200._lowering import describe, describe_mass, flatten

    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    cases = [("normal_iid", dict(dims=1000)), ("source_location", {}),
             ("dense_large", dict(dims=256, data=512)), ("tomography", dict(nx=20, ny=20, rays=1500))]
    report = {}
    ok = True
    for name, kw in cases:
        w = workloads.BUILDERS[name](chains=world * per_rank, **kw)
        plan, mass = flatten(describe(w.posterior)), describe_mass(w.mass_matrix)
        d = w.dims

        def run(q0, offset):
            C = q0.shape[0]
            eng = Engine(plan, mass, C, integrator=w.integrator, amount_of_steps=w.amount_of_steps,
                         device=ctx.local_rank)
            q = torch.as_tensor(q0, dtype=torch.float64).to(dev).contiguous()
            x = eng.misfit(q)
            acc = torch.zeros(proposals, C, dtype=torch.uint8, device=dev)
            eng.run_block(q, x, proposals, stepsize=w.stepsize, randomize_stepsize=True,
                          chain_offset=offset, seed=99, out_accept=acc)
            torch.cuda.synchronize()
            eng.close()
            return torch.cat([q, x[:, None], acc.t().double()], dim=1).contiguous()

        mine = run(w.initial_models[rank * per_rank:(rank + 1) * per_rank], rank * per_rank)
        everything = torch.empty(world * per_rank, d + 1 + proposals, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(everything, mine)
        if rank == 0:
            good = True
            for b in range(1, world):
                lo = b * per_rank - 8
                again = run(w.initial_models[lo:lo + 16], lo)
                good &= bool(torch.equal(again.view(torch.int64), everything[lo:lo + 16].view(torch.int64)))
            report[name] = "ok" if good else "fail"
            ok &= good
    if rank != 0:
        return None
    return {"result": "ok" if ok else "fail", "cases": report,
            "how": f"{world} ranks x {per_rank} chains x {proposals} proposals (device RNG); rank 0 recomputes the 16 "
                   "chains around each shard boundary via chain_offset; q, misfit and accept bits compared bit for bit"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="all",
                    help="`all` (headline configs[2] + per_config sub-records) or one workload name")
    ap.add_argument("--block", type=int, default=0, help="proposals per step (default: per workload)")
    ap.add_argument("--thinning", type=int, default=1)
    ap.add_argument("--chains", type=int, default=0, help="override chains per GPU")
    ap.add_argument("--cpu-seconds", type=float, default=8.0, help="CPU baseline sample per workload")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-peaks", action="store_true", help="skip the in-run fp64 / D2H ceiling measurements")
    args = ap.parse_args()
    if args.impl != "reference" and args.warmup < 3:
        args.warmup = 3   # timing rules: at least 3 warm-up steps (the line reports what was run)
    args.steps = max(1, args.steps)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    names = [HEADLINE] + SUB_WORKLOADS if args.workload == "all" else [args.workload]

    if args.impl == "reference":
        if rank == 0:
            run_reference(args, names)
        return

    # CPU baselines first: the worker processes are forked before CUDA is initialised
    cpu = {}
    if world == 1 and not args.no_cpu_baseline:
        arm = CpuArm(names, args.chains)
        for n in names:
            cpu[n] = arm.baseline(n, args.cpu_seconds)
        arm.close()

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; hmclab_b200 has no CPU path to time")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    ctx = Ctx()
    ctx.torch, ctx.dist, ctx.rank, ctx.world, ctx.local_rank = torch, dist, rank, world, local_rank
    ctx.dev = torch.device("cuda", local_rank)
    ctx.thinning, ctx.e2e = args.thinning, not args.no_e2e
    ctx.peaks = measured_peaks()
    ctx.traffic = committed_json("traffic_r02.json")
    ctx.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=ctx.dev)  # > 126 MB L2
    ctx.fp64 = {}
    ctx.d2h_ceiling = None
    if not args.no_peaks:
        ctx.fp64 = measure_fp64_peak(torch, ctx.dev, local_rank)   # every rank: same load on every GPU
        if ctx.e2e:
            ctx.d2h_ceiling = measure_d2h_ceiling(torch, dist, ctx.dev, world)
    if not ctx.fp64:
        prev = committed_json("fp64_peak_r01.json")
        ctx.fp64 = {"dgemm_tflops_burst": prev.get("dgemm_tflops", 37.0),
                    "dgemm_tflops_sustained": prev.get("dgemm_tflops", 37.0),
                    "dfma_ginst_per_s": prev.get("dfma_ginst_per_s", 16940.0),
                    "how": "NOT measured in this run (--no-peaks): profiles/fp64_peak_r01.json"}

    records = {}
    for n in names:
        head = n == names[0]
        steps = args.steps if head else min(args.steps, SUB_STEPS.get(n, args.steps))
        records[n] = bench_workload(ctx, n, steps, args.warmup, block=args.block if head or args.workload != "all" else 0,
                                    chains=args.chains)
    single = None
    if args.workload == "all" and world == 1:
        # configs[0] as BASELINE.json words it: ONE chain (latency-bound on any GPU)
        e2e_flag, ctx.e2e = ctx.e2e, False
        single = bench_workload(ctx, "dense_small", 10, 3, chains=1)
        ctx.e2e = e2e_flag
    check = sharding_check(ctx) if world > 1 else None

    if rank == 0:
        head = records[names[0]]
        if names[0] in cpu:
            head["cpu_baseline"] = cpu[names[0]]
        line = {
            "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": head["steps"],
            "warmup": head["warmup"], "ms_per_step": head["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": head["workload"], "baseline_config": head["baseline_config"],
                       "chains_per_gpu": head["chains_per_gpu"], "chains_total": head["chains_total"],
                       "dims": head["dims"], "proposals_per_step": head["proposals_per_step"],
                       "online_thinning": head["online_thinning"], "integrator": head["integrator"],
                       "amount_of_steps": head["amount_of_steps"], "path": head["path"],
                       "rng": "on-device Philox4x32-10", "l2": "256 MiB flush write between timed steps"},
            "e2e": head["e2e"], "gpu_launches": head["gpu_launches"], "clocks": head["clocks"],
            "roofline": head["roofline"], "cpu_baseline": head.get("cpu_baseline"),
            "acceptance_rate": head["acceptance_rate"], "wall_s_timed_region": head["wall_s_timed_region"],
            "ms_per_step_by_rank": head["ms_per_step_by_rank"], "step_ms_rank0": head["step_ms_rank0"],
            "fp64_peak": ctx.fp64, "host": {"numa": numa, "cores": os.cpu_count()},
        }
        if ctx.d2h_ceiling:
            line["d2h_ceiling"] = ctx.d2h_ceiling
        if len(names) > 1:
            per = {}
            for n in names[1:]:
                rec = records[n]
                rec["cpu_baseline"] = cpu.get(n)
                per[n] = rec
            if single is not None:
                per["dense_small"]["single_chain"] = {k: single[k] for k in ("value", "unit", "ms_per_step", "chains_per_gpu",
                                                                              "proposals_per_step", "path", "clocks")}
            line["per_config"] = per
            line["gpu_launches_all_configs"] = int(sum(r["gpu_launches"] for r in records.values()))
        if check is not None:
            line["sharding_check"] = check["result"]
            line["sharding_check_detail"] = check
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
