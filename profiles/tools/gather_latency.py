"""Per-block diagnostics gather: cost of one NCCL all_gather of the acceptance counters
next to one block of proposals (run under torchrun, 2+ ranks)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hmclab_b200 import workloads  # noqa: E402
from hmclab_b200._engine import Engine  # noqa: E402
from hmclab_b200._lowering import describe, describe_mass, flatten  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
w = workloads.normal_iid()
eng = Engine(flatten(describe(w.posterior)), describe_mass(w.mass_matrix), w.chains, device=local)
q = torch.as_tensor(w.initial_models).cuda().contiguous()
x = eng.misfit(q)
acc = torch.zeros(w.chains, dtype=torch.int32, device="cuda")
allacc = torch.zeros(world * w.chains, dtype=torch.int32, device="cuda")
samples = torch.empty(10, w.chains, w.dims + 1, dtype=torch.float64, device="cuda")


def timed(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


def block():
    eng.run_block(q, x, 10, stepsize=0.05, seed=1, out_samples=samples, accepted_total=acc)


def gather():
    dist.all_gather_into_tensor(allacc, acc)


side = torch.cuda.Stream()


def both_side_stream():
    block()
    ev = torch.cuda.Event()
    ev.record()
    with torch.cuda.stream(side):
        side.wait_event(ev)
        dist.all_gather_into_tensor(allacc, acc)


res = {"block_ms": timed(block), "gather_ms": timed(gather),
       "block_then_gather_ms": timed(lambda: (block(), gather())),
       "block_gather_side_stream_ms": timed(both_side_stream)}
if rank == 0:
    print(res)
dist.destroy_process_group()
