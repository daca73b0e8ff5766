"""fp64 roofline denominators measured on the GPU box (MEASURED_PEAKS.json has none):
cuBLAS DGEMM through torch.matmul (8192^3, best of 5), plus the DFMA / DMMA microbenchmarks
of fp64_peak.cu.  Writes gpurun_out/fp64_peak.json; the committed copy is
profiles/fp64_peak_r01.json."""
import json
import os
import subprocess
import sys

import torch

here = os.path.dirname(os.path.abspath(__file__))
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda")
b = torch.randn(n, n, dtype=torch.float64, device="cuda")
torch.matmul(a, b)
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    torch.matmul(a, b)
    e.record()
    torch.cuda.synchronize()
    best = min(best, s.elapsed_time(e))
out = {"dgemm_tflops": 2.0 * n ** 3 / best / 1e9, "dgemm_ms": best, "dgemm_n": n,
       "how": "torch.matmul float64 8192^3 best of 5 (cuBLAS DGEMM); fp64_peak.cu DFMA/DMMA loops"}
micro = subprocess.check_output([os.path.join(here, "fp64_peak")], text=True).strip().splitlines()[-1]
out.update(json.loads(micro))
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/fp64_peak.json", "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out))
