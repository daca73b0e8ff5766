#!/bin/bash
# One bench.py line per BASELINE.json config (1 GPU) -> gpurun_out/bench_all.jsonl
# (the committed copy is profiles/bench_all_r01.jsonl).
out=gpurun_out/bench_all.jsonl
: > $out
python bench.py --steps 20 --warmup 3 >> $out
python bench.py --workload dense_small --chains 4096 --steps 10 --warmup 3 >> $out
python bench.py --workload dense_large --steps 3 --warmup 3 --cpu-seconds 10 >> $out
python bench.py --workload dense_large_premult --steps 3 --warmup 3 --cpu-seconds 10 >> $out
python bench.py --workload tomography --steps 3 --warmup 3 --cpu-seconds 10 >> $out
python bench.py --workload source_location --steps 10 --warmup 3 >> $out
python - <<'PY'
import json
for line in open("gpurun_out/bench_all.jsonl"):
    d = json.loads(line)
    r, e, c = d["roofline"], d["e2e"], d["cpu_baseline"]
    print(f'{d["config"]["workload"][:60]:60s} value {d["value"]:.3e} e2e {e["value"] if e else float("nan"):.3e} '
          f'cpu1 {c["value"] if c else float("nan"):.3e} roof {r["bound"]} {r["frac"]:.3f}')
PY
