#!/bin/bash
# per-kernel duration / tensor-pipe activity / L2 and DRAM bytes of the Ozaki launch groups of one config-3 step
set -u
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none -k regex:"oz_|i8_gemm" -s 40 -c 8 --csv --log-file gpurun_out/group_timing.csv \
  python bench.py --workload dense_large --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-peaks > /dev/null 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/group_timing.csv")) if len(r) > 10]
h = rows[0]
ik, im, iv, iu, iid = (h.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
d = {}
for r in rows[1:]:
    d.setdefault(int(r[iid]), {"k": r[ik][:50]})[r[im][:28]] = r[iv] + " " + r[iu]
for k in sorted(d):
    print(d[k])
PY
