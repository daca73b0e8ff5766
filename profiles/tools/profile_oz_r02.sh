#!/bin/bash
# Round-2 profiling of the config-3 headline path (Ozaki-sliced products on tcgen05), run on a B200
# under gpurun: (1) launch list of one bench step, (2) DRAM traffic + duration + tensor-pipe activity of
# every kernel of one gradient evaluation (the launch group bench.py times), (3) `--set full` captures
# of the slice-product kernel at both shapes (profiles/tools/ncu_i8_gemm.sh), (4) SASS opcode counts.
set -u
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-peaks --workload dense_large"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_dense_large_r02.csv $B > /dev/null 2> gpurun_out/launches_dense_large.err
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum \
    --clock-control none -k regex:"oz_|i8_gemm" -s 40 -c 24 --csv --log-file gpurun_out/group_dense_large_oz_r02.csv \
    $B > /dev/null 2> gpurun_out/group_dense_large_oz.err
python - <<'PY'
import csv, json
rows = [r for r in csv.reader(open("gpurun_out/group_dense_large_oz_r02.csv")) if len(r) > 10]
hdr = rows[0]
ik, im, iv, iu, iid = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
launches = {}
for r in rows[1:]:
    d = launches.setdefault(int(r[iid]), {"kernel": r[ik]})
    v = float(r[iv].replace(",", ""))
    unit = r[iu]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(unit, 1.0)
    d[r[im]] = v * scale
seq = [launches[k] for k in sorted(launches)]
want = ["oz_colmax", "oz_slice_chains", "i8_gemm", "ResidualEpi", "oz_slice_chains", "i8_gemm", "UpdateEpi"]
for s in range(len(seq) - 6):
    if all(w in seq[s + j]["kernel"] for j, w in enumerate(want)):
        group = seq[s: s + 7]
        break
else:
    raise SystemExit("no complete gradient evaluation in the capture")
out = {"kernels": [{"kernel": g["kernel"][:70], "ms": g["gpu__time_duration.sum"] * 1e3,
                    "dram_read": g["dram__bytes_read.sum"], "dram_write": g["dram__bytes_write.sum"],
                    "l2_bytes": g["lts__t_bytes.sum"],
                    "tensor_pipe_active_pct": g["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]}
                   for g in group]}
out["dram_bytes_read"] = sum(k["dram_read"] for k in out["kernels"])
out["dram_bytes_write"] = sum(k["dram_write"] for k in out["kernels"])
out["group_ms_under_ncu"] = sum(k["ms"] for k in out["kernels"])
json.dump(out, open("gpurun_out/traffic_dense_large_oz_r02.json", "w"), indent=1)
print(json.dumps(out, indent=1))
PY
bash profiles/tools/ncu_i8_gemm.sh
cuobjdump -sass hmclab_b200/lib/libhmcb.so | grep -oE "^\s+/\*[0-9a-f]+\*/\s+[A-Z0-9_.]+" | awk '{print $2}' | sort | uniq -c | sort -rn \
    | grep -E "UTC|LDTM|UTMA|UBLKCP|DMMA|SYNCS" > gpurun_out/sass_counts_r02.txt
