"""Turn an .ncu-rep (ncu --set full --import-source on) into the small text summary that is
committed under profiles/: headline counters, opcode mix and stall reasons.

    python profiles/tools/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/prof.txt
"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_global_ld.sum"]


def ncu(page):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True,
                          text=True).stdout


rows = list(csv.reader(io.StringIO(ncu("raw"))))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print(f"kernel: {r[4]}\n  grid {r[8]} block {r[7]}")
    for h, u, v in zip(hdr, units, r):
        if h in KEYS:
            print(f"  {h} [{u}] = {v}")

text = ncu("source")
blocks = text.split('"Kernel Name"')
for blk in blocks[1:]:
    rows = list(csv.reader(io.StringIO('"Kernel Name"' + blk)))
    name = rows[0][1] if len(rows[0]) > 1 else "?"
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    ops, samples = collections.Counter(), collections.Counter()
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    stalls = collections.Counter()
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        try:
            n, s = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
        except ValueError:
            continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]])
        op = m.group(2) if m else "?"
        key = ".".join(op.split(".")[:2]) if op.startswith(("MUFU", "LDG", "STG", "LDL", "STL", "LDS", "STS", "DMMA", "LDGSTS")) else op.split(".")[0]
        ops[key] += n
        samples[key] += s
        for h in stall_cols:
            try:
                stalls[h] += int(r[ix[h]])
            except ValueError:
                pass
    tot, tots = sum(ops.values()), max(sum(samples.values()), 1)
    print(f"\nSASS opcode mix of {name[:90]}\n  warp-level instructions executed: {tot}")
    for k, v in ops.most_common(18):
        print(f"  {k:12s} {v:12d} {100 * v / tot:5.1f}%   pc samples {100 * samples[k] / tots:5.1f}%")
    st = max(sum(stalls.values()), 1)
    print("  stall reasons (all samples): " + ", ".join(f"{k[6:]} {100 * v / st:.0f}%" for k, v in stalls.most_common(6)))
