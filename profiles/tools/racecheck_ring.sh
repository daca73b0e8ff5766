#!/bin/bash
# Builds the minimal ring and runs it plain and under compute-sanitizer racecheck / memcheck in both
# modes -> gpurun_out/racecheck_ring_r02.log (committed as profiles/sanitizer_racecheck_ring_r02.log)
set -u
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -o profiles/tools/racecheck_ring profiles/tools/racecheck_ring.cu || exit 1
{
  compute-sanitizer --version | head -3
  for mode in 0 1; do
    echo "=== plain run, mode $mode"; ./profiles/tools/racecheck_ring $mode
    echo "=== racecheck, mode $mode"; timeout 300 compute-sanitizer --tool racecheck --racecheck-report all ./profiles/tools/racecheck_ring $mode 2>&1 | tail -40
    echo "=== memcheck, mode $mode"; timeout 300 compute-sanitizer --tool memcheck ./profiles/tools/racecheck_ring $mode 2>&1 | tail -5
  done
} > gpurun_out/racecheck_ring_r02.log 2>&1
tail -60 gpurun_out/racecheck_ring_r02.log
