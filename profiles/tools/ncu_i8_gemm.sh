#!/bin/bash
# ncu --set full captures of the tcgen05 int8 slice-product kernel at the two config-3 shapes
# (profiles/tools/i8_gemm_rate.py launches each shape 7 times); raw counter pages kept as CSV.
set -u
mkdir -p gpurun_out
for spec in "gq 2" "gtr 9"; do
  set -- $spec
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:i8_gemm -s $2 -c 1 -f \
      -o gpurun_out/ncu_i8_gemm_$1_r02 python profiles/tools/i8_gemm_rate.py > /dev/null 2> gpurun_out/ncu_i8_gemm_$1_r02.err
  python profiles/tools/summarize_ncu.py gpurun_out/ncu_i8_gemm_$1_r02.ncu-rep > gpurun_out/ncu_i8_gemm_$1_r02.txt 2>&1
  ncu -i gpurun_out/ncu_i8_gemm_$1_r02.ncu-rep --page raw --csv > gpurun_out/ncu_i8_gemm_$1_r02_raw.csv 2>/dev/null
  rm -f gpurun_out/ncu_i8_gemm_$1_r02.ncu-rep
done
