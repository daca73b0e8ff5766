#!/bin/bash
# Round-2 profiling pass (run on a B200 under gpurun): per workload the ncu launch list of one bench
# step (gpu__time_duration per launch: shares, not absolutes) and one `--set full` capture of the
# dominant kernel.  Outputs land in gpurun_out/; profiles/tools/summarize_ncu.py turns the reports
# into the text summaries committed under profiles/.
set -u
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-peaks"
for w in dense_large tomography normal_iid source_location dense_small; do
  extra=""; [ $w = dense_small ] && extra="--chains 4096"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
      --log-file gpurun_out/launches_${w}_r02.csv $B --workload $w $extra > /dev/null 2> gpurun_out/launches_${w}.err
done
cap() {  # name workload kernel-regex skip count [extra]; the report is summarised here and dropped
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$3 -s $4 -c $5 -f \
      -o gpurun_out/ncu_$1_r02 $B --workload $2 ${6:-} > /dev/null 2> gpurun_out/ncu_$1.err
  python profiles/tools/summarize_ncu.py gpurun_out/ncu_$1_r02.ncu-rep > gpurun_out/ncu_$1_r02.txt 2>&1
  rm -f gpurun_out/ncu_$1_r02.ncu-rep
}
cap dmma_gemm dense_large dmma_gemm 8 2
cap spmm_block tomography csr_spmm_block 6 2
cap fused_priors normal_iid hmc_fused_priors 3 1
cap fused_srcloc source_location hmc_fused_srcloc 3 1
cap fused_dense dense_small hmc_fused_dense 3 1 "--chains 4096"
ls -la gpurun_out/ncu_*_r02.txt gpurun_out/launches_*_r02.csv
