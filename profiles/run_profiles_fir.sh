#!/bin/bash
# FIR (BASELINE config 3) profiling recipe, run under gpurun on one GPU: for the tensor-core kernel (default) and the
# FP32 FMA kernel (TSDGPU_FIR_TC=0): plain run (must exit 0), launch list with device time and DRAM bytes, one
# --set full capture of one launch of the dominant kernel.  Full BASELINE size: 1024 channels x 64 Ki samples per launch.
set -x
mkdir -p gpurun_out
for tc in 1 0; do
  CMD="env TSDGPU_FIR_TC=$tc python bench.py --workload fir --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
  $CMD > gpurun_out/plain_fir_tc$tc.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none \
      -k regex:"fir_tc|fir_direct|fir_hist" -s 96 -c 32 --csv --log-file gpurun_out/launches_fir_tc$tc.csv $CMD > gpurun_out/ncu_l_fir_tc$tc.log 2>&1
  $CMD > gpurun_out/plain2_fir_tc$tc.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:"fir_tc_kernel|fir_direct" -s 50 -c 1 -f -o gpurun_out/prof_fir_tc$tc $CMD > gpurun_out/ncu_f_fir_tc$tc.log 2>&1
  tail -1 gpurun_out/ncu_f_fir_tc$tc.log
done
