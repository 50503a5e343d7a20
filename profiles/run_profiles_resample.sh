#!/bin/bash
# Resampler (BASELINE config 5) profiling recipe, run under gpurun on one GPU: for the tensor-core kernel (default) and
# the FP32 FMA kernel (TSDGPU_RESAMP_TC=0): plain run (must exit 0), launch list with device time and DRAM bytes, one
# --set full capture of one launch of the dominant kernel.  Full BASELINE size: 512 channels x 8 Mi samples per step.
set -x
mkdir -p gpurun_out
for tc in 1 0; do
  CMD="env TSDGPU_RESAMP_TC=$tc python bench.py --workload resample --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
  $CMD > gpurun_out/plain_resample_tc$tc.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none \
      -k regex:"resamp_" -s 27 -c 18 --csv --log-file gpurun_out/launches_resample_tc$tc.csv $CMD > gpurun_out/ncu_l_resample_tc$tc.log 2>&1
  $CMD > gpurun_out/plain2_resample_tc$tc.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:"resamp_tc_kernel|resamp_banded" -s 14 -c 1 -f -o gpurun_out/prof_resample_tc$tc $CMD > gpurun_out/ncu_f_resample_tc$tc.log 2>&1
  tail -1 gpurun_out/ncu_f_resample_tc$tc.log
done
