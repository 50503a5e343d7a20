#!/bin/bash
# Round-1 profiling recipe (run under gpurun, one GPU).  For each workload: plain run (must exit 0),
# then the ncu launch list of the same command, then one `--set full` capture of the dominant kernel.
set -x
mkdir -p gpurun_out
declare -A KERN=( [ola]=ola64k [fft]=fft64k [fir]=fir_direct [resample]=resamp_banded )
declare -A SCALE=( [ola]=0.0157 [fft]=0.0625 [fir]=0.0625 [resample]=0.125 )
declare -A SKIP=( [ola]=3 [fft]=6 [fir]=48 [resample]=24 )
for w in ola fft fir resample; do
  CMD="python bench.py --workload $w --scale ${SCALE[$w]} --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
  $CMD > gpurun_out/plain_$w.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$w.csv $CMD > gpurun_out/ncu_l_$w.log 2>&1
  $CMD > gpurun_out/plain2_$w.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:${KERN[$w]} -s ${SKIP[$w]} -c 1 -o gpurun_out/prof_$w $CMD > gpurun_out/ncu_f_$w.log 2>&1
  tail -1 gpurun_out/ncu_f_$w.log
done
ls -la gpurun_out
