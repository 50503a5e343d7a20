#!/bin/bash
# Round-2 profiling recipe (run under gpurun, one GPU).  For each workload: plain run (must exit 0), then the ncu launch
# list of the same command, then one `--set full` capture of the dominant kernel (reduced channel count so that ncu's
# ~40 replays of one launch stay short; DRAM traffic at the FULL size comes from profiles/tools/traffic_capture.sh).
# profiles/make_summaries_r02.py turns gpurun_out/r02_* into the tracked summaries under profiles/.
set -x
mkdir -p gpurun_out
declare -A KERN=( [ola]=ols16k [fft]=fft64k [fir]=fir_tc [resample]=resamp_tc )
declare -A SCALE=( [ola]=0.125 [fft]=0.0625 [fir]=0.25 [resample]=0.125 )
declare -A SKIP=( [ola]=3 [fft]=6 [fir]=3 [resample]=3 )
for w in ${1:-ola fft fir resample}; do
  CMD="python bench.py --workload $w --scale ${SCALE[$w]} --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extra"
  timeout 300 $CMD > gpurun_out/r02_plain_$w.log 2>&1 && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_$w.csv $CMD > gpurun_out/r02_ncu_l_$w.log 2>&1
  timeout 300 $CMD > gpurun_out/r02_plain2_$w.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:${KERN[$w]} -s ${SKIP[$w]} -c 1 -f -o gpurun_out/r02_prof_$w $CMD > gpurun_out/r02_ncu_f_$w.log 2>&1
  tail -1 gpurun_out/r02_ncu_f_$w.log
done
ls -la gpurun_out | grep r02_
