#!/bin/bash
# Profiling recipe of the staged FFT-domain filter (run under gpurun, one GPU): plain run (must exit 0), the ncu launch
# list of the same command with per-launch device time and DRAM bytes (caches NOT flushed between launches, so the
# L2-resident scratch behaves as in a real run), then one --set full capture of one launch of each stage kernel.
set -x
mkdir -p gpurun_out
CMD="python bench.py --workload ola --scale 0.0157 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_ola.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none \
    -k regex:"ola64k_stage|carry_update" -s 273 -c 91 --csv --log-file gpurun_out/launches_ola.csv $CMD > gpurun_out/ncu_l_ola.log 2>&1
$CMD > gpurun_out/plain2_ola.log 2>&1 && \
ncu --set full --clock-control none --cache-control none --import-source on -k regex:ola64k_stage -s 303 -c 3 -f -o gpurun_out/prof_ola $CMD > gpurun_out/ncu_f_ola.log 2>&1
tail -2 gpurun_out/ncu_l_ola.log; tail -2 gpurun_out/ncu_f_ola.log; ls -la gpurun_out | head -30
