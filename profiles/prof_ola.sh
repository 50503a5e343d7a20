set -x
CMD="env TSDGPU_OLA_LAG=40 TSDGPU_OLA_RING=128 python bench.py --workload ola --scale 0.0157 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_ola.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_ola.csv $CMD > gpurun_out/ncu_l_ola.log 2>&1
$CMD > gpurun_out/plain_ola2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ola64k -s 3 -c 1 -o gpurun_out/prof_ola $CMD > gpurun_out/ncu_f_ola.log 2>&1
tail -2 gpurun_out/plain_ola.log; tail -5 gpurun_out/ncu_f_ola.log; ls -la gpurun_out
