#!/bin/bash
# All four workloads at BASELINE size on one GPU + the CPU reference arm of the headline workload.
mkdir -p gpurun_out
for w in ola fft fir resample reechan; do
  timeout 500 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/r01_bench_$w.json 2> gpurun_out/r01_bench_$w.err
  tail -c 400 gpurun_out/r01_bench_$w.err
done
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01_bench_reference.json 2>gpurun_out/r01_bench_reference.err
cat gpurun_out/r01_bench_*.json | cut -c1-2000
