#!/bin/bash
# Weak-scaling evidence on N GPUs of one box (channels sharded, no data-path collective): usage run_bench_multi.sh N w1 w2 ...
N=$1; shift
mkdir -p gpurun_out
for w in "$@"; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r01_bench_${w}_${N}gpu.json 2> gpurun_out/r01_bench_${w}_${N}gpu.err
  tail -c 300 gpurun_out/r01_bench_${w}_${N}gpu.err | tail -2
  cut -c1-160 gpurun_out/r01_bench_${w}_${N}gpu.json
done
