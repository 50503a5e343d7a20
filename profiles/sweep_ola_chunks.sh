#!/bin/bash
for cfg in "TSDGPU_OLA_CHUNK=32 TSDGPU_OLA_STREAMS=4" "TSDGPU_OLA_CHUNK=64 TSDGPU_OLA_STREAMS=2" "TSDGPU_OLA_CHUNK=64 TSDGPU_OLA_STREAMS=3" "TSDGPU_OLA_CHUNK=48 TSDGPU_OLA_STREAMS=3" "TSDGPU_OLA_CHUNK=96 TSDGPU_OLA_STREAMS=2" "TSDGPU_OLA_CHUNK=40 TSDGPU_OLA_STREAMS=4" "TSDGPU_OLA_CHUNK=24 TSDGPU_OLA_STREAMS=6" "TSDGPU_OLA_CHUNK=37 TSDGPU_OLA_STREAMS=4" "TSDGPU_OLA_CHUNK=74 TSDGPU_OLA_STREAMS=3"; do
  echo -n "[$cfg] : "
  env $cfg python bench.py --workload ola --scale 0.5 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print(round(d['value'],1), 'Gs/s  frac', round(r['frac'],3))"
done
