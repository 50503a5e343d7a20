#!/bin/bash
for cfg in "TSDGPU_OLA_STREAMS=4" "TSDGPU_OLA_STREAMS=3" "TSDGPU_OLA_STREAMS=5" "TSDGPU_OLA_STREAMS=6" "TSDGPU_OLA_STREAMS=8"; do
  echo -n "[$cfg] : "
  env $cfg python bench.py --workload ola --scale 0.5 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print(round(d['value'],1), 'Gs/s  frac', round(r['frac'],3))"
done
