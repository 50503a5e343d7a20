#!/bin/bash
# A/B runs on ONE box: usage ab.sh <workload> <scale> "ENV=.. ENV=.." "ENV=.." ...   (each config run twice, interleaved)
w=$1; sc=$2; shift 2
for rep in 1 2; do
  for cfg in "$@"; do
    echo -n "$w [$cfg] : "
    env $cfg python bench.py --workload $w --scale $sc --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-extra 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print(round(d['value'],1), 'Gs/s  frac', round(r['frac'],3), ' kernel ms', round(r['kernel_ms_per_step'],2), ' launches', d['gpu_launches'])"
  done
done
