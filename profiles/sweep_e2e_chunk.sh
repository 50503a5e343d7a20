#!/bin/bash
for mb in 48 16 8 24; do for w in ola fir; do echo -n "$w chunk ${mb}MB: "; TSDGPU_HOST_CHUNK_MB=$mb python bench.py --workload $w --scale 0.1 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['e2e']['value'],3))"; done; done
