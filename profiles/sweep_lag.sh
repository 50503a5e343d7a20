run() { timeout 100 python bench.py --workload $1 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/tmp/err.txt | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', '$2', round(d['value'],1), round(d['roofline']['kernel_ms_per_step'],2))"; }
for cfg in "16 64" "12 48" "24 80" "8 32"; do set -- $cfg; TSDGPU_OLA_LAG=$1 TSDGPU_OLA_RING=$2 run ola "lag=$1,ring=$2"; done
for cfg in "24 64" "16 48" "32 64" "12 32"; do set -- $cfg; TSDGPU_FFT_LAG=$1 TSDGPU_FFT_RING=$2 run fft "lag=$1,ring=$2"; done
