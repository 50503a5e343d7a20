run() { timeout 100 python bench.py --workload $1 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/tmp/err.txt | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', '$2', round(d['value'],1), round(d['roofline']['kernel_ms_per_step'],2), d['clocks']['power_w_max'])"; }
for cfg in "24 80" "32 96" "20 64"; do set -- $cfg; TSDGPU_OLA_LAG=$1 TSDGPU_OLA_RING=$2 run ola "lag=$1,ring=$2"; done
for cfg in "24 64" "32 64" "40 80"; do set -- $cfg; TSDGPU_FFT_LAG=$1 TSDGPU_FFT_RING=$2 run fft "lag=$1,ring=$2"; done
