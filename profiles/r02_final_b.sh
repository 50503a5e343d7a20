#!/bin/bash
# round-2 final refresh: GPU tests, smoke, ola profile (launch list + full ncu), ola DRAM traffic, default bench line
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_tests.log 2>&1; tail -2 gpurun_out/r02_tests.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r02_smoke.log 2>&1; tail -3 gpurun_out/r02_smoke.log
bash profiles/run_profiles_r02.sh "ola" > gpurun_out/r02_prof.log 2>&1
bash profiles/tools/traffic_capture.sh ola ols16k 3 2 2>&1 | tail -1 | cut -c1-120
START=$(date +%s); python bench.py > gpurun_out/r02_bench_full.json 2> gpurun_out/r02_bench_full.err; echo bench rc=$? elapsed=$(( $(date +%s) - START ))s
