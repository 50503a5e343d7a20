#!/bin/bash
# round-2 refresh after the tensor-map forms of fir_tc / resamp_tc: GPU tests, launch lists + full ncu captures, DRAM traffic
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_tests.log 2>&1; tail -3 gpurun_out/r02_tests.log
bash profiles/run_profiles_r02.sh "fir resample" > gpurun_out/r02_prof.log 2>&1
bash profiles/tools/traffic_capture.sh fir fir_tc 48 16 2>&1 | tail -2
bash profiles/tools/traffic_capture.sh resample resamp_tc 24 8 2>&1 | tail -2
ls -la gpurun_out | grep -E "r02_prof_(fir|resample)|traffic_(fir|resample).csv"
