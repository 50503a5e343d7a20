#!/bin/bash
# resampler tensor-core kernel: parity of both forms, then A/B on one box
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_ra_kat.py -m gpu -x -q -k "itrp or reechan or resamp or tensor or ra_kat" 2>&1 | tail -5
timeout 900 bash profiles/ab.sh resample 1 "TSDGPU_LIB=profiles/variants/libtsdgpu_span12.so" "A=1"
timeout 600 bash profiles/ab.sh reechan 1 "TSDGPU_LIB=profiles/variants/libtsdgpu_span12.so" "A=1"
