#!/bin/bash
# FIR tensor-core kernel variants: parity first, then A/B on one box
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fir" 2>&1 | tail -5
timeout 900 bash profiles/ab.sh fir 1 "TSDGPU_FIR_TC_VARIANT=1" "TSDGPU_FIR_TC_VARIANT=2" "TSDGPU_FIR_TC_VARIANT=3"
