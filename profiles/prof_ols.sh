#!/bin/bash
# ncu capture of the single-SM overlap-save kernel (profiles/ols_quick.py time <nchan> <log2n>); run under gpurun
set -e
NCH=${1:-128}; LG=${2:-22}
python profiles/ols_quick.py time $NCH $LG > gpurun_out/ols_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ols16k -s 3 -c 1 -o gpurun_out/ols_prof -f python profiles/ols_quick.py time $NCH $LG > gpurun_out/ols_ncu.log 2>&1
cat gpurun_out/ols_plain.log
