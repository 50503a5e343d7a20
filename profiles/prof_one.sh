#!/bin/bash
# usage: prof_one.sh <workload> <kernel regex> <scale> <skip> <tag>   (one --set full capture of one launch; run under gpurun)
w=$1; k=$2; sc=$3; skip=$4; tag=$5
mkdir -p gpurun_out
CMD="python bench.py --workload $w --scale $sc --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_$tag.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/prof_$tag $CMD > gpurun_out/ncu_$tag.log 2>&1
tail -2 gpurun_out/plain_$tag.log | cut -c1-300; tail -3 gpurun_out/ncu_$tag.log
