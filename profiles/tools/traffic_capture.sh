#!/bin/bash
# DRAM traffic of ONE launch of a workload's dominant kernel AT THE BENCHED SIZE (bench.py's own command, full BASELINE
# shape): ncu with the two dram counters only (single pass, no replay of the 64 GiB working set), after the same command
# has run clean without ncu.  Writes gpurun_out/traffic_<workload>.csv; profiles/tools/traffic_json.py turns it into
# profiles/r02_traffic_<workload>.json, which bench.py reports as roofline.traffic.
# Usage (under gpurun): bash profiles/tools/traffic_capture.sh ola ols16k 3 [count]   (count kernels captured, default 2)
set -e
WL=$1; KREGEX=$2; SKIP=${3:-3}; COUNT=${4:-2}
CMD="python bench.py --workload $WL --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extra"
$CMD > gpurun_out/traffic_${WL}_plain.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --cache-control none -k regex:$KREGEX -s $SKIP -c $COUNT \
    --csv --log-file gpurun_out/traffic_${WL}.csv $CMD > gpurun_out/traffic_${WL}_ncu.log 2>&1
tail -1 gpurun_out/traffic_${WL}_plain.log | cut -c1-200
