#!/bin/bash
# Profiling recipe used for profiles/ (run under gpurun): plain run first, then ncu on the identical command line.
# usage: tools_gpu_profile.sh <tag> [kernel-regex ...]
cd "${GRAFT_REPO_ROOT:-/root/repo}"
TAG=${1:-prof}; shift
SMALL="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --queries 20000000"
$SMALL > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv $SMALL > gpurun_out/${TAG}_ncu_launches.log 2>&1
for K in "$@"; do
  $SMALL > gpurun_out/${TAG}_plain_$K.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -o gpurun_out/${TAG}_$K $SMALL > gpurun_out/${TAG}_ncu_$K.log 2>&1
done
