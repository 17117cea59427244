#!/bin/bash
# Profiling recipe used for profiles/ (run under gpurun): plain run first, then ncu on the identical command line.
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
SMALL="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --queries 20000000"
$SMALL > gpurun_out/plain_small.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv $SMALL > gpurun_out/ncu_launches.log 2>&1
$SMALL > gpurun_out/plain_small2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_locate -s 3 -c 1 -o gpurun_out/prof_locate $SMALL > gpurun_out/ncu_locate.log 2>&1
$SMALL > gpurun_out/plain_small3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_fit -s 3 -c 1 -o gpurun_out/prof_fit $SMALL > gpurun_out/ncu_fit.log 2>&1
