#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
CMD="python bench.py --config C3 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --queries 40000000"
$CMD > gpurun_out/c3_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_eval_binned -s 1 -c 1 -o gpurun_out/r1_k_eval_binned_C3 $CMD > gpurun_out/c3_ncu.log 2>&1
tail -2 gpurun_out/c3_ncu.log
