#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
CMD="python tools/bench_dense.py --queries 20000000 --steps 1"
$CMD > gpurun_out/dense_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_eval_dmma -s 3 -c 1 -o gpurun_out/r1_k_eval_dmma $CMD > gpurun_out/dense_ncu.log 2>&1
