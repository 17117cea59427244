#!/bin/bash
# ncu capture of the leaf-fit kernel (K2) on config C2 or C3: plain run first, then the identical command under ncu.
# usage: tools/prof_fit.sh [C2|C3] [tag]
cd "${GRAFT_REPO_ROOT:-/root/repo}"
CFG=${1:-C2}; TAG=${2:-r1_k_fit_$CFG}
CMD="python tools/bench_fits.py $CFG"
$CMD > gpurun_out/fit_plain_$CFG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_fit -s 3 -c 1 -o gpurun_out/$TAG $CMD > gpurun_out/fit_ncu_$CFG.log 2>&1
