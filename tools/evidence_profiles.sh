#!/bin/bash
# Round evidence, part 2: ncu launch list + one --set full capture per hot kernel (each after the identical plain run).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
bash tools_gpu_profile.sh r1 k_locate_only k_eval_binned k_scatter k_unpermute k_fit_chain
bash tools/prof_fit.sh C2 r1_k_fit
bash tools/prof_fit.sh C3 r1_k_fit_C3
bash tools/prof_dense.sh
bash tools/prof_c3_eval.sh
CMD="python tools/bench_c4.py --steps 2"
$CMD > gpurun_out/c4_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_locate_wide -s 2 -c 1 -o gpurun_out/r1_k_locate_wide_C4 $CMD > gpurun_out/c4_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'
