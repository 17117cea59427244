#!/bin/bash
# Round-2 evidence (run under gpurun, one GPU): the ncu launch list of a short bench.py run and one --set full capture per hot
# kernel of the C3 and C2 pipelines, each after the identical plain command exited 0 (B200_PROFILING.md recipe).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --queries 100000000 --no-e2e --no-cpu-baseline --no-secondary"
$BENCH > $O/r2_launch_plain.json 2> $O/r2_launch_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_ncu.csv $BENCH > $O/r2_launch_ncu.log 2>&1
for CFG in C3 C2; do
  CMD="python tools/bench_pipeline.py --config $CFG --queries 50000000 --steps 1"
  KS="k_locate_ring k_scatter k_unpermute"
  [ $CFG = C3 ] && KS="$KS k_eval_binned_mma" || KS="$KS k_eval_binned_exact"
  [ -n "$ONLY" ] && KS="$ONLY"
  export CSP_BIN_STREAMS=1  # one chunk in flight: the captured launch runs alone (C2 overlaps two chunks by default)
  for K in $KS; do
    $CMD > $O/r2_plain_${CFG}_$K.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -o $O/r2_${K}_$CFG $CMD > $O/r2_ncu_${CFG}_$K.log 2>&1
  done
done
unset CSP_BIN_STREAMS
CMD="python tools/bench_fits.py C3"
$CMD > $O/r2_plain_fit.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_fit -s 3 -c 1 -o $O/r2_k_fit_C3 $CMD > $O/r2_ncu_fit.log 2>&1
ls -la $O/*.ncu-rep | awk '{print $5, $9}'
# K3b on C4 and the FP64 microbenchmarks (profiles/r2_k_eval_dmma_C4.*, r2_fp64_peaks.txt, r2_slab_peak.txt)
CMD="python tools/bench_dense.py --queries 20000000 --steps 1"
$CMD > $O/dense_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_eval_dmma -s 3 -c 1 -o $O/r2_k_eval_dmma_C4 $CMD > $O/dense_ncu.log 2>&1
./tools/ubench/fp64_peaks > $O/r2_fp64_peaks.txt
./tools/ubench/slab_peak > $O/r2_slab_peak.txt
