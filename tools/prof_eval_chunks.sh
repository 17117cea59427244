cd "${GRAFT_REPO_ROOT:-/root/repo}"
SMALL="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --queries 20000000"
for c in 1048576 4194304; do
export CSP_BIN_CHUNK=$c
$SMALL > gpurun_out/r1f_plain_$c.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_eval_binned -s 40 -c 2 -o gpurun_out/r1f_eval_$c $SMALL > gpurun_out/r1f_ncu_$c.log 2>&1
done
