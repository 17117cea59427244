P='import json,sys; d=json.loads(sys.stdin.read()); print(round(d["value"]/1e9,3), round(d["locate_eval_ms"],2), {k:(round(v["ms_per_step"],2), v["launches_per_step"]) for k,v in d["roofline"]["kernels"].items()})'
B="python bench.py --config C3 --queries 50000000 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
echo direct; CSP_EVAL_MODE=direct $B 2>&1 | python -c "$P"
for c in 2097152 8388608 16777216 50000000; do echo chunk $c; CSP_BIN_CHUNK=$c $B 2>&1 | python -c "$P"; done
