for c in 524288 1048576 2097152 4194304; do echo chunk $c; CSP_BIN_CHUNK=$c python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']/1e9,2), round(d['locate_eval_ms'],2), {k:(round(v['ms_per_step'],2), v['launches_per_step']) for k,v in d['roofline']['kernels'].items()})"; done
