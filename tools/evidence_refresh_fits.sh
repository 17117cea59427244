#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
python bench.py > $O/ev_bench_n1.json 2> $O/ev_bench_n1.err
python bench.py --config C3 --steps 5 > $O/ev_bench_c3.json 2> $O/ev_bench_c3.err
python tools/bench_fits.py C2 C3 C5 C5det > $O/ev_fits.jsonl 2> $O/ev_fits.err; cut -c1-130 $O/ev_fits.jsonl
bash tools/prof_fit.sh C2 r1_k_fit
bash tools/prof_fit.sh C3 r1_k_fit_C3
python -c "
import json
for f in ('ev_bench_n1','ev_bench_c3'):
    b=json.load(open('gpurun_out/%s.json'%f)); print(f, b['value'], b['ms_per_step'], b['fits']['kernel_ms'], b['fits']['frac'])"
