#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
python bench.py > $O/ev_bench_n1.json 2> $O/ev_bench_n1.err
python bench.py --config C3 --steps 5 > $O/ev_bench_c3.json 2> $O/ev_bench_c3.err
bash tools_gpu_profile.sh r1 k_locate_only
python tools/bench_dense.py > $O/ev_dense_c4.json 2>> $O/ev_dense.err
python tools/bench_dense.py --n 64 > $O/ev_dense_n64.json 2>> $O/ev_dense.err
python -c "
import json
for f in ('ev_bench_n1','ev_bench_c3'):
    b=json.load(open('gpurun_out/%s.json'%f)); print(f, b['value'], b['ms_per_step'], b['roofline']['frac'], b['roofline']['kernel'][:40])"
