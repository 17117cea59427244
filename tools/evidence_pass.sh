#!/bin/bash
# Round evidence, part 1 (plain runs, never under a profiler): GPU tests, smoke, bench lines and the side benchmarks.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/ev_tests.log 2>&1; tail -2 $O/ev_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/ev_smoke.log 2>&1; tail -1 $O/ev_smoke.log
python bench.py --impl reference > $O/ev_bench_reference.json 2> $O/ev_bench_reference.err; tail -c 400 $O/ev_bench_reference.json
python bench.py > $O/ev_bench_n1.json 2> $O/ev_bench_n1.err; tail -c 600 $O/ev_bench_n1.json
python bench.py --config C3 --steps 5 > $O/ev_bench_c3.json 2> $O/ev_bench_c3.err; tail -c 300 $O/ev_bench_c3.json
python tools/bench_dense.py > $O/ev_dense_c4.json 2>> $O/ev_dense.err
python tools/bench_dense.py --n 64 > $O/ev_dense_n64.json 2>> $O/ev_dense.err
python tools/bench_dense.py --n 300 --leaves 1000 --queries 2000000 > $O/ev_dense_n300.json 2>> $O/ev_dense.err
cat $O/ev_dense_c4.json | cut -c1-300
python tools/bench_c4.py > $O/ev_c4.json 2> $O/ev_c4.err; cut -c1-200 $O/ev_c4.json
python tools/bench_fits.py C2 C3 C5 > $O/ev_fits.jsonl 2> $O/ev_fits.err; cut -c1-140 $O/ev_fits.jsonl
python tools/bench_update.py > $O/ev_update.json 2> $O/ev_update.err; cat $O/ev_update.json
tools/ubench/fp64_peaks > $O/ev_fp64_peaks.txt 2>&1; head -3 $O/ev_fp64_peaks.txt
