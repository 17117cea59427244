#!/usr/bin/env python
"""Regenerates profiles/README.md from the committed measurement files (run from the repo root)."""
import csv
import json
from collections import defaultdict

rows = list(csv.reader(open('profiles/r1_launches_ncu.csv')))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[hdr]; ki = H.index('Kernel Name'); vi = H.index('Metric Value')
agg = defaultdict(list)
for r in rows[hdr + 1:]:
    if len(r) > vi:
        agg[r[ki][:70]].append(float(r[vi].replace(',', '')))
mine = {k: v for k, v in agg.items() if not ('at::' in k or 'at_cuda' in k)}
tot = sum(sum(v) for v in mine.values())
lines = [f"| `{k}` | {len(v)} | {sum(v)/len(v)/1e3:.1f} | {100*sum(v)/tot:.1f} % |" for k, v in mine.items()]
b = json.load(open('profiles/r1_bench_n1.json'))
b2 = json.load(open('profiles/r1_bench_n2.json'))
b8 = json.load(open('profiles/r1_bench_n8.json'))
ref = json.load(open('profiles/r1_bench_reference_arm.json'))
d = json.load(open('profiles/r1_dense_c4.json'))
d64 = json.load(open('profiles/r1_dense_n64.json'))
d300 = json.load(open('profiles/r1_dense_n300.json'))
c4 = json.load(open('profiles/r1_c4_end_to_end.json'))
c3 = json.load(open('profiles/r1_bench_c3.json'))
fits = [json.loads(l) for l in open('profiles/r1_fits.jsonl') if l.strip()]
upd = json.load(open('profiles/r1_update_loop.json'))
tr = json.load(open('profiles/r1_traffic.json'))
K = b['roofline']['kernels']
txt = f"""# profiles/ — round 1 evidence (B200, sm_100a, CUDA 12.9, driver 580)

How these were produced: `tools_gpu_profile.sh` / `tools/prof_dense.sh` under `gpurun` (plain run first, then `ncu` on the
identical command line, `--clock-control none`); bench lines are plain runs (never under a profiler).
Workload: BASELINE config C2 — CS2 tree, n = 10, 100 006 leaves, 1e8 resident queries per step (8 GB, larger than L2),
step = re-fit every leaf model + locate/evaluate every query.  Peak = measured copy bandwidth {b['roofline']['peak']} GB/s
(`MEASURED_PEAKS.json`).  Regenerate this file with `python tools/make_profiles_readme.py`.

## Files
| file | what |
|---|---|
| `r1_bench_n1.json`, `r1_bench_n2.json`, `r1_bench_n8.json` | `bench.py` lines at N = 1, 2 and 8 (torchrun, NCCL) on config C2 |
| `r1_bench_c3.json` | `bench.py --config C3 --steps 5` (n = 20, 1 000 021 leaves, 1e8 queries per step) |
| `r1_bench_reference_arm.json` | `bench.py --impl reference` (oracle port on the box's host cores) |
| `r1_launches_ncu.csv` | every launch of `bench.py --queries 20000000 --steps 3` with `gpu__time_duration.sum` |
| `r1_<kernel>.ncu-rep` / `.summary.txt` | `ncu --set full --import-source on` capture of one launch + key metrics and per-source-line stall summary (`tools/ncu_metrics.py`, `tools/ncu_lines.py`); `k_fit` is the C2 launch, `k_fit_C3` the C3 launch, `k_eval_binned_C3` = `k_eval_binned_mma<20>` on one 4e7-query chunk of C3, `k_eval_dmma` and `k_locate_wide_C4` config C4 |
| `r1_traffic.json` | DRAM bytes (read + write) per launch of each capture, per query / per leaf (`tools/make_traffic_json.py`) |
| `r1_dense_c4.json`, `r1_dense_n64.json`, `r1_dense_n300.json` | `tools/bench_dense.py`: grouped dense evaluation on the DMMA kernels (C4: n = 100; n = 64; n = 300 = the panel-streaming kernel), with a cuBLAS DGEMM rate measured in the same run and the issued-DMMA fraction of the measured pipe peak |
| `r1_c4_end_to_end.json` | `tools/bench_c4.py`: the whole C4 path (CS1 tree n = 100 grown by 100 addpoint!s, K1 locate + binning + DMMA evaluation, 2e7 resident uniform queries) |
| `r1_fp64_peaks.txt` | `tools/ubench/fp64_peaks.cu`: raw DMMA.8x8x4 / DFMA throughput of the box (the K3b roofline denominator) |
| `r1_fits.jsonl` | `tools/bench_fits.py C2 C3 C5`: all-leaf fit time, fits/s and fraction of the HBM roofline with SURVEY 8(d)'s algorithmic bytes |
| `r1_update_loop.json` | `tools/bench_update.py`: one optimiser-loop iteration on a C2-sized tree (addpoint, re-flatten, re-upload, 1e4 queries, all-leaf fit) |

## Headline (N = 1)
| quantity | value |
|---|---|
| locate + evaluate, inputs resident (`value`) | **{b['value']:.3e} queries/s** ({b['ms_per_step']:.2f} ms per 1e8-query step incl. the all-leaf re-fit) |
| same through the host-buffer C-ABI call, pinned memory, copies timed (`e2e`) | {b['e2e']['value']:.3e} queries/s — PCIe-bound: 9.3 GB per step |
| all-leaf fits | {b['fits']['value']:.3e} fits/s ({b['fits']['kernel_ms']:.3f} ms for 100 006 leaves, two kernels) |
| CPU baseline (oracle port, {b['cpu_baseline']['cores']} cores, same box) | {b['cpu_baseline']['value']:.3e} queries/s, {b['cpu_baseline']['fits_per_s']:.3e} fits/s |
| reference arm (`--impl reference`) | {ref['value']:.3e} queries/s, {ref['fits']['value']:.3e} fits/s |
| N = 2 (weak scaling, one NCCL broadcast of the tree blob) | {b2['value']:.3e} queries/s |
| N = 8 (`r1_bench_n8.json`, same box type, 10 steps) | {b8['value']:.3e} queries/s ({b8['value']/8/b['value']:.2f} of 8x the N = 1 value; tree broadcast {b8['tree_broadcast_ms']:.1f} ms once) |
| config C3 (n = 20, 1e6 leaves), N = 1 | {c3['value']:.3e} queries/s ({c3['ms_per_step']:.1f} ms per 1e8-query step, of which {c3['fits']['kernel_ms']:.2f} ms re-fit all 1 000 021 leaves: {c3['fits']['value']:.3e} fits/s, {c3['fits']['frac']:.2f} of the HBM roofline) |
| config C4 end to end (CS1 n = 100, 1e4 leaves, locate + DMMA evaluation), N = 1 | {c4['queries_per_s']:.3e} queries/s ({c4['ms_per_step']:.1f} ms per 2e7 queries: locate {c4['kernel_ms_per_step']['locate']:.1f}, evaluation {c4['kernel_ms_per_step']['eval_binned']:.1f}) |
| clocks during the timed region | {b['clocks']['sm_mhz']} / {b['clocks']['sm_max_mhz']} MHz, throttle reasons: {b['clocks']['reasons'] or 'none'} |

## Where a step's time goes (CUDA events around every launch, `csp_timing_*`; 24 chunks of 4 194 304 queries)
| kernel | ms / step | launches / step | share | algorithmic B/query | algorithmic GB/s | frac of peak | DRAM B/query (ncu) |
|---|---|---|---|---|---|---|---|
| `k_locate_only<2>` (K1: find_leaf_at + histogram) | {K['locate']['ms_per_step']:.2f} | {K['locate']['launches_per_step']:.0f} | {100*K['locate']['share']:.0f} % | 85 | {K['locate']['algorithmic_GBps']:.0f} | {K['locate']['frac']:.2f} | {tr['k_locate_only']['dram_bytes_per_query']:.0f} |
| `k_scan` | {K['scan']['ms_per_step']:.2f} | {K['scan']['launches_per_step']:.0f} | {100*K['scan']['share']:.0f} % | – | – | – | – |
| `k_scatter` + `k_unpermute` | {K['scatter']['ms_per_step']:.2f} | {K['scatter']['launches_per_step']:.0f} | {100*K['scatter']['share']:.0f} % | 12 | {K['scatter']['algorithmic_GBps']:.0f} | {K['scatter']['frac']:.2f} | {tr['k_scatter']['dram_bytes_per_query']:.0f} (+ L2 atomics) |
| `k_eval_binned_exact<10>` (K3a: modelvalue) | {K['eval_binned']['ms_per_step']:.2f} | {K['eval_binned']['launches_per_step']:.0f} | {100*K['eval_binned']['share']:.0f} % | 96 | {K['eval_binned']['algorithmic_GBps']:.0f} | {K['eval_binned']['frac']:.2f} | {tr['k_eval_binned']['dram_bytes_per_query']:.0f} |
| whole pipeline | {b['roofline']['pipeline']['kernel_ms_per_step']:.2f} | | | 93 | {b['roofline']['pipeline']['achieved']:.0f} | {b['roofline']['pipeline']['frac']:.2f} | |
| `k_fit_chain` + `k_fit<2>` (K2) | {b['fits']['kernel_ms']:.3f} | 2 | | 5 616 B/leaf | | {b['fits']['frac']:.2f} | {tr['k_fit']['dram_bytes_per_leaf']+tr['k_fit_chain']['dram_bytes_per_leaf']:.0f} B/leaf |

ncu launch list (`r1_launches_ncu.csv`, 2e7 queries per step, cold-cache, serialised — compare shares):

| kernel | launches | mean µs | share |
|---|---|---|---|
""" + "\n".join(lines) + f"""

The shares agree with the event timings (at 2e7 queries per step the fit is a larger part of the step; at 1e8 it is 4 %).

## What bounds each kernel (from the `--set full` captures)
* **K1 `k_locate_only`** — L1/shared-memory bound: `l1tex__throughput` 85 % of peak (the transposed coordinate reads + the
  tree-top records in shared memory, one `LDG.E.256` record per level below the top), DRAM 23 %, issue 49 %.  DRAM traffic
  equals the algorithmic bytes ({tr['k_locate_only']['dram_bytes_per_query']:.0f} vs 85 B/query): nothing is re-read; the cost is ≈ 9 dependent random 32-byte record
  fetches per query.
* **K3a `k_eval_binned_exact<10>`** — DRAM bound on *extra* traffic: {tr['k_eval_binned']['dram_bytes_per_query']:.0f} B/query against 96 algorithmic (63 % of DRAM
  peak while it runs).  The x rows are gathered a second time at 64-byte DRAM granularity (≈ 128 B for an 80-byte row);
  values are written in sorted order (coalesced) and `k_unpermute` gathers them back from L2.  `long_scoreboard` dominates.
* **K1 `k_locate_wide<1>`** (n > 32; config C4, n = 100) — HBM bound: {tr['k_locate_wide_C4']['dram_bytes_per_query']:.0f} B/query of DRAM traffic for 805 algorithmic
  ({tr['k_locate_wide_C4']['dram_bytes_read']/tr['k_locate_wide_C4']['duration_us_under_ncu']/1e3:.0f} GB/s read while it runs): the rows are streamed once for the bounds check and the descent's re-reads mostly hit L1/L2.
* **K3a `k_eval_binned_mma<20>`** (config C3, n = 20) — latency bound at 16 warps/SM (128 registers): DRAM 48 %, DMMA pipe 31 %,
  `long_scoreboard` on the x-row gathers; {tr['k_eval_binned_C3']['dram_bytes_per_query']:.0f} B/query of DRAM traffic for 176 algorithmic (sector-granular second read of x + one
  2 KB record per ~40-query run).  The FMA kernel it replaces was bound by shared-memory read bandwidth (65 % `l1tex`, 18 % issue).
* **`k_scatter`** — L2 atomic latency/throughput (1e8 returning atomics per step; `lg_throttle` + `long_scoreboard`, issue 6 %).
* **K2 `k_fit<2>`** — instruction-issue / latency bound (≈ 60 % issue-active, 64 registers, 32 warps/SM; ≈ 2.0 k warp-instructions
  per leaf at n = 10 and 3.5 k at n = 20; DRAM 2–6 %).  Sibling leaves share the split scan and the off-diagonal coefficients;
  what remains per leaf is sequential per output entry (bit-exactness fixes the summation order).
* **K3b `k_eval_dmma`** (config C4, `r1_dense_c4.json`) — {d['dmma_kernel_useful_tflops']:.1f} useful TFLOP/s = {100*d['frac_of_dgemm_peak_kernel']:.0f} % of the cuBLAS DGEMM
  rate measured in the same run ({d['cublas_dgemm_tflops']:.1f} TFLOP/s; "useful" counts the full 2n^2 + 5n + 2 flop/query although the triangular
  structure lets the kernel issue about half).  In issued flop: {d['dmma_issued_tflops']:.1f} TFLOP/s = {100*d['dmma_pipe_frac']:.0f} % of the measured DMMA pipe peak
  (37.1 TFLOP/s, `r1_fp64_peaks.txt`; ncu: `sm__pipe_tensor_subpipe_dmma_cycles_active` 52 %, `math_pipe_throttle` the top stall).
  SASS `DMMA.8x8x4` (the only FP64 tensor op on sm_100a), `UBLKCP` row gathers double-buffered behind the MMAs;
  {d['queries_per_s']:.2e} queries/s at n = 100, {d64['queries_per_s']:.2e} at n = 64, {d300['queries_per_s']:.2e} at n = 300 (panel-streaming kernel).

## Leaf fits (K2), `r1_fits.jsonl`
| config | n | leaves | ms (chain + fit) | fits/s | algorithmic B/leaf | frac of HBM peak |
|---|---|---|---|---|---|---|
""" + "\n".join(f"| {f['config']} | {f['n']} | {f['leaves']} | {f['fit_ms']:.3f} | {f['fits_per_s']:.3e} | {f['algorithmic_bytes_per_leaf']:.0f} | {f['frac_of_hbm_peak']:.2f} |" for f in fits) + f"""

Optimiser loop on a C2-sized tree (`r1_update_loop.json`, ms per iteration): addpoint {upd['addpoint']}, re-flatten {upd['flatten']},
re-upload {upd['upload']}, 1e4 queries {upd['locate_1e4']}, all-leaf fit (allocate + fit) {upd['fit_all']}.

## Round-1 progression of the headline (C2, one B200, resident)
| version | ms / 1e8 queries | queries/s |
|---|---|---|
| v1: thread-per-query descent (48-byte records) + per-thread model reads | 52.8 | 1.9e9 |
| BFS 32-byte records (`LDG.E.256`), warp-cooperative coalesced evaluation | 22.1 | 4.4e9 |
| uniform packed-A records, 8 lanes per query | 17.9 | 5.4e9 |
| binned (model-stationary) pipeline, unchunked | 17.0 | 5.7e9 |
| + L2-friendly 4 M-query chunks, single-launch scan | 12.6 | 7.6e9 |
| + transposed conflict-free tile, constant-bank bounds, single TMA stage in K1 | 11.6 | 8.2e9 |
| + record staged per warp in shared memory, perm prefetch in the evaluation | 11.1 | 8.6e9 |
| + sorted-order value writes, L2-resident unpermute gather | 10.8 | 9.2e9 |

Other kernels over the round: all-leaf fits C2 0.48 -> 0.34 ms, C3 11.9 -> 5.35 ms (flattened ancestor windows, sibling-shared scan,
record assembled in shared memory, dynamic unit scheduling, per-split coefficient table), C5 (n = 300) 26.7 -> 14.1 ms, C5det 1 593 -> 445 ms; K3b at C4 24.9 -> 11.7 ms per 2e7 queries (software-pipelined DMMA
kernel), n = 300 evaluation 8.0e6 -> 1.4e8 queries/s (panel-streaming DMMA kernel); C3 step 26.8 -> 23.4 ms (tensor-core evaluation at n = 20);
C4 end to end 6.7e8 -> 1.25e9 queries/s (one chunk for the dense path, streaming descent kernel for n > 32).
"""
open('profiles/README.md', 'w').write(txt)
