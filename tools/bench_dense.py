#!/usr/bin/env python
"""Config C4: CS1 tree, n=100, 1e4 leaves, models (c,g,Q,b) supplied per leaf, many queries per leaf, grouped
by leaf -> K3b (FP64 tensor-core DMMA) evaluation.  Reports queries/s, useful TFLOP/s (2n^2+5n+2 flop/query) and the
fraction of a cuBLAS DGEMM peak measured in the same run (8192^3, torch.matmul float64), plus the issued DMMA flop
against the DMMA pipe peak measured by tools/ubench/fp64_peaks.cu.
    python tools/bench_dense.py [--n 100] [--leaves 10000] [--queries 20000000] [--steps 5]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import csp_b200 as cb
from csp_b200.device import LeafModels

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=100)
ap.add_argument("--leaves", type=int, default=10000)
ap.add_argument("--queries", type=int, default=20_000_000)
ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
n, nl, m = a.n, a.leaves, a.queries
rng = np.random.default_rng(0)
c, g, b = rng.standard_normal(nl), rng.standard_normal((nl, n)), 0.1 * rng.standard_normal((nl, n))
Q = np.triu(rng.standard_normal((nl, n, n)) / n)
models = LeafModels.upload(c, g, Q, b)
dev = torch.device("cuda")
X = torch.randn((m, n), dtype=torch.float64, device=dev)
leaf = torch.arange(m, device=dev, dtype=torch.int64).mul_(nl).div_(m, rounding_mode="floor").to(torch.int32)  # grouped by leaf
val = torch.empty(m, dtype=torch.float64, device=dev)


def timed(fn, steps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


ms = timed(lambda: models.evaluate(X, leaf, val=val), a.steps)
from csp_b200 import device as cdev
cdev.timing_read(); cdev.timing_enable(True)
for _ in range(a.steps):
    models.evaluate(X, leaf, val=val)
kt = cdev.timing_read(); cdev.timing_enable(False)
A = torch.randn((8192, 8192), dtype=torch.float64, device=dev)
B = torch.randn((8192, 8192), dtype=torch.float64, device=dev)
gemm_ms = timed(lambda: torch.matmul(A, B), 3)
dgemm_tf = 2 * 8192 ** 3 / (gemm_ms * 1e-3) / 1e12
flop_q = 2 * n * n + 5 * n + 2
kms = (kt["eval_binned"][0] + kt["eval"][0]) / a.steps  # evaluation kernel time per step (DMMA kernel when n is in 33..127)
nt = (n + 1 + 7) // 8
issued_q = sum(2 * (J + 1) for J in range(nt)) * 2 * 512 / 16  # DMMA.8x8x4 flop issued per query row (triangular skipping)
DMMA_PEAK_TF = 37.1  # tools/ubench/fp64_peaks.cu on this pool's B200 (DMMA.8x8x4, shared FP64 pipe)
out = {"config": "C4: n=%d, %d leaves, %d queries grouped by leaf, models supplied" % (n, nl, m), "queries_per_s": m / (ms * 1e-3),
       "ms_per_step": ms, "useful_tflops": m * flop_q / (ms * 1e-3) / 1e12, "dmma_kernel_ms": kms,
       "dmma_kernel_useful_tflops": m * flop_q / (kms * 1e-3) / 1e12, "cublas_dgemm_tflops": dgemm_tf,
       "frac_of_dgemm_peak_kernel": m * flop_q / (kms * 1e-3) / 1e12 / dgemm_tf, "kernels": {k: v for k, v in kt.items() if v[1]},
       "dmma_issued_tflops": m * issued_q / (kms * 1e-3) / 1e12, "dmma_pipe_peak_tflops": DMMA_PEAK_TF,
       "dmma_pipe_frac": m * issued_q / (kms * 1e-3) / 1e12 / DMMA_PEAK_TF, "hbm_bytes_per_query": 8 * n + 12, "sample_val": float(val[12345].item())}
print(json.dumps(out))
