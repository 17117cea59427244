#!/usr/bin/env python
"""Config C4 end to end: CS1 tree, n=100, 100 addpoint!s -> 1e4 leaves, models (c, g, Q, b) supplied per leaf,
resident uniform queries: find_leaf_at (K1) + binning + dense evaluation on the FP64 tensor cores (K3b).
    python tools/bench_c4.py [--queries 20000000] [--steps 5]"""
import argparse, json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import csp_b200 as cb
from csp_b200 import device as cdev
from csp_b200.device import LeafModels

ap = argparse.ArgumentParser()
ap.add_argument("--queries", type=int, default=20_000_000)
ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
root, obj = cb.synthetic.config_tree("C4")
dt = cb.device_tree(root)
n, nl, m = dt.n, dt.n_leaves, a.queries
rng = np.random.default_rng(0)
c, g, b = rng.standard_normal(nl), rng.standard_normal((nl, n)), 0.1 * rng.standard_normal((nl, n))
Q = np.triu(rng.standard_normal((nl, n, n)) / n)
models = LeafModels.upload(c, g, Q, b)
dev = torch.device("cuda")
X = torch.rand((m, n), dtype=torch.float64, device=dev) * 2 - 1  # the world is [-1, 1]^n
leaf = torch.empty(m, dtype=torch.int32, device=dev)
val = torch.empty(m, dtype=torch.float64, device=dev)
status = torch.empty(m, dtype=torch.uint8, device=dev)
run = lambda: dt.locate_eval(models, X, leaf=leaf, val=val, status=status)
for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
cdev.timing_read(); cdev.timing_enable(True)
for _ in range(a.steps):
    run()
kt = cdev.timing_read(); cdev.timing_enable(False)
flop_q = 2 * n * n + 5 * n + 2
print(json.dumps({"config": "C4 end to end: CS1 n=%d, %d leaves, %d uniform resident queries, models supplied" % (n, nl, m),
                  "queries_per_s": m / (ms * 1e-3), "ms_per_step": ms, "useful_tflops": m * flop_q / (ms * 1e-3) / 1e12,
                  "kernel_ms_per_step": {k: v[0] / a.steps for k, v in kt.items() if v[1]},
                  "launches_per_step": {k: v[1] / a.steps for k, v in kt.items() if v[1]},
                  "hbm_GBps_on_8n+13_bytes": m * (8 * n + 13) / (ms * 1e-3) / 1e9,
                  "outside": int((status != 0).sum().item()), "distinct_leaves": int(torch.unique(leaf).numel())}))
