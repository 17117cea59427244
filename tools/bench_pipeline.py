#!/usr/bin/env python
"""A/B of the locate+evaluate pipeline's tuning switches on one GPU: per-kernel-class device times (csp_timing_*) for each setting.
    python tools/bench_pipeline.py [--config C3] [--queries 100000000] [--steps 3] KEY=a,b KEY2=c,d ...
Every KEY=v1,v2 is an environment variable of libcsp_b200.so; the cartesian product of the settings is run
("unset" removes the variable)."""
import argparse, itertools, json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import csp_b200 as cb
from csp_b200 import device as cdev
from csp_b200.device import DeviceTree

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="C3")
ap.add_argument("--queries", type=int, default=100_000_000)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("settings", nargs="*")
a = ap.parse_args()
root, _ = cb.synthetic.config_tree(a.config)
tree = DeviceTree(root.tree.flat())
n, m = tree.n, a.queries
dev = torch.device("cuda")
X = torch.empty((m, n), dtype=torch.float64, device=dev)
gen = torch.Generator(device=dev); gen.manual_seed(5)
for o in range(0, m, 1 << 22):
    k = min(1 << 22, m - o)
    X[o:o + k] = torch.rand((k, n), dtype=torch.float64, device=dev, generator=gen) * 2 - 1
leaf = torch.empty(m, dtype=torch.int32, device=dev)
val = torch.empty(m, dtype=torch.float64, device=dev)
status = torch.empty(m, dtype=torch.uint8, device=dev)
models = tree.fit_models()
keys = [s.split("=")[0] for s in a.settings]
vals = [s.split("=")[1].split(",") for s in a.settings]
ref = None
for combo in itertools.product(*vals) if keys else [()]:
    for k, v in zip(keys, combo):
        if v == "unset":
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    run = lambda: tree.locate_eval(models, X, leaf=leaf, val=val, status=status)
    for _ in range(2):
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
    ck = (int(leaf.sum(dtype=torch.int64).item()), float(val.nan_to_num().sum().item()))
    if ref is None:
        ref = ck
    print(json.dumps({"config": a.config, "queries": m, "settings": dict(zip(keys, combo)), "ms": round(ms, 3), "q_per_s": m / (ms * 1e-3),
                      "kernel_ms": {k: round(v[0] / a.steps, 3) for k, v in kt.items() if v[1]}, "leaf_checksum_equal": ck[0] == ref[0],
                      "value_sum_rel_diff": abs(ck[1] - ref[1]) / max(abs(ref[1]), 1e-300)}), flush=True)
