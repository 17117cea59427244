#!/usr/bin/env python
"""Cost of one optimiser-loop iteration on a C2-sized tree: addpoint! on the host, re-flatten + re-upload (the device copy is
stale after growth), then a small batched query + refit."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import csp_b200 as cb
root, obj = cb.synthetic.config_tree("C2")
cycle = cb.synthetic.round_robin_dimlists(10)
rng = np.random.default_rng(0)
X = cb.synthetic.uniform_queries(root, 10000, seed=1)
cb.find_leaf_at(root, X)
t = {"addpoint": 0.0, "flatten": 0.0, "upload": 0.0, "locate_1e4": 0.0, "fit_all": 0.0}
iters = 10
WARM = 3  # untimed iterations: the allocation pool and the CUDA context reach their steady state
for it in range(-WARM, iters):
    t0 = time.perf_counter(); cb.addpoint(root, rng.random(10) * 2 - 1, cycle[it % 9], obj); t1 = time.perf_counter()
    flat = root.tree.flat(); t2 = time.perf_counter()
    dt = cb.device_tree(root); t3 = time.perf_counter()
    dt.find_leaf(X); t4 = time.perf_counter()
    m = dt.fit_models(); m.close(); t5 = time.perf_counter()
    if it < 0:
        continue
    for k, v in zip(t, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)):
        t[k] += v / iters
print(json.dumps({k: round(v * 1e3, 3) for k, v in t.items()} | {"unit": "ms per iteration", "leaves": root.tree.counts()["n_leaves"]}))
