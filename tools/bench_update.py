#!/usr/bin/env python
"""Cost of one optimiser-loop iteration on a C2-sized tree: addpoint! on the host, bring the device copy up to
date, then a small batched query + a re-fit of every leaf.  Two ways to refresh the device copy:
  full   re-flatten on the host + csp_tree_upload of the whole description (what a stale handle needs);
  live   csp_tree_append_splits: only the new splits cross PCIe, the flattening is re-derived on the device.
The last iteration's results of the two paths are compared bit for bit."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import csp_b200 as cb
import torch

WARM, ITERS = 5, 20
out = {}
keep = {}
for mode in ("full", "live"):
    root, obj = cb.synthetic.config_tree("C2")
    cycle = cb.synthetic.round_robin_dimlists(10)
    rng = np.random.default_rng(0)
    X = cb.synthetic.uniform_queries(root, 10000, seed=1)
    live = mode == "live"
    cb.device_tree(root, live=live).find_leaf(X)
    t = {"addpoint": 0.0, "refresh": 0.0, "locate_1e4": 0.0, "fit_all": 0.0, "total": 0.0}
    for it in range(-WARM, ITERS):  # untimed iterations first: the allocation pool and the CUDA context reach their steady state
        t0 = time.perf_counter()
        cb.addpoint(root, rng.random(10) * 2 - 1, cycle[it % 9], obj)
        t1 = time.perf_counter()
        dt = cb.device_tree(root, live=live)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        leaf, status = dt.find_leaf(X)
        t3 = time.perf_counter()
        m = dt.fit_models()
        torch.cuda.synchronize()
        t4 = time.perf_counter()
        if it == ITERS - 1:
            keep[mode] = (leaf.copy(), m.download())
        m.close()
        if it < 0:
            continue
        for k, v in zip(t, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t4 - t0)):
            t[k] += v / ITERS
    out[mode] = {k: round(v * 1e3, 3) for k, v in t.items()}
    out["leaves"] = root.tree.counts()["n_leaves"]
same = np.array_equal(keep["full"][0], keep["live"][0]) and all(
    np.array_equal(keep["full"][1][k], keep["live"][1][k], equal_nan=True) for k in ("status", "c", "g", "Q", "b"))
print(json.dumps(out | {"unit": "ms per iteration (host wall clock, device synchronised)", "results_bit_identical": bool(same),
                        "note": f"{WARM} untimed warm-up iterations, {ITERS} timed; refresh = device copy brought up to date"}))
