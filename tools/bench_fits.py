#!/usr/bin/env python
"""Leaf-fit throughput (K2) on the benchmark configurations: all-leaf fit_quadratic (models kept resident), and for C5 also the
sparse (tridiagonal) coefficients_p output.  One JSON line per configuration.
    python tools/bench_fits.py [C2 C3 C5 C5det]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import csp_b200 as cb
from csp_b200 import device as cdev

PEAK = 6550.4
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def fit_bytes(n, visited):
    return 8 * (1 + 2 * n + n * (n + 1) // 2) + 64 * n * (n - 1) // 2 + 8 * n * n + 48 * ((n + 1) // 2) + 8 * visited


for name in (sys.argv[1:] or ["C2", "C3", "C5"]):
    if name == "C5det":  # determined variant of C5: every pair appears, ~1.35e5 leaves
        obj = cb.synthetic.quad_objective(300, 5, sigma=0.0)
        root = cb.synthetic.grow(300, 2, 299, obj, seed=1005)
    else:
        root, obj = cb.synthetic.config_tree(name)
    dt = cb.device_tree(root)
    n, nl = dt.n, dt.n_leaves
    t0 = time.perf_counter()
    models = dt.fit_models()
    torch.cuda.synchronize()
    first = time.perf_counter() - t0
    steps = 5 if nl * n * n < 5e9 else 2
    cdev.timing_read(); cdev.timing_enable(True)
    for _ in range(steps):
        dt.refit_models(models)
    kt = cdev.timing_read(); cdev.timing_enable(False)
    ms = kt["fit"][0] / steps
    st = models.download()["status"] if nl * n * n < 3e9 else None
    # mean number of splits the traversal examined (for the algorithmic bytes), on a sample of leaves
    f = root.tree.flat()
    sample = f["leaf_nodes"][:: max(1, nl // 2000)]
    _, vis = dt.coefficients_p(node_ids=sample) if n <= 64 else (None, np.zeros(1))
    out = {"config": name, "n": n, "leaves": nl, "fit_ms": ms, "fits_per_s": nl / (ms * 1e-3), "first_call_s": first,
           "mean_splits_examined": float(vis.mean()), "algorithmic_bytes_per_leaf": fit_bytes(n, float(vis.mean())),
           "frac_of_hbm_peak": nl * fit_bytes(n, float(vis.mean())) / (ms * 1e-3) / 1e9 / PEAK,
           "status_ok_fraction": None if st is None else float((st == 0).mean())}
    if n >= 100:
        tri = [(i, i + 1) for i in range(1, n)]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        vals, tvis = dt.coefficients_p_sparse(tri)
        out["tridiagonal_coefficients_s"] = time.perf_counter() - t0
        out["tridiagonal_mean_splits_examined"] = float(tvis.mean())
        out["tridiagonal_nan_fraction"] = float(np.isnan(vals).mean())
    print(json.dumps(out), flush=True)
    models.close()
