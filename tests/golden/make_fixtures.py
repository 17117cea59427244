#!/usr/bin/env python
"""Regenerates tests/golden/*.npz: outputs of the CPU oracle (oracle/, the restatement pinned to the reference's own
hand-written goldens in tests/test_oracle_*.py) on small seeded trees.  The reference is Julia and cannot run in the build
image, so these vectors are ORACLE-generated, not reference-generated; they freeze the oracle's behaviour (a CPU test
checks it against them) and give the GPU parity tests committed vectors to compare with.
    python tests/golden/make_fixtures.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from helpers import grow_both  # noqa: E402
from oracle import oracle as O  # noqa: E402

CASES = {"cs2_n4": dict(n=4, p=2, npoints=60, seed=41), "cs2_n10": dict(n=10, p=2, npoints=120, seed=42),
         "cs1_n3": dict(n=3, p=1, npoints=80, seed=43)}


def build(case):
    c = CASES[case]
    root, oroot, _ = grow_both(c["n"], c["p"], c["npoints"], seed=c["seed"])
    rng = np.random.default_rng(c["seed"])
    X = rng.random((2000, c["n"])) * 2.2 - 1.1  # some rows fall outside the world [-1, 1]^n
    return root, oroot, X


def oracle_outputs(oroot, X, p):
    leaf, status, _ = O.find_leaf_batch(oroot, X)
    out = dict(X=X, leaf=leaf, status=status)
    lv = O.leaves(oroot)
    if p == 2:
        fit = O.fit_boxes_batch(lv)
        out.update(fit_status=fit["status"], c=fit["c"], g=fit["g"], Q=fit["Q"], b=fit["b"])
    Cp, vis, _ = O.coefficients_p_batch(lv)
    out.update(Cp=Cp, visited=vis)
    return out


if __name__ == "__main__":
    for case in CASES:
        root, oroot, X = build(case)
        out = oracle_outputs(oroot, X, CASES[case]["p"])
        np.savez_compressed(os.path.join(HERE, case + ".npz"), **out)
        print(case, {k: v.shape for k, v in out.items()})
