"""Shared test helpers: build the same synthetic tree in the product's host tree and in the CPU oracle."""
import ctypes as C

import numpy as np

import csp_b200 as cb
from oracle import oracle as O


def oracle_objective(obj):
    """(function pointer, ctx) pair for the oracle from a product NativeObjective (same compiled function)."""
    return O.OBJECTIVE(obj.fn_ptr), C.c_void_p(obj.ctx)


def grow_both(n, p, npoints, seed, sigma=1e-3, lower=-1.0, upper=1.0, base=0.0, cycle=None, objective=None):
    """Grow identical trees (same points, same dimlists, same objective) on both sides.  Returns (root, oroot, objective)."""
    obj = objective if objective is not None else cb.synthetic.quad_objective(n, seed, sigma=sigma)
    rng = np.random.default_rng(1000 + seed)
    lo, up = np.full(n, float(lower)), np.full(n, float(upper))
    pos = np.full(n, float(base)) if np.isscalar(base) else np.asarray(base, np.float64)
    if cycle is None:
        cycle = cb.synthetic.cs1_dimlists(n) if p == 1 else cb.synthetic.round_robin_dimlists(n)
    X = lo + (up - lo) * rng.random((npoints, n))
    root = cb.Box(cb.World(lo, up, pos, obj), p)
    cb.addpoints(root, X, cycle, obj)
    ow = O.World(lo, up, pos, obj(pos))
    oroot = O.Box(ow, p)
    fp, ctx = oracle_objective(obj)
    O.addpoints_cycle(oroot, X, cycle, fp, ctx)
    O.index_tree(oroot)
    return root, oroot, obj


FLAT_KEYS = ("parent", "childindex", "internal_id", "leaf_id", "val", "inode", "dims", "xs", "child", "pos", "lower", "upper")


def assert_flat_equal(a, b):
    for k in ("n", "p", "n_nodes", "n_internal", "n_leaves"):
        assert a[k] == b[k], k
    for k in FLAT_KEYS:
        assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), k
