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


def grow_config_both(name):
    """BASELINE config `name` (cb.synthetic.CONFIGS) grown identically in the product's host tree and in the oracle's pointer tree
    (same seed, points, dimlists and compiled objective as cb.synthetic.config_tree).  Returns (root, oroot, objective)."""
    c = cb.synthetic.CONFIGS[name]
    n, p = c["n"], c["p"]
    root, obj = cb.synthetic.config_tree(name)
    rng = np.random.default_rng(1000)
    lo, up = np.full(n, float(c.get("lower", -1.0))), np.full(n, float(c.get("upper", 1.0)))
    base = c.get("base", 0.0)
    pos = np.full(n, float(base)) if np.isscalar(base) else np.asarray(base, np.float64)
    X = lo + (up - lo) * rng.random((c["npoints"], n))
    cycle = cb.synthetic.cs1_dimlists(n) if p == 1 else cb.synthetic.round_robin_dimlists(n)
    oroot = O.Box(O.World(lo, up, pos, obj(pos)), p)
    fp, ctx = oracle_objective(obj)
    O.addpoints_cycle(oroot, X, cycle, fp, ctx)
    O.index_tree(oroot)
    return root, oroot, obj


WELL_CONDITIONED = 16.0  # a value is "well conditioned" when the summed magnitudes are at most this multiple of |value|


def eval_errors(val, oval, X, leaf, c, g, Q, b):
    """Error summary of evaluated values against the oracle's (finite entries only).  Q: (nl, n, n) in the .data layout or None.
    scale = |c| + |g|.|dx| + |dx|'|Q||dx|/2 bounds the magnitude of the terms modelvalue sums (reference src/cs2.jl:155-159), so
    |err| <= 1e-12 * scale is the backward-stable bar; on well-conditioned values (scale <= 16 |value|) the plain relative error
    |err| / |value| is reported as well."""
    fin = np.isfinite(oval)
    lf = leaf[fin]
    dx = np.abs(X[fin] - b[lf])
    scale = np.abs(c[lf]) + (np.abs(g[lf]) * dx).sum(1)
    if Q is not None:
        step = 200000
        quad = np.empty(len(lf))
        for o in range(0, len(lf), step):
            Qs = np.abs(np.nan_to_num(Q[lf[o:o + step]]))
            Qs = np.triu(Qs) + np.triu(Qs, 1).transpose(0, 2, 1)
            quad[o:o + step] = np.einsum("qi,qij,qj->q", dx[o:o + step], Qs, dx[o:o + step]) / 2
        scale = scale + quad
    err = np.abs(val[fin] - oval[fin])
    scaled = err / np.maximum(scale, 1e-300)
    av = np.abs(oval[fin])
    well = scale <= WELL_CONDITIONED * av
    rel = err[well] / av[well]
    return dict(max_scaled=float(scaled.max()) if err.size else 0.0, max_rel_wellcond=float(rel.max()) if rel.size else 0.0,
                frac_wellcond=float(well.mean()) if err.size else 0.0, n=int(err.size),
                nan_pattern_equal=bool(np.array_equal(np.isfinite(val), fin)))
