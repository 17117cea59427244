"""BASELINE.json's own configurations at FULL size, GPU against the CPU oracle (VERDICT r1 "what's weak" 1).

Bars: leaf ids and status bit-exact; every fitted coefficient of every leaf bit-exact (status, c, g, Q, b, NaN pattern);
evaluated values within 1e-12 of the condition scale and within 1e-12 plain relative error on well-conditioned values
(north_star: "<= 1e-12 relative on well-conditioned boxes"; the reference does not pin modelvalue's summation order)."""
import numpy as np
import pytest

import csp_b200 as cb
from csp_b200.device import LeafModels
from oracle import oracle as O
from helpers import eval_errors, grow_config_both

pytestmark = pytest.mark.gpu
TOL = 1e-12


def same(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


_CACHE = {}


def config(name):
    if name not in _CACHE:
        _CACHE.clear()  # one full-size configuration resident at a time
        root, oroot, obj = grow_config_both(name)
        _CACHE[name] = (root, oroot, obj, cb.device_tree(root))
    return _CACHE[name]


def check_values(val, oval, X, leaf, mo, Q):
    e = eval_errors(val, oval, X, leaf, mo["c"], mo["g"], Q, mo["b"])
    assert e["nan_pattern_equal"]
    assert e["max_scaled"] <= TOL, e
    assert e["max_rel_wellcond"] <= TOL, e
    assert e["frac_wellcond"] > 0.5, e
    return e


def test_c1_cs1_rosenbrock_full():
    """C1: CS1 (p=1), n=2, Rosenbrock in [-2,2]^2, 500 addpoint!s, 1e5 queries: leaf ids bit-exact; the degree-1 model
    c + g.(x - b) with g = coefficients_p(leaf) bit-exact in its coefficients and 1e-12 in its values."""
    root, oroot, _, dt = config("C1")
    olv = O.leaves(oroot)
    assert dt.n_leaves == len(olv) and 900 <= dt.n_leaves <= 1100
    X = cb.synthetic.uniform_queries(root, 100_000, seed=11)
    X[:4] = [[-2.0, -2.0], [2.0, 2.0], [-2.0, 2.0], [-1.2, 1.0]]  # world corners and the base point
    models = dt.fit_models()
    mo = models.download()
    og, ovis, _ = O.coefficients_p_batch(olv, nthreads=O.max_threads())
    assert same(mo["g"], og) and same(mo["c"], [O.value(b) for b in olv]) and same(mo["b"], [O.position(b) for b in olv])
    oleaf, oval, ostatus, _ = O.locate_eval_batch(oroot, X, mo["c"], mo["g"], np.zeros((len(olv), 2, 2)), mo["b"], nthreads=O.max_threads())
    for mode in ("direct", "binned"):
        import os
        os.environ["CSP_EVAL_MODE"] = mode
        try:
            leaf, val, status = dt.locate_eval(models, X)
        finally:
            os.environ.pop("CSP_EVAL_MODE")
        assert same(leaf, oleaf) and same(status, ostatus) and (status == 0).all()
        check_values(val, oval, X, oleaf, mo, None)
    leaf, status = dt.find_leaf(X)
    assert same(leaf, oleaf)


def test_c2_all_fits_and_queries_full():
    """C2: CS2, n=10, 100 006 leaves: ALL leaf fits bit-exact against the oracle; 2e6 queries located bit-exact and evaluated."""
    import torch
    root, oroot, _, dt = config("C2")
    olv = O.leaves(oroot)
    assert dt.n_leaves == len(olv) == 100006
    models = dt.fit_models()
    mo = models.download()
    ofit = O.fit_boxes_batch(olv, nthreads=O.max_threads())
    for k in ("status", "c", "g", "Q", "b"):
        assert same(mo[k], ofit[k]), k
    assert (mo["status"] == 0).mean() > 0.95
    X = cb.synthetic.uniform_queries(root, 2_000_000, seed=12)
    oleaf, oval, ostatus, _ = O.locate_eval_batch(oroot, X, ofit["c"], ofit["g"], ofit["Q"], ofit["b"], nthreads=O.max_threads())
    Xd = torch.from_numpy(X).cuda()
    leaf, val, status = dt.locate_eval(models, Xd)  # the resident (bench) path: binned pipeline
    torch.cuda.synchronize()
    assert same(leaf.cpu().numpy(), oleaf) and same(status.cpu().numpy(), ostatus)
    e = check_values(val.cpu().numpy(), oval, X, oleaf, mo, ofit["Q"])
    print("C2 value errors:", e)
    k = 300_000  # host-buffer path (a batch this small is evaluated inside the locate kernel: another summation order)
    leaf_h, val_h, status_h = dt.locate_eval(models, X[:k])
    assert same(leaf_h, oleaf[:k]) and same(status_h, ostatus[:k])
    check_values(val_h, oval[:k], X[:k], oleaf[:k], mo, ofit["Q"])


def test_c3_all_fits_and_queries_full():
    """C3: CS2, n=20, 1 000 021 leaves (the north-star configuration): ALL leaf fits bit-exact (status, c, g, Q, b), 1e6 queries
    located bit-exact and evaluated."""
    import torch
    root, oroot, _, dt = config("C3")
    olv = O.leaves(oroot)
    nl = len(olv)
    assert dt.n_leaves == nl == 1000021
    models = dt.fit_models()
    mo = models.download()
    step = 125_000
    oc, og, ob = np.empty(nl), np.empty((nl, 20)), np.empty((nl, 20))
    for o in range(0, nl, step):
        ofit = O.fit_boxes_batch(olv[o:o + step], nthreads=O.max_threads())
        sl = slice(o, min(o + step, nl))
        for k in ("status", "c", "g", "Q", "b"):
            assert same(mo[k][sl], ofit[k]), (k, o)
        oc[sl], og[sl], ob[sl] = ofit["c"], ofit["g"], ofit["b"]
    assert (mo["status"] == 0).mean() > 0.95
    X = cb.synthetic.uniform_queries(root, 1_000_000, seed=13)
    oleaf, ostatus, _ = O.find_leaf_batch(oroot, X, nthreads=O.max_threads())
    # the oracle's evaluation needs the models of the touched leaves only: compact them
    touched, inv = np.unique(oleaf, return_inverse=True)
    oval, _ = O.eval_batch(oc[touched], og[touched], mo["Q"][touched], ob[touched], X, inv.astype(np.int32), nthreads=O.max_threads())
    Xd = torch.from_numpy(X).cuda()
    for mode in ("direct", "binned"):
        import os
        os.environ["CSP_EVAL_MODE"] = mode
        try:
            leaf, val, status = dt.locate_eval(models, Xd)
            torch.cuda.synchronize()
        finally:
            os.environ.pop("CSP_EVAL_MODE")
        assert same(leaf.cpu().numpy(), oleaf) and same(status.cpu().numpy(), ostatus)
        e = check_values(val.cpu().numpy(), oval, X, oleaf, mo, mo["Q"])
        print("C3 value errors (%s):" % mode, e)


def test_c4_cs1_n100_full():
    """C4: CS1, n=100, 100 addpoint!s -> 10 001 leaves, models (c, g, Q, b) supplied per leaf: locate bit-exact (the
    streaming descent kernel) and the dense evaluation on the FP64 tensor cores within 1e-12."""
    root, oroot, _, dt = config("C4")
    nl, n = dt.n_leaves, 100
    assert nl == 10001
    rng = np.random.default_rng(9)
    c, g, b = rng.standard_normal(nl), rng.standard_normal((nl, n)), 0.1 * rng.standard_normal((nl, n))
    Q = np.triu(rng.standard_normal((nl, n, n)) / n)
    models = LeafModels.upload(c, g, Q, b)
    X = cb.synthetic.uniform_queries(root, 400_000, seed=5)
    oleaf, oval, ostatus, _ = O.locate_eval_batch(oroot, X, c, g, Q, b, nthreads=O.max_threads())
    leaf, val, status = dt.locate_eval(models, X)
    assert same(leaf, oleaf) and same(status, ostatus)
    e = check_values(val, oval, X, oleaf, dict(c=c, g=g, b=b), Q)
    print("C4 value errors:", e)
    # degree-1 model of the reference for CS1 trees: gradient = coefficients_p, all leaves
    g1, vis = dt.coefficients_p()
    og1, ovis, _ = O.coefficients_p_batch(O.leaves(oroot), nthreads=O.max_threads())
    assert same(g1, og1) and same(vis, ovis)


def test_c5_cs2_n300_all_fits_full():
    """C5: CS2, n=300, 23 addpoint!s -> 10 351 leaves, structurally under-determined: ALL leaf fits, identical NaN
    pattern and bit-exact finite entries, in batches (a dense Q is 720 KB per leaf)."""
    root, oroot, _, dt = config("C5")
    olv = O.leaves(oroot)
    nl = len(olv)
    assert dt.n_leaves == nl == 10351
    f = root.tree.flat()
    step = 1000
    nfinite = 0
    for o in range(0, nl, step):
        ids = f["leaf_nodes"][o:o + step]
        fit = dt.fit_boxes(node_ids=ids)
        ofit = O.fit_boxes_batch(olv[o:o + step], nthreads=O.max_threads())
        for k in ("status", "c", "g", "Q", "b"):
            assert same(fit[k], ofit[k]), (k, o)
        nfinite += int(np.isfinite(ofit["Q"]).sum())
    assert nfinite > 0
    # the all-leaf resident path gives the same records
    mo = dt.fit_models().download()
    sel = np.linspace(0, nl - 1, 300).astype(int)
    ofit = O.fit_boxes_batch([olv[i] for i in sel], nthreads=O.max_threads())
    for k in ("status", "c", "g", "Q", "b"):
        assert same(mo[k][sel], ofit[k]), k
