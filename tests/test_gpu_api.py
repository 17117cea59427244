"""GPU parity for the rest of the reference's API on the path: find_leaf_at from any box with the general boxbounds check
(src/tree.jl:146-171, 351-360), degree p = 3 (src/types.jl:194-201, src/polynomials.jl:182-230, test/runtests.jl:593-648), the
callback / used-split counts of coefficients_p (test/runtests.jl:652-700), fit_poly1's firstmatch (src/cs2.jl:173,191-193) and
minimum / maximum / extrema (src/CoordinateSplittingPTrees.jl:64-68).  Everything bit-exact against the oracle."""
import itertools
import math

import numpy as np
import pytest

import csp_b200 as cb
from oracle import oracle as O
from helpers import grow_both

pytestmark = pytest.mark.gpu


def same(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


def oracle_find_from(obox, X):
    """(leaf rank or -1, status) of find_leaf_at(obox, x) per row, status 1 where the reference throws."""
    leaf, status = np.full(len(X), -1, np.int32), np.zeros(len(X), np.uint8)
    for i, x in enumerate(X):
        try:
            leaf[i] = O.leaf_rank(O.find_leaf_at(obox, x))
        except O.JlError:
            status[i] = 1
    return leaf, status


@pytest.mark.parametrize("n,p,npoints", [(1, 1, 40), (2, 1, 120), (2, 2, 80), (5, 2, 60), (10, 2, 120)])
def test_find_leaf_from_any_box(n, p, npoints):
    root, oroot, _ = grow_both(n, p, npoints, seed=300 + n + p)
    dt = cb.device_tree(root)
    onodes = O.collect(oroot)
    rng = np.random.default_rng(n)
    f = root.tree.flat()
    picks = [0] + sorted(set(int(v) for v in rng.integers(1, len(onodes), 12)))
    X = cb.synthetic.uniform_queries(root, 400, seed=2)
    for v in picks:
        ob = onodes[v]
        assert O.preorder_id(ob) == v
        lo, hi = dt.boxbounds(v)
        obb = O.boxbounds(ob)
        assert same(lo, [b[0] for b in obb]) and same(hi, [b[1] for b in obb]), v
        # queries: uniform in the world, inside the box, on its edges (lower edge inclusive, upper edge exclusive unless the world's)
        Xin = lo + (hi - lo) * rng.random((200, n))
        Xe = np.vstack([X, Xin, lo[None], hi[None], Xin[:n] * 0 + np.where(np.eye(n, dtype=bool), hi, Xin[:n]),
                        Xin[:n] * 0 + np.where(np.eye(n, dtype=bool), lo, Xin[:n])])
        leaf, status = dt.find_leaf(np.ascontiguousarray(Xe), start=v)
        oleaf, ostatus = oracle_find_from(ob, Xe)
        assert same(status, ostatus), v
        assert same(leaf[ostatus == 0], oleaf[ostatus == 0]), v
        assert (ostatus[len(X):len(X) + 200] == 0).all()
        if v:
            import torch
            ld, sd = dt.find_leaf(torch.from_numpy(np.ascontiguousarray(Xe)).cuda(), start=v)
            assert same(ld.cpu().numpy(), leaf) and same(sd.cpu().numpy(), status)
    # through the reference-shaped API: a Box in, a Box out; outside the box raises like src/tree.jl:359
    boxes = cb.collect(root)
    inner = next(b for b in boxes if not cb.isleaf(b) and b.id != 0)
    lo, hi = dt.boxbounds(cb.node_id(inner))
    xin = lo + (hi - lo) * 0.37
    got = cb.find_leaf_at(inner, xin if n > 1 else float(xin[0]))
    ref = O.find_leaf_at(onodes[cb.node_id(inner)], xin)
    assert cb.leaf_id(got) == O.leaf_rank(ref)
    outside = [w for w in X if not ((lo <= w) & (w < hi)).all()]
    if outside:
        with pytest.raises(cb.ErrorException):
            cb.find_leaf_at(inner, outside[0])


def bprod(x, B):
    """B contracted with x along every axis, over p! (test/runtests.jl:594-606 for p = 1, 2, 3)."""
    out = B
    for _ in range(B.ndim):
        out = out @ x
    return float(out) / math.factorial(B.ndim)


def grow_degree(n, p, npoints, seed):
    """The same tree on both sides for any degree p: random dimension groupings per point (test/functions.jl:84-91), cubic /
    quadratic / linear form as the objective (test/runtests.jl:594-606)."""
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((n,) * p)
    B = np.empty_like(A)
    for I in itertools.product(range(n), repeat=p):
        B[I] = A[tuple(sorted(I))]
    f = lambda y: bprod(np.asarray(y), B)
    lo, up = np.full(n, -3.0), np.full(n, 3.0)
    root = cb.Box(cb.World(lo, up, np.zeros(n), f), p)
    oroot = O.Box(O.World(lo, up, np.zeros(n), f), p)
    for _ in range(npoints):
        order = rng.permutation(n) + 1
        dimlists = [tuple(int(d) for d in order[i:i + p]) for i in range(0, n, p)]
        x = rng.uniform(-3, 3, n)
        cb.addpoint(root, x, dimlists, f)
        O.addpoint(oroot, x, dimlists, f)
    O.index_tree(oroot)
    return root, oroot, B


@pytest.mark.parametrize("n,p,npoints", [(6, 3, 25), (7, 3, 25), (3, 3, 10), (5, 4, 8)])
def test_degree_p_descent_and_coefficients(n, p, npoints):
    """Box{p} with p > 2: the flattener, find_leaf_at and coefficients_p (dense p-dimensional SymmetricArray) against the oracle."""
    root, oroot, B = grow_degree(n, p, npoints, seed=40 + n + p)
    dt = cb.device_tree(root)
    olv = O.leaves(oroot)
    assert dt.p == p and dt.n_leaves == len(olv)
    X = cb.synthetic.uniform_queries(root, 3000, seed=1)
    X[0] = root.world.upper
    X[1, 0] = 3.5
    leaf, status = dt.find_leaf(X)
    oleaf, ostatus, _ = O.find_leaf_batch(oroot, X)
    assert same(leaf, oleaf) and same(status, ostatus)
    for skip in (0, 2):
        Cp, vis, used = dt.coefficients_p(skip=skip, counts=True)
        for l in range(0, len(olv), max(1, len(olv) // 60)):
            oCp, ovis, oused = O.coefficients_p(olv[l], skip=skip, stats=True)
            assert same(Cp[l], oCp) and vis[l] == ovis and used[l] == oused, (l, skip)
    # internal boxes too
    f = root.tree.flat()
    ids = f["inode"][:: max(1, len(f["inode"]) // 10)]
    Cp, vis, used = dt.coefficients_p(node_ids=ids, counts=True)
    onodes = O.collect(oroot)
    for k, v in enumerate(ids):
        oCp, ovis, oused = O.coefficients_p(onodes[v], stats=True)
        assert same(Cp[k], oCp) and vis[k] == ovis and used[k] == oused
    if n >= 6 and p == 3:  # enough points: the mixed third-order coefficients of the cubic form are recovered (runtests.jl:640-647)
        api = cb.coefficients_p(cb.leaves(root)[-1])
        found = [I for I in itertools.combinations(range(1, n + 1), 3) if not math.isnan(api[I])]
        assert found
        for I in found:
            assert api[I] == pytest.approx(B[tuple(i - 1 for i in I)], rel=1e-6, abs=1e-8)
    with pytest.raises(cb.CspError):
        dt.fit_models()


@pytest.mark.parametrize("n,p,npoints", [(6, 2, 80), (5, 2, 50), (4, 1, 60)])
def test_generic_coefficient_kernel_equals_fast_kernel(n, p, npoints, monkeypatch):
    root, oroot, _ = grow_both(n, p, npoints, seed=70 + n)
    dt = cb.device_tree(root)
    fast = dt.coefficients_p(counts=True)
    monkeypatch.setenv("CSP_COEF_GENERIC", "1")
    slow = dt.coefficients_p(counts=True)
    for a, b in zip(fast, slow):
        assert same(a, b)
    olv = O.leaves(oroot)
    for l in range(0, len(olv), 7):
        oCp, ovis, oused = O.coefficients_p(olv[l], stats=True)
        assert same(fast[0][l], oCp) and fast[1][l] == ovis and fast[2][l] == oused


def test_callback_counts_like_reference():  # runtests.jl:652-700: 24 (dense), 12 (SymTridiagonal), 16 (sparse) callback invocations
    fsparse = lambda x: float(np.sum(np.diff(x) ** 2) / 2)
    x0 = np.array([0.1, 0.2, 0.3, 0.4])
    root = cb.Box(cb.World(x0, fsparse), 2)
    cb.addpoint(root, 2 * x0, [(1, 3), (2, 4)], fsparse)
    cb.addpoint(root, 3 * x0, [(2, 3), (1, 4)], fsparse)
    cb.addpoint(root, 4 * x0, [(1, 4), (2, 3)], fsparse)
    box = cb.addpoint(root, 5 * x0, [(1, 2), (3, 4)], fsparse)
    counts = []
    Q = cb.coefficients_p(box, callback=counts.append)
    assert counts == [24] and Q[1, 2] == pytest.approx(-1)
    cb.coefficients_p_sparse(box, "tridiagonal", callback=counts.append)
    cb.coefficients_p_sparse(box, [(1, 2), (2, 3), (3, 4), (1, 4)], callback=counts.append)
    assert counts == [24, 12, 16]


@pytest.mark.parametrize("n,npoints", [(2, 30), (7, 120), (8, 150), (10, 300)])
def test_used_counts_and_firstmatch_bit_exact(n, npoints):
    root, oroot, _ = grow_both(n, 2, npoints, seed=500 + n)
    dt = cb.device_tree(root)
    olv = O.leaves(oroot)
    f = root.tree.flat()
    sel = np.arange(0, len(olv), max(1, len(olv) // 80))
    ids = f["leaf_nodes"][sel]
    Cp, vis, used = dt.coefficients_p(node_ids=ids, counts=True)
    fit = dt.fit_boxes(node_ids=ids, native=True)
    pat = [(i, i + 1) for i in range(1, n)]
    sv, svis, sused = dt.coefficients_p_sparse(pat, node_ids=ids, counts=True) if n > 1 else (None, None, None)
    nchecked = 0
    for k, l in enumerate(sel):
        oCp, ovis, oused = O.coefficients_p(olv[l], stats=True)
        assert vis[k] == ovis and used[k] == oused
        if n > 1:
            _, ovis2, oused2, _ = O.coefficients_p_pattern_batch([olv[l]], pat)
            assert svis[k] == ovis2[0] and sused[k] == oused2[0]
        if fit["status"][k] == 0:
            fm = O.fit_quadratic_native(olv[l])[6]
            assert same(fit["firstmatch"][k], fm), (k, fit["firstmatch"][k], fm)
            nchecked += 1
    assert nchecked > len(sel) // 2
    # the reference-shaped API returns firstmatch from the device
    ok = int(sel[np.flatnonzero(fit["status"] == 0)[0]])
    out = cb.fit_quadratic_native(cb.leaves(root)[ok])
    assert same(out[6], O.fit_quadratic_native(olv[ok])[6])


def jl_isless(a, b):
    if a != a:
        return False
    if b != b:
        return True
    return a < b or (a == b and math.copysign(1, a) < 0 < math.copysign(1, b))


def test_extrema_on_device():
    vals = iter(np.random.default_rng(3).standard_normal(100000))
    special = {5: float("nan"), 11: -0.0, 12: 0.0, 20: -7.5, 33: -7.5, 40: 9.25, 41: 9.25}
    calls = [0]

    def f(y):  # values with ties, NaN and signed zeros in known places
        calls[0] += 1
        return special.get(calls[0], float(next(vals)))
    n = 3
    lo, up = np.full(n, -1.0), np.full(n, 1.0)
    root = cb.Box(cb.World(lo, up, np.zeros(n), 0.5), 2)
    rng = np.random.default_rng(8)
    cyc = cb.synthetic.round_robin_dimlists(n)
    for i in range(60):
        cb.addpoint(root, rng.uniform(-1, 1, n), cyc[i % len(cyc)], f)
    dt = cb.device_tree(root)
    flat = root.tree.flat()
    for v in [0] + [int(x) for x in flat["inode"][1::7]] + [int(flat["leaf_nodes"][3])]:
        box = root._mk(flat["box_of_pre"][v])
        lv = cb.leaves(box)
        bmin, bmax = lv[0], lv[0]
        for b in lv[1:]:  # Julia's minimum / maximum over isless: the first of equal elements is kept
            if jl_isless(cb.value(b), cb.value(bmin)):
                bmin = b
            if jl_isless(cb.value(bmax), cb.value(b)):
                bmax = b
        lmin, vmin, lmax, vmax = dt.extrema(v)
        assert lmin == cb.leaf_id(bmin) and lmax == cb.leaf_id(bmax), v
        assert same(vmin, cb.value(bmin)) and same(vmax, cb.value(bmax))
        gmin, gmax = cb.extrema(box)
        assert gmin == bmin and gmax == bmax and cb.minimum(box) == bmin and cb.maximum(box) == bmax


DERIVED = ("node records", "descent records", "descent leaf table", "split-box coordinates", "child values", "leaf nodes",
           "path offsets", "ancestor paths", "fit units")


@pytest.mark.parametrize("n,p,npoints", [(1, 1, 30), (2, 1, 200), (3, 2, 0), (2, 2, 150), (5, 2, 90), (10, 2, 400), (7, 3, 20), (40, 1, 25)])
def test_device_builder_equals_host_builder(n, p, npoints, monkeypatch):
    """Every array the library derives at upload (tree_build.cu, on the device, straight from the blob sections) equals the
    host-side construction bit for bit -- from a descriptor, from a host blob and from a device blob (the NCCL broadcast buffer)."""
    import torch
    from csp_b200.device import DeviceTree
    if p == 3:
        root, _, _ = grow_degree(n, p, npoints, seed=5)
    else:
        root, _, _ = grow_both(n, p, npoints, seed=600 + n)
    flat = root.tree.flat()
    blob = cb.pack_blob(flat)
    monkeypatch.setenv("CSP_UPLOAD_HOST", "1")
    ref = DeviceTree(flat)
    want = [ref.export_derived(k) for k in range(9)]
    monkeypatch.delenv("CSP_UPLOAD_HOST")
    for how in ("desc", "host blob", "device blob"):
        dt = DeviceTree(flat) if how == "desc" else DeviceTree(blob=blob if how == "host blob" else torch.from_numpy(blob).cuda())
        assert (dt.n, dt.p, dt.n_nodes, dt.n_internal, dt.n_leaves, dt.max_depth) == (ref.n, ref.p, ref.n_nodes, ref.n_internal, ref.n_leaves, ref.max_depth)
        for k in range(9):
            if p > 2 and k in (1, 2):
                continue  # no packed descent records for p > 2
            assert np.array_equal(dt.export_derived(k), want[k]), (how, DERIVED[k])
        dt.close()
    ref.close()


def test_device_builder_rejects_bad_descriptions():
    from csp_b200.device import DeviceTree
    root, _, _ = grow_both(4, 2, 30, seed=9)
    good = root.tree.flat()
    for key, idx, val in (("parent", 5, 7), ("child", 3, 1), ("internal_id", 0, 3), ("dims", 2, 99), ("leaf_id", 4, 77)):
        bad = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in good.items()}
        bad[key][idx] = val
        with pytest.raises(cb.CspError):
            DeviceTree(bad)


RAW_SECTIONS = ("parent", "childindex", "internal_id", "leaf_id", "val", "inode", "dims", "xs", "child", "pos")


@pytest.mark.parametrize("n,p,npoints", [(1, 1, 25), (2, 1, 60), (3, 2, 40), (6, 2, 60), (10, 2, 120), (7, 3, 12)])
def test_live_tree_equals_full_upload(n, p, npoints):
    """a device tree grown from the split log (csp_tree_append_splits, flattening re-derived on the device) is, after
    every addpoint!, bit for bit the tree a full host re-flatten + upload produces: derived arrays, ids, query and fit results."""
    from csp_b200.device import DeviceTree
    if p == 3:
        root, _, _ = grow_degree(n, p, 3, seed=11)
        f = lambda y: float(np.sum(np.asarray(y) ** 3))
        cyc = None
    else:
        root, _, f = grow_both(n, p, 5, seed=700 + n)
        cyc = cb.synthetic.cs1_dimlists(n) if p == 1 else cb.synthetic.round_robin_dimlists(n)
    rng = np.random.default_rng(n + p)
    X = cb.synthetic.uniform_queries(root, 3000, seed=4)
    live = cb.device_tree(root, live=True)
    for it in range(npoints):
        x = rng.uniform(root.world.lower, root.world.upper)
        if p == 3:
            order = rng.permutation(n) + 1
            dimlists = [tuple(int(d) for d in order[i:i + p]) for i in range(0, n, p)]
        else:
            dimlists = cyc[it % len(cyc)]
        cb.addpoint(root, x, dimlists, f)
        live = cb.device_tree(root, live=True)          # appends the new splits only
        if it % 7 and it != npoints - 1:
            continue
        flat = root.tree.flat()
        ref = DeviceTree(flat)
        assert (live.n_nodes, live.n_internal, live.n_leaves, live.max_depth) == (ref.n_nodes, ref.n_internal, ref.n_leaves, ref.max_depth)
        for k in range(9):
            if p > 2 and k in (1, 2):
                continue
            assert np.array_equal(live.export_derived(k), ref.export_derived(k)), (it, DERIVED[k])
        node_of_box, box_of_node = live.live_maps()
        assert same(node_of_box, flat["pre_of_box"]) and same(box_of_node, flat["box_of_pre"])
        l1, s1 = live.find_leaf(X)
        l2, s2 = ref.find_leaf(X)
        assert same(l1, l2) and same(s1, s2)
        if p == 2:
            a, b = live.fit_models().download(), ref.fit_models().download()
            for key in ("status", "c", "g", "Q", "b"):
                assert same(a[key], b[key]), key
        else:
            a, b = live.coefficients_p(counts=True), ref.coefficients_p(counts=True)
            for u, w in zip(a, b):
                assert same(u, w)
        ref.close()


def test_live_tree_rejects_bad_log():
    from csp_b200.device import DeviceTree
    t = DeviceTree.live(2, 2, [-1, -1], [1, 1], [0, 0], 0.5)
    assert (t.n_nodes, t.n_leaves) == (1, 1)
    t.append_splits([0], [1, 2], [0.5, 0.5], [0.5, 1, 2, 3])
    assert (t.n_nodes, t.n_internal, t.n_leaves) == (5, 1, 4)
    with pytest.raises(cb.CspError):
        t.append_splits([9], [1, 2], [0.1, 0.1], [0, 0, 0, 0])   # box 9 does not exist yet
    with pytest.raises(cb.CspError):
        t.append_splits([0], [1, 2], [0.1, 0.1], [0, 0, 0, 0])   # box 0 is split already
    assert (t.n_nodes, t.n_internal, t.n_leaves) == (5, 1, 4)     # the handle survived
    leaf, status = t.find_leaf(np.array([[0.9, 0.9], [-0.9, -0.9]]))
    assert (status == 0).all() and leaf[0] != leaf[1]


@pytest.mark.parametrize("n,nl,m", [(40, 9, 20000), (10, 7, 300000), (6, 5, 2000)])
def test_eval_leaf_ids_out_of_range_give_nan(n, nl, m):
    """csp_eval[_dev] takes caller-supplied leaf ids: an id outside [0, n_models) -- negative or too large -- gives NaN on every
    evaluation path (grouped dense n > 32, binned n <= 32, one-thread-per-query), never an out-of-bounds access."""
    import torch
    from csp_b200.device import LeafModels
    rng = np.random.default_rng(n)
    c, g, b = rng.standard_normal(nl), rng.standard_normal((nl, n)), 0.1 * rng.standard_normal((nl, n))
    Q = np.triu(rng.standard_normal((nl, n, n)) / n)
    models = LeafModels.upload(c, g, Q, b)
    X = rng.uniform(-1, 1, (m, n))
    leaf = rng.integers(0, nl, m).astype(np.int32)
    bad = rng.random(m) < 0.1
    leaf[bad] = rng.choice(np.array([-1, -5, nl, nl + 3, 2 ** 30], np.int32), int(bad.sum()))
    oval, _ = O.eval_batch(c, g, Q, b, X[~bad], leaf[~bad])
    for where in ("host", "device"):
        if where == "host":
            val = models.evaluate(X, leaf)
        else:
            val = models.evaluate(torch.from_numpy(X).cuda(), torch.from_numpy(leaf).cuda()).cpu().numpy()
        assert np.isnan(val[bad]).all() and np.isfinite(val[~bad]).all()
        assert np.allclose(val[~bad], oval, rtol=1e-12, atol=1e-12 * np.abs(oval).max())


def test_locate_eval_very_large_n_small_batch():
    """n beyond what the fused locate+evaluate kernel can tile in shared memory: a small batch falls back to the
    streaming descent kernel followed by the evaluate-from-leaf-ids path instead of failing."""
    from csp_b200.device import LeafModels
    n = 900
    root, oroot, _ = grow_both(n, 1, 1, seed=31, sigma=0.0)
    dt = cb.device_tree(root)
    nl = dt.n_leaves
    rng = np.random.default_rng(2)
    c, g, b = rng.standard_normal(nl), rng.standard_normal((nl, n)), 0.1 * rng.standard_normal((nl, n))
    models = LeafModels.upload(c, g, None, b)
    X = cb.synthetic.uniform_queries(root, 500, seed=8)
    X[3, 7] = 5.0
    leaf, val, status = dt.locate_eval(models, X)
    oleaf, ostatus, _ = O.find_leaf_batch(oroot, X)
    assert same(leaf, oleaf) and same(status, ostatus) and status[3] == 1 and np.isnan(val[3])
    ok = oleaf >= 0
    want = c[oleaf[ok]] + np.einsum("qi,qi->q", g[oleaf[ok]], X[ok] - b[oleaf[ok]])
    assert np.allclose(val[ok], want, rtol=1e-11, atol=1e-11 * np.abs(want).max())
