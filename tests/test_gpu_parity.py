"""GPU parity tests (run on the B200 box with `-m gpu`): the CUDA path, called through the C ABI, against the CPU oracle
on the same seeded inputs.  Bars: leaf ids / status / fitted coefficients BIT-EXACT (the fit kernels evaluate every
expression in the reference's order with FMA contraction off); evaluated values within 1e-12 relative (the reference
does not pin the summation order of modelvalue, src/cs2.jl:155-159)."""
import os

import numpy as np
import pytest

import csp_b200 as cb
from csp_b200.device import DeviceTree, LeafModels
from oracle import oracle as O
from helpers import oracle_objective, grow_both

pytestmark = pytest.mark.gpu

EVAL_RTOL = 1e-12


def same(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


def edge_queries(root, oroot, rng, m):
    """Random queries plus the reference's edge cases: split midpoints (ties), leaf positions, world corners/edges,
    out-of-world rows and NaN."""
    w = root.world
    n = root.n
    X = w.lower + (w.upper - w.lower) * rng.random((m, n))
    f = root.tree.flat()
    ni = f["n_internal"]
    if ni:
        # rows sitting exactly on a split midpoint along one split dimension, elsewhere at the split's own position
        k = min(ni, m // 4)
        I = rng.integers(0, ni, k)
        pos = f["pos"].reshape(ni, n)[I].copy()
        for r, i in enumerate(I):
            d = f["dims"][i * f["p"]]
            if d <= n:
                pos[r, d - 1] = (pos[r, d - 1] + f["xs"][i * f["p"]]) / 2
        X[:k] = pos
        X[k:2 * k] = f["pos"].reshape(ni, n)[I]  # exact evaluation points
    X[-1] = w.lower
    X[-2] = w.upper
    X[-3] = w.upper
    X[-3, 0] = np.nextafter(w.upper[0], np.inf)  # just outside
    X[-4] = w.lower
    X[-4, n - 1] = np.nextafter(w.lower[n - 1], -np.inf)
    X[-5, 0] = np.nan
    return np.ascontiguousarray(X)


TREES = [  # (n, p, npoints)
    (1, 1, 50), (2, 1, 500), (5, 1, 200), (1, 2, 30), (2, 2, 200), (3, 2, 150), (7, 2, 200), (10, 2, 400), (20, 2, 120),
    (33, 2, 40), (40, 1, 60), (64, 2, 20), (101, 1, 15),  # n > 32: the streaming kernel (k_locate_wide)
]


@pytest.mark.parametrize("kernel", ["default", "tiled", "wide", "ring"])
@pytest.mark.parametrize("n,p,npoints", TREES)
def test_find_leaf_bit_exact(n, p, npoints, kernel, monkeypatch):
    """Every descent kernel (TMA-staged transposed CTA tile / streaming / warp-granular ring) on every tree, whatever the default
    choice for its n (the ring kernel takes n <= 64; above that the request falls through to the tiled kernel)."""
    if kernel != "default":
        monkeypatch.setenv("CSP_LOCATE_WIDE", "1" if kernel == "wide" else "0")
        monkeypatch.setenv("CSP_LOCATE_RING", "1" if kernel == "ring" else "0")
    root, oroot, _ = grow_both(n, p, npoints, seed=100 + n + p)
    rng = np.random.default_rng(n * 7 + p)
    X = edge_queries(root, oroot, rng, 20000)
    oleaf, ostatus, _ = O.find_leaf_batch(oroot, X, nthreads=O.max_threads())
    dt = cb.device_tree(root)
    leaf, status = dt.find_leaf(X)  # host-buffer path
    assert same(leaf, oleaf) and same(status, ostatus)
    assert (ostatus == 1).sum() >= 3 and (ostatus == 0).sum() > 19000
    import torch
    Xd = torch.from_numpy(X).cuda()
    leafd, statusd = dt.find_leaf(Xd)  # resident path
    torch.cuda.synchronize()
    assert same(leafd.cpu().numpy(), oleaf) and same(statusd.cpu().numpy(), ostatus)
    # ragged / unaligned / tiny batches (odd row counts defeat the 16-byte TMA granularity; offset base pointer)
    for m in (0, 1, 3, 127, 129, 1001):
        l2, s2 = dt.find_leaf(Xd[5:5 + m].contiguous() if m else Xd[:0])
        torch.cuda.synchronize()
        assert same(l2.cpu().numpy(), oleaf[5:5 + m])
    flat = Xd.reshape(-1)
    if n % 2 == 1:  # a view starting at an odd element offset is only 8-byte aligned
        Xo = flat[n:n + 999 * n].view(999, n)
        l3, _ = dt.find_leaf(Xo)
        torch.cuda.synchronize()
        assert same(l3.cpu().numpy(), oleaf[1:1000])


def test_find_leaf_reference_goldens_on_gpu():  # runtests.jl:87-96, 110-114, 242-251 through the public API
    world = cb.World([0], [1], [0], 10)
    root = cb.Box(world, 1)
    box = cb.split(root, 1, 1, 20)[0]
    box1 = cb.split(box, 1, 0.5, 30)[0]
    assert cb.find_leaf_at(root, [0.3]) == cb.getleaf(root)
    assert cb.find_leaf_at(root, [0.5]) == box1
    assert cb.find_leaf_at(root, [0.8]) == cb.getleaf(box)
    with pytest.raises(cb.ErrorException):
        cb.find_leaf_at(root, [1.2])
    with pytest.raises(cb.DimensionMismatch):
        cb.find_leaf_at(root, [0.1, 0.2])
    world = cb.World([0], [1], [1], 10)
    root = cb.Box(world, 1)
    cb.split(root, 1, 0, 20)
    assert cb.find_leaf_at(root, 0.5) == cb.getleaf(root)  # scalar x; tie goes to the closed-below side
    root = cb.Box(cb.World([0], [1], [0], 10), 1)
    b1 = cb.split(root, 1, 1, 20)[0]
    assert cb.find_leaf_at(root, 0.5) == b1
    # a tree without splits: the root is the only leaf
    root = cb.Box(cb.World([0, 0], [1, 1], [0.5, 0.5], 1.0), 2)
    assert cb.find_leaf_at(root, [0.2, 0.9]) == root
    # staleness: growing the tree re-flattens the device copy
    leaf = cb.addpoint(root, [0.25, 0.75], [(1, 2)], lambda y: float(y.sum()))
    assert cb.find_leaf_at(root, [0.2, 0.9]) == leaf


FIT_TREES = [(2, 60), (3, 80), (4, 100), (7, 200), (8, 200), (10, 400), (20, 150), (33, 70)]


@pytest.mark.parametrize("n,npoints", FIT_TREES)
@pytest.mark.parametrize("skip", [0, 3])
def test_fit_leaves_bit_exact(n, npoints, skip):
    root, oroot, _ = grow_both(n, 2, npoints, seed=200 + n)
    dt = cb.device_tree(root)
    olv = O.leaves(oroot)
    fit = dt.fit_boxes(skip=skip, native=True)
    ofit = O.fit_boxes_batch(olv, skip=skip, nthreads=O.max_threads(), native=True)
    assert same(fit["status"], ofit["status"])
    assert (ofit["status"] == 0).sum() > 0.5 * len(olv)
    for k in ("c", "g", "Q", "b", "gnat", "y"):
        assert same(fit[k], ofit[k]), k
    ok = ofit["status"] == 0
    assert same(fit["prec"][ok], ofit["prec"][ok])
    # off-diagonals alone (coefficients_p) and the number of splits the traversal examined
    Cp, vis = dt.coefficients_p(skip=skip)
    oCp, ovis, _ = O.coefficients_p_batch(olv, skip=skip, nthreads=O.max_threads())
    assert same(Cp, oCp.reshape(len(olv), n, n).transpose(0, 2, 1))
    assert same(vis, ovis)


def _deep_tree(n, depth_points, seed):
    """A tree whose leaves near x* sit under more than 32 ancestors: every point lands in the box the previous one created."""
    obj = cb.synthetic.quad_objective(n, seed, sigma=1e-3)
    rng = np.random.default_rng(seed)
    xstar = rng.random(n) * 0.5 + 0.1
    delta = rng.random(n) * 0.3 + 0.1
    X = np.array([xstar + delta * 0.8 ** k for k in range(depth_points)])
    lo, up, pos = np.full(n, -1.0), np.full(n, 1.0), np.zeros(n)
    cycle = cb.synthetic.round_robin_dimlists(n)
    root = cb.Box(cb.World(lo, up, pos, obj), 2)
    cb.addpoints(root, X, cycle, obj)
    oroot = O.Box(O.World(lo, up, pos, obj(pos)), 2)
    fp, ctx = oracle_objective(obj)
    O.addpoints_cycle(oroot, X, cycle, fp, ctx)
    O.index_tree(oroot)
    return root, oroot


@pytest.mark.parametrize("case", ["n2", "n6_failures", "n10", "n20", "n40_smem_limit", "n50_global_q", "deep_n2", "deep_n4"])
def test_all_leaf_model_fit_bit_exact(case, monkeypatch):
    """csp_models_fit (the resident all-leaf fit: sibling leaves share the split scan and the off-diagonal coefficients, the
    record is assembled in shared memory) equals the per-box path and the oracle bit for bit -- including trees where
    chaintop fails for some siblings only, records too large for shared memory, and leaves under more than 32 ancestors
    (several ancestor windows per traversal)."""
    if case.startswith("deep"):
        n = int(case[-1])
        root, oroot = _deep_tree(n, 70 if n == 2 else 45, seed=77 + n)
    elif case == "n6_failures":
        rng = np.random.default_rng(5)
        cycle = []
        for _ in range(40):
            perm = rng.permutation(6) + 1
            cycle.append([(int(perm[i]), int(perm[i + 1])) for i in range(0, 6, 2)])
        root, oroot, _ = grow_both(6, 2, 160, seed=9, cycle=cycle)
    else:
        n = int(case.split("_")[0][1:])
        root, oroot, _ = grow_both(n, 2, {2: 300, 10: 400, 20: 150, 40: 30, 50: 30}[n], seed=600 + n)
    n = root.n
    dt = cb.device_tree(root)
    olv = O.leaves(oroot)
    if case.startswith("deep"):
        assert dt.max_depth > 32, dt.max_depth
    ofit = O.fit_boxes_batch(olv, nthreads=O.max_threads())
    if case == "n6_failures":
        assert set(np.unique(ofit["status"])) >= {0, 1}
    iu = np.triu_indices(n)
    for shared in ("1", "0", "2"):  # default policy, never, always (also for records that stay in global memory)
        monkeypatch.setenv("CSP_FIT_SHARED", shared)
        mo = dt.fit_models().download()
        assert same(mo["status"], ofit["status"])
        assert same(mo["c"], ofit["c"]) and same(mo["g"], ofit["g"]) and same(mo["b"], ofit["b"])
        # download returns Q in the SymmetricArray .data layout (values at [i, j], i <= j)
        assert same(mo["Q"][:, iu[0], iu[1]], ofit["Q"][:, iu[0], iu[1]])
    box = dt.fit_boxes()
    for k in ("c", "g", "b", "status"):
        assert same(box[k], ofit[k]), k
    assert same(box["Q"][:, iu[0], iu[1]], ofit["Q"][:, iu[0], iu[1]])


def test_fit_internal_boxes_and_failures():
    """Boxes that are not leaves (fit_poly1 uses getleaf(box), the traversals start at the box itself), and trees grown
    with random pairings where chaintop fails: status codes must match the oracle's exceptions."""
    rng = np.random.default_rng(5)
    n = 6
    cycle = []
    for _ in range(40):
        perm = rng.permutation(n) + 1
        cycle.append([(int(perm[i]), int(perm[i + 1])) for i in range(0, n, 2)])
    root, oroot, _ = grow_both(n, 2, 160, seed=9, cycle=cycle)
    dt = cb.device_tree(root)
    oall = O.collect(oroot)
    ids = np.arange(len(oall), dtype=np.int32)
    keep = [i for i in ids if not O.isfake(oall[i])]
    fit = dt.fit_boxes(node_ids=keep, native=True)
    ofit = O.fit_boxes_batch([oall[i] for i in keep], nthreads=O.max_threads(), native=True)
    assert same(fit["status"], ofit["status"])
    assert set(np.unique(ofit["status"])) >= {0, 1}
    for k in ("c", "g", "Q", "b", "gnat", "y"):
        assert same(fit[k], ofit[k]), k
    Cp, vis = dt.coefficients_p(node_ids=keep)
    oCp, ovis, _ = O.coefficients_p_batch([oall[i] for i in keep], nthreads=O.max_threads())
    assert same(Cp, oCp.reshape(len(keep), n, n).transpose(0, 2, 1)) and same(vis, ovis)


def test_fit_underdetermined_nan_pattern():
    """few addpoint!s at larger n leave most pairs unseen -> NaN pattern identical to the reference's."""
    root, oroot, _ = grow_both(40, 2, 5, seed=11, sigma=0.0)
    dt = cb.device_tree(root)
    olv = O.leaves(oroot)
    fit = dt.fit_boxes()
    ofit = O.fit_boxes_batch(olv, nthreads=O.max_threads())
    assert np.isnan(ofit["Q"]).any() and same(fit["status"], ofit["status"])
    for k in ("c", "g", "Q", "b"):
        assert same(fit[k], ofit[k]), k


def test_fit_nonfinite_samples_sequential_path():
    """Non-finite objective values make computed coefficients NaN; the reference then re-fills those slots from later
    splits while counting them (polynomials.jl:211-227).  The kernel's exact sequential path must reproduce that."""
    n = 6
    calls = [0]
    base = cb.synthetic.quad_objective(n, 3, sigma=0.0)

    def f(y):
        calls[0] += 1
        return float("nan") if calls[0] % 17 == 0 else base(y)
    rng = np.random.default_rng(1)
    lo, up = -np.ones(n), np.ones(n)
    X = lo + (up - lo) * rng.random((60, n))
    cycle = cb.synthetic.round_robin_dimlists(n)
    root = cb.Box(cb.World(lo, up, np.zeros(n), 0.0), 2)
    oroot = O.Box(O.World(lo, up, np.zeros(n), 0.0), 2)
    for i, x in enumerate(X):
        calls[0] = 1000 * i
        cb.addpoint(root, x, cycle[i % len(cycle)], f)
        calls[0] = 1000 * i
        O.addpoint(oroot, x, cycle[i % len(cycle)], f)
    O.index_tree(oroot)
    dt = cb.device_tree(root)
    olv = O.leaves(oroot)
    Cp, vis = dt.coefficients_p()
    oCp, ovis, _ = O.coefficients_p_batch(olv)
    assert np.isnan(root.tree.flat()["val"]).any()
    assert same(Cp, oCp.reshape(len(olv), n, n).transpose(0, 2, 1)) and same(vis, ovis)
    fit = dt.fit_boxes()
    ofit = O.fit_boxes_batch(olv)
    for k in ("status", "c", "g", "Q", "b"):
        assert same(fit[k], ofit[k]), k


@pytest.mark.parametrize("n,npoints", [(1, 40), (2, 500), (6, 100), (100, 12)])
def test_cs1_gradient_bit_exact(n, npoints):
    """CS1 trees: coefficients_p returns the gradient (polynomials.jl:216,232-233)."""
    root, oroot, _ = grow_both(n, 1, npoints, seed=300 + n)
    dt = cb.device_tree(root)
    olv = O.leaves(oroot)
    Cp, vis = dt.coefficients_p()
    oCp, ovis, _ = O.coefficients_p_batch(olv, nthreads=O.max_threads())
    assert same(Cp, oCp) and same(vis, ovis)
    # degree-1 models: c = value(leaf), g = gradient, b = position(leaf)
    mo = dt.fit_models().download()
    assert same(mo["g"], oCp)
    assert same(mo["c"], np.array([O.value(b) for b in olv]))
    assert same(mo["b"], np.array([O.position(b) for b in olv]))


@pytest.mark.parametrize("mode", ["direct", "binned"])
@pytest.mark.parametrize("n,p,npoints", [(2, 2, 200), (3, 2, 150), (10, 2, 400), (13, 2, 60), (16, 2, 100), (20, 2, 120), (24, 2, 50),
                                         (25, 2, 40), (30, 2, 40), (2, 1, 300), (40, 1, 30)])
def test_locate_eval_parity(n, p, npoints, mode, monkeypatch):
    """Both evaluation strategies (model fetched per query inside the locate kernel / queries binned by leaf and evaluated
    model-stationary) against the oracle."""
    monkeypatch.setenv("CSP_EVAL_MODE", mode)
    root, oroot, _ = grow_both(n, p, npoints, seed=400 + n)
    dt = cb.device_tree(root)
    olv = O.leaves(oroot)
    rng = np.random.default_rng(n)
    X = edge_queries(root, oroot, rng, 30000)
    models = dt.fit_models()
    mo = models.download()
    if p == 2:
        ofit = O.fit_boxes_batch(olv, nthreads=O.max_threads())
        for k in ("status", "c", "g", "Q", "b"):
            assert same(mo[k], ofit[k]), k
        Q = ofit["Q"]
    else:
        Q = np.zeros((len(olv), n, n))
    oleaf, oval, ostatus, _ = O.locate_eval_batch(oroot, X, mo["c"], mo["g"], Q, mo["b"], nthreads=O.max_threads())
    leaf, val, status = dt.locate_eval(models, X)
    assert same(leaf, oleaf) and same(status, ostatus)
    fin = np.isfinite(oval)
    assert same(np.isfinite(val), fin)
    # condition-aware scale: |c| + |g|.|dx| + |dx|'|Q||dx|/2 bounds the magnitude of the summed terms
    dx = np.abs(X[fin] - mo["b"][oleaf[fin]])
    Qs = np.abs(np.nan_to_num(np.triu(Q) + np.triu(Q, 1).transpose(0, 2, 1)))[oleaf[fin]]
    scale = np.abs(mo["c"][oleaf[fin]]) + (np.abs(mo["g"][oleaf[fin]]) * dx).sum(1) + np.einsum("qi,qij,qj->q", dx, Qs, dx) / 2
    err = np.abs(val[fin] - oval[fin])
    assert (err <= EVAL_RTOL * np.maximum(scale, 1e-300)).all(), (err / scale).max()
    # resident path and eval-only path give the same bits as the host path
    import torch
    Xd = torch.from_numpy(X).cuda()
    leafd, vald, _ = dt.locate_eval(models, Xd)
    torch.cuda.synchronize()
    assert same(leafd.cpu().numpy(), leaf) and same(vald.cpu().numpy(), val)
    ok = oleaf >= 0
    v2 = models.evaluate(X[ok], oleaf[ok])
    assert (np.abs(v2[fin[ok]] - oval[ok][fin[ok]]) <= EVAL_RTOL * np.maximum(scale, 1e-300)).all()


@pytest.mark.parametrize("n", [10, 12, 20])
def test_binned_pipeline_one_and_two_streams_agree(n, monkeypatch):
    """The chunked binned pipeline with one chunk in flight and with two (two internal streams, the descent of one chunk sharing
    the SMs with the evaluation of the other: the default for 8 <= n <= 12): same leaf ids, status and value bits, several
    chunks with a ragged last one, against the oracle's leaf ids."""
    import torch
    root, oroot, _ = grow_both(n, 2, 150, seed=900 + n)
    dt = cb.device_tree(root)
    models = dt.fit_models()
    rng = np.random.default_rng(n + 1)
    X = edge_queries(root, oroot, rng, 300000)
    oleaf, ostatus, _ = O.find_leaf_batch(oroot, X, nthreads=O.max_threads())
    Xd = torch.from_numpy(X).cuda()
    monkeypatch.setenv("CSP_EVAL_MODE", "binned")
    monkeypatch.setenv("CSP_BIN_CHUNK", "70000")  # 5 chunks, the last one ragged
    out = {}
    for streams in ("1", "2"):
        monkeypatch.setenv("CSP_BIN_STREAMS", streams)
        for rep in range(2):  # the second call re-uses the streams and the pooled workspaces
            leaf, val, status = dt.locate_eval(models, Xd)
            torch.cuda.synchronize()
        out[streams] = (leaf.cpu().numpy(), val.cpu().numpy(), status.cpu().numpy())
    assert same(out["1"][0], oleaf) and same(out["1"][2], ostatus)
    for a, b in zip(out["1"], out["2"]):
        assert same(a, b)
    monkeypatch.delenv("CSP_BIN_STREAMS")  # the default choice for this n
    leaf, val, status = dt.locate_eval(models, Xd)
    torch.cuda.synchronize()
    assert same(leaf.cpu().numpy(), oleaf) and same(val.cpu().numpy(), out["1"][1])


def test_models_upload_roundtrip_and_supplied_eval():
    """Config-C4 style: models (c, g, Q, b) supplied per leaf; only the evaluation arithmetic is pinned."""
    rng = np.random.default_rng(3)
    n, nl, m = 12, 50, 5000
    c, g, b = rng.standard_normal(nl), rng.standard_normal((nl, n)), rng.standard_normal((nl, n))
    Q = rng.standard_normal((nl, n, n))
    Qu = np.triu(Q) + np.tril(np.full((n, n), np.nan), -1)  # SymmetricArray .data: lower triangle is fill
    mo = LeafModels.upload(c, g, Qu, b)
    d = mo.download()
    assert same(d["c"], c) and same(d["g"], g) and same(d["b"], b) and same(d["Q"], Qu)
    X = rng.standard_normal((m, n))
    leaf = rng.integers(0, nl, m).astype(np.int32)
    val = mo.evaluate(X, leaf)
    oval, _ = O.eval_batch(c, g, Qu, b, X, leaf)
    assert np.allclose(val, oval, rtol=1e-12, atol=1e-12)
    Qs = np.triu(Q) + np.triu(Q, 1).transpose(0, 2, 1)
    dx = X - b[leaf]
    direct = c[leaf] + (g[leaf] * dx).sum(1) + np.einsum("qi,qij,qj->q", dx, Qs[leaf], dx) / 2
    assert np.allclose(val, direct, rtol=1e-11, atol=1e-11)
    assert cb.modelvalue(c[0], g[0], cb.SymmetricArray(Qu[0]), b[0], X[0]) == pytest.approx(
        c[0] + g[0] @ (X[0] - b[0]) + (X[0] - b[0]) @ Qs[0] @ (X[0] - b[0]) / 2, rel=1e-12)


def test_blob_broadcast_format_on_device():
    """The wire format: pack on the host, upload from a host pointer and from a DEVICE pointer (as after an NCCL
    broadcast); all three trees answer identically."""
    import torch
    root, oroot, _ = grow_both(5, 2, 100, seed=21)
    flat = root.tree.flat()
    blob = cb.pack_blob(flat)
    X = cb.synthetic.uniform_queries(root, 5000, seed=2)
    ref, _ = cb.device_tree(root).find_leaf(X)
    t1 = DeviceTree(blob=blob)
    t2 = DeviceTree(blob=torch.from_numpy(blob).cuda())
    for t in (t1, t2):
        leaf, _ = t.find_leaf(X)
        assert same(leaf, ref)
        assert (t.n, t.p, t.n_leaves) == (5, 2, flat["n_leaves"])


@pytest.mark.parametrize("n", [2, 3, 7, 8])
def test_cs2_full_models_like_reference(n):
    """runtests.jl:824-859 through the public API on the GPU: Q ~ B, g ~ Q*b, value identities."""
    rng = np.random.default_rng(60 + n)
    B = rng.standard_normal((n, n))
    B = (B + B.T) / 2
    f = cb.quadnoise_objective(B, sigma=0.0)
    x0 = rng.standard_normal(n)
    world = cb.World(np.full(n, -np.inf), np.full(n, np.inf), x0, f)
    root = cb.Box(world, 2)
    cycle = cb.synthetic.round_robin_dimlists(n)
    for i in range(2 * len(cycle) + 2):
        box = cb.addpoint(root, rng.standard_normal(n), cycle[i % len(cycle)], f)
    Cp = cb.coefficients_p(box)
    assert not np.isnan(Cp.data[np.triu_indices(n, 1)]).any()
    c, g, Q, b = cb.fit_quadratic(box)
    Qf = Q.full()
    assert np.allclose(Qf, f.B, rtol=1e-6, atol=1e-8)
    assert np.allclose(g, Qf @ b, rtol=1e-6, atol=1e-8)
    x = cb.position(box)
    assert cb.modelvalue(c, g, Q, b, x) == pytest.approx(f(x), rel=1e-8, abs=1e-10)
    assert abs(c - g @ b + b @ Qf @ b / 2) < 1e-8
    cn, gn, Qn, bn, y, prec, firstmatch = cb.fit_quadratic_native(box)
    assert same(Qn.data, Q.data) and same(bn, b) and cn == c
    assert Q[1, n] == Q[n, 1] == Qf[0, n - 1]  # SymmetricArray indexing, 1-based, order-insensitive
    full = cb.polynomial_full(box)
    assert same(full[2].data, Q.data)
    mn = cb.polynomial_minimal(box)
    assert np.isnan(np.diag(mn[2].data)).all()
    with pytest.raises(TypeError):
        cb.fit_quadratic(cb.Box(cb.World([0.0], [1.0], [0.5], 0.0), 1))


def test_full_size_properties_c2():
    """BASELINE config C2 (CS2, n=10, 1e5 leaves) at full tree size through size-independent properties:
    every leaf's own evaluation point is located in that leaf; the fitted model reproduces the stored value there;
    leaf ids are a valid range; locate is idempotent under re-query; sharded batches agree with one big batch."""
    import torch
    root, obj = cb.synthetic.config_tree("C2")
    f = root.tree.flat()
    assert f["n_leaves"] == 100006
    dt = cb.device_tree(root)
    # leaf positions: parent's position with the displaced dims replaced
    n, p = 10, 2
    par = f["parent"][f["leaf_nodes"]]
    I = f["internal_id"][par]
    P = f["pos"].reshape(-1, n)[I].copy()
    ci = f["childindex"][f["leaf_nodes"]]
    for k in range(p):
        d = f["dims"].reshape(-1, p)[I, k] - 1
        sel = ((ci >> k) & 1).astype(bool)
        P[np.flatnonzero(sel), d[sel]] = f["xs"].reshape(-1, p)[I, k][sel]
    leaf, status = dt.find_leaf(P)
    assert (status == 0).all() and same(leaf, np.arange(f["n_leaves"], dtype=np.int32))
    models = dt.fit_models()
    mo = models.download()
    okfit = mo["status"] == 0
    assert okfit.mean() > 0.95
    leaf2, val, _ = dt.locate_eval(models, P)
    assert same(leaf2, leaf)
    # the canonical model interpolates the data it was built from up to the noise level sigma = 1e-3
    truth = f["val"][f["leaf_nodes"]]
    fin = np.isfinite(val) & okfit
    assert np.median(np.abs(val[fin] - truth[fin])) < 0.05
    Xd = (torch.rand(2_000_000, n, dtype=torch.float64, device="cuda") * 2 - 1)
    l_all, v_all, s_all = dt.locate_eval(models, Xd)
    torch.cuda.synchronize()
    assert int(l_all.min()) >= 0 and int(l_all.max()) < f["n_leaves"] and int(s_all.max()) == 0
    halves = [dt.locate_eval(models, Xd[i::2].contiguous()) for i in range(2)]
    torch.cuda.synchronize()
    for i in range(2):
        assert torch.equal(halves[i][0], l_all[i::2]) and torch.equal(halves[i][1].nan_to_num(), v_all[i::2].nan_to_num())
    sub = slice(0, 20000)
    v_again = models.evaluate(Xd[sub].contiguous(), l_all[sub].contiguous())
    torch.cuda.synchronize()
    assert torch.allclose(v_again.nan_to_num(), v_all[sub].nan_to_num(), rtol=1e-12, atol=1e-12)


def test_c5_highdim_fits_global_pair_table():
    """Config C5 shape (CS2, n=300, a handful of addpoint!s): the pair table (n(n-1)/2 slots) no longer fits in shared memory
    and lives in global scratch; most pairs are unseen so the NaN pattern must match the oracle exactly."""
    root, oroot, _ = grow_both(300, 2, 3, seed=55, sigma=0.0)
    dt = cb.device_tree(root)
    olv = O.leaves(oroot)
    sel = np.linspace(0, len(olv) - 1, 40).astype(int)
    f = root.tree.flat()
    ids = f["leaf_nodes"][sel]
    fit = dt.fit_boxes(node_ids=ids)
    ofit = O.fit_boxes_batch([olv[i] for i in sel], nthreads=O.max_threads())
    assert same(fit["status"], ofit["status"])
    for k in ("c", "g", "Q", "b"):
        assert same(fit[k], ofit[k]), k
    Cp, vis = dt.coefficients_p(node_ids=ids)
    oCp, ovis, _ = O.coefficients_p_batch([olv[i] for i in sel], nthreads=O.max_threads())
    assert same(Cp, oCp.reshape(len(sel), 300, 300).transpose(0, 2, 1)) and same(vis, ovis)
    # point location at n = 300 (small tiles, bounds from shared memory)
    X = cb.synthetic.uniform_queries(root, 3000, seed=3)
    leaf, status = dt.find_leaf(X)
    oleaf, ostatus, _ = O.find_leaf_batch(oroot, X, nthreads=O.max_threads())
    assert same(leaf, oleaf) and same(status, ostatus)


def test_c4_cs1_n100_supplied_models():
    """Config C4 shape: CS1 tree, n=100, models (c, g, Q, b) supplied per leaf; locate bit-exact, values 1e-12."""
    n = 100
    root, oroot, _ = grow_both(n, 1, 20, seed=77, sigma=0.0)
    dt = cb.device_tree(root)
    nl = dt.n_leaves
    rng = np.random.default_rng(9)
    c, g, b = rng.standard_normal(nl), rng.standard_normal((nl, n)), 0.1 * rng.standard_normal((nl, n))
    Q = np.triu(rng.standard_normal((nl, n, n)) / n)
    models = LeafModels.upload(c, g, Q, b)
    X = cb.synthetic.uniform_queries(root, 4000, seed=5)
    for mode in ("direct", "binned"):
        os.environ["CSP_EVAL_MODE"] = mode
        try:
            leaf, val, status = dt.locate_eval(models, X)
        finally:
            os.environ.pop("CSP_EVAL_MODE")
        oleaf, oval, ostatus, _ = O.locate_eval_batch(oroot, X, c, g, Q, b, nthreads=O.max_threads())
        assert same(leaf, oleaf) and same(status, ostatus)
        assert np.allclose(val, oval, rtol=1e-12, atol=1e-12 * np.abs(oval).max())


def test_full_size_properties_c3():
    """BASELINE config C3 (CS2, n=20, 1e6 leaves) at full tree size: every leaf's own evaluation point is located in that
    leaf, fitted models reproduce the stored values at those points up to the noise level, both evaluation strategies
    agree, and the fit status distribution is sane."""
    import torch
    root, obj = cb.synthetic.config_tree("C3")
    f = root.tree.flat()
    n, p = 20, 2
    assert f["n_leaves"] == 1000021
    dt = cb.device_tree(root)
    par = f["parent"][f["leaf_nodes"]]
    I = f["internal_id"][par]
    P = f["pos"].reshape(-1, n)[I].copy()
    ci = f["childindex"][f["leaf_nodes"]]
    for k in range(p):
        d = f["dims"].reshape(-1, p)[I, k] - 1
        sel = ((ci >> k) & 1).astype(bool)
        P[np.flatnonzero(sel), d[sel]] = f["xs"].reshape(-1, p)[I, k][sel]
    leaf, status = dt.find_leaf(P)
    assert (status == 0).all() and same(leaf, np.arange(f["n_leaves"], dtype=np.int32))
    models = dt.fit_models()
    st = models.download()["status"]
    assert (st == 0).mean() > 0.95
    Pd = torch.from_numpy(P).cuda()
    out = {}
    for mode in ("direct", "binned"):
        os.environ["CSP_EVAL_MODE"] = mode
        try:
            l2, v2, _ = dt.locate_eval(models, Pd)
            torch.cuda.synchronize()
        finally:
            os.environ.pop("CSP_EVAL_MODE")
        out[mode] = v2.cpu().numpy()
        assert same(l2.cpu().numpy(), leaf)
    fin = np.isfinite(out["direct"])
    assert same(np.isfinite(out["binned"]), fin)
    scale = np.abs(out["direct"][fin]).max()
    assert np.allclose(out["binned"][fin], out["direct"][fin], rtol=1e-9, atol=1e-11 * scale)
    truth = f["val"][f["leaf_nodes"]]
    ok = fin & (st == 0)
    assert np.median(np.abs(out["direct"][ok] - truth[ok])) < 0.1


@pytest.mark.parametrize("n,nl,m", [(33, 11, 20000), (40, 7, 20000), (64, 300, 20000), (64, 3, 150000), (71, 3, 20000), (72, 50, 20000),
                                    (95, 6, 20000), (96, 2, 20000), (100, 5, 20000), (103, 40, 20000), (127, 9, 20000), (128, 4, 20000), (130, 12, 20000),
                                    (200, 150, 20000), (300, 3, 30000), (301, 7, 8000), (400, 2, 8000), (600, 3, 4000), (720, 2, 2000)])
def test_k3b_dense_dmma_eval(n, nl, m):
    """K3b: grouped evaluation on the FP64 tensor cores (mma.sync DMMA) for n > 32, models supplied per leaf (C4 shape).
    Long runs (few leaves, several tiles per CTA: the double-buffered pipeline) and short runs (segments inside a tile), every
    column-tile count 5..16, even n (TMA row gathers) and odd n (cp.async gathers), evaluate-only and through the binned pipeline.
    n > 127 runs the panel-streaming kernel (64-, 32- and 16-row tiles as shared memory allows); n = 720 is past it (generic kernel)."""
    import torch
    rng = np.random.default_rng(n + nl)
    c, g, b = rng.standard_normal(nl), rng.standard_normal((nl, n)), 0.1 * rng.standard_normal((nl, n))
    Q = np.triu(rng.standard_normal((nl, n, n)) / n)
    models = LeafModels.upload(c, g, Q, b)
    X = rng.standard_normal((m, n))
    leaf = rng.integers(0, nl, m).astype(np.int32)
    leaf[::97] = -1  # evaluate-only: negative leaf ids give NaN
    oval, _ = O.eval_batch(c, g, Q, b, X, np.maximum(leaf, 0), nthreads=O.max_threads())
    oval[leaf < 0] = np.nan
    Qs = np.triu(Q) + np.triu(Q, 1).transpose(0, 2, 1)
    scale = np.abs(c[leaf]) + (np.abs(g[leaf]) * np.abs(X - b[leaf])).sum(1) + np.einsum("qi,qij,qj->q", np.abs(X - b[leaf]), np.abs(Qs[leaf]), np.abs(X - b[leaf])) / 2
    for mode in ("binned", "direct"):
        os.environ["CSP_EVAL_MODE"] = mode
        try:
            val = models.evaluate(X, leaf)                       # host path
            vd = models.evaluate(torch.from_numpy(X).cuda(), torch.from_numpy(leaf).cuda())  # resident path
            torch.cuda.synchronize()
        finally:
            os.environ.pop("CSP_EVAL_MODE")
        for v in (val, vd.cpu().numpy()):
            assert same(np.isnan(v), np.isnan(oval))
            ok = ~np.isnan(oval)
            assert (np.abs(v[ok] - oval[ok]) <= 1e-12 * scale[ok]).all(), mode


@pytest.mark.parametrize("n", [2, 3, 7, 8])
def test_native_model_reproduces_every_leaf(n):
    """runtests.jl:839-844: the CS2-native model (fit_quadratic_native) evaluated with modelvalue(c,g,Q,b,y,prec,x) reproduces
    f at every leaf position (rtol 1e-7); the GPU evaluation equals the oracle's restatement bit for bit."""
    rng = np.random.default_rng(70 + n)
    B = rng.standard_normal((n, n))
    f = cb.quadnoise_objective((B + B.T) / 2, sigma=0.0)
    x0 = rng.standard_normal(n)
    world = cb.World(np.full(n, -np.inf), np.full(n, np.inf), x0, f)
    root = cb.Box(world, 2)
    cycle = cb.synthetic.round_robin_dimlists(n)
    for i in range(2 * len(cycle) + 2):
        box = cb.addpoint(root, rng.standard_normal(n), cycle[i % len(cycle)], f)
    c, g, Q, b, y, prec, firstmatch = cb.fit_quadratic_native(box)
    P = np.array([cb.position(l) for l in cb.leaves(root)])
    vals = cb.modelvalue_native(c, g, Q, b, y, prec, P)
    truth = np.array([f(p) for p in P])
    assert np.allclose(vals, truth, rtol=1e-7, atol=1e-9)
    ovals = np.array([O.modelvalue_native(c, g, Q.data, b, y, prec, p) for p in P])
    assert same(vals, ovals)
    assert cb.modelvalue_native(c, g, Q, b, y, prec, P[0]) == ovals[0]


@pytest.mark.parametrize("n,npoints", [(4, 12), (10, 300), (21, 60), (300, 4)])
def test_sparse_pattern_coefficients_bit_exact(n, npoints):
    """coefficients_p! into SymTridiagonal / sparse symmetric outputs (src/polynomials.jl:244-271,
    runtests.jl:652-700): values and the number of splits examined bit-exact vs the oracle, for leaves and internal boxes."""
    root, oroot, _ = grow_both(n, 2, npoints, seed=500 + n, sigma=0.0)
    dt = cb.device_tree(root)
    oall = [b for b in O.collect(oroot) if not O.isfake(b)]
    rng = np.random.default_rng(n)
    pick = rng.choice(len(oall), size=min(300, len(oall)), replace=False)
    boxes = [oall[i] for i in pick]
    ids = np.array([O.preorder_id(b) for b in boxes], np.int32)
    pats = [O.tridiagonal_pattern(n)]
    allp = [(i, j) for i in range(1, n + 1) for j in range(i + 1, n + 1)]
    sel = rng.choice(len(allp), size=min(len(allp), 3 * n), replace=False)
    pats.append([allp[k] for k in sel])
    pats.append([])  # an empty pattern determines nothing and examines at most one split
    for pat in pats:
        for skip in (0, 2):
            vals, vis = dt.coefficients_p_sparse(pat, node_ids=ids, skip=skip)
            ovals, ovis, _, _ = O.coefficients_p_pattern_batch(boxes, pat, skip=skip, nthreads=O.max_threads())
            assert same(vals, ovals), (len(pat), skip)
            assert same(vis, ovis), (len(pat), skip)


def test_sparse_hessian_reference_case():  # runtests.jl:652-700 through the public API
    fsparse = lambda x: float(np.sum(np.diff(x) ** 2) / 2)
    x0 = np.array([0.1, 0.2, 0.3, 0.4])
    root = cb.Box(cb.World(x0, fsparse), 2)
    cb.addpoint(root, 2 * x0, [(1, 3), (2, 4)], fsparse)
    cb.addpoint(root, 3 * x0, [(2, 3), (1, 4)], fsparse)
    cb.addpoint(root, 4 * x0, [(1, 4), (2, 3)], fsparse)
    box = cb.addpoint(root, 5 * x0, [(1, 2), (3, 4)], fsparse)
    ev = cb.coefficients_p_sparse(box, "tridiagonal")
    assert np.allclose(ev, -1)
    sp = cb.coefficients_p_sparse(box, [(1, 2), (2, 3), (3, 4), (1, 4)])
    assert all(abs(sp[(i, i + 1)] + 1) < 1e-12 for i in (1, 2, 3)) and abs(sp[(1, 4)]) < 1e-8
    Q = cb.coefficients_p(box)
    assert all(Q[i, i + 1] == sp[(i, i + 1)] for i in (1, 2, 3)) and np.isnan(Q[1, 1])
    vals, vis = cb.device_tree(root).coefficients_p_sparse([(1, 2), (2, 3), (3, 4)], [cb.node_id(box)])
    dense_vis = cb.device_tree(root).coefficients_p([cb.node_id(box)])[1]
    assert vis[0] <= dense_vis[0]


def test_possemidef_goldens_and_batch():
    """possemidef (src/posdef.jl:41-68) on the GPU: the reference's exact cases (runtests.jl:933-948) and a batch of
    fitted leaf models, bit-exact vs the oracle."""
    assert np.array_equal(cb.possemidef(np.array([[4.0, 1], [1, 4]])), [[4, 1], [1, 4]])
    assert np.array_equal(cb.possemidef(np.array([[1.0, 4], [4, 1]])), [[4, 4], [4, 4]])
    assert np.array_equal(cb.possemidef(np.array([[4.5, 3], [3, 0.5]])), [[9, 3], [3, 1]])
    Qp = cb.possemidef(np.array([[0.0, 2], [2, 1]]))
    assert np.isfinite(Qp).all() and Qp[0, 1] == 2
    np.linalg.cholesky(Qp)
    for n, npts in ((7, 60), (10, 200), (40, 6)):
        root, oroot, _ = grow_both(n, 2, npts, seed=600 + n)
        models = cb.device_tree(root).fit_models()
        before = models.download()
        after = models.possemidef().download()
        ok = before["status"] == 0
        oQp = O.possemidef(before["Q"][ok])
        got = after["Q"][ok]
        got_full = np.triu(got) + np.triu(got, 1).transpose(0, 2, 1)
        assert same(got_full, oQp)
        for k in ("c", "g", "b"):
            assert same(after[k], before[k])
        fin = np.isfinite(oQp).all(axis=(1, 2))
        if fin.any():
            ev = np.linalg.eigvalsh(oQp[fin])
            assert ev.min() > -1e-8 * np.abs(oQp[fin]).max()
    leaf = cb.minimum(root)
    assert cb.value(leaf) == min(cb.value(l) for l in cb.leaves(root))


@pytest.mark.parametrize("n,npoints,skip", [(2, 40, 0), (3, 60, 0), (7, 150, 0), (8, 150, 2), (10, 300, 0), (20, 100, 0)])
def test_fit_gdiag_bit_exact(n, npoints, skip):
    """coefficients_p + fit_quadratic_gdiag! (src/cs2.jl:332-384), the chain-free fit, bit-exact vs the oracle for
    every leaf and some internal boxes."""
    root, oroot, _ = grow_both(n, 2, npoints, seed=700 + n)
    dt = cb.device_tree(root)
    oall = [b for b in O.collect(oroot) if not O.isfake(b)]
    rng = np.random.default_rng(n)
    pick = rng.choice(len(oall), size=min(400, len(oall)), replace=False)
    boxes = [oall[i] for i in pick]
    ids = np.array([O.preorder_id(b) for b in boxes], np.int32)
    fit = dt.fit_gdiag(node_ids=ids, skip=skip)
    ofit = O.fit_gdiag_batch(boxes, skip=skip, nthreads=O.max_threads())
    for k in ("c", "g", "Q", "b"):
        assert same(fit[k], ofit[k]), k
    assert np.isfinite(fit["g"]).mean() > 0.5


def test_fit_gdiag_like_reference_and_no_chain_fallback():
    """runtests.jl:851-856: for a quadratic objective the gdiag fit recovers Q ~ B and g ~ Q*b; and it succeeds on leaves where
    fit_quadratic raises 'no chain found' (SURVEY M7)."""
    rng = np.random.default_rng(11)
    n = 6
    cycle = []
    for _ in range(40):
        perm = rng.permutation(n) + 1
        cycle.append([(int(perm[i]), int(perm[i + 1])) for i in range(0, n, 2)])
    B = rng.standard_normal((n, n)); B = (B + B.T) / 2
    f = cb.quadnoise_objective(B, sigma=0.0)
    root, oroot, _ = grow_both(n, 2, 200, seed=11, cycle=cycle, objective=f)  # this tree has leaves without a chain (oracle-checked)
    dt = cb.device_tree(root)
    st = dt.fit_boxes()["status"]
    assert (st == 1).any()
    g = dt.fit_gdiag()
    det = np.isfinite(g["Q"][:, np.triu_indices(n)[0], np.triu_indices(n)[1]]).all(1) & np.isfinite(g["g"]).all(1)
    assert det[st == 1].any()
    for l in np.flatnonzero(det)[:200]:
        Qf = np.triu(g["Q"][l]) + np.triu(g["Q"][l], 1).T
        assert np.allclose(Qf, f.B, rtol=1e-5, atol=1e-7)
        assert np.allclose(g["g"][l], Qf @ g["b"][l], rtol=1e-5, atol=1e-6)
    leaf = cb.leaves(root)[int(np.flatnonzero(det & (st == 1))[0])]
    with pytest.raises(cb.ErrorException):
        cb.fit_quadratic(leaf)
    c, gg, Q, b = cb.fit_quadratic_gdiag(leaf)
    assert np.allclose(Q.full(), f.B, rtol=1e-5, atol=1e-7) and c == cb.value(leaf) and same(b, cb.position(leaf))


@pytest.mark.parametrize("case", ["cs2_n4", "cs2_n10", "cs1_n3"])
def test_committed_fixtures_on_gpu(case):
    """The CUDA path against the committed vectors of tests/golden/ (oracle-generated, see its README): leaf ids, status,
    coefficients_p, visited counts and the full CS2 fit, bit for bit, without running the oracle."""
    import importlib.util
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_fixtures", os.path.join(here, "golden", "make_fixtures.py"))
    mf = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mf)
    gold = np.load(os.path.join(here, "golden", case + ".npz"))
    root, _, X = mf.build(case)
    assert same(X, gold["X"])
    dt = cb.device_tree(root)
    leaf, status = dt.find_leaf(X)
    assert same(leaf, gold["leaf"]) and same(status, gold["status"])
    Cp, vis = dt.coefficients_p()
    n = root.n
    if mf.CASES[case]["p"] == 2:
        assert same(Cp, gold["Cp"].reshape(-1, n, n).transpose(0, 2, 1))
        fit = dt.fit_boxes()
        assert same(fit["status"], gold["fit_status"])
        for k in ("c", "g", "Q", "b"):
            assert same(fit[k], gold[k]), k
        mo = dt.fit_models().download()
        iu = np.triu_indices(n)
        assert same(mo["status"], gold["fit_status"]) and same(mo["c"], gold["c"]) and same(mo["g"], gold["g"])
        assert same(mo["Q"][:, iu[0], iu[1]], gold["Q"][:, iu[0], iu[1]])
    else:
        assert same(Cp, gold["Cp"])
    assert same(vis, gold["visited"])
