"""Pins the CPU oracle against the reference's tests for the model-fitting half of the hot path:
/root/reference/test/runtests.jl "Split iteration" (:451-518, exact orders), "Chains" (:538-591),
"p-th order direct coefficients" (:593-648), "Sparse Hessians" dense case (:652-678), "Qcoef" (:752-822) and
"CS2 full models" (:824-859).  Randomised reference tests are unseeded; here they are seeded and assert the same
properties against analytic truth."""
import itertools
import math

import numpy as np
import pytest

from oracle import oracle as O

Inf = math.inf


def dims_of(boxes):
    return [b.split_dims for b in boxes]


def test_split_iteration_n3_fictive():  # runtests.jl:455-479
    rng = np.random.default_rng(1)
    f = lambda y: rng.random()
    n = 3
    world = O.World([-Inf] * n, [Inf] * n, [0] * n, rng.random())
    root = O.Box(world, 2)
    box = O.addpoint(root, np.ones(n), [(1, 2), (3,)], f)
    assert list(O.position(box)) == [1, 1, 1]
    box = O.addpoint(root, [0.6, 2, 2], [(2, 3), (1,)], f)
    assert list(O.position(box)) == [0.6, 2, 2]
    assert O.boxbounds(box) == [(0.5, 0.8), (1.5, Inf), (1.5, Inf)]
    assert dims_of(O.splits(box)) == [(1, 4), (2, 3), (3, 4), (1, 2)]
    box = box.parent
    assert dims_of(O.splits(box)) == [(1, 4), (2, 3), (3, 4), (1, 2)]
    box = box.parent
    assert dims_of(O.splits(box)) == [(2, 3), (1, 4), (3, 4), (1, 2)]
    box = box.parent
    assert dims_of(O.splits(box)) == [(3, 4), (2, 3), (1, 4), (1, 2)]
    box = box.parent
    assert dims_of(O.splits(box)) == [(1, 2), (3, 4), (2, 3), (1, 4)]


def test_split_iteration_n6():  # runtests.jl:481-504
    rng = np.random.default_rng(2)
    f = lambda y: rng.random()
    n = 6
    world = O.World([-Inf] * n, [Inf] * n, [0] * n, rng.random())
    root = O.Box(world, 2)
    O.addpoint(root, np.ones(n), [(1, 2), (3, 4), (5, 6)], f)
    O.addpoint(root, [0.8, 0.8, 0.8, 0.2, 0.2, 0.2], [(1, 3), (2, 5), (4, 6)], f)
    O.addpoint(root, [0.2] * n, [(1, 5), (3, 6), (2, 4)], f)
    box = O.find_leaf_at(root, [0.2, 0, 0.2, 0, 0.2, 0])
    assert dims_of(O.splits(box)) == [(2, 4), (3, 6), (1, 5), (3, 4), (1, 3), (2, 5), (4, 6), (5, 6), (1, 2)]
    box = box.parent
    assert dims_of(O.splits(box)) == [(3, 6), (2, 4), (1, 5), (3, 4), (1, 3), (2, 5), (4, 6), (5, 6), (1, 2)]
    box = O.find_leaf_at(root, [0.8, 0.8, 0.8, 0.2, 0.2, 0.2])
    assert dims_of(O.splits(box)) == [(4, 6), (2, 5), (1, 3), (5, 6), (3, 4), (1, 5), (3, 6), (2, 4), (1, 2)]
    box = O.find_leaf_at(root, np.ones(6))
    assert dims_of(O.splits(box)) == [(5, 6), (1, 3), (2, 5), (4, 6), (3, 4), (1, 5), (3, 6), (2, 4), (1, 2)]
    box = O.find_leaf_at(root, [1, 1, 0, 0, 0, 0])
    assert dims_of(O.splits(box)) == [(1, 3), (2, 5), (4, 6), (5, 6), (3, 4), (1, 5), (3, 6), (2, 4), (1, 2)]
    assert dims_of(O.splits(root)) == [(1, 2), (1, 5), (3, 6), (2, 4), (3, 4), (1, 3), (2, 5), (4, 6), (5, 6)]


def test_split_iteration_n8_unique():  # runtests.jl:506-517
    rng = np.random.default_rng(3)
    f = lambda y: rng.random()
    n = 8
    world = O.World([-Inf] * n, [Inf] * n, [0] * n, rng.random())
    for _ in range(20):
        root = O.Box(world, 2)
        O.addpoint(root, rng.standard_normal(n), [(1, 2), (3, 4), (5, 6), (7, 8)], f)
        O.addpoint(root, rng.standard_normal(n), [(3, 7), (4, 8), (1, 5), (2, 6)], f)
        O.addpoint(root, rng.standard_normal(n), [(6, 8), (1, 3), (2, 4), (5, 7)], f)
        O.addpoint(root, rng.standard_normal(n), [(4, 5), (1, 6), (2, 7), (3, 8)], f)
        lv = O.leaves(root)
        box = min(lv, key=O.value)  # minimum(root), CoordinateSplittingPTrees.jl:64-66
        s = O.splits(box)
        assert len(s) == 16 and len(set(b.h for b in s)) == 16
        # every leaf sees every split exactly once
        for leaf in lv:
            ss = O.splits(leaf)
            assert len(ss) == 16 and len(set(b.h for b in ss)) == 16


@pytest.mark.parametrize("n,dimlistss", [
    (3, ([(1, 2), (3,)], [(3, 1), (2,)])),
    (4, ([(1, 2), (3, 4)], [(3, 1), (2, 4)])),
    (4, ([(1, 2), (3, 4)], [(1, 2), (3, 4)])),
])
def test_chains(n, dimlistss):  # runtests.jl:538-591
    rng = np.random.default_rng(4)
    f = lambda y: rng.random()
    x0 = rng.standard_normal(n)
    world = O.World([-Inf] * n, [Inf] * n, x0, f(x0))
    root = box = O.Box(world, 2)
    chaintops = []
    repeats = dimlistss[1] == dimlistss[0]
    for dimlists in dimlistss:
        x = O.position(box) + 1
        bx = box
        box = O.addpoint(box, x, dimlists, f)
        for _ in range(math.ceil(n / 2)):
            chaintops.append(bx)
            if repeats:
                bx = bx.others[-1]
    if repeats:
        chaintops[-1] = chaintops[-1].parent
    i = len(chaintops)
    while not O.isroot(box):
        top, success = O.chaintop(box)
        assert success
        assert top == chaintops[i - 1]
        ch = O.chain(top)
        assert len(ch) == 2
        assert ch[0] == top and ch[1] == top.others[-1]
        for child in box.parent.others:
            top, success = O.chaintop(child)
            assert success
            assert top == chaintops[i - 1]
        i -= 1
        box = box.parent


def bprod(x, B):
    if B.ndim == 1:
        return float(B @ x)
    if B.ndim == 2:
        return float(x @ B @ x / 2)
    return float(np.einsum("ijk,i,j,k->", B, x, x, x) / 6)


def symmetrized(A):
    """SymmetricArray(randn(sz)): reads go through sorted indices, so only the sorted-index entries matter."""
    n, p = A.shape[0], A.ndim
    B = np.empty_like(A)
    for I in itertools.product(range(n), repeat=p):
        B[I] = A[tuple(sorted(I))]
    return B


def dimpair(order, p):  # test/functions.jl:84-91
    return [tuple(int(d) for d in order[i:i + p]) for i in range(0, len(order), p)]


@pytest.mark.parametrize("p", [1, 2, 3])
@pytest.mark.parametrize("n", [6, 7])
def test_direct_coefficients(p, n):  # runtests.jl:593-648
    rng = np.random.default_rng(100 * p + n)
    world = O.World([-Inf] * n, [Inf] * n, np.zeros(n), 0.0)
    B = symmetrized(rng.standard_normal((n,) * p))
    f = lambda y: bprod(y, B)
    root = O.Box(world, p)
    filled = set()
    target = set(itertools.combinations(range(1, n + 1), p))
    while filled != target:
        while True:
            pairings = dimpair(rng.permutation(n) + 1, p)
            isnovel = True
            for pr in pairings:
                if len(pr) == p:
                    isnovel = tuple(sorted(pr)) not in filled  # (last full tuple decides, as in the reference)
            if isnovel:
                break
        for pr in pairings:
            if len(pr) == p:
                filled.add(tuple(sorted(pr)))
        O.addpoint(root, rng.standard_normal(n), pairings, f)
    for box in O.leaves(root):
        Cp = O.coefficients_p(box)
        for I in target:
            assert O.sym_get(Cp, *I) == pytest.approx(B[tuple(i - 1 for i in I)], rel=1e-6, abs=1e-8)


def test_sparse_hessian_dense_case():  # runtests.jl:652-678 (dense SymmetricArray output; callback count 24 = 6 used splits x 4)
    fsparse = lambda x: float(np.sum(np.diff(x) ** 2) / 2)
    x0 = np.array([0.1, 0.2, 0.3, 0.4])
    world = O.World.unbounded(x0, fsparse)
    root = O.Box(world, 2)
    O.addpoint(root, 2 * x0, [(1, 3), (2, 4)], fsparse)
    O.addpoint(root, 3 * x0, [(2, 3), (1, 4)], fsparse)
    O.addpoint(root, 4 * x0, [(1, 4), (2, 3)], fsparse)
    box = O.addpoint(root, 5 * x0, [(1, 2), (3, 4)], fsparse)
    Q, visited, used = O.coefficients_p(box, stats=True)
    for i in (1, 2, 3):
        assert O.sym_get(Q, i, i + 1) == pytest.approx(-1)
        assert O.sym_get(Q, i + 1, i) == pytest.approx(-1)
        assert math.isnan(O.sym_get(Q, i, i))
    assert abs(O.sym_get(Q, 1, 3)) < 1e-8 and abs(O.sym_get(Q, 2, 4)) < 1e-8 and abs(O.sym_get(Q, 1, 4)) < 1e-8
    assert used * 4 == 24
    # SymmetricArray storage: values live in the sorted-index (upper) triangle, the rest keeps the NaN fill
    assert np.all(np.isnan(Q[np.tril_indices(4)]))


@pytest.mark.parametrize("n", [7, 8])
def test_qcoef(n):  # runtests.jl:752-822
    rng = np.random.default_rng(n)
    dims = rng.permutation(n) + 1
    seen = np.zeros(n, bool)
    prec = np.zeros((n, n), bool)
    for i in range(0, n, 2):
        d1 = d2 = dims[i]
        prec[:, d1 - 1] = seen
        if i + 1 < n:
            d2 = dims[i + 1]
            prec[:, d2 - 1] = seen
        seen[d1 - 1] = True
        if i + 1 < n:
            seen[d2 - 1] = True
    b = rng.standard_normal(n)
    y = rng.standard_normal(n)
    x = b.copy()
    chi = lambda j, i: O.chi(j, i, b, y, prec)
    for i in range(0, n, 2):
        d1 = d2 = int(dims[i])
        dx1 = dx2 = y[d1 - 1] - b[d1 - 1]
        if i + 1 < n:
            d2 = int(dims[i + 1])
            dx2 = y[d2 - 1] - b[d2 - 1]
        for k in range(1, n + 1):
            if k in (d1, d2):
                continue
            assert O.Qcoef_lineq(d1, k, dx1, x[k - 1], b[k - 1], y[k - 1], prec[d1 - 1, k - 1]) == 0
            assert O.Qcoef_lineq(d2, k, dx2, x[k - 1], b[k - 1], y[k - 1], prec[d2 - 1, k - 1]) == 0
        x[d1 - 1] = y[d1 - 1]
        if i + 1 < n:
            x[d2 - 1] = y[d2 - 1]
        for k in range(1, n + 1):
            if k in (d1, d2):
                continue
            for a, c in ((k, d1), (d1, k), (k, d2), (d2, k)):
                assert O.Qcoef_value(a, c, x, b, y, prec) == 0
            for d in (d1, d2):
                assert abs((x[d - 1] - b[d - 1]) * (x[k - 1] - chi(k, d)) / 2 + (x[k - 1] - b[k - 1]) * (x[d - 1] - chi(d, k)) / 2) < 1e-12
        assert O.Qcoef_value(d1, d1, x, b, y, prec) == 0
        assert O.Qcoef_value(d2, d2, x, b, y, prec) == 0
        if d1 != d2:
            assert O.Qcoef_value(d1, d2, x, b, y, prec) == pytest.approx(
                (x[d1 - 1] - b[d1 - 1]) * (x[d2 - 1] - chi(d2, d1)) / 4 + (x[d2 - 1] - b[d2 - 1]) * (x[d1 - 1] - chi(d1, d2)) / 4)
    x = rng.standard_normal(n)
    for j in range(1, n + 1):
        for i in range(1, n + 1):
            assert O.Qcoef_value(i, j, x, b, y, prec) == pytest.approx(
                (x[i - 1] - b[i - 1]) * (x[j - 1] - chi(j, i)) / 4 + (x[j - 1] - b[j - 1]) * (x[i - 1] - chi(i, j)) / 4)


def test_qcoef_lineq_precedence_quirk():
    """cs2.jl:317 parses as  A || (B && return coef): case A falls through to the general expression."""
    b, y = 0.3, 0.9
    assert O.Qcoef_lineq(1, 1, b - y, y, b, y, True) == 0.0                      # case B: exact zero returned
    assert O.Qcoef_lineq(1, 1, y - b, b, b, y, True) == (y - b) / 2 + b - (b + y) / 2  # case A: evaluated in floating point


def random_pairings(n, rng):
    return dimpair(rng.permutation(n) + 1, 2)


@pytest.mark.parametrize("n", [2, 3, 7, 8])
def test_cs2_full_models(n):  # runtests.jl:824-859 (dimlists chosen at random instead of by BlossomV)
    rng = np.random.default_rng(50 + n)
    B = symmetrized(rng.standard_normal((n, n)))
    f = lambda x: float(x @ B @ x / 2)
    x0 = rng.standard_normal(n)
    world = O.World([-Inf] * n, [Inf] * n, x0, f(x0))
    nfit = 0
    for it in range(10):
        root = O.Box(world, 2)
        while True:
            box = O.addpoint(root, rng.standard_normal(n), random_pairings(n, rng), f)
            Cp = O.coefficients_p(box)
            if not np.any(np.isnan(Cp[np.triu_indices(n, 1)])):
                break
        box = O.addpoint(root, rng.standard_normal(n), random_pairings(n, rng), f)
        try:
            c, g, Q, b, y, prec, firstmatch = O.fit_quadratic_native(box)
        except O.JlError as e:  # chaintop can fail on trees not grown with matched dimlists
            assert "no chain found" in str(e) or "AssertionError" in str(e)
            continue
        if np.any(np.isnan(np.diag(Q))):
            continue
        nfit += 1
        for leaf in O.leaves(root):
            x = O.position(leaf)
            assert O.modelvalue_native(c, g, Q, b, y, prec, x) == pytest.approx(f(x), rel=1e-7, abs=1e-9)
            assert O.modelvalue_chi(c, g, Q, b, y, prec, x) == pytest.approx(f(x), rel=1e-7, abs=1e-9)
        c, g, Q, b = O.fit_quadratic(box)
        Qs = O.symmetrize(Q)
        assert np.allclose(Qs, B, rtol=1e-6, atol=1e-8)
        assert np.allclose(g, Qs @ b, rtol=1e-6, atol=1e-8)
        x = O.position(box)
        assert c + g @ (x - b) + (x - b) @ Qs @ (x - b) / 2 == pytest.approx(f(x), rel=1e-8, abs=1e-10)
        assert abs(c - g @ b + (b @ Qs @ b) / 2) < 1e-8
        assert O.modelvalue(c, g, Q, b, x) == pytest.approx(f(x), rel=1e-8, abs=1e-10)
        # gdiag variant (runtests.jl:853-856)
        Q0 = O.coefficients_p(box)
        c2, g2, Q2, b2 = O.fit_quadratic_gdiag(Q0, box)
        Q2s = O.symmetrize(Q2)
        assert np.allclose(Q2s, B, rtol=1e-6, atol=1e-8)
        assert np.allclose(g2, Q2s @ b2, rtol=1e-6, atol=1e-8)
    assert nfit >= 5


def test_sparse_hessian_patterns():  # runtests.jl:679-699: SymTridiagonal (hesscount 12) and sparse SymmetricArray (16) outputs
    fsparse = lambda x: float(np.sum(np.diff(x) ** 2) / 2)
    x0 = np.array([0.1, 0.2, 0.3, 0.4])
    world = O.World.unbounded(x0, fsparse)
    root = O.Box(world, 2)
    O.addpoint(root, 2 * x0, [(1, 3), (2, 4)], fsparse)
    O.addpoint(root, 3 * x0, [(2, 3), (1, 4)], fsparse)
    O.addpoint(root, 4 * x0, [(1, 4), (2, 3)], fsparse)
    box = O.addpoint(root, 5 * x0, [(1, 2), (3, 4)], fsparse)
    vals, vis, used, _ = O.coefficients_p_pattern_batch([box], O.tridiagonal_pattern(4))
    assert np.allclose(vals[0], -1) and used[0] * 4 == 12
    # sparse([1,1,2,2,3,3,4,1], [1,2,2,3,3,4,4,4], ...): above-diagonal stored entries (1,2),(2,3),(3,4),(1,4); nnz_offdiag_sym == 4
    pat = [(1, 2), (2, 3), (3, 4), (1, 4)]
    vals, vis, used, _ = O.coefficients_p_pattern_batch([box], pat)
    assert np.allclose(vals[0][:3], -1) and abs(vals[0][3]) < 1e-8 and used[0] * 4 == 16
    # the pattern output agrees with the dense output on the stored entries and stops earlier
    Q, dvis, dused = O.coefficients_p(box, stats=True)
    assert all(vals[0][k] == O.sym_get(Q, *pat[k]) for k in range(4)) and vis[0] <= dvis


def test_possemidef_goldens():  # runtests.jl:933-948 ("posdef"), exact equality as in the reference
    assert np.array_equal(O.possemidef(np.array([[4.0, 1], [1, 4]])), [[4, 1], [1, 4]])
    assert np.array_equal(O.possemidef(np.array([[1.0, 4], [4, 1]])), [[4, 4], [4, 4]])
    assert np.array_equal(O.possemidef(np.array([[4.5, 3], [3, 0.5]])), [[9, 3], [3, 1]])
    Qp = O.possemidef(np.array([[0.0, 2], [2, 1]]))
    assert np.isfinite(Qp).all() and Qp[0, 1] == 2
    np.linalg.cholesky(Qp)
    # only the SymmetricArray (sorted-index) triangle is read
    assert np.array_equal(O.possemidef(np.array([[1.0, 4], [np.nan, 1]])), [[4, 4], [4, 4]])
    # NaN off-diagonals are treated as 0 (and stay NaN in the copy); random matrices become positive semi-definite
    rng = np.random.default_rng(0)
    for n in (3, 6, 11):
        A = rng.standard_normal((n, n)); A = (A + A.T) / 2
        Qp = O.possemidef(np.triu(A))
        assert np.allclose(Qp - np.diag(np.diag(Qp)), A - np.diag(np.diag(A)))
        assert np.linalg.eigvalsh(Qp).min() > -1e-9 * np.abs(Qp).max()


@pytest.mark.parametrize("case", ["cs2_n4", "cs2_n10", "cs1_n3"])
def test_committed_fixtures_match_oracle(case):
    """tests/golden/*.npz (oracle outputs on small seeded trees, tests/golden/make_fixtures.py) still equal what the oracle
    computes today, bit for bit: the vectors the GPU parity tests compare with cannot drift silently."""
    import importlib.util
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_fixtures", os.path.join(here, "golden", "make_fixtures.py"))
    mf = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mf)
    gold = np.load(os.path.join(here, "golden", case + ".npz"))
    root, oroot, X = mf.build(case)
    assert np.array_equal(X, gold["X"])
    out = mf.oracle_outputs(oroot, X, mf.CASES[case]["p"])
    assert set(out) == set(gold.files)
    for k in gold.files:
        assert np.array_equal(out[k], gold[k], equal_nan=True), k
