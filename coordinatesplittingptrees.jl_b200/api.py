"""The reference's function API for the hot path, routed to the GPU library.

    find_leaf_at(root, x)                 reference src/tree.jl:351-388
    coefficients_p(box; skip)             reference src/polynomials.jl:127-230
    fit_quadratic(box; skip)              reference src/cs2.jl:89-93
    fit_quadratic_native(box; skip)       reference src/cs2.jl:115-121
    modelvalue(c, g, Q, b, x)             reference src/cs2.jl:155-159
    polynomial_full / polynomial_minimal  named in the reference README only (README.md:98-99): thin aliases

Single-box calls keep the reference's calling convention and exceptions; the batched forms (2-D X, lists of boxes) are
what the accelerator is for.  The device copy of a tree is cached per tree and re-flattened when the tree has grown.
"""
import numpy as np

from . import _native as N
from . import tree as T
from .device import DeviceTree, LeafModels, FIT_OK, FIT_NO_CHAIN, FIT_CHAIN_ASSERT, FIT_CHAIN_UNDEF, Q_OUTSIDE


class SymmetricArray:
    """Mirror of the reference's SymmetricArray (src/types.jl:432-454): `.data` is the dense backing store; indexing sorts
    the (1-based) index tuple, so values live at data[i,j] with i <= j and the rest keeps its fill value (NaN)."""

    def __init__(self, data):
        self.data = np.asarray(data)

    def __getitem__(self, idx):
        idx = idx if isinstance(idx, tuple) else (idx,)
        return self.data[tuple(i - 1 for i in sorted(idx))]

    @property
    def shape(self):
        return self.data.shape

    def full(self):
        """Dense symmetric matrix (2-D only)."""
        U = np.triu(self.data)
        return U + np.triu(self.data, 1).T

    def __array__(self, dtype=None, copy=None):
        return self.full() if self.data.ndim == 2 else self.data


def device_tree(root, live=False):
    """The tree of `root`, resident on the current GPU.  live=False: re-flattened on the host and re-uploaded whenever
    addpoint!/split changed it.  live=True: the device copy grows with the tree -- only the splits made since the last
    call are sent (the split log) and the flattening is re-derived on the device (csp_tree_append_splits): what an optimiser loop
    that alternates addpoint! with queries and fits uses.  Node / leaf ids are preorder ids of the current version either way."""
    t = root.tree
    if live:
        if t._live is None:
            w = t.world
            t._live = DeviceTree.live(t.n, t.p, w.lower, w.upper, w.position, w.meta)
            t._live_splits = 0
        ns = t.split_count()
        if ns > t._live_splits:
            t._live.append_splits(*t.split_log(t._live_splits, ns))
            t._live_splits = ns
        return t._live
    if t._device is None or t._device_version != t.version:
        if t._device is not None:
            t._device.close()
        t._device = DeviceTree(t.flat())
        t._device_version = t.version
    return t._device


def find_leaf_at(root, x):
    """find_leaf_at(box, x), `box` being the root or any other box of the tree (the reference accepts any, src/tree.jl:351-360).
    1-D x (or a scalar for 1-d trees): returns the leaf Box, raising like the reference.  2-D X of shape (m, n): returns the int32
    array of leaf ids (ranks in leaves(getroot(box))); rows outside boxbounds(box) raise."""
    X = np.asarray(x, np.float64)
    single = X.ndim <= 1
    X = np.ascontiguousarray(np.atleast_1d(X).reshape(1, -1) if single else X)
    n = root.n
    if X.shape[1] != n:
        raise N.DimensionMismatch(f"tree has {n} dimensions, got {X.shape[1]}")
    leaf, status = device_tree(root).find_leaf(X, start=T.node_id(root) if root.id != 0 else None)
    bad = np.flatnonzero(status == Q_OUTSIDE)
    if bad.size:
        i = int(bad[0])
        raise N.ErrorException(f"{X[i].tolist()} not within the bounds of the box (query {i})")
    if single:
        f = root.tree.flat()
        return root._mk(f["box_of_pre"][f["leaf_nodes"][leaf[0]]])
    return leaf


_FIT_ERRORS = {FIT_NO_CHAIN: (N.ErrorException, "no chain found at {}"),
               FIT_CHAIN_ASSERT: (AssertionError, "!(wassplit[d1]) && (d2 > n || !(wassplit[d2])) at {}"),
               FIT_CHAIN_UNDEF: (N.UndefRefError, "access to undefined reference (box.split) while walking the chain of {}")}


def _raise_fit(status, box):
    if status != FIT_OK:
        exc, msg = _FIT_ERRORS[int(status)]
        raise exc(msg.format(box))


def _require_cs2(box):
    if box.p != 2:
        raise TypeError(f"no method matching fit_quadratic(::Box{{{box.p}}}): the reference defines it for Box{{2}} only")


def fit_quadratic_native(box, skip=0):
    """(c, g, Q, b, y, prec, firstmatch) of the CS2-native model (reference src/cs2.jl:115-121)."""
    _require_cs2(box)
    r = device_tree(box).fit_boxes([T.node_id(box)], skip=skip, native=True)
    _raise_fit(r["status"][0], box)
    return (float(r["c"][0]), r["gnat"][0], SymmetricArray(r["Q"][0]), r["b"][0], r["y"][0], r["prec"][0],
            r["firstmatch"][0].astype(np.int64))


def fit_quadratic(box, skip=0):
    """(c, g, Q, b): m(x) = c + g'dx + dx'Q dx/2, dx = x - b (reference src/cs2.jl:89-93)."""
    _require_cs2(box)
    r = device_tree(box).fit_boxes([T.node_id(box)], skip=skip)
    _raise_fit(r["status"][0], box)
    return float(r["c"][0]), r["g"][0], SymmetricArray(r["Q"][0]), r["b"][0]


def coefficients_p(box, callback=None, skip=0):
    """Highest-order coefficients (reference src/polynomials.jl:127-230): SymmetricArray of off-diagonal Hessian entries
    for p=2 (diagonal NaN), the gradient vector for p=1, the p-dimensional array of mixed coefficients for p>2.
    The reference's `callback(child, value)` runs once per child of every split that supplies a coefficient; the kernels do not
    call back into the host, so a callback here receives the total number of such invocations, callback(count)."""
    Cp, _, used = device_tree(box).coefficients_p([T.node_id(box)], skip=skip, counts=True)
    if callback is not None:
        callback(int(used[0]) << box.p)
    return SymmetricArray(Cp[0])


def fit_quadratic_gdiag(box, skip=0):
    """(c, g, Q, b) from Q = coefficients_p(box) followed by fit_quadratic_gdiag!(Q, box) (reference src/cs2.jl:332-384): the
    chain-free alternative to fit_quadratic (no "no chain found" failures; b = position(box))."""
    _require_cs2(box)
    r = device_tree(box).fit_gdiag([T.node_id(box)], skip=skip)
    return float(r["c"][0]), r["g"][0], SymmetricArray(r["Q"][0]), r["b"][0]


def coefficients_p_sparse(box, pattern, callback=None, skip=0):
    """coefficients_p!(Cp, box) with a sparse symmetric Cp (reference src/polynomials.jl:161-230, :244-271): `pattern` is either
    the string "tridiagonal" (SymTridiagonal: returns its off-diagonal vector ev) or a list of above-diagonal positions (i, j),
    1-based, i < j (returns {(i, j): value}).  Only those coefficients are determined and the traversal stops once they are."""
    tri = isinstance(pattern, str) and pattern == "tridiagonal"
    pat = [(i, i + 1) for i in range(1, box.n)] if tri else [tuple(sorted(p)) for p in pattern]
    vals, _, used = device_tree(box).coefficients_p_sparse(pat, [T.node_id(box)], skip=skip, counts=True)
    if callback is not None:  # total number of reference callback invocations (4 per split that supplied a coefficient)
        callback(int(used[0]) << box.p)
    return vals[0] if tri else {p: float(v) for p, v in zip(pat, vals[0])}


def polynomial_full(box, skip=0):
    """Full-degree model at `box`.  p=2: fit_quadratic.  p=1: the degree-1 model (c, g, b) with c = value(box),
    g = coefficients_p(box), b = position(box) (the reference has no CS1 quadratic fit)."""
    if box.p == 2:
        return fit_quadratic(box, skip=skip)
    g = coefficients_p(box, skip=skip)
    return T.value(box), g.data, T.position(box)


def polynomial_minimal(box, skip=0):
    """Builder-defined (no reference definition; parity unpinned): the chain-only part of the CS2 model --
    (c, g_native, Q_offdiag, b, y, prec) with Q's diagonal left NaN."""
    _require_cs2(box)
    c, g, Q, b, y, prec, _ = fit_quadratic_native(box, skip=skip)
    Qd = Q.data.copy()
    Qd[np.diag_indices(box.n)] = np.nan
    return c, g, SymmetricArray(Qd), b, y, prec


def modelvalue(c, g, Q, b, x):
    """modelvalue(c, g, Q, b, x) = c + g'(x-b) + (x-b)'Q(x-b)/2 (reference src/cs2.jl:155-159), evaluated on the GPU."""
    x = np.ascontiguousarray(np.atleast_1d(np.asarray(x, np.float64)))
    Qd = Q.data if isinstance(Q, SymmetricArray) else np.asarray(Q, np.float64)
    mo = LeafModels.upload(np.array([c], np.float64), np.asarray(g, np.float64)[None], Qd[None], np.asarray(b, np.float64)[None])
    try:
        return float(mo.evaluate(x[None], np.zeros(1, np.int32))[0])
    finally:
        mo.close()


def modelvalue_native(c, g, Q, b, y, prec, x):
    """modelvalue(c, g, Q, b, y, prec, x) for the CS2-native model (reference src/cs2.jl:142-153), evaluated on the GPU in the
    reference's loop order.  x: one point (n,) -> float, or (m, n) -> array."""
    import ctypes as C
    X = np.ascontiguousarray(np.asarray(x, np.float64))
    single = X.ndim == 1
    X = np.ascontiguousarray(X.reshape(1, -1) if single else X)
    m, n = X.shape
    Qd = np.ascontiguousarray(np.asarray(Q.data if isinstance(Q, SymmetricArray) else Q, np.float64).T)  # column-major .data
    P = np.ascontiguousarray(np.asarray(prec, bool).T.astype(np.uint8))
    g, b, y = (np.ascontiguousarray(v, np.float64) for v in (g, b, y))
    out = np.empty(m)
    N.dchk(N.dev().csp_modelvalue_native(n, float(c), g.ctypes.data, Qd.ctypes.data, b.ctypes.data, y.ctypes.data, P.ctypes.data,
                                         X.ctypes.data, m, out.ctypes.data))
    return float(out[0]) if single else out


def possemidef(Q):
    """possemidef(Q) (reference src/posdef.jl:41-68): the full symmetric matrix with the off-diagonals of Q and diagonals raised as
    needed to make it positive semi-definite; NaN entries count as 0.  Q: SymmetricArray or an n x n array whose sorted-index
    (upper) triangle holds the values.  Runs on the GPU."""
    Qd = np.asarray(Q.data if isinstance(Q, SymmetricArray) else Q, np.float64)
    n = Qd.shape[0]
    mo = LeafModels.upload(np.zeros(1), np.zeros((1, n)), Qd[None], np.zeros((1, n)))
    try:
        out = mo.possemidef().download()["Q"][0]
    finally:
        mo.close()
    U = np.triu(out)
    return U + np.triu(out, 1).T


def _leaf_box(box, leaf):
    f = box.tree.flat()
    return box._mk(f["box_of_pre"][f["leaf_nodes"][leaf]])


def extrema(box):
    """extrema(box) = (minimum(box), maximum(box)) over leaves(box) (reference src/CoordinateSplittingPTrees.jl:64-68: isless on
    value, the first of equal leaves): a reduction on the device over the resident values."""
    lmin, _, lmax, _ = device_tree(box).extrema(T.node_id(box))
    return _leaf_box(box, lmin), _leaf_box(box, lmax)


def minimum(box):
    """minimum(box): the leaf with the smallest value (reference src/CoordinateSplittingPTrees.jl:66)."""
    return extrema(box)[0]


def maximum(box):
    """maximum(box): the leaf with the largest value (reference src/CoordinateSplittingPTrees.jl:67)."""
    return extrema(box)[1]


def fit_leaves(root, skip=0):
    """Batched fit of every leaf, models kept on the device (LeafModels)."""
    return device_tree(root).fit_models(skip=skip)


def locate_eval(root, models, X):
    """Batched find_leaf_at + modelvalue: (leaf ids, values).  Raises like find_leaf_at for out-of-world rows."""
    leaf, val, status = device_tree(root).locate_eval(models, X)
    st = status.cpu().numpy() if hasattr(status, "cpu") else status
    bad = np.flatnonzero(st == Q_OUTSIDE)
    if bad.size:
        raise N.ErrorException(f"query {int(bad[0])} is not within the world bounds")
    return leaf, val
