"""ctypes binding of the CPU oracle (oracle/libcsp_oracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module; the product
package never does.  The API mirrors the reference's Julia names (World, Box, addpoint!, find_leaf_at, position,
boxbounds, meta/value, leaves, splits, chaintop, coefficients_p, fit_quadratic[_native], modelvalue ...), so the
golden tests in tests/test_oracle_*.py read like /root/reference/test/runtests.jl.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcsp_oracle.so")

OBJECTIVE = C.CFUNCTYPE(C.c_double, C.POINTER(C.c_double), C.c_int, C.c_void_p)


class JlError(Exception):
    """A Julia exception the reference would throw; .kind names the class."""
    KINDS = {1: "ErrorException", 2: "DimensionMismatch", 3: "AssertionError", 4: "UndefRefError", 5: "ArgumentError",
             6: "BoundsError"}

    def __init__(self, code, msg):
        self.kind = self.KINDS.get(-code, "ErrorException")
        super().__init__(f"{self.kind}: {msg}")


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("capi.cpp", "csp_oracle.hpp")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.cspo_last_error.restype = C.c_char_p
        for name in ("cspo_parent", "cspo_getroot", "cspo_getleaf"):
            getattr(_lib, name).restype = C.c_void_p
            getattr(_lib, name).argtypes = [C.c_void_p]
        for name in ("cspo_collect", "cspo_leaves", "cspo_splits"):
            getattr(_lib, name).restype = C.c_long
            getattr(_lib, name).argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_long]
        for name in ("cspo_modelvalue_native", "cspo_modelvalue_chi", "cspo_modelvalue_canonical", "cspo_Qcoef_value",
                     "cspo_Qcoef_lineq", "cspo_chi"):
            getattr(_lib, name).restype = C.c_double
        _lib.cspo_Qcoef_lineq.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int]
    return _lib


def _chk(rc):
    if rc != 0:
        raise JlError(rc, lib().cspo_last_error().decode())


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _up(a):
    return a.ctypes.data_as(C.POINTER(C.c_ubyte))


def _f64(x):
    return np.ascontiguousarray(np.atleast_1d(np.asarray(x, dtype=np.float64)))


def as_objective(f):
    """Python callable f(y::ndarray)->float  ->  C objective pointer (kept alive by the caller)."""
    def tramp(yp, n, ctx):
        return float(f(np.ctypeslib.as_array(yp, shape=(n,)).copy()))
    return OBJECTIVE(tramp)


class World:
    def __init__(self, lower, upper, position, meta):
        lo, up, pos = _f64(lower), _f64(upper), _f64(position)
        if not (len(lo) == len(up) == len(pos)):
            raise JlError(-2, "lower, upper, and position must have equal length")
        self.n = len(lo)
        if callable(meta):  # metagen(f::Function, x) = f(x), types.jl:56
            meta = meta(pos.copy())
        self.lower, self.upper, self.position, self.meta = lo, up, pos, float(meta)
        h = C.c_void_p()
        _chk(lib().cspo_world_new(self.n, _dp(lo), _dp(up), _dp(pos), C.c_double(self.meta), C.byref(h)))
        self.h = h

    @classmethod
    def unbounded(cls, position, meta):  # World(position, meta), types.jl:49-53
        pos = _f64(position)
        return cls(np.full(len(pos), -np.inf), np.full(len(pos), np.inf), pos, meta)


class Box:
    """Handle to a node of the oracle's pointer tree.  Box(world, p) creates a root (`Box{p}(world)`)."""
    __slots__ = ("h", "world", "p")

    def __init__(self, world, p, _h=None):
        self.world, self.p = world, p
        if _h is None:
            h = C.c_void_p()
            _chk(lib().cspo_root_new(world.h, p, C.byref(h)))
            _h = h.value
        self.h = _h

    def _mk(self, h):
        return Box(self.world, self.p, h)

    def __eq__(self, o):
        return isinstance(o, Box) and o.h == self.h

    def __hash__(self):
        return hash(self.h)

    @property
    def n(self):
        return self.world.n

    @property
    def parent(self):
        return self._mk(lib().cspo_parent(C.c_void_p(self.h)))

    @property
    def childindex(self):
        return lib().cspo_childindex(C.c_void_p(self.h))

    def child(self, idx):  # getchild(box.split, idx); idx 0 = self
        h = C.c_void_p()
        _chk(lib().cspo_getchild(C.c_void_p(self.h), idx, C.byref(h)))
        return self._mk(h.value)

    @property
    def split_self(self):
        return self.child(0)

    @property
    def others(self):  # box.split.others.children
        return [self.child(i) for i in range(1, 1 << self.p)]

    @property
    def split_dims(self):
        d = np.zeros(self.p, np.int32)
        x = np.zeros(self.p)
        _chk(lib().cspo_split_info(C.c_void_p(self.h), _ip(d), _dp(x)))
        return tuple(int(v) for v in d)

    @property
    def split_xs(self):
        d = np.zeros(self.p, np.int32)
        x = np.zeros(self.p)
        _chk(lib().cspo_split_info(C.c_void_p(self.h), _ip(d), _dp(x)))
        return tuple(float(v) for v in x)


def split(parent, dims, xs, metas):
    """Box(parent, splitdims, xs, metas) -> tuple of the 2^p-1 'other' children (types.jl:270-298)."""
    dims = np.atleast_1d(np.asarray(dims, np.int32))
    xs, metas = _f64(xs), _f64(metas)
    out = (C.c_void_p * ((1 << parent.p) - 1))()
    _chk(lib().cspo_split(C.c_void_p(parent.h), len(dims), _ip(dims), _dp(xs), len(metas), _dp(metas), out))
    return [parent._mk(h) for h in out]


def _dimlists_flat(dimlists):
    flat = np.asarray([d for dl in dimlists for d in dl], np.int32)
    lens = np.asarray([len(dl) for dl in dimlists], np.int32)
    return flat, lens


def addpoint(box, x, dimlists, metagen):
    """addpoint!(box, x, dimlists, metagen) (CoordinateSplittingPTrees.jl:94-142).  metagen: python callable or a
    (function-pointer, ctx) pair from the objectives library."""
    x = _f64(x)
    flat, lens = _dimlists_flat(dimlists)
    if callable(metagen):
        fp, ctx, keep = as_objective(metagen), None, None
    else:
        fp, ctx = metagen
    out = C.c_void_p()
    _chk(lib().cspo_addpoint(C.c_void_p(box.h), _dp(x), len(x), len(lens), _ip(flat), _ip(lens), fp, ctx, C.byref(out)))
    return box._mk(out.value)


def addpoints_cycle(root, X, cycle, fp, ctx):
    """One addpoint! per row of X (m x n, C-contiguous), dimlists taken cyclically from `cycle`; fp/ctx native objective."""
    X = np.ascontiguousarray(X, np.float64)
    per = np.asarray([len(c) for c in cycle], np.int32)
    flat = np.asarray([d for c in cycle for dl in c for d in dl], np.int32)
    lens = np.asarray([len(dl) for c in cycle for dl in c], np.int32)
    _chk(lib().cspo_addpoints_cycle(C.c_void_p(root.h), _dp(X), C.c_long(X.shape[0]), X.shape[1], len(cycle), _ip(per), _ip(flat),
                                    _ip(lens), fp, ctx))


def isleaf(b):
    return bool(lib().cspo_isleaf(C.c_void_p(b.h)))


def isroot(b):
    return bool(lib().cspo_isroot(C.c_void_p(b.h)))


def isfake(b):
    return bool(lib().cspo_isfake(C.c_void_p(b.h)))


def getroot(b):
    return b._mk(lib().cspo_getroot(C.c_void_p(b.h)))


def getleaf(b):
    return b._mk(lib().cspo_getleaf(C.c_void_p(b.h)))


def position(b, d=None):
    if d is None:
        out = np.zeros(b.n)
        _chk(lib().cspo_position(C.c_void_p(b.h), _dp(out)))
        return out
    v = C.c_double()
    _chk(lib().cspo_position_d(C.c_void_p(b.h), int(d), C.byref(v)))
    return v.value


def boxbounds(b, d=None):
    if d is None:
        out = np.zeros(2 * b.n)
        _chk(lib().cspo_boxbounds(C.c_void_p(b.h), _dp(out)))
        return [(float(out[2 * i]), float(out[2 * i + 1])) for i in range(b.n)]
    out = np.zeros(2)
    _chk(lib().cspo_boxbounds_d(C.c_void_p(b.h), int(d), _dp(out)))
    return (float(out[0]), float(out[1]))


def meta(b):
    v = C.c_double()
    _chk(lib().cspo_meta(C.c_void_p(b.h), C.byref(v)))
    return v.value


value = meta


def find_leaf_at(root, x):
    x = _f64(x)
    out = C.c_void_p()
    _chk(lib().cspo_find_leaf_at(C.c_void_p(root.h), _dp(x), len(x), C.byref(out)))
    return root._mk(out.value)


def _boxlist(fn, b):
    cnt = fn(C.c_void_p(b.h), None, 0)
    buf = (C.c_void_p * max(cnt, 1))()
    fn(C.c_void_p(b.h), buf, cnt)
    return [b._mk(buf[i]) for i in range(cnt)]


def collect(root):
    return _boxlist(lib().cspo_collect, root)


def leaves(root):
    return _boxlist(lib().cspo_leaves, root)


def splits(box):
    """[the box owning each split yielded by splits(box)], in iteration order (tree.jl:461-540)."""
    return _boxlist(lib().cspo_splits, box)


def chain(box):
    cnt = C.c_long()
    cap = box.n + 2
    buf = (C.c_void_p * cap)()
    _chk(lib().cspo_chain(C.c_void_p(box.h), buf, cap, C.byref(cnt)))
    return [box._mk(buf[i]) for i in range(cnt.value)]


def chaintop(box):
    top, ok = C.c_void_p(), C.c_int()
    _chk(lib().cspo_chaintop(C.c_void_p(box.h), C.byref(top), C.byref(ok)))
    return box._mk(top.value), bool(ok.value)


def index_tree(root):
    lib().cspo_index_tree(C.c_void_p(root.h))


def preorder_id(b):
    return lib().cspo_preorder_id(C.c_void_p(b.h))


def leaf_rank(b):
    return lib().cspo_leaf_rank(C.c_void_p(b.h))


def coefficients_p(box, skip=0, stats=False):
    """Dense SymmetricArray .data (shape (n,)*p, Julia column-major == symmetric-sorted storage)."""
    n, p = box.n, box.p
    out = np.zeros(n ** p)
    vis, used = C.c_long(), C.c_long()
    _chk(lib().cspo_coefficients_p(C.c_void_p(box.h), skip, _dp(out), C.byref(vis), C.byref(used)))
    arr = out.reshape((n,) * p, order="F")
    return (arr, vis.value, used.value) if stats else arr


def sym_get(data, *idx):
    """SymmetricArray getindex (types.jl:440-443): 1-based indices, sorted."""
    return data[tuple(i - 1 for i in sorted(idx))]


def fit_quadratic_native(box, skip=0):
    n = box.n
    c = C.c_double()
    g, Q, b, y = np.zeros(n), np.zeros(n * n), np.zeros(n), np.zeros(n)
    prec = np.zeros(n * n, np.uint8)
    fm = np.zeros(n, np.int32)
    _chk(lib().cspo_fit_quadratic_native(C.c_void_p(box.h), skip, C.byref(c), _dp(g), _dp(Q), _dp(b), _dp(y), _up(prec), _ip(fm)))
    return c.value, g, Q.reshape(n, n, order="F"), b, y, prec.reshape(n, n, order="F").astype(bool), fm


def fit_quadratic(box, skip=0):
    n = box.n
    c = C.c_double()
    g, Q, b = np.zeros(n), np.zeros(n * n), np.zeros(n)
    _chk(lib().cspo_fit_quadratic(C.c_void_p(box.h), skip, C.byref(c), _dp(g), _dp(Q), _dp(b)))
    return c.value, g, Q.reshape(n, n, order="F"), b


def fit_quadratic_gdiag(Qdata, box, skip=0):
    n = box.n
    Q = np.asfortranarray(Qdata, dtype=np.float64).reshape(-1, order="F").copy()
    c = C.c_double()
    g, b = np.zeros(n), np.zeros(n)
    _chk(lib().cspo_fit_quadratic_gdiag(C.c_void_p(box.h), skip, _dp(Q), C.byref(c), _dp(g), _dp(b)))
    return c.value, g, Q.reshape(n, n, order="F"), b


def symmetrize(Qdata):
    """Full symmetric matrix from SymmetricArray .data (upper triangle holds the values)."""
    U = np.triu(Qdata)
    return U + np.triu(Qdata, 1).T


def _flatF(a, dt=np.float64):
    return np.asfortranarray(a, dtype=dt).reshape(-1, order="F").copy()


def modelvalue_native(c, g, Q, b, y, prec, x):
    n = len(x)
    return lib().cspo_modelvalue_native(n, C.c_double(c), _dp(_f64(g)), _dp(_flatF(Q)), _dp(_f64(b)), _dp(_f64(y)),
                                        _up(_flatF(prec, np.uint8)), _dp(_f64(x)))


def modelvalue_chi(c, g, Q, b, y, prec, x):
    n = len(x)
    return lib().cspo_modelvalue_chi(n, C.c_double(c), _dp(_f64(g)), _dp(_flatF(Q)), _dp(_f64(b)), _dp(_f64(y)),
                                     _up(_flatF(prec, np.uint8)), _dp(_f64(x)))


def modelvalue(c, g, Q, b, x):
    n = len(x)
    return lib().cspo_modelvalue_canonical(n, C.c_double(c), _dp(_f64(g)), _dp(_flatF(Q)), _dp(_f64(b)), _dp(_f64(x)))


def Qcoef_value(i, j, x, b, y, prec):
    n = len(x)
    return lib().cspo_Qcoef_value(n, i, j, _dp(_f64(x)), _dp(_f64(b)), _dp(_f64(y)), _up(_flatF(prec, np.uint8)))


def Qcoef_lineq(i, k, dxi, xk, bk, yk, precik):
    return lib().cspo_Qcoef_lineq(i, k, dxi, xk, bk, yk, int(bool(precik)))


def chi(j, i, b, y, prec):
    n = len(b)
    return lib().cspo_chi(n, j, i, _dp(_f64(b)), _dp(_f64(y)), _up(_flatF(prec, np.uint8)))


def tree_counts(root):
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    _chk(lib().cspo_tree_counts(C.c_void_p(root.h), C.byref(a), C.byref(b), C.byref(c)))
    return a.value, b.value, c.value


def export_flat(root):
    """Preorder flat arrays (the csp_tree_desc fields) computed with the oracle's own position/meta walks."""
    n, p = root.n, root.p
    nn, ni, nl = tree_counts(root)
    Cn = 1 << p
    d = dict(n=n, p=p, n_nodes=nn, n_internal=ni, n_leaves=nl, lower=root.world.lower.copy(), upper=root.world.upper.copy(),
             parent=np.zeros(nn, np.int32), childindex=np.zeros(nn, np.uint8), internal_id=np.zeros(nn, np.int32),
             leaf_id=np.zeros(nn, np.int32), val=np.zeros(nn), inode=np.zeros(max(ni, 1), np.int32)[:ni],
             dims=np.zeros(ni * p, np.int32), xs=np.zeros(ni * p), child=np.zeros(ni * Cn, np.int32), pos=np.zeros(ni * n))
    _chk(lib().cspo_export_flat(C.c_void_p(root.h), _ip(d["parent"]), _up(d["childindex"]), _ip(d["internal_id"]), _ip(d["leaf_id"]),
                                _dp(d["val"]), _ip(d["inode"]), _ip(d["dims"]), _dp(d["xs"]), _ip(d["child"]), _dp(d["pos"])))
    return d


def _boxarr(boxes):
    arr = (C.c_void_p * len(boxes))()
    for i, b in enumerate(boxes):
        arr[i] = b.h
    return arr


def find_leaf_batch(root, X, nthreads=1):
    """leaf ranks (index_tree must have been called) + status (1 = reference throws) + seconds."""
    X = np.ascontiguousarray(X, np.float64)
    m, n = X.shape
    leaf = np.zeros(m, np.int32)
    st = np.zeros(m, np.uint8)
    sec = C.c_double()
    _chk(lib().cspo_find_leaf_batch(C.c_void_p(root.h), _dp(X), C.c_long(m), n, _ip(leaf), _up(st), nthreads, C.byref(sec)))
    return leaf, st, sec.value


def fit_boxes_batch(boxes, skip=0, nthreads=1, native=False):
    nb = len(boxes)
    n = boxes[0].n
    c, g, Q, b = np.zeros(nb), np.zeros((nb, n)), np.zeros((nb, n * n)), np.zeros((nb, n))
    st = np.zeros(nb, np.uint8)
    gn = np.zeros((nb, n)) if native else None
    y = np.zeros((nb, n)) if native else None
    prec = np.zeros((nb, n * n), np.uint8) if native else None
    sec = C.c_double()
    _chk(lib().cspo_fit_boxes_batch(_boxarr(boxes), C.c_long(nb), skip, _dp(c), _dp(g), _dp(Q), _dp(b),
                                    _dp(gn) if native else None, _dp(y) if native else None, _up(prec) if native else None,
                                    _up(st), nthreads, C.byref(sec)))
    out = dict(c=c, g=g, Q=Q.reshape(nb, n, n).transpose(0, 2, 1), b=b, status=st, seconds=sec.value)  # Q[l][i][j] = data[i,j]
    if native:
        out.update(gnat=gn, y=y, prec=prec.reshape(nb, n, n).transpose(0, 2, 1).astype(bool))
    return out


def coefficients_p_batch(boxes, skip=0, nthreads=1):
    nb = len(boxes)
    n, p = boxes[0].n, boxes[0].p
    out = np.zeros((nb, n ** p))
    vis = np.zeros(nb, np.int64)
    sec = C.c_double()
    _chk(lib().cspo_coefficients_p_batch(_boxarr(boxes), C.c_long(nb), skip, _dp(out), vis.ctypes.data_as(C.POINTER(C.c_long)),
                                         nthreads, C.byref(sec)))
    return out, vis, sec.value


def coefficients_p_pattern_batch(boxes, pattern, skip=0, nthreads=1):
    """coefficients_p! into a sparse symmetric output with the given above-diagonal pattern [(i, j), ...] (1-based, i < j).
    Returns (values (nb, npat), visited, used, seconds)."""
    nb, npat = len(boxes), len(pattern)
    pi = np.ascontiguousarray([p[0] for p in pattern], np.int32)
    pj = np.ascontiguousarray([p[1] for p in pattern], np.int32)
    out = np.zeros((nb, max(npat, 1)))[:, :npat].copy()
    vis, used = np.zeros(nb, np.int64), np.zeros(nb, np.int64)
    sec = C.c_double()
    _chk(lib().cspo_coefficients_p_pattern_batch(_boxarr(boxes), C.c_long(nb), skip, npat, _ip(pi), _ip(pj), _dp(out),
                                                 vis.ctypes.data_as(C.POINTER(C.c_long)), used.ctypes.data_as(C.POINTER(C.c_long)),
                                                 nthreads, C.byref(sec)))
    return out, vis, used, sec.value


def fit_gdiag_batch(boxes, skip=0, nthreads=1):
    nb, n = len(boxes), boxes[0].n
    c, g, Q, b = np.zeros(nb), np.zeros((nb, n)), np.zeros((nb, n * n)), np.zeros((nb, n))
    _chk(lib().cspo_fit_gdiag_batch(_boxarr(boxes), C.c_long(nb), skip, _dp(c), _dp(g), _dp(Q), _dp(b), nthreads))
    return dict(c=c, g=g, Q=Q.reshape(nb, n, n).transpose(0, 2, 1), b=b)


def possemidef(Qdata):
    """possemidef (posdef.jl:41-68) of matrices given in the SymmetricArray .data layout: Qdata (nb, n, n) with Qdata[l][i][j] =
    data[i,j] (values at i <= j).  Returns the full adjusted matrices (nb, n, n)."""
    Q = np.asarray(Qdata, np.float64)
    single = Q.ndim == 2
    Q = Q[None] if single else Q
    nb, n, _ = Q.shape
    Qf = np.ascontiguousarray(Q.transpose(0, 2, 1))
    out = np.zeros_like(Qf)
    _chk(lib().cspo_possemidef_batch(n, C.c_long(nb), _dp(Qf), _dp(out)))
    out = out.transpose(0, 2, 1)
    return out[0] if single else out


def tridiagonal_pattern(n):
    return [(i, i + 1) for i in range(1, n)]


def locate_eval_batch(root, X, c, g, Qdata, b, nthreads=1, Qf=None):
    """Qdata: (nl, n, n) with Qdata[l][i][j] = SymmetricArray .data[i,j] (as returned by fit_boxes_batch); or pass Qf, the same
    already in the library's column-major per-leaf layout (Qf[l][j][i] = .data[i,j]), to skip the conversion copy."""
    X = np.ascontiguousarray(X, np.float64)
    m, n = X.shape
    if Qf is None:
        Qf = np.ascontiguousarray(np.asarray(Qdata).transpose(0, 2, 1))  # back to column-major per leaf
    assert Qf.flags.c_contiguous and Qf.dtype == np.float64
    leaf, val, st = np.zeros(m, np.int32), np.zeros(m), np.zeros(m, np.uint8)
    sec = C.c_double()
    _chk(lib().cspo_locate_eval_batch(C.c_void_p(root.h), _dp(X), C.c_long(m), n, _dp(np.ascontiguousarray(c)),
                                      _dp(np.ascontiguousarray(g)), _dp(Qf), _dp(np.ascontiguousarray(b)), _ip(leaf), _dp(val),
                                      _up(st), nthreads, C.byref(sec)))
    return leaf, val, st, sec.value


def eval_batch(c, g, Qdata, b, X, leaf, nthreads=1):
    X = np.ascontiguousarray(X, np.float64)
    m, n = X.shape
    Qf = np.ascontiguousarray(np.asarray(Qdata).transpose(0, 2, 1))
    leaf = np.ascontiguousarray(leaf, np.int32)
    val = np.zeros(m)
    sec = C.c_double()
    _chk(lib().cspo_eval_batch(n, _dp(np.ascontiguousarray(c)), _dp(np.ascontiguousarray(g)), _dp(Qf), _dp(np.ascontiguousarray(b)),
                               _dp(X), _ip(leaf), C.c_long(m), _dp(val), nthreads, C.byref(sec)))
    return val, sec.value


def max_threads():
    return lib().cspo_max_threads()
