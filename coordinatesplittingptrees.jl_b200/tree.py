"""Host-side mirror of the reference's tree API (World, Box{p}, addpoint!, position, value, leaves ...).

In the reference these are Julia types and functions (src/types.jl, src/tree.jl, src/CoordinateSplittingPTrees.jl) and
they stay on the host in this design; a Julia wrapper keeps using the reference's own `Box` graph and only flattens it
(see INTEGRATION.md).  Julia is not installed here, so this module plays the host role on top of libcsp_host.so, with
the same names, argument meaning and error behaviour (Julia's `addpoint!` is spelled `addpoint`; dimensions are
1-based as in the reference).  Tree *growth* happens here; every batched query / fit / evaluation goes to the GPU
library through `device.py`.
"""
import ctypes as C

import numpy as np

from . import _native as N
from ._native import c_f64p, c_i32p, c_u8p


def _f64(x):
    return np.ascontiguousarray(np.atleast_1d(np.asarray(x, dtype=np.float64)))


def _p(a, t):
    return a.ctypes.data_as(t)


class NativeObjective:
    """A compiled objective `double f(const double* y, int n, void* ctx)` (see csrc/objectives.cpp)."""

    def __init__(self, fn_ptr, ctx=None, keepalive=None, pyfunc=None):
        self.fn_ptr, self.ctx, self._keep, self._py = fn_ptr, ctx, keepalive, pyfunc

    def __call__(self, y):
        if self._py is not None:
            return self._py(y)
        y = _f64(y)
        return N.OBJECTIVE(self.fn_ptr)(_p(y, c_f64p), len(y), self.ctx)


def quadnoise_objective(B, sigma=0.0, seed=0):
    """f(x) = x'Bx/2 + sigma*eta(x) with B symmetrised; eta is a deterministic hash noise in [-1, 1)."""
    B = np.asarray(B, np.float64)
    B = np.ascontiguousarray((B + B.T) / 2)
    n = B.shape[0]
    L = N.host()
    ctx = L.csph_quadnoise_new(n, _p(B, c_f64p), float(sigma), int(seed))
    obj = NativeObjective(L.csph_objective_ptr(0), ctx)
    obj.B, obj.sigma = B, sigma
    return obj


def rosenbrock_objective():
    return NativeObjective(N.host().csph_objective_ptr(1), None)


def _as_objective(f):
    if isinstance(f, NativeObjective):
        return f.fn_ptr, f.ctx, f
    if not callable(f):
        raise TypeError("metagen must be callable")

    def tramp(yp, n, ctx):
        return float(f(np.ctypeslib.as_array(yp, shape=(n,)).copy()))
    cb = N.OBJECTIVE(tramp)
    return C.cast(cb, C.c_void_p).value, None, cb


class World:
    """World(lower, upper, position, meta) (reference src/types.jl:7-53).  `meta` may be a number or a function of the
    base position (metagen, :55-56).  Coordinates are promoted to Float64 (boxcoordtype, :1-2)."""

    def __init__(self, lower, upper=None, position=None, meta=None):
        if position is None and meta is None and upper is not None:
            # World(position, meta): unbounded domain (src/types.jl:49-53)
            position, meta = lower, upper
            pos = _f64(position)
            lower, upper = np.full(len(pos), -np.inf), np.full(len(pos), np.inf)
        lo, up, pos = _f64(lower), _f64(upper), _f64(position)
        if not (len(lo) == len(up) == len(pos)):
            raise N.DimensionMismatch(f"lower, upper, and position must have equal length (got {len(lo)}, {len(up)}, {len(pos)})")
        for l, x, u in zip(lo, pos, up):
            if not (l <= x <= u):
                raise N.ArgumentError(f"position {x} is not within bounds [{l}, {u}]")
        self.lower, self.upper, self.position = lo, up, pos
        self.meta = float(meta(pos.copy())) if callable(meta) else float(meta)

    def __len__(self):
        return len(self.lower)

    ndims = property(lambda self: len(self.lower))


class _Tree:
    """Owner of one native host tree; shared by all Box handles of that tree."""

    def __init__(self, world, p):
        self.world, self.p, self.n = world, p, len(world)
        h = C.c_void_p()
        N.hchk(N.host().csph_tree_new(self.n, p, _p(world.lower, c_f64p), _p(world.upper, c_f64p), _p(world.position, c_f64p),
                                      world.meta, C.byref(h)))
        self.h = h
        self.version = 0       # bumped by every split: device copies are re-flattened when stale
        self._flat = None
        self._device = None
        self._device_version = -1
        self._live = None      # device copy grown from the split log (api.device_tree(root, live=True))
        self._live_splits = 0  # splits already appended to it
        self._keep = []

    def __del__(self):
        try:
            if self._device is not None:
                self._device.close()
            if self._live is not None:
                self._live.close()
            N.host().csph_tree_free(self.h)
        except Exception:
            pass

    def counts(self):
        v = [C.c_int32() for _ in range(5)]
        N.hchk(N.host().csph_counts(self.h, *[C.byref(x) for x in v]))
        return dict(n=v[0].value, p=v[1].value, n_nodes=v[2].value, n_internal=v[3].value, n_leaves=v[4].value)

    def split_count(self):
        v = C.c_int32()
        N.hchk(N.host().csph_split_count(self.h, C.byref(v)))
        return v.value

    def split_log(self, s0, s1):
        """Entries [s0, s1) of the split log (creation order): (split_box, dims, xs, child_val) as csp_tree_append_splits takes them."""
        k, p, Cn = s1 - s0, self.p, 1 << self.p
        sb, d, x, cv = np.zeros(k, np.int32), np.zeros(k * p, np.int32), np.zeros(k * p), np.zeros(k * Cn)
        if k:
            N.hchk(N.host().csph_split_log(self.h, s0, s1, _p(sb, c_i32p), _p(d, c_i32p), _p(x, c_f64p), _p(cv, c_f64p)))
        return sb, d, x, cv

    def flat(self):
        """Preorder SoA arrays == the fields of csp_tree_desc (include/csp_b200.h), plus the box<->node id maps."""
        if self._flat is not None and self._flat["version"] == self.version:
            return self._flat
        c = self.counts()
        n, p, nn, ni = c["n"], c["p"], c["n_nodes"], c["n_internal"]
        Cn = 1 << p
        f = dict(c, version=self.version, lower=np.zeros(n), upper=np.zeros(n), parent=np.zeros(nn, np.int32),
                 childindex=np.zeros(nn, np.uint8), internal_id=np.zeros(nn, np.int32), leaf_id=np.zeros(nn, np.int32),
                 val=np.zeros(nn), inode=np.zeros(ni, np.int32), dims=np.zeros(ni * p, np.int32), xs=np.zeros(ni * p),
                 child=np.zeros(ni * Cn, np.int32), pos=np.zeros(ni * n), pre_of_box=np.zeros(nn, np.int32),
                 box_of_pre=np.zeros(nn, np.int32))
        N.hchk(N.host().csph_flatten(self.h, _p(f["lower"], c_f64p), _p(f["upper"], c_f64p), _p(f["parent"], c_i32p),
                                     _p(f["childindex"], c_u8p), _p(f["internal_id"], c_i32p), _p(f["leaf_id"], c_i32p),
                                     _p(f["val"], c_f64p), _p(f["inode"], c_i32p), _p(f["dims"], c_i32p), _p(f["xs"], c_f64p),
                                     _p(f["child"], c_i32p), _p(f["pos"], c_f64p), _p(f["pre_of_box"], c_i32p),
                                     _p(f["box_of_pre"], c_i32p)))
        f["leaf_nodes"] = np.flatnonzero(f["leaf_id"] >= 0).astype(np.int32)  # node id of each leaf, in leaves(root) order
        self._flat = f
        return f

    def info(self, b):
        par, ci, fc, fake = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        val = C.c_double()
        pos, dims, xs = np.zeros(self.n), np.zeros(self.p, np.int32), np.zeros(self.p)
        N.hchk(N.host().csph_box_info(self.h, b, C.byref(par), C.byref(ci), C.byref(fc), C.byref(fake), C.byref(val), _p(pos, c_f64p),
                                      _p(dims, c_i32p), _p(xs, c_f64p)))
        return dict(parent=par.value, childindex=ci.value, first_child=fc.value, isfake=bool(fake.value), val=val.value, pos=pos,
                    dims=tuple(int(d) for d in dims), xs=tuple(float(x) for x in xs))


class Box:
    """Box(world, p) creates the root of a CSp-tree (`Box{p}(world)`, reference src/types.jl:194-201,248); other boxes are
    handles (tree, id) returned by split / addpoint / find_leaf_at."""
    __slots__ = ("tree", "id")

    def __init__(self, world, p=None, _tree=None, _id=0):
        if _tree is None:
            if p is None:
                raise TypeError("Box(world, p): the degree p is required")
            if p > 8:
                raise N.ErrorException("degree must be less than or equal to 8")
            _tree = _Tree(world, p)
        self.tree, self.id = _tree, _id

    def _mk(self, i):
        return Box(None, _tree=self.tree, _id=int(i))

    def __eq__(self, o):
        return isinstance(o, Box) and o.tree is self.tree and o.id == self.id

    def __hash__(self):
        return hash((id(self.tree), self.id))

    def __repr__(self):
        return f"Box{{{self.tree.p}}}#{self.id}({value(self)!r}@{position(self).tolist()})"

    world = property(lambda self: self.tree.world)
    n = property(lambda self: self.tree.n)
    p = property(lambda self: self.tree.p)

    @property
    def parent(self):
        return self._mk(self.tree.info(self.id)["parent"])

    @property
    def childindex(self):
        return self.tree.info(self.id)["childindex"]

    def child(self, k):
        fc = self.tree.info(self.id)["first_child"]
        if fc < 0:
            raise N.UndefRefError("access to undefined reference (box.split of a leaf)")
        if not 0 <= k < (1 << self.p):
            raise N.BoundsError(f"child index {k}")
        return self._mk(fc + k)

    @property
    def split_dims(self):
        i = self.tree.info(self.id)
        if i["first_child"] < 0:
            raise N.UndefRefError("access to undefined reference (box.split of a leaf)")
        return i["dims"]

    @property
    def split_xs(self):
        i = self.tree.info(self.id)
        if i["first_child"] < 0:
            raise N.UndefRefError("access to undefined reference (box.split of a leaf)")
        return i["xs"]

    @property
    def split_self(self):
        return self.child(0)

    @property
    def others(self):
        return [self.child(k) for k in range(1, 1 << self.p)]


def ndims(box):
    return box.n


def degree(box):
    return box.p


def isroot(box):
    return box.tree.info(box.id)["parent"] == box.id


def isleaf(box):
    return box.tree.info(box.id)["first_child"] < 0


def isfake(box):
    return box.tree.info(box.id)["isfake"]


def getroot(box):
    return box._mk(0)


def getleaf(box):
    while not isleaf(box):
        box = box.child(0)
    return box


def position(box, d=None):
    """position(box) / position(box, d) (reference src/tree.jl:71-93)."""
    pos = box.tree.info(box.id)["pos"]
    if d is None:
        return pos
    if not 1 <= d <= box.n:
        raise N.BoundsError(f"dimension {d}")
    return float(pos[d - 1])


def meta(box):
    return box.tree.info(box.id)["val"]


value = meta  # value(box) = value(meta(box)) with Real metadata (reference src/CoordinateSplittingPTrees.jl:61-62)


def boxbounds(box, d=None):
    """boxbounds(box[, d]) (reference src/tree.jl:136-216)."""
    out = np.zeros(2)
    if d is None:
        res = []
        for k in range(1, box.n + 1):
            N.hchk(N.host().csph_boxbounds(box.tree.h, box.id, k, _p(out, c_f64p)))
            res.append((float(out[0]), float(out[1])))
        return res
    N.hchk(N.host().csph_boxbounds(box.tree.h, box.id, int(d), _p(out, c_f64p)))
    return (float(out[0]), float(out[1]))


def split(parent, dims, xs, metas):
    """Box(parent, splitdims, xs, metas) (reference src/types.jl:203-298): returns the 2^p-1 non-self children."""
    dims = np.ascontiguousarray(np.atleast_1d(np.asarray(dims, np.int32)))
    xs, metas = _f64(xs), _f64(metas)
    if len(xs) != len(dims):
        raise N.DimensionMismatch("splitdims and xs must have the same length")
    out = np.zeros((1 << parent.p) - 1, np.int32)
    N.hchk(N.host().csph_split(parent.tree.h, parent.id, len(dims), _p(dims, c_i32p), _p(xs, c_f64p), len(metas), _p(metas, c_f64p),
                               _p(out, c_i32p)))
    parent.tree.version += 1
    return [parent._mk(i) for i in out]


def _flat_dimlists(dimlists):
    flat = np.ascontiguousarray([d for dl in dimlists for d in dl], np.int32)
    lens = np.ascontiguousarray([len(dl) for dl in dimlists], np.int32)
    return flat, lens


def addpoint(box, x, dimlists, metagen):
    """addpoint!(box, x, dimlists, metagen) (reference src/CoordinateSplittingPTrees.jl:94-142): returns the final leaf,
    whose position is x.  The 3-argument automatic form needs BlossomV and is out of scope."""
    x = _f64(x)
    flat, lens = _flat_dimlists(dimlists)
    fp, ctx, keep = _as_objective(metagen)
    out = C.c_int32()
    rc = N.host().csph_addpoint(box.tree.h, box.id, _p(x, c_f64p), len(x), len(lens), _p(flat, c_i32p), _p(lens, c_i32p), fp, ctx,
                                C.byref(out))
    box.tree.version += 1
    N.hchk(rc)
    return box._mk(out.value)


def addpoints(root, X, dimlist_cycle, metagen):
    """Bulk growth: one addpoint!(root, X[i], dimlist_cycle[i % len], metagen) per row of X, in native code."""
    X = np.ascontiguousarray(X, np.float64)
    if X.ndim != 2 or X.shape[1] != root.n:
        raise N.DimensionMismatch("X must be (npoints, n)")
    per = np.ascontiguousarray([len(c) for c in dimlist_cycle], np.int32)
    flat = np.ascontiguousarray([d for c in dimlist_cycle for dl in c for d in dl], np.int32)
    lens = np.ascontiguousarray([len(dl) for c in dimlist_cycle for dl in c], np.int32)
    fp, ctx, keep = _as_objective(metagen)
    rc = N.host().csph_addpoints_cycle(root.tree.h, _p(X, c_f64p), X.shape[0], X.shape[1], len(per), _p(per, c_i32p), _p(flat, c_i32p),
                                       _p(lens, c_i32p), fp, ctx)
    root.tree.version += 1
    N.hchk(rc)


def collect(root):
    """collect(root): every box below root in preorder (reference src/tree.jl:393-415)."""
    if root.id != 0:
        out, stack = [], [root]
        while stack:
            b = stack.pop()
            out.append(b)
            if not isleaf(b):
                stack.extend(reversed([b.child(k) for k in range(1 << b.p)]))
        return out
    f = root.tree.flat()
    return [root._mk(i) for i in f["box_of_pre"]]


def leaves(root):
    """leaves(root): non-fake leaf boxes in depth-first order (reference src/tree.jl:423)."""
    if root.id != 0:
        return [b for b in collect(root) if isleaf(b) and not isfake(b)]
    f = root.tree.flat()
    return [root._mk(f["box_of_pre"][v]) for v in f["leaf_nodes"]]


def length(root):
    return len(collect(root))


def node_id(box):
    """Preorder node id of a box in the current flattening (what the C ABI calls a node id)."""
    return int(box.tree.flat()["pre_of_box"][box.id])


def leaf_id(box):
    """Rank of a leaf box in leaves(root) (what K1 returns), or -1."""
    f = box.tree.flat()
    return int(f["leaf_id"][f["pre_of_box"][box.id]])
