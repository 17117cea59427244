"""Device-side objects: the flattened tree resident in HBM, per-leaf models, and the batched operations.

Everything here calls libcsp_b200.so through the C ABI of include/csp_b200.h.  Host buffers are numpy arrays; resident
buffers are torch CUDA tensors (torch is used only for device memory and streams -- plumbing, not compute).  There is
no CPU fallback: without the built library or without a CUDA device these calls raise.
"""
import ctypes as C

import numpy as np

from . import _native as N
from ._native import TreeDesc, c_f64p, c_i32p, c_i64p, c_u8p

Q_OK, Q_OUTSIDE = 0, 1
FIT_OK, FIT_NO_CHAIN, FIT_CHAIN_ASSERT, FIT_CHAIN_UNDEF = 0, 1, 2, 3


def _is_cuda_tensor(x):
    return hasattr(x, "data_ptr") and getattr(x, "is_cuda", False)


def _ptr(a):
    if a is None:
        return None
    if _is_cuda_tensor(a):
        return a.data_ptr()
    return a.ctypes.data


_NP_OF = {"int32": np.int32, "float64": np.float64, "uint8": np.uint8}


def _check_buf(buf, m, dtype, like, what):
    """Validate a caller-supplied in/out buffer before its pointer crosses the C ABI: element type, contiguity, length >= m,
    and residency matching the query array (CUDA tensor on the same device, or numpy)."""
    if buf is None:
        return None
    if _is_cuda_tensor(like):
        import torch
        if not _is_cuda_tensor(buf) or buf.device != like.device:
            raise TypeError(f"{what} must be a CUDA tensor on {like.device} (the queries are resident there)")
        if buf.dtype != getattr(torch, dtype) or not buf.is_contiguous():
            raise TypeError(f"{what} must be a contiguous {dtype} tensor")
        if buf.numel() < m:
            raise N.DimensionMismatch(f"{what} holds {buf.numel()} elements, {m} needed")
    else:
        if not isinstance(buf, np.ndarray) or buf.dtype != _NP_OF[dtype] or not buf.flags.c_contiguous or not buf.flags.writeable:
            raise TypeError(f"{what} must be a writable C-contiguous numpy {dtype} array")
        if buf.size < m:
            raise N.DimensionMismatch(f"{what} holds {buf.size} elements, {m} needed")
    return buf


def _stream_ptr(stream):
    if stream is None:
        try:
            import torch
            return torch.cuda.current_stream().cuda_stream
        except Exception:
            return None
    return getattr(stream, "cuda_stream", stream)


def make_desc(flat):
    """csp_tree_desc over the arrays of a flat dict (tree._Tree.flat() or the oracle's export_flat)."""
    d = TreeDesc()
    d.n, d.p, d.n_nodes, d.n_internal, d.n_leaves = flat["n"], flat["p"], flat["n_nodes"], flat["n_internal"], flat["n_leaves"]
    keep = {}
    for name, dt, ct in (("lower", np.float64, c_f64p), ("upper", np.float64, c_f64p), ("parent", np.int32, c_i32p),
                         ("childindex", np.uint8, c_u8p), ("internal_id", np.int32, c_i32p), ("leaf_id", np.int32, c_i32p),
                         ("val", np.float64, c_f64p), ("inode", np.int32, c_i32p), ("dims", np.int32, c_i32p),
                         ("xs", np.float64, c_f64p), ("child", np.int32, c_i32p), ("pos", np.float64, c_f64p)):
        a = np.ascontiguousarray(flat[name], dt)
        keep[name] = a
        setattr(d, name, a.ctypes.data_as(ct) if a.size else C.cast(None, ct))
    return d, keep


def pack_blob(flat):
    """The flattened tree as one relocatable byte blob (numpy uint8): the wire / on-disk format."""
    d, keep = make_desc(flat)
    sz = C.c_size_t()
    N.dchk(N.dev().csp_tree_blob_size(C.byref(d), C.byref(sz)))
    blob = np.zeros(sz.value, np.uint8)
    N.dchk(N.dev().csp_tree_pack(C.byref(d), blob.ctypes.data, sz.value))
    return blob


def save_blob(flat, path):
    """Write the flattened tree to `path` in the wire format (csp_tree_pack): the persistent on-disk form of a tree."""
    pack_blob(flat).tofile(path)


def load_blob(path):
    """Read a blob written by save_blob (numpy uint8); pass it to DeviceTree(blob=...)."""
    return np.fromfile(path, dtype=np.uint8)


class DeviceTree:
    """Flattened CSp-tree resident on one GPU (csp_tree)."""

    def __init__(self, flat=None, blob=None, device=None):
        L = N.dev()
        if device is not None:
            N.dchk(L.csp_set_device(int(device)))
        h = C.c_void_p()
        if blob is not None:
            nbytes = blob.numel() if _is_cuda_tensor(blob) else blob.nbytes
            N.dchk(L.csp_tree_upload_blob(_ptr(blob), nbytes, C.byref(h)))
        else:
            d, keep = make_desc(flat)
            N.dchk(L.csp_tree_upload(C.byref(d), C.byref(h)))
        self.h = h
        info = (C.c_int64 * 8)()
        N.dchk(L.csp_tree_info(h, info))
        self.n, self.p, self.n_nodes, self.n_internal, self.n_leaves, self.device, self.max_depth, self.kib = [int(v) for v in info]

    def export_derived(self, which):
        """Diagnostic (csp_tree_export_derived): one of the device-derived arrays as raw bytes (numpy uint8)."""
        sz = C.c_size_t()
        N.dchk(N.dev().csp_tree_export_derived(self.h, int(which), None, 0, C.byref(sz)))
        out = np.zeros(sz.value, np.uint8)
        if sz.value:
            N.dchk(N.dev().csp_tree_export_derived(self.h, int(which), out.ctypes.data, sz.value, C.byref(sz)))
        return out

    # ---- incremental residency (live trees: csp_tree_create_live / csp_tree_append_splits) -----------------------------
    @classmethod
    def live(cls, n, p, lower, upper, root_position, root_value, device=None):
        """A tree that grows on the device from its split log (include/csp_b200.h, "incremental residency")."""
        L = N.dev()
        if device is not None:
            N.dchk(L.csp_set_device(int(device)))
        lo, up, pos = (np.ascontiguousarray(a, np.float64) for a in (lower, upper, root_position))
        h = C.c_void_p()
        N.dchk(L.csp_tree_create_live(int(n), int(p), _ptr(lo), _ptr(up), _ptr(pos), float(root_value), C.byref(h)))
        t = cls.__new__(cls)
        t.h = h
        t._refresh()
        return t

    def _refresh(self):
        info = (C.c_int64 * 8)()
        N.dchk(N.dev().csp_tree_info(self.h, info))
        self.n, self.p, self.n_nodes, self.n_internal, self.n_leaves, self.device, self.max_depth, self.kib = [int(v) for v in info]

    def append_splits(self, split_box, dims, xs, child_val):
        """Append splits (creation order) to a live tree and re-derive the flattening on the device."""
        sb = np.ascontiguousarray(split_box, np.int32)
        d = np.ascontiguousarray(dims, np.int32)
        x = np.ascontiguousarray(xs, np.float64)
        cv = np.ascontiguousarray(child_val, np.float64)
        N.dchk(N.dev().csp_tree_append_splits(self.h, len(sb), _ptr(sb), _ptr(d), _ptr(x), _ptr(cv)))
        self._refresh()

    def live_maps(self):
        """(node_of_box, box_of_node) of a live tree: creation-order box ids <-> current preorder node ids."""
        a, b = np.empty(self.n_nodes, np.int32), np.empty(self.n_nodes, np.int32)
        N.dchk(N.dev().csp_tree_live_maps(self.h, _ptr(a), _ptr(b)))
        return a, b

    def close(self):
        if self.h:
            N.dev().csp_tree_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- K1 -------------------------------------------------------------------------------------------------
    def find_leaf(self, X, leaf=None, status=None, stream=None, start=None):
        """Batched find_leaf_at.  X: (m, n) float64 numpy array (host path) or torch CUDA tensor (resident path).
        start: node id of the box the search starts from (None / 0: the root; any other box uses the general boxbounds check of
        the reference, csp_find_leaf_from).  Returns (leaf ids int32, status uint8)."""
        m = self._check_X(X)
        _check_buf(leaf, m, "int32", X, "leaf")
        _check_buf(status, m, "uint8", X, "status")
        if _is_cuda_tensor(X):
            import torch
            leaf = torch.empty(m, dtype=torch.int32, device=X.device) if leaf is None else leaf
            status = torch.empty(m, dtype=torch.uint8, device=X.device) if status is None else status
            if start:
                N.dchk(N.dev().csp_find_leaf_from_dev(self.h, int(start), _ptr(X), m, _ptr(leaf), _ptr(status), _stream_ptr(stream)))
            else:
                N.dchk(N.dev().csp_find_leaf_dev(self.h, _ptr(X), m, _ptr(leaf), _ptr(status), _stream_ptr(stream)))
        else:
            X = np.ascontiguousarray(X, np.float64)
            leaf = np.empty(m, np.int32) if leaf is None else leaf
            status = np.empty(m, np.uint8) if status is None else status
            if start:
                N.dchk(N.dev().csp_find_leaf_from(self.h, int(start), _ptr(X), m, _ptr(leaf), _ptr(status)))
            else:
                N.dchk(N.dev().csp_find_leaf(self.h, _ptr(X), m, _ptr(leaf), _ptr(status)))
        return leaf, status

    def boxbounds(self, node_id):
        """boxbounds(box) of any node, computed on the device (csp_boxbounds): (lower[n], upper[n])."""
        lo, hi = np.empty(self.n), np.empty(self.n)
        N.dchk(N.dev().csp_boxbounds(self.h, int(node_id), _ptr(lo), _ptr(hi)))
        return lo, hi

    def extrema(self, node_id=0):
        """(leaf_min, value_min, leaf_max, value_max) over leaves(box) (csp_tree_extrema, isless order, first leaf wins ties)."""
        lmin, lmax = C.c_int32(), C.c_int32()
        vmin, vmax = C.c_double(), C.c_double()
        N.dchk(N.dev().csp_tree_extrema(self.h, int(node_id), C.byref(lmin), C.byref(vmin), C.byref(lmax), C.byref(vmax)))
        return lmin.value, vmin.value, lmax.value, vmax.value

    def _check_X(self, X):
        shape = tuple(X.shape)
        if len(shape) != 2 or shape[1] != self.n:
            raise N.DimensionMismatch(f"tree has {self.n} dimensions, got a query array of shape {shape}")
        if _is_cuda_tensor(X):
            import torch
            if X.dtype != torch.float64 or not X.is_contiguous():
                raise TypeError("resident queries must be a contiguous float64 CUDA tensor")
        return shape[0]

    # ---- K2 -------------------------------------------------------------------------------------------------
    def fit_boxes(self, node_ids=None, skip=0, native=False):
        """fit_quadratic for the listed node ids (None: every leaf).  Returns a dict of host arrays:
        c (nb,), g (nb,n), Q (nb,n,n) with Q[l,i,j] = SymmetricArray .data[i,j], b (nb,n), status (nb,);
        with native=True also gnat, y, prec (nb,n,n bool), firstmatch (nb,n int32)."""
        n = self.n
        if node_ids is None:
            nb, ids = self.n_leaves, None
        else:
            ids = np.ascontiguousarray(node_ids, np.int32)
            nb = len(ids)
        c, g, Q, b = np.empty(nb), np.empty((nb, n)), np.empty((nb, n * n)), np.empty((nb, n))
        st = np.empty(nb, np.uint8)
        gn = np.empty((nb, n)) if native else None
        y = np.empty((nb, n)) if native else None
        prec = np.empty((nb, n * n), np.uint8) if native else None
        fm = np.empty((nb, n), np.int32) if native else None
        N.dchk(N.dev().csp_fit_boxes(self.h, _ptr(ids), nb, skip, _ptr(c), _ptr(g), _ptr(Q), _ptr(b), _ptr(gn), _ptr(y), _ptr(prec),
                                     _ptr(fm), _ptr(st)))
        out = dict(c=c, g=g, Q=Q.reshape(nb, n, n).transpose(0, 2, 1), b=b, status=st)
        if native:
            out.update(gnat=gn, y=y, prec=prec.reshape(nb, n, n).transpose(0, 2, 1).astype(bool), firstmatch=fm)
        return out

    def fit_gdiag(self, node_ids=None, skip=0):
        """coefficients_p + fit_quadratic_gdiag! (chain-free fit) for the listed node ids (None: every leaf): dict c, g, Q, b."""
        n = self.n
        if node_ids is None:
            nb, ids = self.n_leaves, None
        else:
            ids = np.ascontiguousarray(node_ids, np.int32)
            nb = len(ids)
        c, g, Q, b = np.empty(nb), np.empty((nb, n)), np.empty((nb, n * n)), np.empty((nb, n))
        N.dchk(N.dev().csp_fit_gdiag(self.h, _ptr(ids), nb, skip, _ptr(c), _ptr(g), _ptr(Q), _ptr(b)))
        return dict(c=c, g=g, Q=Q.reshape(nb, n, n).transpose(0, 2, 1), b=b)

    def coefficients_p(self, node_ids=None, skip=0, counts=False):
        """coefficients_p for the listed node ids (None: every leaf): (Cp, visited) -- with counts=True (Cp, visited, used), used =
        number of splits that supplied a coefficient (the reference's callback runs 2^p times for each).  Cp is (nb,n) for p=1,
        (nb,n,n) (.data layout, Cp[l,i,j] = data[i,j]) for p=2 and (nb,)+(n,)*p with Cp[l,i1,..,ip] = data[i1,..,ip] for p>2."""
        n = self.n
        if node_ids is None:
            nb, ids = self.n_leaves, None
        else:
            ids = np.ascontiguousarray(node_ids, np.int32)
            nb = len(ids)
        per = n ** self.p
        Cp = np.empty((nb, per))
        vis, used = np.empty(nb, np.int64), np.empty(nb, np.int64)
        N.dchk(N.dev().csp_coefficients_p(self.h, _ptr(ids), nb, skip, _ptr(Cp), _ptr(vis), _ptr(used)))
        if self.p >= 2:  # column-major .data per box -> numpy index order
            Cp = Cp.reshape((nb,) + (n,) * self.p).transpose((0,) + tuple(range(self.p, 0, -1)))
        return (Cp, vis, used) if counts else (Cp, vis)

    def coefficients_p_sparse(self, pattern, node_ids=None, skip=0, counts=False):
        """coefficients_p! into a sparse symmetric output: pattern = [(i, j), ...] above-diagonal positions (1-based, i < j).
        Returns (values (nb, npat) in pattern order, visited (nb,)) and, with counts=True, also used (nb,)."""
        if node_ids is None:
            nb, ids = self.n_leaves, None
        else:
            ids = np.ascontiguousarray(node_ids, np.int32)
            nb = len(ids)
        pi = np.ascontiguousarray([p[0] for p in pattern], np.int32)
        pj = np.ascontiguousarray([p[1] for p in pattern], np.int32)
        vals = np.empty((nb, len(pattern)))
        vis, used = np.empty(nb, np.int64), np.empty(nb, np.int64)
        N.dchk(N.dev().csp_coefficients_p_sparse(self.h, _ptr(ids), nb, skip, len(pattern), _ptr(pi), _ptr(pj), _ptr(vals), _ptr(vis),
                                                 _ptr(used)))
        return (vals, vis, used) if counts else (vals, vis)

    def fit_models(self, skip=0):
        """Fit every leaf and keep the canonical models on the device (csp_models_fit)."""
        h = C.c_void_p()
        N.dchk(N.dev().csp_models_fit(self.h, skip, C.byref(h)))
        return LeafModels(h)

    def alloc_models(self):
        """An unfitted LeafModels for this tree (csp_models_alloc), to be filled by refit_models / refit_models_range."""
        h = C.c_void_p()
        N.dchk(N.dev().csp_models_alloc(self.h, C.byref(h)))
        return LeafModels(h)

    def refit_models_range(self, models, leaf_lo, leaf_hi, skip=0, stream=None):
        """Re-fit only the leaves [leaf_lo, leaf_hi) into `models` (the per-GPU piece of a leaf-sharded fit)."""
        N.dchk(N.dev().csp_models_refit_range(self.h, skip, models.h, int(leaf_lo), int(leaf_hi), _stream_ptr(stream)))
        return models

    def refit_models(self, models, skip=0, stream=None):
        """Re-fit every leaf into an existing LeafModels (asynchronous, no allocation)."""
        N.dchk(N.dev().csp_models_refit(self.h, skip, models.h, _stream_ptr(stream)))
        return models

    # ---- K3 -------------------------------------------------------------------------------------------------
    def locate_eval(self, models, X, leaf=None, val=None, status=None, stream=None):
        """Fused find_leaf_at + modelvalue.  Returns (leaf ids, values, status)."""
        m = self._check_X(X)
        _check_buf(leaf, m, "int32", X, "leaf")
        _check_buf(val, m, "float64", X, "val")
        _check_buf(status, m, "uint8", X, "status")
        if _is_cuda_tensor(X):
            import torch
            leaf = torch.empty(m, dtype=torch.int32, device=X.device) if leaf is None else leaf
            val = torch.empty(m, dtype=torch.float64, device=X.device) if val is None else val
            status = torch.empty(m, dtype=torch.uint8, device=X.device) if status is None else status
            N.dchk(N.dev().csp_locate_eval_dev(self.h, models.h, _ptr(X), m, _ptr(leaf), _ptr(val), _ptr(status), _stream_ptr(stream)))
        else:
            X = np.ascontiguousarray(X, np.float64)
            leaf = np.empty(m, np.int32) if leaf is None else leaf
            val = np.empty(m, np.float64) if val is None else val
            status = np.empty(m, np.uint8) if status is None else status
            N.dchk(N.dev().csp_locate_eval(self.h, models.h, _ptr(X), m, _ptr(leaf), _ptr(val), _ptr(status)))
        return leaf, val, status


class LeafModels:
    """Per-leaf canonical models (c, g, Q, b) resident on one GPU (csp_models)."""

    def __init__(self, h):
        self.h = h
        info = (C.c_int64 * 4)()
        N.dchk(N.dev().csp_models_info(h, info))
        self.n, self.n_models, self.degree, self.device = [int(v) for v in info]

    @classmethod
    def upload(cls, c, g, Q, b, device=None):
        """Caller-supplied models: c (nl,), g (nl,n), Q (nl,n,n) with Q[l,i,j] = .data[i,j] (upper triangle used) or None
        for degree-1 models, b (nl,n)."""
        if device is not None:
            N.dchk(N.dev().csp_set_device(int(device)))
        c = np.ascontiguousarray(c, np.float64)
        g = np.ascontiguousarray(g, np.float64)
        b = np.ascontiguousarray(b, np.float64)
        nl, n = g.shape
        Qf = None if Q is None else np.ascontiguousarray(np.asarray(Q, np.float64).transpose(0, 2, 1))
        h = C.c_void_p()
        N.dchk(N.dev().csp_models_upload(n, nl, _ptr(c), _ptr(g), _ptr(Qf), _ptr(b), C.byref(h)))
        return cls(h)

    def download(self):
        nl, n = self.n_models, self.n
        c, g, b = np.empty(nl), np.empty((nl, n)), np.empty((nl, n))
        Q = np.empty((nl, n * n)) if self.degree == 2 else None
        st = np.empty(nl, np.uint8)
        N.dchk(N.dev().csp_models_download(self.h, _ptr(c), _ptr(g), _ptr(Q), _ptr(b), _ptr(st)))
        out = dict(c=c, g=g, b=b, status=st)
        if Q is not None:
            out["Q"] = Q.reshape(nl, n, n).transpose(0, 2, 1)
        return out

    def download_rows(self, leaf_ids):
        """The records of the listed leaves only (csp_models_download_rows): dict like download()."""
        ids = np.ascontiguousarray(leaf_ids, np.int32)
        nb, n = len(ids), self.n
        c, g, b = np.empty(nb), np.empty((nb, n)), np.empty((nb, n))
        Q = np.empty((nb, n * n)) if self.degree == 2 else None
        st = np.empty(nb, np.uint8)
        N.dchk(N.dev().csp_models_download_rows(self.h, _ptr(ids), nb, _ptr(c), _ptr(g), _ptr(Q), _ptr(b), _ptr(st)))
        out = dict(c=c, g=g, b=b, status=st)
        if Q is not None:
            out["Q"] = Q.reshape(nb, n, n).transpose(0, 2, 1)
        return out

    def possemidef(self, stream=None):
        """possemidef (reference src/posdef.jl:41-68) of every model's Q, in place on the device."""
        N.dchk(N.dev().csp_models_possemidef(self.h, _stream_ptr(stream)))
        return self

    def evaluate(self, X, leaf, val=None, stream=None):
        """modelvalue of model leaf[i] at X[i] (K3, leaf ids supplied)."""
        shape = tuple(X.shape)
        if len(shape) != 2 or shape[1] != self.n:
            raise N.DimensionMismatch(f"models have {self.n} dimensions, got a query array of shape {shape}")
        m = shape[0]
        _check_buf(val, m, "float64", X, "val")
        if _is_cuda_tensor(X):
            import torch
            if X.dtype != torch.float64 or not X.is_contiguous():
                raise TypeError("resident queries must be a contiguous float64 CUDA tensor")
            if not _is_cuda_tensor(leaf):
                leaf = torch.as_tensor(np.ascontiguousarray(leaf, np.int32), device=X.device)
            elif leaf.dtype != torch.int32 or not leaf.is_contiguous() or leaf.device != X.device:
                leaf = leaf.to(device=X.device, dtype=torch.int32).contiguous()
            if leaf.numel() < m:
                raise N.DimensionMismatch(f"leaf holds {leaf.numel()} ids, {m} needed")
            val = torch.empty(m, dtype=torch.float64, device=X.device) if val is None else val
            N.dchk(N.dev().csp_eval_dev(self.h, _ptr(X), _ptr(leaf), m, _ptr(val), _stream_ptr(stream)))
        else:
            X = np.ascontiguousarray(X, np.float64)
            leaf = np.ascontiguousarray(leaf, np.int32)
            if leaf.size < m:
                raise N.DimensionMismatch(f"leaf holds {leaf.size} ids, {m} needed")
            val = np.empty(m, np.float64) if val is None else val
            N.dchk(N.dev().csp_eval(self.h, _ptr(X), _ptr(leaf), m, _ptr(val)))
        return val

    def close(self):
        if self.h:
            N.dev().csp_models_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def kernel_launches():
    return int(N.dev().csp_kernel_launches())


def device_count():
    c = C.c_int()
    N.dchk(N.dev().csp_device_count(C.byref(c)))
    return c.value


KERNEL_CLASSES = ("locate", "locate_eval_fused", "scan", "scatter", "eval_binned", "eval", "fit")


def timing_enable(on=True):
    N.dchk(N.dev().csp_timing_enable(1 if on else 0))


def timing_read():
    """{kernel class: (milliseconds, launches)} accumulated since the last read (synchronizes the device)."""
    ms = (C.c_double * 7)()
    cnt = (C.c_int64 * 7)()
    N.dchk(N.dev().csp_timing_read(ms, cnt))
    return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(KERNEL_CLASSES)}
