"""Multi-GPU plumbing.  The path shards with no exchange step; the collectives are the library's own, behind the
C ABI (csrc/comm.cu, NCCL over NVLink):

  C1  csp_shard_tree_broadcast   the flattened tree blob, ONCE per tree version, uploaded on every GPU from the device buffer;
  --  csp_shard_fit              leaf fits sharded by contiguous leaf-id range, fitted records all-gathered;
  --  queries split into contiguous ranges: no communication;
  C2  csp_shard_locate_eval_dev  (leaf id int32, value f64) gathered to one rank, pipelined behind the next chunk's kernels.

Two ways to form the clique: `init_single_process(ndev)` (one process drives every GPU -- what a Julia host does) and
`init_from_torch_distributed()` (one process per GPU under torchrun: torch.distributed only carries the 128-byte NCCL id and
the barriers; the data path never goes through it).  Nothing in this file computes.

The torch.distributed helpers at the end (`broadcast_blob`, `gather_results`) are the host-side equivalents used by the CPU
(gloo) tests of the partition logic; the GPU path does not call them.
"""
import ctypes as C

import numpy as np

from . import _native as N
from .device import DeviceTree, LeafModels, _is_cuda_tensor, _ptr, _stream_ptr


def shard_range(total, world, rank):
    """Contiguous [lo, hi) of `total` items owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_counts(total, world):
    return [shard_range(total, world, r)[1] - shard_range(total, world, r)[0] for r in range(world)]


# ---- clique ---------------------------------------------------------------------------------------------------------------
def init_single_process(ndev, devices=None):
    """csp_init: GPUs devices[i] (default 0..ndev-1) become ranks 0..ndev-1, all driven by this process."""
    devs = None if devices is None else (C.c_int32 * ndev)(*devices)
    N.dchk(N.dev().csp_init(int(ndev), devs))
    return comm_info()


def init_from_torch_distributed(device):
    """csp_init_rank with the world / rank of the initialised torch.distributed group (NCCL or gloo); the 128-byte NCCL unique id
    travels from rank 0 by a torch.distributed broadcast."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    ident = (C.c_char * 128)()
    if world > 1:
        if rank == 0:
            N.dchk(N.dev().csp_comm_unique_id(ident))
        backend = dist.get_backend()
        t = torch.frombuffer(bytearray(ident.raw), dtype=torch.uint8).clone()
        if backend == "nccl":
            t = t.cuda(device)
        dist.broadcast(t, 0)
        ident = (C.c_char * 128).from_buffer_copy(bytes(t.cpu().numpy().tobytes()))
    N.dchk(N.dev().csp_init_rank(world, rank, int(device), ident))
    return comm_info()


def shutdown():
    N.dchk(N.dev().csp_shutdown())


def comm_info():
    info = (C.c_int64 * 4)()
    N.dchk(N.dev().csp_comm_info(info))
    return dict(world=int(info[0]), local=int(info[1]), first_rank=int(info[2]), nccl_version=int(info[3]))


def synchronize():
    N.dchk(N.dev().csp_shard_synchronize())


TIMING_KEYS = ("blob_h2d_ms", "tree_broadcast_ms", "upload_ms", "fit_ms", "allgather_ms", "gather_ms")


def timing():
    ms = (C.c_double * 6)()
    N.dchk(N.dev().csp_shard_timing(ms))
    return {k: float(ms[i]) for i, k in enumerate(TIMING_KEYS)}


class _RawCuda:
    """A device buffer owned by the library, exposed through __cuda_array_interface__ so that torch.as_tensor can view it."""

    def __init__(self, ptr, count, typestr):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def gather_alloc(capacity, root=0, device=None):
    """csp_shard_gather_alloc: library-owned gather buffers on rank `root` (collective).  Returns (leaf_all, val_all) as torch views
    on the process that owns the root rank, (None, None) elsewhere.  Gathers to that root then move their blocks with the copy
    engines (peer DMA over NVLink) at the element offset `gather_base` the calls are given."""
    import torch
    pl, pv = C.c_void_p(), C.c_void_p()
    N.dchk(N.dev().csp_shard_gather_alloc(int(capacity), int(root), C.byref(pl), C.byref(pv)))
    if not pl.value:
        return None, None
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    with torch.cuda.device(dev):
        leaf = torch.as_tensor(_RawCuda(pl.value, capacity, "<i4"), device=dev)
        val = torch.as_tensor(_RawCuda(pv.value, capacity, "<f8"), device=dev)
    return leaf, val


def defer_join(on=True):
    """csp_shard_defer_join: resident locate+evaluate calls stop waiting for their own last gather step; join() does it once."""
    N.dchk(N.dev().csp_shard_defer_join(1 if on else 0))


def join(streams=None):
    """csp_shard_join: the given streams (one per local rank; None: the library's) wait for every gather issued so far."""
    nlocal = comm_info()["local"]
    N.dchk(N.dev().csp_shard_join(_stream_array(streams, nlocal)))


def gather_free():
    N.dchk(N.dev().csp_shard_gather_free())


def _ptr_array(items):
    if items is None:
        return None
    arr = (C.c_void_p * len(items))()
    for i, it in enumerate(items):
        arr[i] = None if it is None else (_ptr(it) if hasattr(it, "shape") else it)
    return arr


def _stream_array(streams, nlocal):
    if streams is None:
        return None
    arr = (C.c_void_p * nlocal)()
    for i, s in enumerate(streams):
        arr[i] = _stream_ptr(s)
    return arr


class _Borrowed:
    """A DeviceTree / LeafModels view over a handle owned by a sharded object."""

    @staticmethod
    def tree(h):
        t = DeviceTree.__new__(DeviceTree)
        t.h = h
        info = (C.c_int64 * 8)()
        N.dchk(N.dev().csp_tree_info(h, info))
        t.n, t.p, t.n_nodes, t.n_internal, t.n_leaves, t.device, t.max_depth, t.kib = [int(v) for v in info]
        t.close = lambda: None
        return t

    @staticmethod
    def models(h):
        m = LeafModels(h)
        m.close = lambda: None
        return m


class ShardedTree:
    """csp_shard_tree: the tree replicated on every GPU of the clique by one NCCL broadcast (collective constructor)."""

    def __init__(self, blob=None, root=0):
        info = comm_info()
        self.nlocal, self.world, self.first_rank = info["local"], info["world"], info["first_rank"]
        h = C.c_void_p()
        if blob is not None:
            blob = np.ascontiguousarray(blob, np.uint8)
            N.dchk(N.dev().csp_shard_tree_broadcast(blob.ctypes.data, blob.nbytes, int(root), C.byref(h)))
        else:
            N.dchk(N.dev().csp_shard_tree_broadcast(None, 0, int(root), C.byref(h)))
        self.h = h
        self.local = []
        for i in range(self.nlocal):
            th = C.c_void_p()
            N.dchk(N.dev().csp_shard_tree_get(h, i, C.byref(th)))
            self.local.append(_Borrowed.tree(th))
        t0 = self.local[0]
        self.n, self.p, self.n_leaves = t0.n, t0.p, t0.n_leaves

    def fit(self, models=None, skip=0, streams=None):
        """Leaf fits sharded by leaf-id range over the ranks + all-gather of the records (collective).  models=None allocates."""
        h = models.h if models is not None else C.c_void_p()
        N.dchk(N.dev().csp_shard_fit(self.h, skip, C.byref(h), _stream_array(streams, self.nlocal)))
        return models if models is not None else ShardedModels(h, self.nlocal)

    def locate_eval_resident(self, models, X, counts, leaf, val, status=None, gather_root=-1, leaf_all=None, val_all=None, chunk=0,
                             streams=None, gather_base=0):
        """Per local rank i: locate + evaluate the resident shard X[i] (counts[rank] queries) into leaf[i] / val[i] / status[i];
        with gather_root >= 0 also gather every rank's (leaf, val) into leaf_all / val_all on that rank (collective) -- or, when
        gather_alloc() buffers exist, into those at element offset gather_base (copy-engine path)."""
        cnt = (C.c_int64 * self.world)(*[int(c) for c in counts])
        N.dchk(N.dev().csp_shard_locate_eval_dev(self.h, models.h, _ptr_array(X), cnt, _ptr_array(leaf), _ptr_array(val),
                                                 _ptr_array(status) if status is not None else None, int(gather_root),
                                                 None if leaf_all is None else _ptr(leaf_all), None if val_all is None else _ptr(val_all),
                                                 int(gather_base), int(chunk), _stream_array(streams, self.nlocal)))

    def gather_resident(self, leaf, val, counts, gather_root, leaf_all, val_all, streams=None, gather_base=0):
        cnt = (C.c_int64 * self.world)(*[int(c) for c in counts])
        N.dchk(N.dev().csp_shard_gather_dev(_ptr_array(leaf), _ptr_array(val), cnt, int(gather_root),
                                            None if leaf_all is None else _ptr(leaf_all), None if val_all is None else _ptr(val_all),
                                            int(gather_base), _stream_array(streams, self.nlocal)))

    def locate_eval(self, models, X, leaf=None, val=None, status=None):
        """Host buffers: this process's queries split over its local GPUs; results in host arrays (leaf, val, status)."""
        X = np.ascontiguousarray(X, np.float64)
        if X.ndim != 2 or X.shape[1] != self.n:
            raise N.DimensionMismatch(f"tree has {self.n} dimensions, got a query array of shape {X.shape}")
        m = X.shape[0]
        leaf = np.empty(m, np.int32) if leaf is None else leaf
        val = np.empty(m, np.float64) if val is None else val
        status = np.empty(m, np.uint8) if status is None else status
        N.dchk(N.dev().csp_shard_locate_eval(self.h, models.h, X.ctypes.data, m, leaf.ctypes.data, val.ctypes.data, status.ctypes.data))
        return leaf, val, status

    def close(self):
        if self.h:
            N.dev().csp_shard_tree_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ShardedModels:
    """csp_shard_models: every GPU holds every leaf model; each rank fitted its own leaf-id range."""

    def __init__(self, h, nlocal):
        self.h = h
        self.local = []
        for i in range(nlocal):
            mh = C.c_void_p()
            N.dchk(N.dev().csp_shard_models_get(h, i, C.byref(mh)))
            self.local.append(_Borrowed.models(mh))

    def close(self):
        if self.h:
            N.dev().csp_shard_models_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- torch.distributed equivalents (CPU / gloo tests of the partition logic) ---------------------------------------------------
def broadcast_blob(blob, src=0, device=None):
    """Broadcast a packed tree blob (numpy uint8 on `src`, None elsewhere) with torch.distributed.  Returns a torch uint8 tensor on
    `device` (CPU for gloo) holding the blob on every rank."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank()
    dev = torch.device(device) if device is not None else torch.device("cpu")
    size = torch.tensor([blob.size if rank == src else 0], dtype=torch.int64, device=dev)
    dist.broadcast(size, src)
    if rank == src:
        buf = torch.from_numpy(np.ascontiguousarray(blob)).to(dev)
    else:
        buf = torch.empty(int(size.item()), dtype=torch.uint8, device=dev)
    dist.broadcast(buf, src)
    return buf


def gather_results(leaf, val, dst=0):
    """Gather per-rank (leaf, val) tensors of possibly different lengths to `dst`; returns (leaf, val) there, else None."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [torch.zeros(1, dtype=torch.int64, device=leaf.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([leaf.numel()], dtype=torch.int64, device=leaf.device))
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes)
    pl = torch.zeros(mx, dtype=leaf.dtype, device=leaf.device)
    pv = torch.zeros(mx, dtype=val.dtype, device=val.device)
    pl[: leaf.numel()] = leaf
    pv[: val.numel()] = val
    gl = [torch.empty_like(pl) for _ in range(world)] if rank == dst else None
    gv = [torch.empty_like(pv) for _ in range(world)] if rank == dst else None
    dist.gather(pl, gl, dst)
    dist.gather(pv, gv, dst)
    if rank != dst:
        return None
    return torch.cat([g[:s] for g, s in zip(gl, sizes)]), torch.cat([g[:s] for g, s in zip(gv, sizes)])
