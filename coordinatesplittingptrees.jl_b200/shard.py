"""Multi-GPU plumbing (one process per GPU, torch.distributed): the path shards with no exchange step.

  C1  the flattened tree blob is broadcast ONCE from the rank that grew the tree (NCCL over NVLink on GPUs; the same code
      runs over gloo with CPU tensors in the CPU tests);
  --  queries are split into contiguous ranges, leaves into contiguous leaf-id ranges: no communication;
  C2  results (leaf id int32 + value f64 per query) are gathered to rank 0 on request.

Nothing here computes: kernels are reached through device.py.
"""
import numpy as np


def shard_range(total, world, rank):
    """Contiguous [lo, hi) of `total` items owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_blob(blob, src=0, device=None):
    """Broadcast a packed tree blob (numpy uint8 on `src`, None elsewhere).  Returns a torch uint8 tensor on `device`
    (CUDA for NCCL, CPU for gloo) holding the blob on every rank."""
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
