"""Worker of tests/test_gpu_shard.py::test_multiprocess_clique (run under torchrun, one rank per GPU): the sharded path through
csp_init_rank / csp_shard_* equals the single-GPU answer computed by every rank for itself."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import csp_b200 as cb
    from csp_b200 import shard
    from csp_b200.device import DeviceTree
    from helpers import grow_both

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    info = shard.init_from_torch_distributed(local)
    assert info["world"] == world and info["local"] == 1 and info["first_rank"] == rank
    root, _, _ = grow_both(8, 2, 400, seed=77)  # every rank grows the same tree: its own single-GPU reference
    same = lambda a, b: np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)
    dt = DeviceTree(root.tree.flat(), device=local)
    ref_models = dt.fit_models()
    ref = ref_models.download()
    blob = cb.pack_blob(root.tree.flat()) if rank == 0 else None
    st = shard.ShardedTree(blob, root=0)
    sm = st.fit()
    got = sm.local[0].download()
    for k in ("status", "c", "g", "Q", "b"):
        assert same(got[k], ref[k]), k
    m = 200_003
    X = cb.synthetic.uniform_queries(root, m, seed=3)
    counts = shard.shard_counts(m, world)
    lo, hi = shard.shard_range(m, world, rank)
    dev = torch.device("cuda", local)
    Xd = torch.from_numpy(X[lo:hi]).to(dev)
    leaf = torch.empty(hi - lo, dtype=torch.int32, device=dev)
    val = torch.empty(hi - lo, dtype=torch.float64, device=dev)
    status = torch.empty(hi - lo, dtype=torch.uint8, device=dev)
    leaf_all = torch.empty(m, dtype=torch.int32, device=dev) if rank == 0 else None
    val_all = torch.empty(m, dtype=torch.float64, device=dev) if rank == 0 else None
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream()
    for _ in range(2):  # twice: the second pass reuses every buffer and communicator state
        sm = st.fit(models=sm, streams=[stream])
        st.locate_eval_resident(sm, [Xd], counts, [leaf], [val], [status], gather_root=0, leaf_all=leaf_all, val_all=val_all, chunk=30_000,
                                streams=[stream])
    torch.cuda.synchronize()
    rleaf, rval, rstatus = dt.locate_eval(ref_models, X)
    assert same(leaf.cpu().numpy(), rleaf[lo:hi]) and same(val.cpu().numpy(), rval[lo:hi])
    if rank == 0:
        assert same(leaf_all.cpu().numpy(), rleaf) and same(val_all.cpu().numpy(), rval)
    # the same gather through library-owned buffers and the copy engines (CUDA IPC mapping of rank 0's arrays)
    gl, gv = shard.gather_alloc(m + 500, root=0, device=local)
    if rank == 0:
        gl.fill_(-7)
    torch.cuda.synchronize()
    dist.barrier()
    for _ in range(2):
        st.locate_eval_resident(sm, [Xd], counts, [leaf], [val], [status], gather_root=0, chunk=40_000, streams=[stream], gather_base=500)
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        assert same(gl[500:].cpu().numpy(), rleaf) and same(gv[500:].cpu().numpy(), rval) and int(gl[:500].max()) == -7
    dist.barrier()
    shard.gather_free()
    dist.barrier()
    sm.close()
    st.close()
    shard.shutdown()
    dist.destroy_process_group()
    if rank == 0:
        print("mp_shard_worker ok")


if __name__ == "__main__":
    main()
