"""Multi-GPU layer behind the C ABI (csp_init / csp_shard_*, csrc/comm.cu) against the single-GPU answer, bit for bit.
Runs on whatever the box has: with one GPU the clique has one rank (no NCCL traffic, same code path); with >= 2 GPUs the
broadcast, the record all-gather and the result gather go over NCCL, in single-process mode here and in one-process-per-GPU mode
(torchrun) in test_multiprocess_clique."""
import os
import subprocess
import sys

import numpy as np
import pytest

import csp_b200 as cb
from csp_b200 import shard
from helpers import grow_both

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def same(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


@pytest.mark.parametrize("n,npoints", [(6, 300), (10, 700), (45, 20)])
def test_leaf_range_fits_equal_full_fit(n, npoints, monkeypatch):
    """csp_models_refit_range over a partition of the leaves reproduces the all-leaf fit exactly (work units whose sibling leaves
    straddle a range boundary are processed by both owners, each writing only its own records)."""
    root, _, _ = grow_both(n, 2, npoints, seed=900 + n)
    dt = cb.device_tree(root)
    full = dt.fit_models().download()
    nl = dt.n_leaves
    for parts in (2, 3, 8):
        mo = dt.alloc_models()
        for r in range(parts):
            lo, hi = shard.shard_range(nl, parts, r)
            dt.refit_models_range(mo, lo, hi)
        got = mo.download()
        for k in ("status", "c", "g", "Q", "b"):
            assert same(got[k], full[k]), (k, parts)
    rows = np.array([0, nl - 1, nl // 2, 3, 3], np.int32)
    part = dt.fit_models().download_rows(rows)
    for k in ("status", "c", "g", "Q", "b"):
        assert same(part[k], full[k][rows]), k


def test_single_process_clique_matches_single_gpu(monkeypatch):
    import torch
    # per-query evaluation inside the locate kernel: the summation order does not depend on how the batch is split, so the
    # sharded / chunked results can be compared with the one-batch answer bit for bit
    monkeypatch.setenv("CSP_EVAL_MODE", "direct")
    ndev = min(cb.device_count(), 8)
    root, _, _ = grow_both(8, 2, 400, seed=77)
    dt = cb.device_tree(root)
    ref_models = dt.fit_models()
    ref = ref_models.download()
    blob = cb.pack_blob(root.tree.flat())
    info = shard.init_single_process(ndev)
    try:
        assert info["world"] == ndev and info["local"] == ndev
        st = shard.ShardedTree(blob, root=ndev - 1)  # the tree starts on the LAST rank's GPU
        sm = st.fit()
        for i in range(ndev):
            got = sm.local[i].download()
            for k in ("status", "c", "g", "Q", "b"):
                assert same(got[k], ref[k]), (k, i)
        sm = st.fit(models=sm)  # in-place re-fit
        shard.synchronize()
        assert same(sm.local[ndev - 1].download()["Q"], ref["Q"])
        # resident shards + gather to rank 0
        m = 300_001
        X = cb.synthetic.uniform_queries(root, m, seed=3)
        X[5, 0] = 7.0  # outside the world
        counts = shard.shard_counts(m, ndev)
        offs = np.concatenate([[0], np.cumsum(counts)])
        Xd, leaf, val, status = [], [], [], []
        for i in range(ndev):
            dev = torch.device("cuda", st.local[i].device)
            Xd.append(torch.from_numpy(X[offs[i]:offs[i + 1]]).to(dev))
            leaf.append(torch.empty(counts[i], dtype=torch.int32, device=dev))
            val.append(torch.empty(counts[i], dtype=torch.float64, device=dev))
            status.append(torch.empty(counts[i], dtype=torch.uint8, device=dev))
        dev0 = torch.device("cuda", st.local[0].device)
        leaf_all = torch.full((m,), -7, dtype=torch.int32, device=dev0)
        val_all = torch.zeros(m, dtype=torch.float64, device=dev0)
        for d in range(ndev):
            torch.cuda.synchronize(d)
        st.locate_eval_resident(sm, Xd, counts, leaf, val, status, gather_root=0, leaf_all=leaf_all, val_all=val_all, chunk=50_000)
        shard.synchronize()
        rleaf, rval, rstatus = dt.locate_eval(ref_models, X)
        assert same(leaf_all.cpu().numpy(), rleaf) and same(val_all.cpu().numpy(), rval)
        assert same(np.concatenate([s.cpu().numpy() for s in status]), rstatus) and rstatus[5] == 1
        # the gather alone, to another root
        root2 = ndev - 1
        devr = torch.device("cuda", st.local[root2].device)
        la, va = torch.empty(m, dtype=torch.int32, device=devr), torch.empty(m, dtype=torch.float64, device=devr)
        st.gather_resident(leaf, val, counts, root2, la, va)
        shard.synchronize()
        assert same(la.cpu().numpy(), rleaf) and same(va.cpu().numpy(), rval)
        t = shard.timing()
        assert t["fit_ms"] > 0
        # library-owned gather buffers: the same gather through the copy engines (peer DMA), at an offset into the buffers
        gl, gv = shard.gather_alloc(m + 1000, root=root2, device=st.local[root2].device)
        gl.fill_(-7)
        st.locate_eval_resident(sm, Xd, counts, leaf, val, status, gather_root=root2, chunk=70_000, gather_base=1000)
        shard.synchronize()
        assert same(gl[1000:].cpu().numpy(), rleaf) and same(gv[1000:].cpu().numpy(), rval) and int(gl[:1000].max()) == -7
        # deferred join: two calls over the two halves of every shard, one join at the end
        shard.defer_join(True)
        gl.fill_(-7)
        h1 = [c // 2 for c in counts]
        h2 = [c - a for c, a in zip(counts, h1)]
        st.locate_eval_resident(sm, Xd, h1, leaf, val, status, gather_root=root2, chunk=40_000, gather_base=0)
        st.locate_eval_resident(sm, [x[a:] for x, a in zip(Xd, h1)], h2, [l[a:] for l, a in zip(leaf, h1)], [v[a:] for v, a in zip(val, h1)],
                                [t_[a:] for t_, a in zip(status, h1)], gather_root=root2, chunk=40_000, gather_base=sum(h1))
        shard.join()
        shard.defer_join(False)
        shard.synchronize()
        o1, o2 = np.concatenate([[0], np.cumsum(h1)]), np.concatenate([[0], np.cumsum(h2)])
        got = gl.cpu().numpy()
        for r in range(ndev):  # the first call's blocks, then the second call's, each in rank order
            assert same(got[o1[r]:o1[r + 1]], rleaf[offs[r]:offs[r] + h1[r]])
            assert same(got[sum(h1) + o2[r]: sum(h1) + o2[r + 1]], rleaf[offs[r] + h1[r]:offs[r + 1]])
        gl.fill_(-7)
        st.gather_resident(leaf, val, counts, root2, None, None, gather_base=0)
        shard.synchronize()
        assert same(gl[:m].cpu().numpy(), rleaf) and same(gv[:m].cpu().numpy(), rval)
        shard.gather_free()
        # host buffers, split over the local GPUs by one thread per GPU
        hleaf, hval, hstatus = st.locate_eval(sm, X)
        assert same(hleaf, rleaf) and same(hval, rval) and same(hstatus, rstatus)
        sm.close()
        st.close()
    finally:
        shard.shutdown()


def test_multiprocess_clique():
    """One process per GPU (torchrun, csp_init_rank): needs >= 2 GPUs."""
    ndev = min(cb.device_count(), 4)
    if ndev < 2:
        pytest.skip("needs at least two GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", CSP_EVAL_MODE="direct")  # batch-split-independent summation order
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={ndev}", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", os.path.join(ROOT, "tests", "mp_shard_worker.py")], capture_output=True, text=True,
                       env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "mp_shard_worker ok" in r.stdout
