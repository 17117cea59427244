"""World-size-2 test of the multi-GPU HOST logic on CPU (gloo): the partition arithmetic (shard_range / shard_counts), the wire
format surviving a broadcast (rank 1 only ever sees the broadcast blob and decodes the same header and bytes), and the ordered
gather of unequal per-rank results.  Nothing is computed here: the per-rank "results" are a deterministic function of the global
query index, so what is checked is placement and order.  The kernels and the library's own NCCL collectives (csrc/comm.cu) need
GPUs and are covered by tests/test_gpu_shard.py."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["CSP_ROOT"]); sys.path.insert(0, os.path.join(os.environ["CSP_ROOT"], "tests"))
import csp_b200 as cb
from csp_b200 import shard
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
blob = None
if rank == 0:
    root, _ = cb.synthetic.config_tree("C1")
    blob = cb.pack_blob(root.tree.flat())
buf = shard.broadcast_blob(blob, src=0)
b = buf.numpy()
assert bytes(b[:7]) == b"CSPB200"
# every rank decodes the header + sections identically
hdr = np.frombuffer(b[:32].tobytes(), dtype=np.int32)
n, p, nn, ni, nl = [int(v) for v in hdr[3:8]]
assert (n, p) == (2, 1) and nn == 1 + 2 * ni and nl == ni + 1
# shard 1001 queries: contiguous, disjoint, covering
m = 1001
lo, hi = shard.shard_range(m, world, rank)
cover = torch.zeros(m, dtype=torch.int32); cover[lo:hi] = 1
dist.all_reduce(cover)
assert bool((cover == 1).all())
# per-rank results (here: a deterministic function of the global query index) gathered to rank 0 in order
leaf = torch.arange(lo, hi, dtype=torch.int32)
val = leaf.double() * 0.5
res = shard.gather_results(leaf, val, dst=0)
if rank == 0:
    L, V = res
    assert torch.equal(L, torch.arange(m, dtype=torch.int32)) and torch.equal(V, torch.arange(m).double() * 0.5)
    ck = int(np.frombuffer(b.tobytes(), dtype=np.uint8).astype(np.uint64).sum())
else:
    ck = int(np.frombuffer(b.tobytes(), dtype=np.uint8).astype(np.uint64).sum())
cks = [None] * world
dist.all_gather_object(cks, ck)
assert len(set(cks)) == 1
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_two_rank_gloo_broadcast_shard_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, CSP_ROOT=ROOT, MASTER_ADDR="127.0.0.1", MASTER_PORT="29541", WORLD_SIZE="2")
    procs = []
    for r in range(2):
        e = dict(env, RANK=str(r), LOCAL_RANK=str(r))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} failed:\n{o}"
        assert f"rank {r} ok" in o


def test_shard_range_properties():
    from csp_b200 import shard
    for total in (0, 1, 7, 100, 1001):
        for world in (1, 2, 3, 8):
            r = [shard.shard_range(total, world, k) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1
            assert shard.shard_counts(total, world) == sizes
