"""CPU test of the derivation K2 rests on: with internal nodes numbered in preorder (I) and subtree internal
counts (S), the reference's splits(box) iterator (src/tree.jl:461-540, restated literally in the oracle) yields, for each
ancestor a of the box bottom-up, the id ranges [I_a+1, I_c) and [I_c+S_c, I_a+S_a) ascending and then I_a -- for leaves
and for internal boxes, with and without fictive dimensions."""
import numpy as np
import pytest

from oracle import oracle as O
from helpers import grow_both


def ranges_order(flat, node):
    nn = flat["n_nodes"]
    internal = flat["internal_id"] >= 0
    istart = np.concatenate([[0], np.cumsum(internal)[:-1]])
    icount = internal.astype(np.int64).copy()
    for v in range(nn - 1, 0, -1):
        icount[flat["parent"][v]] += icount[v]
    out = []
    c = node
    if internal[c]:
        out += list(range(istart[c], istart[c] + icount[c]))
    while flat["parent"][c] != c:
        a = flat["parent"][c]
        out += list(range(istart[a] + 1, istart[c]))
        out += list(range(istart[c] + icount[c], istart[a] + icount[a]))
        out.append(istart[a])
        c = a
    return out


@pytest.mark.parametrize("n,p,npoints", [(2, 1, 40), (5, 1, 30), (3, 2, 25), (6, 2, 30), (7, 2, 20)])
def test_splits_iterator_equals_preorder_ranges(n, p, npoints):
    root, oroot, _ = grow_both(n, p, npoints, seed=3 * n + p)
    flat = root.tree.flat()
    boxes = O.collect(oroot)
    assert len(boxes) == flat["n_nodes"]
    for node, box in enumerate(boxes):
        if O.isfake(box):
            continue
        owners = O.splits(box)
        lit = [int(flat["internal_id"][O.preorder_id(b)]) for b in owners]
        assert lit == ranges_order(flat, node), node
