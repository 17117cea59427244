"""Pins the CPU oracle against the reference's own hand-written goldens for tree geometry, point location and
iteration: /root/reference/test/runtests.jl "Geometry and iteration, CS1" (:49-216), "... CS2" (:218-449),
plus the randomised generate_randboxes/record_geometry! cross-check (test/functions.jl:28-82) with a seeded RNG."""
import math

import numpy as np
import pytest

from oracle import oracle as O

Inf = math.inf


def approx_t(t1, t2):
    return t1[0] == pytest.approx(t2[0]) and t1[1] == pytest.approx(t2[1])


def test_cs1_1d_goldens():  # runtests.jl:55-100
    world = O.World([0], [1], [0], 10)
    root = O.Box(world, 1)
    for bad in (-0.5, 1.5, 0):  # out-of-bounds x2, coincides with parent
        with pytest.raises(O.JlError, match="ErrorException"):
            O.split(root, 1, bad, 20)
    box = O.split(root, 1, 1, 20)[0]
    assert O.meta(box) == 20
    assert list(O.position(box)) == [1]
    assert O.position(box, 1) == 1
    with pytest.raises(O.JlError, match="BoundsError"):
        O.position(box, 2)
    assert O.boxbounds(box) == [(0.5, 1)]
    assert O.boxbounds(box, 1) == (0.5, 1)
    with pytest.raises(O.JlError, match="BoundsError"):
        O.boxbounds(box, 2)
    with pytest.raises(O.JlError, match="AssertionError"):
        O.split(root, 1, 0.5, 30)  # can only split leaves
    with pytest.raises(O.JlError, match="ErrorException"):
        O.split(O.getleaf(root), 1, 0.5, 30)  # can't eval at upper edge
    assert O.collect(root) == [root, root.split_self, box]
    assert O.collect(box) == [box]
    assert O.leaves(root) == [root.split_self, box]
    box1 = O.split(box, 1, 0.5, 30)[0]
    assert O.boxbounds(box1, 1) == (0.5, 0.75)
    assert O.boxbounds(O.getleaf(box), 1) == (0.75, 1.0)
    assert O.boxbounds(O.getleaf(root), 1) == (0.0, 0.5)
    assert O.boxbounds(box1) == [(0.5, 0.75)]
    assert O.boxbounds(O.getleaf(box)) == [(0.75, 1.0)]
    assert O.boxbounds(O.getleaf(root)) == [(0.0, 0.5)]
    leaf = O.find_leaf_at(root, [0.3])
    assert O.isleaf(leaf) and leaf == O.getleaf(root)
    leaf = O.find_leaf_at(root, [0.5])
    assert O.isleaf(leaf) and leaf == box1
    leaf = O.find_leaf_at(root, [0.8])
    assert O.isleaf(leaf) and leaf == O.getleaf(box)
    with pytest.raises(O.JlError, match="ErrorException"):
        O.find_leaf_at(root, [1.2])
    with pytest.raises(O.JlError, match="DimensionMismatch"):
        O.find_leaf_at(root, [0.1, 0.2])
    assert len(O.collect(root)) == 5
    assert len(O.leaves(root)) == 3


def test_cs1_infinite_and_flipped():  # runtests.jl:102-114
    world = O.World([0], [Inf], [0], 10)
    root = O.Box(world, 1)
    box = O.split(root, 1, 1, 20)[0]
    assert O.boxbounds(box) == [(0.5, Inf)]
    assert O.boxbounds(box, 1) == (0.5, Inf)
    world = O.World([0], [1], [1], 10)
    root = O.Box(world, 1)
    O.split(root, 1, 0, 20)
    assert O.find_leaf_at(root, 0.5) == O.getleaf(root)  # scalar x accepted for 1-d


def test_cs1_2d_goldens():  # runtests.jl:124-196
    world = O.World([0, -Inf], [Inf, Inf], [1, 1], 0.0)
    root = O.Box(world, 1)
    box1 = O.split(root, 1, 2, 0.0)[0]
    box2 = O.split(O.getleaf(root), 2, 2, 0.0)[0]
    b = O.getleaf(root)
    assert O.position(b, 1) == O.position(b, 2) == 1
    assert list(O.position(b)) == [1, 1]
    assert O.boxbounds(b, 1) == (0, 1.5)
    assert O.boxbounds(b, 2) == (-Inf, 1.5)
    assert O.boxbounds(b) == [(0, 1.5), (-Inf, 1.5)]
    b = box2
    assert (O.position(b, 1), O.position(b, 2)) == (1, 2)
    assert O.boxbounds(b) == [(0, 1.5), (1.5, Inf)]
    b = box1
    assert list(O.position(b)) == [2, 1]
    assert O.boxbounds(b, 1) == (1.5, Inf)
    assert O.boxbounds(b, 2) == (-Inf, Inf)  # never split along 2
    assert O.boxbounds(b) == [(1.5, Inf), (-Inf, Inf)]

    b = O.getleaf(root)
    assert b.parent.split_dims == (2,)
    b2 = O.split(b, 1, 1 / 3, 0.0)[0]
    assert O.position(b2, 1) == pytest.approx(1 / 3) and O.position(b2, 2) == 1
    assert approx_t(O.boxbounds(b2, 1), (0, 2 / 3))
    assert O.boxbounds(b2, 2) == (-Inf, 1.5)
    b1 = b.split_self
    assert list(O.position(b1)) == [1, 1]
    assert approx_t(O.boxbounds(b1, 1), (2 / 3, 1.5))
    assert O.boxbounds(b1, 2) == (-Inf, 1.5)
    b2_2 = O.split(b2, 2, -1, 0.0)[0]
    assert np.allclose(O.position(b2_2), [1 / 3, -1])
    assert approx_t(O.boxbounds(b2_2, 1), (0, 2 / 3))
    assert O.boxbounds(b2_2, 2) == (-Inf, 0)
    b2_1 = b2.split_self
    assert list(O.position(b2_1)) == list(O.position(b2))
    assert approx_t(O.boxbounds(b2_1, 1), (0, 2 / 3))
    assert O.boxbounds(b2_1, 2) == (0, 1.5)
    b1_2 = O.split(b1, 2, -1, 0.0)[0]
    assert list(O.position(b1_2)) == [1, -1]
    assert approx_t(O.boxbounds(b1_2, 1), (2 / 3, 1.5))
    assert O.boxbounds(b1_2, 2) == (-Inf, 0)
    b1_1 = b1.split_self
    assert list(O.position(b1_1)) == list(O.position(b1))
    assert approx_t(O.boxbounds(b1_1, 1), (2 / 3, 1.5))
    assert O.boxbounds(b1_1, 2) == (0, 1.5)
    bbs = O.boxbounds(b1_1)
    assert approx_t(bbs[0], (2 / 3, 1.5)) and bbs[1] == (0, 1.5)


def test_cs2_fictive_1d():  # runtests.jl:222-239  (addpoint!(root, x, f) with n<=2 picks [(1,)] -- cs2.jl:12-20)
    world = O.World([-Inf], [Inf], [1], 0)
    root = O.Box(world, 2)
    f0 = lambda y: 0.0
    O.addpoint(root, [1.2], [(1,)], f0)
    ch = [root.child(i) for i in range(4)]
    assert [O.isfake(c) for c in ch] == [False, False, True, True]
    assert len(O.leaves(root)) == 2
    box1 = O.addpoint(root, [0.75], [(1,)], f0)
    assert len(O.leaves(root)) == 3
    box2 = O.addpoint(root, [0.5], [(1,)], f0)
    assert len(O.leaves(root)) == 4
    assert O.boxbounds(box2, 1) == (-Inf, 5 / 8)
    assert O.boxbounds(box1, 1) == (-Inf, 7 / 8)
    assert O.boxbounds(O.getleaf(box1), 1) == (5 / 8, 7 / 8)


def test_find_leaf_midpoint_ties():  # runtests.jl:242-251
    world = O.World([0], [1], [0], 10)
    root = O.Box(world, 1)
    box1 = O.split(root, 1, 1, 20)[0]
    assert O.find_leaf_at(root, 0.5) == box1
    world = O.World([0], [1], [1], 10)
    root = O.Box(world, 1)
    O.split(root, 1, 0, 20)
    assert O.find_leaf_at(root, 0.5) == O.getleaf(root)


def test_cs2_2d_goldens():  # runtests.jl:253-279
    world = O.World([0, -Inf], [Inf, Inf], [1, 1], 0.0)
    root = O.Box(world, 2)
    boxes1 = O.split(root, (1, 2), (2, 2), (0.0, 0.0, 0.0))
    b = O.getleaf(root)
    assert list(O.position(b)) == [1, 1]
    assert O.boxbounds(b) == [(0, 1.5), (-Inf, 1.5)]
    b = boxes1[1]
    assert list(O.position(b)) == [1, 2]
    assert O.boxbounds(b) == [(0, 1.5), (1.5, Inf)]
    b = boxes1[0]
    assert list(O.position(b)) == [2, 1]
    assert O.boxbounds(b, 1) == (1.5, Inf)
    assert O.boxbounds(b, 2) == (-Inf, 1.5)
    assert O.boxbounds(b) == [(1.5, Inf), (-Inf, 1.5)]


def test_cs2_3d_fake_iteration():  # runtests.jl:281-292
    world = O.World([-Inf] * 3, [Inf] * 3, [0] * 3, 1)
    root = O.Box(world, 2)
    assert len(O.collect(root)) == 1 and len(O.leaves(root)) == 1
    boxes1 = O.split(root, (1, 2), (1, 1), (2, 3, 4))
    assert len(O.collect(root)) == 5 and len(O.leaves(root)) == 4
    O.split(boxes1[-1], (3, 4), (1, 1), (5, 6, 7))
    assert len(O.leaves(root)) == 5
    assert [O.value(l) for l in O.leaves(root)] == [1, 2, 3, 4, 5]


def test_cs2_4d_goldens():  # runtests.jl:294-439
    world = O.World([0, -Inf, -5, -Inf], [Inf, Inf, 50, 20], [1] * 4, 1)
    root = O.Box(world, 2)
    boxes1 = O.split(root, (1, 2), (2, 2), (2, 3, 4))
    assert len(O.collect(root)) == 5 and len(O.leaves(root)) == 4

    def check(b, pos, bbs, m):
        assert list(O.position(b)) == pos
        for d in range(4):
            assert O.position(b, d + 1) == pos[d]
            assert O.boxbounds(b, d + 1) == bbs[d]
        assert O.boxbounds(b) == bbs
        assert O.meta(b) == m

    check(O.getleaf(root), [1, 1, 1, 1], [(0, 1.5), (-Inf, 1.5), (-5, 50), (-Inf, 20)], 1)
    check(boxes1[0], [2, 1, 1, 1], [(1.5, Inf), (-Inf, 1.5), (-5, 50), (-Inf, 20)], 2)
    check(boxes1[1], [1, 2, 1, 1], [(0, 1.5), (1.5, Inf), (-5, 50), (-Inf, 20)], 3)
    check(boxes1[2], [2, 2, 1, 1], [(1.5, Inf), (1.5, Inf), (-5, 50), (-Inf, 20)], 4)

    boxes2 = O.split(boxes1[1], (3, 4), (2, 3), (5, 6, 7))
    assert len(O.collect(root)) == 9 and len(O.leaves(root)) == 7
    check(O.getleaf(boxes1[1]), [1, 2, 1, 1], [(0, 1.5), (1.5, Inf), (-5, 1.5), (-Inf, 2)], 3)
    check(boxes2[0], [1, 2, 2, 1], [(0, 1.5), (1.5, Inf), (1.5, 50), (-Inf, 2)], 5)
    check(boxes2[1], [1, 2, 1, 3], [(0, 1.5), (1.5, Inf), (-5, 1.5), (2, 20)], 6)
    check(boxes2[2], [1, 2, 2, 3], [(0, 1.5), (1.5, Inf), (1.5, 50), (2, 20)], 7)

    boxes3 = O.split(boxes2[2], (3, 1), (1.8, 1.3), (8, 9, 10))
    assert len(O.collect(root)) == 13 and len(O.leaves(root)) == 10
    check(O.getleaf(boxes2[2]), [1, 2, 2, 3], [(0, 1.15), (1.5, Inf), (1.9, 50), (2, 20)], 7)
    check(boxes3[0], [1, 2, 1.8, 3], [(0, 1.15), (1.5, Inf), (1.5, 1.9), (2, 20)], 8)
    check(boxes3[1], [1.3, 2, 2, 3], [(1.15, 1.5), (1.5, Inf), (1.9, 50), (2, 20)], 9)
    check(boxes3[2], [1.3, 2, 1.8, 3], [(1.15, 1.5), (1.5, Inf), (1.5, 1.9), (2, 20)], 10)


def generate_randboxes(p, n, nboxes, rng, callback=None):
    """test/functions.jl:28-52 with a seeded RNG."""
    nc = (1 << p) - 1
    world = O.World([0.0] * n, [1.0] * n, [0.5] * n, rng.random())
    root = O.Box(world, p)
    lvs = O.leaves(root)
    while len(lvs) < nboxes:
        box = lvs[rng.integers(len(lvs))]
        sd = rng.integers(1, n + 1, size=p)
        while len(set(sd.tolist())) < p:
            sd = rng.integers(1, n + 1, size=p)
        xs = []
        for d in sd:
            bb = O.boxbounds(box, int(d))
            xs.append(bb[0] + (bb[1] - bb[0]) * rng.random())
        newboxes = O.split(box, tuple(int(d) for d in sd), xs, rng.random(nc))
        if callback:
            callback(box, [int(d) for d in sd], xs, newboxes)
        lvs = O.leaves(root)
    return root


def record_geometry(data, top, splitdims, xs, newboxes):
    """test/functions.jl:55-82: an independent top-down computation of (position, boxbounds)."""
    b = O.position(top)
    p = b.copy()
    for sd, x in zip(splitdims, xs):
        p[sd - 1] = x
    bbs = O.boxbounds(top)
    bbb, bbp = list(bbs), list(bbs)
    for sd in splitdims:
        mid = (b[sd - 1] + p[sd - 1]) / 2
        lower, upper = (bbb[sd - 1][0], mid), (mid, bbp[sd - 1][1])
        bbb[sd - 1] = lower if b[sd - 1] < p[sd - 1] else upper
        bbp[sd - 1] = upper if b[sd - 1] < p[sd - 1] else lower
    data[O.getleaf(top)] = (b.copy(), list(bbb))
    for i in range(1, len(newboxes) + 1):
        xbx = b.copy()
        bbx = list(bbb)
        for j, sd in enumerate(splitdims):
            if (i >> j) & 1:
                xbx[sd - 1] = p[sd - 1]
                bbx[sd - 1] = bbp[sd - 1]
        data[newboxes[i - 1]] = (xbx, bbx)


@pytest.mark.parametrize("p,n,nboxes", [(1, 1, 10), (1, 2, 10), (1, 5, 20), (2, 5, 100)])
def test_randboxes_geometry(p, n, nboxes):  # runtests.jl:116-122, 200-215, 441-448
    rng = np.random.default_rng(1000 * p + n)
    geom = {}
    root = generate_randboxes(p, n, nboxes, rng, lambda *a: record_geometry(geom, *a))
    for leaf in O.leaves(root):
        pos, bbs = geom[leaf]
        assert list(O.position(leaf)) == list(pos)
        assert O.boxbounds(leaf) == bbs
        # and the property find_leaf_at relies on: a point strictly inside the box is located in it
        x = [lo + (hi - lo) * 0.37 for lo, hi in bbs]
        assert O.find_leaf_at(root, x) == leaf
        # lower edges are closed (tree.jl:370-383)
        assert O.find_leaf_at(root, [lo for lo, hi in bbs]) == leaf
