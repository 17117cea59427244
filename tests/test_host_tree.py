"""CPU tests of the host side: the product's flat host tree (libcsp_host.so + tree.py) against the CPU oracle's pointer
tree on identical growth sequences, the reference's error behaviour, and the C-ABI library's exported symbols."""
import ctypes as C
import math
import os
import re

import numpy as np
import pytest

import csp_b200 as cb
from csp_b200 import _native as N
from oracle import oracle as O
from helpers import assert_flat_equal, grow_both

Inf = math.inf
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n,p,npoints", [(2, 1, 200), (5, 1, 60), (2, 2, 100), (3, 2, 80), (7, 2, 120), (10, 2, 300)])
def test_flatten_matches_oracle(n, p, npoints):
    root, oroot, _ = grow_both(n, p, npoints, seed=n * 10 + p)
    assert_flat_equal(root.tree.flat(), O.export_flat(oroot))


def test_geometry_matches_oracle_boxbounds():
    root, oroot, _ = grow_both(5, 2, 60, seed=7)
    lv, olv = cb.leaves(root), O.leaves(oroot)
    assert len(lv) == len(olv)
    for a, b in zip(lv, olv):
        assert np.array_equal(cb.position(a), O.position(b))
        assert cb.boxbounds(a) == O.boxbounds(b)
        assert cb.value(a) == O.value(b)
    allb, oall = cb.collect(root), O.collect(oroot)
    assert len(allb) == len(oall) == cb.length(root)
    for a, b in zip(allb, oall):
        assert cb.isleaf(a) == O.isleaf(b) and cb.isfake(a) == O.isfake(b)
        assert cb.boxbounds(a) == O.boxbounds(b)


def test_reference_goldens_on_host_tree():  # runtests.jl:55-98, 294-439 on the product's own host tree
    world = cb.World([0], [1], [0], 10)
    root = cb.Box(world, 1)
    for bad in (-0.5, 1.5, 0):
        with pytest.raises(cb.ErrorException):
            cb.split(root, 1, bad, 20)
    box = cb.split(root, 1, 1, 20)[0]
    assert cb.meta(box) == 20 and list(cb.position(box)) == [1] and cb.boxbounds(box) == [(0.5, 1)]
    with pytest.raises(cb.BoundsError):
        cb.position(box, 2)
    with pytest.raises(AssertionError):
        cb.split(root, 1, 0.5, 30)
    with pytest.raises(cb.ErrorException):
        cb.split(cb.getleaf(root), 1, 0.5, 30)
    assert cb.collect(root) == [root, root.split_self, box]
    assert cb.leaves(root) == [root.split_self, box]
    box1 = cb.split(box, 1, 0.5, 30)[0]
    assert cb.boxbounds(box1, 1) == (0.5, 0.75)
    assert cb.boxbounds(cb.getleaf(box), 1) == (0.75, 1.0)
    assert cb.boxbounds(cb.getleaf(root), 1) == (0.0, 0.5)
    assert cb.length(root) == 5 and len(cb.leaves(root)) == 3

    world = cb.World([0, -Inf, -5, -Inf], [Inf, Inf, 50, 20], [1] * 4, 1)
    root = cb.Box(world, 2)
    boxes1 = cb.split(root, (1, 2), (2, 2), (2, 3, 4))
    boxes2 = cb.split(boxes1[1], (3, 4), (2, 3), (5, 6, 7))
    boxes3 = cb.split(boxes2[2], (3, 1), (1.8, 1.3), (8, 9, 10))
    assert cb.length(root) == 13 and len(cb.leaves(root)) == 10
    b = cb.getleaf(boxes2[2])
    assert list(cb.position(b)) == [1, 2, 2, 3] and cb.meta(b) == 7
    assert cb.boxbounds(b) == [(0, 1.15), (1.5, Inf), (1.9, 50), (2, 20)]
    b = boxes3[2]
    assert list(cb.position(b)) == [1.3, 2, 1.8, 3] and cb.meta(b) == 10
    assert cb.boxbounds(b) == [(1.15, 1.5), (1.5, Inf), (1.5, 1.9), (2, 20)]


def test_fictive_dimensions_and_addpoint_errors():  # runtests.jl:222-239, CoordinateSplittingPTrees.jl:96-104
    world = cb.World([-Inf], [Inf], [1], 0)
    root = cb.Box(world, 2)
    f0 = lambda y: 0.0
    cb.addpoint(root, [1.2], [(1,)], f0)
    assert [cb.isfake(root.child(k)) for k in range(4)] == [False, False, True, True]
    assert len(cb.leaves(root)) == 2
    box1 = cb.addpoint(root, [0.75], [(1,)], f0)
    box2 = cb.addpoint(root, [0.5], [(1,)], f0)
    assert len(cb.leaves(root)) == 4
    assert cb.boxbounds(box2, 1) == (-Inf, 5 / 8) and cb.boxbounds(box1, 1) == (-Inf, 7 / 8)
    assert cb.boxbounds(cb.getleaf(box1), 1) == (5 / 8, 7 / 8)
    root = cb.Box(cb.World([-Inf] * 3, [Inf] * 3, [0] * 3, 0), 2)
    with pytest.raises(cb.ErrorException):
        cb.addpoint(root, [1, 1, 1], [(1, 2), (2,)], f0)  # duplicated
    with pytest.raises(cb.ErrorException):
        cb.addpoint(root, [1, 1, 1], [(1, 2)], f0)  # not covered
    with pytest.raises(cb.ErrorException):
        cb.addpoint(root, [0, 1, 1], [(1, 2), (3,)], f0)  # identical to the parent along 1
    with pytest.raises(cb.ArgumentError):
        cb.World([0, 0], [1, 1], [7.0, 0.5], 0)


def test_round_robin_covers_all_pairs():
    for n in (2, 5, 6, 10, 21):
        rounds = cb.synthetic.round_robin_dimlists(n)
        seen = set()
        for r in rounds:
            assert sorted(d for dl in r for d in dl) == list(range(1, n + 1))
            seen |= {tuple(sorted(dl)) for dl in r if len(dl) == 2}
        assert len(seen) == n * (n - 1) // 2


def test_cabi_library_exports_every_declared_symbol():
    """include/csp_b200.h <-> libcsp_b200.so: every declared entry point is exported (no compute calls here)."""
    hdr = open(os.path.join(ROOT, "include", "csp_b200.h")).read()
    declared = set(re.findall(r"\b(csp_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(N.EXPORTS), declared ^ set(N.EXPORTS)
    L = N.dev()
    for name in declared:
        assert hasattr(L, name), name
    assert L.csp_version() == 200


def test_cabi_fails_loudly_without_device():
    """No CPU fallback: on a box without a GPU every compute entry point must fail with CSP_ERR_NO_DEVICE."""
    cnt = C.c_int()
    N.dev().csp_device_count(C.byref(cnt))
    if cnt.value > 0:
        pytest.skip("a CUDA device is present")
    root, _, _ = grow_both(2, 2, 10, seed=1)
    with pytest.raises(cb.NoDeviceError):
        cb.find_leaf_at(root, [0.1, 0.2])
    with pytest.raises(cb.NoDeviceError):
        cb.fit_quadratic(cb.leaves(root)[0])


def test_blob_roundtrip_host():
    """The wire format packs on the host without a GPU and validates its input."""
    root, _, _ = grow_both(3, 2, 20, seed=2)
    flat = root.tree.flat()
    blob = cb.pack_blob(flat)
    assert bytes(blob[:7]) == b"CSPB200"
    bad = dict(flat)
    bad["parent"] = flat["parent"].copy()
    bad["parent"][3] = 0 if flat["parent"][3] != 0 else 1
    with pytest.raises(cb.CspError):
        cb.pack_blob(bad)


def test_blob_file_roundtrip(tmp_path):
    """The blob is the persistent on-disk format: save, load, bit-identical; header fields decode."""
    root, _, _ = grow_both(4, 2, 30, seed=3)
    flat = root.tree.flat()
    path = tmp_path / "tree.cspb"
    cb.save_blob(flat, str(path))
    blob = cb.load_blob(str(path))
    assert np.array_equal(blob, cb.pack_blob(flat))
    hdr = np.frombuffer(blob[:32].tobytes(), dtype=np.int32)
    assert [int(v) for v in hdr[3:8]] == [flat["n"], flat["p"], flat["n_nodes"], flat["n_internal"], flat["n_leaves"]]


def test_julia_binding_covers_every_export():
    """julia/CSpB200.jl (the reference-side binding; not executable here, Julia is absent) binds every entry point the header
    declares and nothing the header does not."""
    hdr = open(os.path.join(ROOT, "include", "csp_b200.h")).read()
    declared = set(re.findall(r"\b(csp_[a-z_0-9]+)\s*\(", hdr))
    jl = open(os.path.join(ROOT, "julia", "CSpB200.jl")).read()
    bound = set(re.findall(r"\(:(csp_[a-z_0-9]+), LIB\)", jl))
    assert bound == declared, bound ^ declared
