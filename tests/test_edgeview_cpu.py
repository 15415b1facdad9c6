"""tpose/detail/edgeview.h (host only): the unique-edge / corner view the step kernels work on,
checked against a brute-force numpy construction -- grid meshes, open meshes, edges shared by
three triangles, repeated vertices inside a triangle, isolated vertices, bad indices."""
import ctypes as C

import numpy as np
import pytest

import hostlib as H


def edgeview(tri, np_):
    L = H.ours()
    tri = np.ascontiguousarray(tri, np.int32)
    nt = tri.shape[0]
    edges = np.zeros((3 * nt, 2), np.int32); te = np.zeros((nt, 4), np.int32)
    off = np.zeros(np_ + 1, np.int32); cor = np.zeros(3 * nt, np.int32)
    ne = L.f("edgeview")(H._p(tri), nt, np_, H._p(edges), H._p(te), H._p(off), H._p(cor))
    return ne, edges[:max(ne, 0)], te, off, cor


def check(tri, np_):
    tri = np.asarray(tri, np.int32)
    nt = tri.shape[0]
    ne, edges, te, off, cor = edgeview(tri, np_)
    # unique undirected edges, stored (lo, hi), sorted, no duplicates
    want = sorted({(min(int(tri[t, j]), int(tri[t, (j + 1) % 3])), max(int(tri[t, j]), int(tri[t, (j + 1) % 3]))) for t in range(nt) for j in range(3)})
    assert ne == len(want) and [tuple(e) for e in edges.tolist()] == want
    # every side points at its edge, with the reversed bit exactly when it runs hi -> lo
    for t in range(nt):
        for j in range(3):
            x, y = int(tri[t, j]), int(tri[t, (j + 1) % 3])
            packed = int(te[t, j])
            e = packed & 0x7FFFFFFF
            assert tuple(edges[e]) == (min(x, y), max(x, y))
            assert (packed < 0) == (x > y)
        assert te[t, 3] == 0
    # corners: a CSR over vertices that holds every (triangle, corner) exactly once, under its vertex
    assert off[0] == 0 and off[-1] == 3 * nt and np.all(np.diff(off) >= 0)
    seen = set()
    for v in range(np_):
        for c in cor[off[v]:off[v + 1]]:
            t, j = int(c) >> 2, int(c) & 3
            assert j < 3 and int(tri[t, j]) == v
            seen.add((t, j))
    assert len(seen) == 3 * nt


def test_grid_meshes_and_random_soups():
    import importlib, os, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import __graft_entry__ as entry
    entry.load_package()
    synth = importlib.import_module("tpose_b200.synth")
    for ntri in (150, 1000):
        pts, tri, _ = synth.mesh_for(ntri, 1.5)
        check(tri, len(pts))
        ne = edgeview(tri, len(pts))[0]
        assert ne == ntri + len(pts) - 1                 # Euler for a disc: E = T + V - 1
    rng = np.random.default_rng(0)
    for _ in range(20):
        np_ = int(rng.integers(3, 40)); nt = int(rng.integers(1, 60))
        tri = np.zeros((nt, 4), np.int32)
        tri[:, :3] = rng.integers(0, np_, (nt, 3))       # repeated vertices, shared edges, isolated vertices
        tri[:, 3] = rng.integers(-5, 5, nt)              # .w is ignored
        check(tri, np_)


def test_special_cases():
    check([[0, 1, 2, 0]], 3)
    check([[0, 1, 2, 0], [2, 1, 3, 0]], 4)                       # the reference's initial mesh (triangulation.hpp:42-69)
    check([[0, 1, 2, 0], [1, 0, 3, 0], [0, 1, 4, 0]], 6)         # an edge shared by three triangles, vertex 5 isolated
    check([[2, 2, 2, 0], [0, 0, 1, 0]], 3)                       # degenerate triangles
    ne, *_ = edgeview(np.array([[0, 1, 7, 0]], np.int32), 3)     # vertex index out of range
    assert ne == -1
    ne, *_ = edgeview(np.array([[0, -1, 2, 0]], np.int32), 3)
    assert ne == -1
