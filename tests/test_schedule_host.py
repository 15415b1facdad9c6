"""The sparse per-frame edit loops of include/tpose/schedule (what the history replay runs) against
the plain loops of the reference, on the host only: random meshes, several frames of drifting
points, a conservative candidate mask computed -- like the device does -- against the topology of
the LAST synchronisation (stale while the host keeps flipping), and the `dirty` set carried from
frame to frame.  After every frame both meshes must be identical."""
import ctypes as C

import numpy as np
import pytest

import hostlib as H


def conservative_mask(tri, he, pts, ratio):
    """numpy twin of watch_triangle (csrc/eval.cuh) without the convexity refinement: a superset of
    the triangles on which prune / the 0.8*pi flip / collapse can act."""
    NT = tri.shape[0]
    p = pts[tri[:, :3]].astype(np.float32)                       # (NT, 3, 2)
    flag = np.zeros(NT, bool)
    for k in range(3):
        a, b, c = p[:, k], p[:, (k + 1) % 3], p[:, (k + 2) % 3]
        u, v = a - c, b - c
        lu = np.sqrt((u * u).sum(1)); lv = np.sqrt((v * v).sum(1))
        with np.errstate(divide="ignore", invalid="ignore"):
            cs = (u * v).sum(1) / lu / lv
        flag |= (cs < -0.80) & (he.reshape(NT, 3)[:, k] >= 0)
        flag |= np.sqrt(((b - a) ** 2).sum(1)) <= 0.0101
    on_b = (np.abs(p[..., 0]) >= np.float32(ratio)) | (np.abs(p[..., 1]) >= 1.0)
    flag |= on_b.all(1) & (he.reshape(NT, 3) < 0).any(1)
    return flag


def pack32(flag):
    nw = (len(flag) + 31) // 32
    bits = np.zeros(nw * 32, np.uint8); bits[:len(flag)] = flag
    return np.packbits(bits, bitorder="little").view(np.uint32).copy()


def frame_edits(L, t, mask32, dirty):
    counts = (C.c_long * 3)()
    m = None if mask32 is None else mask32.ctypes.data_as(C.c_void_p)
    upd = L.f("frame_edits")(t.h, m, 0 if mask32 is None else len(mask32), dirty.ctypes.data_as(C.c_void_p), len(dirty), counts)
    return bool(upd), tuple(counts)


@pytest.mark.parametrize("seed", range(8))
def test_sparse_frame_edits_equal_the_plain_loops(seed):
    import importlib, os, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import __graft_entry__ as entry
    entry.load_package()
    synth = importlib.import_module("tpose_b200.synth")
    L = H.ours()
    rng = np.random.default_rng(100 + seed)
    ratio = 1.5
    L.set_ratio(ratio)
    ntri = (150, 1000)[seed % 2]
    pts, tri, he = synth.mesh_for(ntri, ratio)
    pts = pts.copy()
    inner = (np.abs(pts[:, 0]) < ratio) & (np.abs(pts[:, 1]) < 1)
    cell = 2.0 / np.sqrt(ntri / 2)
    pts[inner] += rng.normal(0, 0.45 * cell, (inner.sum(), 2)).astype(np.float32)      # slivers, inverted cells, short edges
    # flat boundary triangles (prune candidates): pull the inner corner of a few border cells onto the border
    r32 = np.float32(ratio)
    on_right = pts[tri[:, :3], 0] >= r32
    for t_ in np.flatnonzero(on_right.sum(1) == 2)[: 3 + seed % 3]:
        pts[tri[t_, :3][~on_right[t_]][0], 0] = r32
    plain, sparse = H.Tri(L), H.Tri(L)
    for t in (plain, sparse):
        t.set(tri, he, pts, pts)
    dirty = np.zeros(64, np.uint64)
    seen_tri, seen_he = tri.copy(), he.copy()          # the topology the "device" holds
    total = [0, 0, 0]
    for frame in range(10):
        st = sparse.get()
        cur = st["pts"]
        if st["tri"].shape == seen_tri.shape:
            mask = pack32(conservative_mask(seen_tri, seen_he, cur, ratio))
        else:
            raise AssertionError("a re-indexed mesh must have been re-synchronised")
        up_p, cnt_p = frame_edits(L, plain, None, np.zeros(64, np.uint64))
        up_s, cnt_s = frame_edits(L, sparse, mask, dirty)
        assert (up_p, cnt_p) == (up_s, cnt_s), (frame, cnt_p, cnt_s)
        a, b = plain.get(), sparse.get()
        ok, where = H.same_state(a, b)
        assert ok, (frame, where)
        for i in range(3):
            total[i] += cnt_p[i]
        if up_s:                                       # prune / collapse: upload + watch, like the schedule does
            dirty[:] = 0
            seen_tri, seen_he = b["tri"].copy(), b["he"].copy()
        elif frame % 4 == 3:                           # end of a run with flips only: watch
            dirty[:] = 0
            seen_tri, seen_he = b["tri"].copy(), b["he"].copy()
        # the optimiser moves the points a little before the next frame
        nxt = b["pts"].copy()
        inner = (np.abs(nxt[:, 0]) < ratio) & (np.abs(nxt[:, 1]) < 1)
        inner[:4] = False
        nxt[inner] += rng.normal(0, 0.08 * cell, (inner.sum(), 2)).astype(np.float32)
        plain.set_points(nxt); sparse.set_points(nxt)
    assert total[0] > 0                                # the frames did flip things
    plain.close(); sparse.close()


@pytest.mark.parametrize("seed", range(6))
def test_flip_set_matches_the_std_set_semantics_of_the_reference(seed):
    """detail::flip_set (stable sort + unique) against the literal restatement of the reference's
    std::set / std::map logic (oracle/schedule_oracle._flip_set, triangulate/main.cpp:233-281), with
    many tied energies -- the comparator's tie rule is behaviour."""
    import importlib, os, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import __graft_entry__ as entry
    entry.load_package()
    synth = importlib.import_module("tpose_b200.synth")
    from oracle import schedule_oracle as S
    L = H.ours()
    rng = np.random.default_rng(seed)
    ratio = 1.5
    L.set_ratio(ratio)
    pts, tri, he = synth.mesh_for((150, 1000)[seed % 2], ratio)
    nt = tri.shape[0]
    t = H.Tri(L)
    t.set(tri, he, pts, pts)
    # few distinct values -> many equal sums; a couple of huge ones whose int sum wraps like the reference's
    terr = rng.integers(0, 6 if seed < 4 else 2_000_000, nt).astype(np.int32)
    terr[rng.integers(0, nt, 3)] = 2_000_000_000
    L.f("set_terr")(terr.ctypes.data_as(C.c_void_p), nt)
    out_h = np.zeros(3 * nt, np.int32); out_e = np.zeros(3 * nt, np.float32)
    n = L.f("flip_set")(t.h, H._p(out_h), H._p(out_e))
    with np.errstate(over="ignore"):
        want = S._flip_set(terr, he, nt)
    assert n == len(want) and n > 0
    assert out_h[:n].tolist() == list(want.keys())
    assert np.array_equal(out_e[:n], np.array(list(want.values()), np.float32))
    t.close()


def test_prefilters_decide_like_the_reference_predicates():
    """detail::very_obtuse == (triangulation::angle(h) > 0.8*PI) and detail::collapse_candidate picks the
    half-edge the reference's collapse loop would try (or -1 when collapse() would refuse it), on
    triangles built to sit around both thresholds -- the reference's own angle()/hlength() (unmodified
    headers, oracle/_ref) are the authority."""
    import os
    if not os.path.exists(H.REF_SO):
        pytest.skip("oracle/_ref not built")
    L, R = H.ours(), H.ref()
    rng = np.random.default_rng(17)
    ratio = 1.5
    L.set_ratio(ratio); R.set_ratio(ratio)
    n = 4000
    # apex angle swept through 0.8*pi (144 deg) and side lengths through 0.01
    ang = np.deg2rad(np.concatenate([rng.uniform(143.9, 144.1, n // 2), rng.uniform(100, 179.9, n // 2)]))
    la = np.concatenate([rng.uniform(0.0099, 0.0101, n // 4), rng.uniform(0.002, 0.5, n - n // 4)])
    lb = rng.uniform(0.005, 0.5, n)
    rot = rng.uniform(0, 2 * np.pi, n)
    apex = np.stack([rng.uniform(-1, 1, n), rng.uniform(-0.6, 0.6, n)], 1)
    a = apex + la[:, None] * np.stack([np.cos(rot), np.sin(rot)], 1)
    b = apex + lb[:, None] * np.stack([np.cos(rot + ang), np.sin(rot + ang)], 1)
    pts = np.concatenate([a, b, apex]).astype(np.float32)
    tri = np.zeros((n, 4), np.int32)
    tri[:, 0] = np.arange(n); tri[:, 1] = np.arange(n) + n; tri[:, 2] = np.arange(n) + 2 * n     # half-edge 0 faces the apex
    he = np.full(3 * n, -1, np.int32)
    to, tr_ = H.Tri(L), H.Tri(R)
    for t in (to, tr_):
        t.set(tri, he, pts, pts)
    pi32 = float(np.float32(3.14159265))                  # tpose::PI
    hits = 0
    for t in range(n):
        for k in range(3):
            want = float(np.float32(tr_.angle(3 * t + k))) > 0.8 * pi32
            assert bool(L.f("very_obtuse")(to.h, 3 * t + k)) == want, (t, k)
            hits += want
        # the reference's loop: the first strictly shortest side; collapse() refuses it if longer than 0.01
        ls = [float(np.float32(tr_.hlength(3 * t + k))) for k in range(3)]
        ha = 0
        if ls[1] < ls[ha]: ha = 1
        if ls[ha + 1 if ha < 2 else 2] < ls[ha] and ha < 2: ha = ha + 1
        got = L.f("collapse_candidate")(to.h, t)
        if ls[ha] > 0.01:
            assert got in (-1, 3 * t + ha), (t, got, ls)          # -1, or a candidate collapse() will refuse itself
        else:
            assert got == 3 * t + ha, (t, got, ls)
    assert 500 < hits < 3 * n
    to.close(); tr_.close()
