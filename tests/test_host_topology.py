"""Host layer (include/tpose/*) against the reference's OWN headers compiled unmodified
(oracle/_ref/libtpose_ref.so): identical edit sequences, byte-for-byte identical meshes, files and
error reductions.  CPU only."""
import os

import numpy as np
import pytest

import hostlib as H

pytestmark = pytest.mark.skipif(not os.path.exists(H.REF_SO), reason="oracle/_ref not built (needs /root/reference at build time)")


@pytest.fixture(scope="module")
def libs():
    if not os.path.exists(H.HOST_SO):
        import __graft_entry__ as entry
        entry.build()
    return H.ref(), H.ours()


def _pair(libs, ratio):
    R, O = libs
    R.set_ratio(ratio); O.set_ratio(ratio)
    return H.Tri(R), H.Tri(O)


def _check(a, b, what):
    ok, where = H.same_state(a.get(), b.get())
    assert ok, f"{what}: {where} differs"


def _jitter_interior(rng, a, b, ratio, scale):
    s = a.get()
    pts = s["pts"].copy()
    if len(pts) > 4:
        d = rng.normal(0, scale, size=(len(pts) - 4, 2)).astype(np.float32)
        pts[4:] += d
        # sometimes push a point exactly onto / beyond the boundary
        if rng.random() < 0.3:
            k = int(rng.integers(4, len(pts)))
            pts[k, int(rng.integers(0, 2))] = np.float32(rng.choice([-1.0, 1.0, ratio, -ratio, 1.2]))
    a.set_points(pts); b.set_points(pts)


@pytest.mark.parametrize("seed", range(6))
def test_random_edit_sequences_match_reference(libs, seed):
    rng = np.random.default_rng(seed)
    ratio = float(np.float32(rng.choice([1.0, 1.5, 16 / 9])))
    a, b = _pair(libs, ratio)
    _check(a, b, "fresh")
    for it in range(160):
        nt, npnt = a.counts()
        op = rng.choice(["split", "split", "flip0", "flipd", "jitter", "collapse", "prune", "optimize", "short", "eraset"],
                        p=[.22, .1, .18, .1, .14, .08, .05, .05, .06, .02])
        if op == "split":
            t = int(rng.integers(0, nt))
            assert a.split(t) == b.split(t)
        elif op in ("flip0", "flipd"):
            h = int(rng.integers(-1, 3 * nt))
            ma = 0.0 if op == "flip0" else None
            assert a.flip(h, ma) == b.flip(h, ma)
        elif op == "jitter":
            _jitter_interior(rng, a, b, ratio, 0.02)
        elif op == "short" and npnt > 5:
            # make an edge shorter than the collapse threshold, then try to collapse every edge
            s = a.get()
            t = int(rng.integers(0, nt)); k = int(rng.integers(0, 3))
            i, j = int(s["tri"][t, k]), int(s["tri"][t, (k + 1) % 3])
            if i >= 4:
                pts = s["pts"].copy(); pts[i] = pts[j] + np.float32([0.003, -0.002])
                a.set_points(pts); b.set_points(pts)
            for h in rng.permutation(3 * nt)[:6]:
                nt2, _ = a.counts()
                if h < 3 * nt2:
                    assert a.collapse(int(h)) == b.collapse(int(h))
                    _check(a, b, f"collapse {h} @ {it}")
        elif op == "collapse":
            h = int(rng.integers(-1, 3 * nt))
            assert a.collapse(h) == b.collapse(h)
        elif op == "prune":
            t = int(rng.integers(0, nt))
            assert a.prune(t) == b.prune(t)
        elif op == "optimize" and nt > 2:
            assert a.optimize() == b.optimize()
        elif op == "eraset" and nt > 6:
            t = int(rng.integers(0, nt))
            assert a.eraset(t, True) == b.eraset(t, True)
        _check(a, b, f"{op} @ {it}")
        if a.counts()[0] < 1:
            break
    # queries on the final mesh
    nt, _ = a.counts()
    for h in range(3 * nt):
        assert np.float32(a.angle(h)).view(np.uint32) == np.float32(b.angle(h)).view(np.uint32) or (np.isnan(a.angle(h)) and np.isnan(b.angle(h)))
        assert a.hlength(h) == b.hlength(h)
    for t in range(nt):
        assert a.boundary(t) == b.boundary(t)
    a.close(); b.close()


def _grown_pair(libs, rng, ratio, nsplit):
    a, b = _pair(libs, ratio)
    for _ in range(nsplit):
        t = int(rng.integers(0, a.counts()[0]))
        a.split(t); b.split(t)
        if rng.random() < 0.5:
            h = int(rng.integers(0, 3 * a.counts()[0]))
            a.flip(h, 0.0); b.flip(h, 0.0)
    return a, b


def test_warp_and_reversewarp_match_reference(libs):
    rng = np.random.default_rng(5)
    ratio = 1.5
    a, b = _grown_pair(libs, rng, ratio, 60)
    a.origin_from_points(); b.origin_from_points()
    _jitter_interior(rng, a, b, ratio, 0.01)
    q = np.stack([rng.uniform(-ratio * 1.05, ratio * 1.05, 400), rng.uniform(-1.05, 1.05, 400)], 1).astype(np.float32)
    s = a.get()
    q[:20] = s["opts"][rng.integers(0, len(s["opts"]), 20)]          # exactly on vertices
    q[20:30] = 0.5 * (s["opts"][s["tri"][:10, 0]] + s["opts"][s["tri"][:10, 1]])   # on edges
    for rev in (False, True):
        wa, wb = a.warp(q, rev), b.warp(q, rev)
        assert np.array_equal(wa.view(np.uint32), wb.view(np.uint32))
        assert not np.array_equal(wa, q)
    a.close(); b.close()


def test_tri_files_are_byte_identical_and_cross_readable(libs, tmp_path):
    rng = np.random.default_rng(9)
    ratio = float(np.float32(16 / 9))
    a, b = _grown_pair(libs, rng, ratio, 25)
    fa, fb = str(tmp_path / "ref.tri"), str(tmp_path / "ours.tri")
    nrec = 3
    for lvl in range(nrec):
        for _ in range(12):
            t = int(rng.integers(0, a.counts()[0])); a.split(t); b.split(t)
        s = a.get()
        col = rng.integers(0, 256, size=(len(s["tri"]), 4)).astype(np.int32)
        a.set(s["tri"], s["he"], s["pts"], None, col); b.set(s["tri"], s["he"], s["pts"], None, col)
        a.origin_from_points(); b.origin_from_points()
        _jitter_interior(rng, a, b, ratio, 0.01)
        a.write(fa); b.write(fb)
    a.io_close(); b.io_close()
    da, db = open(fa, "rb").read(), open(fb, "rb").read()
    assert da == db and len(da) > 0
    # size formula of the format: 12 + 36*NT + 16*NP per record
    # each library reads the OTHER one's file, level by level, with and without dowarp
    for dowarp in (False, True):
        R, O = libs
        R.set_ratio(1.0); O.set_ratio(1.0)
        ra, rb = H.Tri(R), H.Tri(O)
        for lvl in range(nrec):
            assert ra.read(fb, dowarp) == 1 and rb.read(fa, dowarp) == 1
            assert R.get_ratio() == O.get_ratio() == ratio          # read() overwrites the global RATIO
            _check(ra, rb, f"read level {lvl} dowarp={dowarp}")
            assert (ra.get()["col"][:, 3] == 1).all() and (ra.get()["tri"][:, 3] == 0).all()
        assert ra.read(fb, dowarp) == 0 and rb.read(fa, dowarp) == 0   # EOF -> false, stream closed
        ra.close(); rb.close()
    a.close(); b.close()


def test_record_size_formula(libs, tmp_path):
    R, O = libs
    O.set_ratio(1.5)
    t = H.Tri(O)
    for _ in range(10):
        t.split(0)
    t.origin_from_points()                                         # like triangulate/main.cpp:226 before io::write
    nt, npnt = t.counts()
    f = str(tmp_path / "one.tri")
    t.write(f); t.io_close()
    assert os.path.getsize(f) == 12 + 36 * nt + 16 * npnt
    t.close()


def test_error_reductions_match_reference(libs):
    R, O = libs
    rng = np.random.default_rng(3)
    a, b = _grown_pair(libs, rng, 1.5, 40)
    nt = a.counts()[0]
    for trial in range(20):
        terr = rng.integers(0, 2 ** 31 - 1, size=13 * nt).astype(np.int32)
        if trial % 4 == 1:
            terr[rng.integers(0, nt, 3)] *= -1                     # wrapped accumulators
        if trial % 5 == 2:
            terr[:nt] = terr[0]                                    # ties: first max wins
        for L in (R, O):
            L.f("set_terr")(terr.ctypes.data_as(H.C.c_void_p), len(terr))
        R.f("set_toterr")(H.C.c_float(1.0 + trial)); O.f("set_toterr")(H.C.c_float(1.0 + trial))
        assert a.geterr() == b.geterr()
        assert R.f("get_toterr")() == O.f("get_toterr")()
        assert a.gettoterr() == b.gettoterr()
        assert R.f("get_relerr")() == O.f("get_relerr")()
        assert a.maxerrid() == b.maxerrid()
    a.close(); b.close()


def test_utility_matches_reference(libs):
    R, O = libs
    rng = np.random.default_rng(1)
    verts = rng.uniform(-1, 1, size=(30, 2)).astype(np.float32)
    verts[3] = verts[4]                                            # zero-length side
    verts[7] = 0.5 * (verts[5] + verts[6])                         # flat triangle
    out = [np.zeros(3, np.float32), np.zeros(3, np.float32)]
    for trial in range(400):
        t4 = np.array(list(rng.integers(0, 30, 3)) + [0], np.int32)
        if trial % 50 == 0:
            t4[:3] = [3, 4, 9]
        if trial % 50 == 1:
            t4[:3] = [5, 6, 7]
        p = rng.uniform(-1, 1, 2).astype(np.float32)
        for L, o in zip((R, O), out):
            L.f("barycentric")(H.C.c_float(p[0]), H.C.c_float(p[1]), H._p(t4), H._p(verts), 30, H._p(o))
        assert np.array_equal(out[0].view(np.uint32), out[1].view(np.uint32))
        assert R.f("intriangle")(H.C.c_float(p[0]), H.C.c_float(p[1]), H._p(t4), H._p(verts), 30) == \
               O.f("intriangle")(H.C.c_float(p[0]), H.C.c_float(p[1]), H._p(t4), H._p(verts), 30)
    # polynomial helpers: the real roots of random sextics (and of near-quintics, a6 tiny)
    for trial in range(60):
        roots = rng.uniform(-3, 3, 6)
        a7 = np.poly(roots)[::-1].copy() if trial % 2 == 0 else rng.normal(size=7)
        if trial % 6 == 3:
            a7[6] *= 1e-6
        ra, rb = np.zeros(6), np.zeros(6)
        na = R.f("realroots")(H._p(a7), H._p(ra)); nb = O.f("realroots")(H._p(a7), H._p(rb))
        assert na == nb
        assert np.allclose(np.sort(ra[:na]), np.sort(rb[:nb]), rtol=1e-9, atol=1e-12)
        x = float(rng.normal())
        assert R.f("horner")(H.C.c_double(x), H._p(a7), 7) == O.f("horner")(H.C.c_double(x), H._p(a7), 7)


def test_readmatches_matches_reference(libs, tmp_path):
    R, O = libs
    f = tmp_path / "m.txt"
    f.write_text("1 2 3 4\n5.5 6 7 8e1\nbad line\n9 10 11 12\n13 14 15 16")   # last line has no newline
    a = [np.zeros((8, 2), np.float32) for _ in range(4)]
    na = R.f("io_readmatches")(str(f).encode(), H._p(a[0]), H._p(a[1]), 8)
    nb = O.f("io_readmatches")(str(f).encode(), H._p(a[2]), H._p(a[3]), 8)
    assert na == nb == 3
    assert np.array_equal(a[0], a[2]) and np.array_equal(a[1], a[3])
    assert O.f("io_readmatches")(str(tmp_path / "missing.txt").encode(), H._p(a[2]), H._p(a[3]), 8) == -1
