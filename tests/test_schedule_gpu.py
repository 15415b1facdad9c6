"""The program loops (include/tpose/schedule, driven by tpose_cuda_run with device-side checks)
against the frame-by-frame CPU restatement of software/triangulate and software/warp
(oracle/schedule_oracle.py: reference headers unmodified + per-fragment step oracle)."""
import ctypes as C
import os

import numpy as np
import pytest

import hostlib as H

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.exists(H.REF_SO), reason="oracle/_ref not built")]


@pytest.fixture(scope="module")
def dev(cuda_pkg):
    L = H.ours()
    L.f("init")(0)
    yield L
    L.f("quit")()


def _triangulate_ours(L, img, ratio, exportlist, path, max_frames, raster=None):
    h, w = img.shape[:2]
    rw, rh = (w, h) if raster is None else raster
    L.set_ratio(ratio)
    L.f("set_toterr")(C.c_float(1.0))
    L.f("texture")(0, img.ctypes.data_as(C.c_void_p), w, h, rw, rh)
    tr = H.Tri(L)
    ex = (C.c_int * len(exportlist))(*exportlist)
    rep = (C.c_long * 7)()
    L.f("triangulate")(tr.h, ex, len(exportlist), path.encode(), C.c_long(max_frames), C.c_double(1e-4), rep)
    tr.io_close()
    st = tr.get()
    tr.close()
    return dict(frames=rep[0], host_frames=rep[1], runs=rep[2], splits=rep[3], exports=rep[4], finished=bool(rep[6]), tr=st)


@pytest.mark.parametrize("w,h,exportlist,seed", [(96, 64, [30, 20, 10], 4), (240, 160, [80, 40, 20], 5), (160, 160, [50, 25], 6)])
def test_triangulate_schedule_equals_frame_by_frame_reference(dev, synth, tmp_path, w, h, exportlist, seed):
    from oracle import schedule_oracle as S
    img = synth.make_image(w, h, seed=seed)
    ratio = float(np.float32(w) / np.float32(h))
    fo, fr = str(tmp_path / "ours.tri"), str(tmp_path / "ref.tri")
    ref = S.triangulate_loop(H.ref(), H.Tri, img, ratio, exportlist, fr, 30000)
    got = _triangulate_ours(dev, img, ratio, exportlist, fo, 30000)
    assert ref["finished"] and got["finished"]
    assert got["frames"] == ref["frames"]
    assert open(fo, "rb").read() == open(fr, "rb").read()
    ok, where = H.same_state(got["tr"], ref["tr"])
    assert ok, where
    # the point of the device-side checks: the host body ran on a fraction of the frames
    assert got["host_frames"] < got["frames"]


def test_triangulate_schedule_frame_cap(dev, synth, tmp_path):
    """max_frames stops both loops at the same state (mid-level, nothing exported yet)."""
    from oracle import schedule_oracle as S
    w, h = 200, 120
    img = synth.make_image(w, h, seed=8)
    ratio = float(np.float32(w) / np.float32(h))
    for cap in (1, 37, 300):
        fo, fr = str(tmp_path / f"o{cap}.tri"), str(tmp_path / f"r{cap}.tri")
        ref = S.triangulate_loop(H.ref(), H.Tri, img, ratio, [1000], fr, cap)
        got = _triangulate_ours(dev, img, ratio, [1000], fo, cap)
        assert got["frames"] == ref["frames"] == cap
        ok, where = H.same_state(got["tr"], ref["tr"])
        assert ok, f"cap {cap}: {where}"


def test_config2_schedule_prefix_at_1920x1080(dev, synth, tmp_path):
    """configs[1] at its stated size: the topology schedule from the two-triangle mesh on the 1920x1080 synthetic
    image, native raster -- the first 60 frames (converged frames with their flip sets and splits included) leave
    the same mesh as the frame-by-frame reference loop.  (The whole schedule cannot be compared: with the
    reference's step constant it does not get past a handful of triangles at this raster)"""
    from oracle import schedule_oracle as S
    w, h = 1920, 1080
    img = synth.make_image(w, h, seed=1000)
    ratio = float(np.float32(w) / np.float32(h))
    fo, fr = str(tmp_path / "o.tri"), str(tmp_path / "r.tri")
    cap = 60
    ref = S.triangulate_loop(H.ref(), H.Tri, img, ratio, [1000], fr, cap)
    got = _triangulate_ours(dev, img, ratio, [1000], fo, cap)
    assert got["frames"] == ref["frames"] == cap
    assert got["tr"]["counts"][0] > 2, "no split within the prefix: lengthen it"
    ok, where = H.same_state(got["tr"], ref["tr"])
    assert ok, where


def test_warp_schedule_equals_frame_by_frame_reference(dev, synth, tmp_path):
    from oracle import schedule_oracle as S
    w, h = 192, 128
    a, b = synth.make_pair(w, h, seed=12)
    ratio = float(np.float32(w) / np.float32(h))
    ex = [60, 30, 12]
    ta, tb = str(tmp_path / "a.tri"), str(tmp_path / "b.tri")
    ga = _triangulate_ours(dev, a, ratio, ex, ta, 30000)
    gb = _triangulate_ours(dev, b, ratio, ex, tb, 30000)
    assert ga["finished"] and gb["finished"] and ga["exports"] == gb["exports"] == 3
    # reference warp loop on copies of the stacks
    ra, rb = str(tmp_path / "ra.tri"), str(tmp_path / "rb.tri")
    for s, d in ((ta, ra), (tb, rb)):
        open(d, "wb").write(open(s, "rb").read())
    ref = S.warp_loop(H.ref(), H.Tri, a, b, ra, rb, 60000)
    # ours
    L = dev
    L.f("set_toterr")(C.c_float(1.0))
    L.f("texture")(0, a.ctypes.data_as(C.c_void_p), w, h, w, h)
    L.f("texture")(1, b.ctypes.data_as(C.c_void_p), w, h, w, h)
    rep = (C.c_long * 7)()
    L.f("warp")(ta.encode(), tb.encode(), C.c_long(60000), C.c_double(1e-6), rep)
    assert ref["finished"] and rep[6] == 1
    assert rep[0] == ref["frames"] and rep[5] == ref["levels"] == 3
    assert open(ta + ".warp", "rb").read() == open(ra + ".warp", "rb").read()
    assert open(tb + ".warp", "rb").read() == open(rb + ".warp", "rb").read()
    # the warp moved something
    assert open(ta + ".warp", "rb").read() != open(ta, "rb").read()


def test_ported_programs_write_the_same_files_as_the_library_calls(dev, synth, tmp_path):
    """software/triangulate.cpp + software/warp.cpp (PPM in, .tri / .tri.warp out) end to end."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe_t = os.path.join(root, "t-pose_b200", "bin", "triangulate")
    exe_w = os.path.join(root, "t-pose_b200", "bin", "warp")
    if not (os.path.exists(exe_t) and os.path.exists(exe_w)):
        pytest.fail("t-pose_b200/bin/* not built (run __graft_entry__.build())")
    w, h = 192, 128
    a, b = synth.make_pair(w, h, seed=3)
    paths = []
    for name, img in (("a", a), ("b", b)):
        p = str(tmp_path / f"{name}.ppm")
        with open(p, "wb") as f:
            f.write(b"P6\n%d %d\n255\n" % (w, h)); f.write(img[..., :3].tobytes())
        paths.append(p)
        out = subprocess.run([exe_t, "-i", p, "-native", "-frames", "3000"], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0 and "steps/s" in out.stdout, out.stdout + out.stderr
    # same job through the library entry point the other tests validate against the oracle
    ratio = float(np.float32(w) / np.float32(h))
    ex = [1000, 900, 800, 700, 600, 500, 400, 300, 200, 100, 50]
    lib_a = str(tmp_path / "lib_a.tri")
    _triangulate_ours(dev, a, ratio, ex, lib_a, 3000)
    assert open(paths[0] + ".tri", "rb").read() == open(lib_a, "rb").read()
    assert os.path.getsize(lib_a) > 0
    out = subprocess.run([exe_w, "-ia", paths[0], "-ib", paths[1], "-ta", paths[0] + ".tri", "-tb", paths[1] + ".tri", "-native"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "levels" in out.stdout, out.stdout + out.stderr
    assert os.path.getsize(paths[0] + ".tri.warp") == os.path.getsize(paths[0] + ".tri")
    # the same program on the tile-streaming kernels (no row-prefix table there, so converged frames take the
    # reference's two full re-evaluations instead of tpose_cuda_eval_list): same file
    tiles_in = str(tmp_path / "tiles.ppm")
    import shutil
    shutil.copy(paths[0], tiles_in)
    out = subprocess.run([exe_t, "-i", tiles_in, "-native", "-frames", "3000"], capture_output=True, text=True, timeout=300,
                         env=dict(os.environ, TPOSE_CUDA_PATH="tiles"))
    assert out.returncode == 0 and "steps/s" in out.stdout, out.stdout + out.stderr
    assert open(tiles_in + ".tri", "rb").read() == open(paths[0] + ".tri", "rb").read()
    # extension -t: continue from the finest level of an existing stack.  io::read leaves `colors` at that
    # level's NT; every level exported afterwards must still carry mean colours (bytes), not raw sums.
    import shutil, struct
    start = str(tmp_path / "start.tri")
    shutil.copy(paths[0] + ".tri", start)
    cont = str(tmp_path / "cont.ppm")
    shutil.copy(paths[0], cont)
    out = subprocess.run([exe_t, "-i", cont, "-native", "-t", start, "-frames", "8000"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "steps/s" in out.stdout, out.stdout + out.stderr
    buf, o, nrec = open(cont + ".tri", "rb").read(), 0, 0
    while o < len(buf):
        _, nt = struct.unpack_from("<fi", buf, o); o += 8
        rec = np.frombuffer(buf, np.int32, 9 * nt, o).reshape(nt, 9); o += 36 * nt
        (npnt,) = struct.unpack_from("<i", buf, o); o += 4 + 16 * npnt
        assert rec[:, 6:9].min() >= 0 and rec[:, 6:9].max() <= 255, "exported colours are not means"
        nrec += 1
    assert nrec >= 1, "the continuation exported no level: lengthen -frames"
    # software/view, headless: one picture per level
    exe_v = os.path.join(root, "t-pose_b200", "bin", "view")
    out = subprocess.run([exe_v, "-t", paths[0] + ".tri", "-o", str(tmp_path / "shot"), "-height", "120"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "level 0" in out.stdout, out.stdout + out.stderr
    shot = open(str(tmp_path / "shot.0.ppm"), "rb").read()
    assert shot.startswith(b"P6\n180 120\n255\n") and len(shot) == len(b"P6\n180 120\n255\n") + 180 * 120 * 3
    assert len(set(shot[-180 * 120 * 3:])) > 4                  # not a blank frame


def test_device_warp_equals_reference_warp_bit_for_bit(cuda_pkg, synth):
    """tpose_cuda_carry (csrc/carry.cuh) against the reference's own triangulation::warp /
    reversewarp (source/triangulation.hpp:465-520, compiled unmodified): random points, points on
    the boundary, on vertices and on edges (which stay where they are), degenerate triangles."""
    R = H.ref()
    rng = np.random.default_rng(5)
    for ntri, ratio in ((150, 1.5), (1000, 16 / 9), (3000, 16 / 9)):
        R.set_ratio(ratio)
        opts, tri, _ = synth.mesh_for(ntri, ratio)
        pts = opts.copy()
        inner = np.arange(len(pts)) >= 4
        pts[inner] += rng.normal(0, 0.004, (inner.sum(), 2)).astype(np.float32)
        pts[:, 0] = np.clip(pts[:, 0], -np.float32(ratio), np.float32(ratio)); pts[:, 1] = np.clip(pts[:, 1], -1, 1)
        tri = tri.copy()
        tri[5, :3] = (tri[5, 0], tri[5, 0], tri[5, 2])          # a zero-length side
        pts[tri[7, 1]] = pts[tri[7, 0]]                         # two coincident vertices
        # slivers around barycentric()'s |det| < 1E-8 cut and just above it (ill-conditioned inverses whose
        # barycentrics are rounding noise: whatever the reference header decides, the device decides too)
        for k, eps in zip(range(20, 20 + 8 * 9, 9), (0.0, 1e-9, 3e-9, 1e-8, 3e-8, 1e-7, 1e-6, 1e-5)):
            a, b = pts[tri[k, 0]].astype(np.float64), pts[tri[k, 1]].astype(np.float64)
            nrm = np.array([-(b - a)[1], (b - a)[0]]) / max(np.linalg.norm(b - a), 1e-12)
            both = (0.5 * (a + b) + eps * nrm).astype(np.float32)
            pts[tri[k, 2]] = both; opts[tri[k, 2]] = both        # degenerate in both embeddings (warp and reversewarp)
        q = np.stack([rng.uniform(-ratio, ratio, 4000), rng.uniform(-1, 1, 4000)], 1).astype(np.float32)
        q[:50, 0] = np.float32(ratio); q[50:100, 1] = -1.0       # on the boundary: never moved
        q[100:200] = opts[rng.integers(0, len(opts), 100)]       # exactly on vertices of the source embedding
        e = rng.integers(0, len(tri), 100)
        q[200:300] = (0.5 * (opts[tri[e, 0]] + opts[tri[e, 1]])).astype(np.float32)   # on edges (up to rounding)
        t = H.Tri(R)
        t.set(tri, np.full(3 * len(tri), -1, np.int32), pts, opts)
        for reverse in (False, True):
            want = t.warp(q, reverse=reverse)
            frm, to = (pts, opts) if reverse else (opts, pts)
            got = cuda_pkg.carry(tri, frm, to, q, float(np.float32(ratio)))
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (ntri, reverse, np.abs(got - want).max())
            assert not np.array_equal(got, q)                    # something moved
        t.close()
