"""BASELINE.json configs 3, 4 and 5 in one process: the batched pair pipeline (include/tpose/pairs through
t-pose_b200/pairs.py) against the frame-by-frame CPU restatement of software/warp (oracle/schedule_oracle.py:
reference headers unmodified + per-fragment step oracle) and the numpy multiview oracle.

The input stacks are synthetic grid stacks (jittered grids coloured by the device's colour pass): the
reference's own triangulate schedule does not settle at these rasters, and the warp
loop only needs *a* coarse-to-fine stack.  Topology is frozen during the warp, so what is compared is the
per-level frame count and the bytes of both .warp stacks."""
import os
import shutil

import numpy as np
import pytest

import hostlib as H
from oracle import mview_oracle as M

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(H.REF_SO), reason="oracle/_ref not built")]


@pytest.fixture(scope="module")
def P(cuda_pkg):
    import importlib
    return importlib.import_module("tpose_b200.pairs")


def _matches_numpy(r, tris_a, tris_b, ratio):
    """tests/compute_fundamental_mat/main.cpp:137-167 with every triangle selected."""
    f = np.float32
    cx, cy = f(0.5) / f(ratio), f(-0.5) / f(ratio)
    T = lambda p: np.stack([(cx * p[:, 0] + f(0) * p[:, 1]) + f(0), (f(0) * p[:, 0] + cy * p[:, 1]) + f(0)], 1).astype(f)
    ua, ub = np.unique(tris_a[:, :3]), np.unique(tris_b[:, :3])
    X = np.concatenate([T(r["optsA"][ua]), T(r["ptsB"][ub])])
    Y = np.concatenate([T(r["ptsA"][ua]), T(r["optsB"][ub])])
    return X, Y


def test_pair_set_equals_single_pair_warp_and_oracle_multiview(cuda_pkg, synth, P, tmp_path):
    """Two different pairs advanced together give, per pair, the .warp bytes of tpose::schedule::warp run on
    that pair alone (itself byte-equal to the oracle loop, test_schedule_gpu.py); F and X for both pairs come
    out of two batched launches and agree with the numpy oracle on the same matches."""
    import ctypes as C
    w, h = 192, 128
    ratio = float(np.float32(w) / np.float32(h))
    sizes = (150, 500)
    ps = P.PairSet(0)
    imgs, stacks, lv = [], [], None
    for i, seed in enumerate((12, 13)):
        a, b = synth.make_pair(w, h, seed=seed)
        ta, tb = str(tmp_path / f"a{i}.tri"), str(tmp_path / f"b{i}.tri")
        lv = P.grid_stack(cuda_pkg, synth, ta, a, ratio, sizes); P.grid_stack(cuda_pkg, synth, tb, b, ratio, sizes)
        for s in (ta, tb):
            shutil.copy(s, s + ".solo")
        imgs.append((a, b)); stacks.append((ta, tb))
        ps.add(a, b, ta, tb, max_triangles=512, max_points=512)
    frames, levels = ps.run(tol=1e-6, level_cap=-1)
    assert levels == [2, 2]
    L = H.ours()
    L.f("init")(0)
    for i, ((a, b), (ta, tb)) in enumerate(zip(imgs, stacks)):
        L.f("set_toterr")(C.c_float(1.0))
        L.f("texture")(0, a.ctypes.data_as(C.c_void_p), w, h, w, h)
        L.f("texture")(1, b.ctypes.data_as(C.c_void_p), w, h, w, h)
        rep = (C.c_long * 7)()
        L.f("warp")((ta + ".solo").encode(), (tb + ".solo").encode(), C.c_long(-1), C.c_double(1e-6), rep)
        assert rep[0] == frames[i] and rep[5] == levels[i]
        assert open(ta + ".warp", "rb").read() == open(ta + ".solo.warp", "rb").read()
        assert open(tb + ".warp", "rb").read() == open(tb + ".solo.warp", "rb").read()
        r = ps.result(i)
        X, Y = _matches_numpy(r, lv[-1][1], lv[-1][1], ratio)
        assert len(X) == len(r["X"])
        refF = M.f_8point(X, Y)
        assert np.abs(r["F"] - refF).max() <= 1e-4 * np.abs(refF).max()
        sub = np.arange(0, len(X), 7)
        refX = M.triangulate(r["F"], M.camera(), X[sub], Y[sub])
        assert np.abs(r["X"][sub] - refX).max() <= 1e-4 * np.abs(refX).max()
    ps.close()


def test_config3_warp_loop_1280x720_four_levels_against_the_oracle_loop(cuda_pkg, synth, P, tmp_path):
    """configs[2]: 1280x720 pair, levels 150 / 500 / 1500 / 3000, every level run to geterr < 1e-6: the same
    number of frames per level and the same bytes in both .warp stacks as the frame-by-frame oracle loop."""
    from oracle import schedule_oracle as S
    w, h = 1280, 720
    ratio = float(np.float32(w) / np.float32(h))
    a, b = synth.make_pair(w, h, seed=1000)
    ta, tb = str(tmp_path / "a.tri"), str(tmp_path / "b.tri")
    P.grid_stack(cuda_pkg, synth, ta, a, ratio); P.grid_stack(cuda_pkg, synth, tb, b, ratio)
    ra, rb = str(tmp_path / "ra.tri"), str(tmp_path / "rb.tri")
    shutil.copy(ta, ra); shutil.copy(tb, rb)
    cap = 2500                                    # (a level that has not converged by then ends there, on both sides)
    ps = P.PairSet(0)
    ps.add(a, b, ta, tb, max_triangles=3000, max_points=3000)
    frames, levels = ps.run(tol=1e-6, level_cap=cap, mview=False)
    ps.close()
    ref = S.warp_loop(H.ref(), H.Tri, a, b, ra, rb, 4 * cap + 10, tol=1e-6, level_cap=cap)
    assert ref["finished"] and levels == [4] and ref["levels"] == 4
    assert frames[0] == ref["frames"], (frames, ref["per_level"])
    assert min(ref["per_level"][:2]) < cap, "the two coarsest levels are meant to end on convergence, not on the cap"
    assert open(ta + ".warp", "rb").read() == open(ra + ".warp", "rb").read()
    assert open(tb + ".warp", "rb").read() == open(rb + ".warp", "rb").read()


def test_config4_f_and_points_at_1080p_3000_triangles(cuda_pkg, synth, P, tmp_path):
    """configs[3]: warp + F + DLT at 1920x1080 / 3000 triangles.  The warp is time-boxed (300 frames per
    level); F_8Point and triangulate(F,K,A,B) on the 3 162 matches of the result agree with the oracle to 1e-4."""
    w, h = 1920, 1080
    ratio = float(np.float32(w) / np.float32(h))
    a, b = synth.make_pair(w, h, seed=1000)
    ta, tb = str(tmp_path / "a.tri"), str(tmp_path / "b.tri")
    lv = P.grid_stack(cuda_pkg, synth, ta, a, ratio); P.grid_stack(cuda_pkg, synth, tb, b, ratio)
    ps = P.PairSet(0)
    ps.add(a, b, ta, tb, max_triangles=3000, max_points=3000)
    frames, levels = ps.run(tol=1e-6, level_cap=300)
    assert levels == [4] and frames[0] <= 1200
    r = ps.result(0)
    ps.close()
    X, Y = _matches_numpy(r, lv[-1][1], lv[-1][1], ratio)
    assert len(X) == 2 * 1581 == len(r["X"])
    refF = M.f_8point(X, Y)
    assert np.abs(r["F"] - refF).max() <= 1e-4 * np.abs(refF).max()
    sub = np.arange(0, len(X), 16)
    refX = M.triangulate(r["F"], M.camera(), X[sub], Y[sub])
    assert np.abs(r["X"][sub] - refX).max() <= 1e-4 * np.abs(refX).max()
    assert np.abs(r["ptsA"] - r["optsA"]).max() > 0          # the warp moved A's mesh
