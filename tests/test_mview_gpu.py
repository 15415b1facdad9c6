"""Multiview kernels (csrc/mview.cuh) through the C ABI against the numpy oracle.
Tolerance: 1e-4 relative, the bound BASELINE.json's north_star states for floating point."""
import ctypes as C
import os

import numpy as np
import pytest

import hostlib as H
from oracle import mview_oracle as M

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sfm_matches.npz")
RTOL = 1e-4


def _close(got, ref, rtol=RTOL):
    ref = np.asarray(ref, np.float64); got = np.asarray(got, np.float64)
    return np.abs(got - ref).max() <= rtol * np.abs(ref).max()


def _synthetic_pair(rng, n, noise):
    """matches of random 3-D points seen by two cameras, plus noise."""
    X = np.c_[rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), rng.uniform(4, 8, n)]
    th = rng.uniform(-0.1, 0.1)
    R = np.array([[np.cos(th), 0, np.sin(th)], [0, 1, 0], [-np.sin(th), 0, np.cos(th)]])
    t = np.array([0.5, rng.uniform(-0.1, 0.1), 0.05])
    a = X[:, :2] / X[:, 2:3]
    Xb = X @ R.T + t
    b = Xb[:, :2] / Xb[:, 2:3]
    return (a + rng.normal(0, noise, a.shape)).astype(np.float32) + np.float32(0.5), (b + rng.normal(0, noise, b.shape)).astype(np.float32) + np.float32(0.5)


def test_f8point_golden_and_batched(cuda_pkg):
    g = np.load(GOLD)
    rng = np.random.default_rng(0)
    As, Bs = [g["A"]], [g["B"]]
    for n in (8, 9, 50, 1581, 3162, 10000):
        a, b = _synthetic_pair(rng, n, 1e-3)
        As.append(a); Bs.append(b)
    As.append(np.zeros((0, 2), np.float32)); Bs.append(np.zeros((0, 2), np.float32))   # empty -> identity
    F = cuda_pkg.f8point(As, Bs)
    assert np.array_equal(F[-1], np.eye(3, dtype=np.float32))
    for i in range(len(As) - 1):
        ref = M.f_8point(As[i], Bs[i])
        assert _close(F[i], ref), (i, np.abs(F[i] - ref).max() / np.abs(ref).max())
        assert F[i][2, 2] == 1.0
    # the pin itself: OpenCV's 8-point on the reference's match list
    assert np.abs(F[0] - g["F_cv"]).max() < 1e-4 * np.abs(g["F_cv"]).max()
    # a batch gives the same numbers as one-at-a-time calls
    solo = cuda_pkg.f8point([As[3]], [Bs[3]])[0]
    assert np.array_equal(solo, F[3])


def test_fsampson_golden_batched_weighted(cuda_pkg):
    """mview::F_Sampson (multiview.hpp:187-299).  The iteration re-derives float32 weights from a
    float32 F a hundred times; between two consecutive iterations of the oracle itself F/F(2,2)
    moves by ~2e-5; with the normalisation summed like the reference (float32, index order) the device
    iterates on the very same normalised points as the oracle and the north_star bound of 1e-4 holds."""
    g = np.load(GOLD)
    rng = np.random.default_rng(1)
    As, Bs = [g["A"]], [g["B"]]
    for n in (12, 200, 1581):
        a, b = _synthetic_pair(rng, n, 2e-3)
        As.append(a); Bs.append(b)
    F = cuda_pkg.fsampson(As, Bs)
    assert _close(F[0], g["F_sampson"], 1e-4), np.abs(F[0] - g["F_sampson"]).max() / np.abs(g["F_sampson"]).max()
    for i in range(1, len(As)):
        ref = M.f_sampson(As[i], Bs[i])
        assert _close(F[i], ref, 1e-4), (i, np.abs(F[i] - ref).max() / np.abs(ref).max())
    # zero iterations is F_8Point
    assert np.array_equal(cuda_pkg.fsampson(As, Bs, iterations=0), cuda_pkg.f8point(As, Bs))
    # unit weights change nothing; real weights follow the three-argument overload
    ones = [np.ones(len(a), np.float32) for a in As]
    assert np.array_equal(cuda_pkg.fsampson(As, Bs, weights=ones), F)
    w = [rng.uniform(0.2, 1.0, len(a)).astype(np.float32) for a in As]
    Fw = cuda_pkg.fsampson(As, Bs, weights=w)
    for i in (0, 2):
        ref = M.f_sampson(As[i], Bs[i], w=w[i])
        assert _close(Fw[i], ref, 1e-4), (i, np.abs(Fw[i] - ref).max() / np.abs(ref).max())
    # empty list -> identity, like F_8Point
    Fe = cuda_pkg.fsampson([np.zeros((0, 2), np.float32)], [np.zeros((0, 2), np.float32)])
    assert np.array_equal(Fe[0], np.eye(3, dtype=np.float32))
    # the C++ header drives the same kernel
    L = H.ours()
    F9 = np.zeros(9, np.float32)
    L.f("fsampson")(H._p(g["A"]), len(g["A"]), H._p(g["B"]), len(g["B"]), None, H._p(F9))
    assert np.array_equal(F9.reshape(3, 3), F[0])
    L.f("fsampson")(H._p(g["A"]), len(g["A"]), H._p(g["B"]), len(g["B"]), H._p(w[0]), H._p(F9))
    assert np.array_equal(F9.reshape(3, 3), Fw[0])


def test_triangulate_golden(cuda_pkg):
    g = np.load(GOLD)
    F = g["F"].astype(np.float32)
    X, Ac, Bc = cuda_pkg.triangulate([F], g["K"], [g["A"]], [g["B"]], correct=True)
    assert np.abs(Ac[0] - g["Ac"]).max() < 1e-6 and np.abs(Bc[0] - g["Bc"]).max() < 1e-6
    assert _close(X[0], g["X"])
    Xn, An, Bn = cuda_pkg.triangulate([F], g["K"], [g["A"]], [g["B"]], correct=False)
    assert np.array_equal(An[0], g["A"]) and np.array_equal(Bn[0], g["B"])
    assert _close(Xn[0], g["X_nc"])
    assert (X[0][:, 3] == 1.0).all()


def test_triangulate_batched_synthetic(cuda_pkg):
    rng = np.random.default_rng(5)
    K = np.eye(3, dtype=np.float32)
    As, Bs, Fs = [], [], []
    for n in (30, 200, 777):
        a, b = _synthetic_pair(rng, n, 2e-4)
        As.append(a); Bs.append(b)
        Fs.append(M.f_8point(a, b).astype(np.float32))
    X, Ac, Bc = cuda_pkg.triangulate(Fs, K, As, Bs, correct=True)
    for i in range(len(As)):
        ref = M.triangulate(Fs[i], K, As[i], Bs[i])
        # compare points through their reprojections as well: depth along the baseline is ill-conditioned
        assert np.abs(X[i] - ref).max() <= 2e-3 * np.abs(ref).max(), i
        F64 = Fs[i].astype(np.float64)
        xa = np.c_[Ac[i], np.ones(len(Ac[i]))]; xb = np.c_[Bc[i], np.ones(len(Bc[i]))]
        raw = np.abs(np.einsum("ni,ij,nj->n", np.c_[Bs[i], np.ones(len(Bs[i]))], F64, np.c_[As[i], np.ones(len(As[i]))])).mean()
        cor = np.abs(np.einsum("ni,ij,nj->n", xb, F64, xa)).mean()
        assert cor < 0.05 * raw + 1e-6          # corrected matches lie on their epipolar lines


def test_hdlt_and_host_header_path(cuda_pkg):
    g = np.load(GOLD)
    K = g["K"].astype(np.float64)
    F = g["F"].astype(np.float32)
    E = K.T @ F.astype(np.float64) @ K
    R1, R2, t = M.get_pose(E)
    KPA = (K @ np.hstack([np.eye(3), np.zeros((3, 1))])).astype(np.float32)
    KPB = (K @ np.hstack([R2, -t[:, None]])).astype(np.float32)
    X = cuda_pkg.hdlt(KPA, KPB, g["A"], g["B"]).astype(np.float64)
    assert _close(X / X[:, 3:4], g["X_nc"])
    # the C++ header (tpose/multiview) drives the same kernels
    L = H.ours()
    F9 = np.zeros(9, np.float32)
    L.f("f8point")(H._p(g["A"]), len(g["A"]), H._p(g["B"]), len(g["B"]), H._p(F9))
    assert _close(F9.reshape(3, 3), g["F"])
    L.f("f8point")(H._p(g["A"]), 5, H._p(g["B"]), 6, H._p(F9))           # size mismatch -> identity
    assert np.array_equal(F9.reshape(3, 3), np.eye(3, dtype=np.float32))
    X4 = np.zeros((len(g["A"]), 4), np.float32)
    L.f("triangulate_points")(H._p(F.reshape(9)), H._p(g["K"].reshape(9)), H._p(g["A"]), H._p(g["B"]), len(g["A"]), H._p(X4))
    assert _close(X4, g["X"])
    a, b = g["A"][3].copy(), g["B"][3].copy()
    L.f("correct_match")(H._p(F.reshape(9)), H._p(a), H._p(b))
    assert np.abs(a - g["Ac"][3]).max() < 1e-6 and np.abs(b - g["Bc"][3]).max() < 1e-6
    Xs = np.zeros(4, np.float32)
    L.f("hdlt")(H._p(KPA.reshape(12)), H._p(KPB.reshape(12)), H._p(g["A"][0]), H._p(g["B"][0]), H._p(Xs))
    assert _close(Xs.astype(np.float64) / Xs[3], g["X_nc"][0])


def test_select_matches_two_way_inconsistency(cuda_pkg, synth):
    """tests/compute_fundamental_mat/model.h:1-82 (dead code in the reference, restated loop for loop by
    oracle/mview_oracle.select_matches): same matches, same order, same float32 weights -- bit for bit --
    and the weights drive F_Sampson's three-argument overload."""
    rng = np.random.default_rng(5)
    ratio = 1920 / 1080
    pts, tri, _ = synth.mesh_for(3000, ratio)
    NP = len(pts)
    f = np.float32
    originA = pts.copy()
    shift = (rng.normal(0, 0.01, pts.shape)).astype(f)
    pointsA = (pts + shift).astype(f)
    pointsA[:4] = pts[:4]
    # the other direction's stack disagrees by a little, by nothing, or by a lot
    noise = rng.choice([0.0, 1e-7, 3e-4, 5e-2], size=(NP, 1)).astype(f) * rng.normal(0, 1, pts.shape).astype(f)
    worigA = (pointsA + noise).astype(f)
    pointsB = (pts - shift).astype(f); originB = pts.copy()
    worigB = (pointsB + noise[::-1]).astype(f)
    for max_dist in (0.0000002, 1e-3, 10.0):
        X, Y, W, D = cuda_pkg.select_matches(originA, pointsA, worigA, pointsB, originB, worigB, ratio, max_dist)
        rX, rY, rW, rD = M.select_matches(originA, pointsA, worigA, pointsB, originB, worigB, ratio, max_dist)
        assert X.shape == rX.shape and len(X) > 0
        assert np.array_equal(X.view(np.uint32), rX.view(np.uint32)) and np.array_equal(Y.view(np.uint32), rY.view(np.uint32))
        assert np.array_equal(W.view(np.uint32), rW.view(np.uint32)) and np.array_equal(D.view(np.uint32), rD.view(np.uint32))
    assert np.isinf(W).any()                      # d = 0 -> weight 1/0, as coded
    # finite weights into the weighted Sampson iteration
    X, Y, W, _ = cuda_pkg.select_matches(originA, pointsA, worigA, pointsB, originB, worigB, ratio, 1e-3)
    ok = np.isfinite(W)
    X, Y, W = X[ok], Y[ok], np.minimum(W[ok], f(1e6))
    Fw = cuda_pkg.fsampson([X], [Y], weights=[W])[0]
    ref = M.f_sampson(X, Y, w=W)
    assert _close(Fw, ref, 1e-4), np.abs(Fw - ref).max() / np.abs(ref).max()
    # C++ header entry point (tpose::mview::select_matches) builds the same vectors
    L = H.ours()
    if hasattr(L.lib, "tph_select_matches"):
        n = C.c_int()
        oX = np.zeros((2 * NP, 2), f); oY = np.zeros((2 * NP, 2), f); oW = np.zeros(2 * NP, f)
        L.f("select_matches")(H._p(originA), H._p(pointsA), H._p(worigA), NP, H._p(pointsB), H._p(originB), H._p(worigB), NP,
                              C.c_float(ratio), C.c_double(1e-3), H._p(oX), H._p(oY), H._p(oW), C.byref(n))
        rX, rY, rW, _ = M.select_matches(originA, pointsA, worigA, pointsB, originB, worigB, ratio, 1e-3)
        assert n.value == len(rX) and np.array_equal(oX[:n.value].view(np.uint32), rX.view(np.uint32))
        assert np.array_equal(oW[:n.value].view(np.uint32), rW.view(np.uint32))
