"""Multiview path, CPU side: the numpy oracle against its golden fixture and OpenCV, and the shared
host/device math (include/tpose/detail/smallmath.h, compiled for the host in libtpose_host.so)
against the oracle.  The CUDA kernels call the very same functions; tests/test_mview_gpu.py checks
them on the device."""
import ctypes as C
import os

import numpy as np
import pytest

import hostlib as H
from oracle import mview_oracle as M

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sfm_matches.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.fixture(scope="module")
def L():
    if not os.path.exists(H.HOST_SO):
        import __graft_entry__ as entry
        entry.build()
    return H.ours()


def _d(a):
    return np.ascontiguousarray(a, np.float64)


def test_oracle_reproduces_golden_and_opencv(gold):
    F = M.f_8point(gold["A"], gold["B"])
    assert np.allclose(F, gold["F"], rtol=1e-9, atol=1e-12)
    # the pin: OpenCV's normalised 8-point on the reference's own match list
    assert np.abs(F - gold["F_cv"]).max() < 5e-5
    try:
        import cv2
        Fcv, _ = cv2.findFundamentalMat(gold["A"].astype(np.float64), gold["B"].astype(np.float64), cv2.FM_8POINT)
        assert np.abs(F - Fcv).max() < 5e-5
    except ImportError:
        pass
    # structure of the literal in tests/compute_fundamental_mat/model.h:90-93 (same magnitudes/signs)
    lit = np.array([[-0.936791, 2.24, -26.1548], [1.03284, 1.77101, -13.5179], [26.3407, 10.5485, 1]])
    assert (np.sign(F) == np.sign(lit)).all() and np.abs(F - lit).max() < 2.0
    # rank 2, and the epipolar constraint holds for the matches
    assert abs(np.linalg.det(F)) < 1e-6 * np.abs(F).max() ** 3
    xa = np.c_[gold["A"], np.ones(len(gold["A"]))]; xb = np.c_[gold["B"], np.ones(len(gold["B"]))]
    r = np.einsum("ni,ij,nj->n", xb, F, xa)
    assert np.abs(r).mean() < 0.05
    X = M.triangulate(F.astype(np.float32), gold["K"], gold["A"], gold["B"])
    assert np.allclose(X, gold["X"], rtol=1e-7, atol=1e-9)


def test_normalize_properties():
    rng = np.random.default_rng(0)
    p = rng.normal(size=(500, 2)).astype(np.float32) * 3 + 7
    q, Hm = M.normalize(p)
    assert np.abs(q.mean(0)).max() < 1e-5
    assert abs(np.linalg.norm(q, axis=1).mean() - np.sqrt(2)) < 1e-5
    back = (Hm.astype(np.float64) @ np.c_[p, np.ones(len(p))].T).T[:, :2]
    assert np.abs(back - q).max() < 1e-4


def test_host_normalize_and_camera_match_oracle(L):
    rng = np.random.default_rng(1)
    p = rng.uniform(0, 1, size=(300, 2)).astype(np.float32)
    q, Hm = M.normalize(p)
    mine = p.copy(); H9 = np.zeros(9, np.float32)
    L.f("normalize")(H._p(mine), len(mine), H._p(H9))
    assert np.array_equal(mine.view(np.uint32), q.view(np.uint32))
    assert np.array_equal(H9.reshape(3, 3).view(np.uint32), Hm.view(np.uint32))
    K9 = np.zeros(9, np.float32)
    L.f("camera")(H._p(K9))
    assert np.array_equal(K9.reshape(3, 3), M.camera())


def test_shared_math_svd_gauge_and_pose(L):
    rng = np.random.default_rng(2)
    for trial in range(50):
        Fm = M._rank2(rng.normal(size=(3, 3)))
        U, S, V = M.svd3_proper(Fm)
        u, s, v = np.zeros(9), np.zeros(3), np.zeros(9)
        L.f("sm_svd3")(H._p(_d(Fm)), H._p(u), H._p(s), H._p(v))
        u, v = u.reshape(3, 3), v.reshape(3, 3)
        assert np.allclose(s[:2], S[:2], rtol=1e-10) and s[2] < 1e-7
        assert abs(np.linalg.det(u) - 1) < 1e-9 and abs(np.linalg.det(v) - 1) < 1e-9
        assert np.allclose(u[:, 2], U[:, 2], atol=1e-7) and np.allclose(v[:, 2], V[:, 2], atol=1e-7)
        assert np.allclose(u @ np.diag(s) @ v.T, Fm, atol=1e-9)
        # essential matrix with equal singular values: R1, R2, t are gauge independent
        R = np.linalg.qr(rng.normal(size=(3, 3)))[0]; R *= np.sign(np.linalg.det(R))
        t = rng.normal(size=3); t /= np.linalg.norm(t)
        E = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]]) @ R
        r1, r2, tt = np.zeros(9), np.zeros(9), np.zeros(3)
        L.f("sm_get_pose")(H._p(_d(E)), H._p(r1), H._p(r2), H._p(tt))
        R1, R2, T = M.get_pose(E)
        assert np.allclose(r1.reshape(3, 3), R1, atol=1e-7) and np.allclose(r2.reshape(3, 3), R2, atol=1e-7)
        assert np.allclose(tt, T, atol=1e-7)
        assert np.allclose(np.abs(tt), np.abs(t), atol=1e-7)


def test_shared_math_rank2_and_jacobi9(L):
    rng = np.random.default_rng(3)
    for trial in range(20):
        Fm = rng.normal(size=(3, 3))
        mine = _d(Fm).copy().reshape(9)
        L.f("sm_rank2")(H._p(mine))
        assert np.allclose(mine.reshape(3, 3), M._rank2(Fm), atol=1e-12)
        A = rng.normal(size=(40, 9)); G = A.T @ A
        g, v = _d(G).copy().reshape(81), np.zeros(81)
        L.f("sm_jacobi9")(H._p(g), H._p(v))
        lam = np.diag(g.reshape(9, 9)); vec = v.reshape(9, 9)
        w, q = np.linalg.eigh(G)
        assert np.allclose(np.sort(lam), w, rtol=1e-10)
        k = int(np.argmin(lam))
        assert min(np.abs(vec[:, k] - q[:, 0]).max(), np.abs(vec[:, k] + q[:, 0]).max()) < 1e-9


def test_shared_math_correction_and_hdlt_match_oracle_on_golden(L, gold):
    F = gold["F"].astype(np.float32)
    Fd = _d(F.astype(np.float64)).reshape(9)
    worst = 0.0
    for i in range(len(gold["A"])):
        a, b = _d(gold["A"][i]).copy(), _d(gold["B"][i]).copy()
        nr = L.f("sm_correct_match")(H._p(Fd), H._p(a), H._p(b))
        assert nr >= 1
        worst = max(worst, np.abs(a - gold["Ac"][i]).max(), np.abs(b - gold["Bc"][i]).max())
    assert worst < 1e-6, worst
    # corrected matches satisfy the epipolar constraint (that is what the correction is for)
    K = gold["K"].astype(np.float64)
    E = K.T @ F.astype(np.float64) @ K
    R1, R2, t = M.get_pose(E)
    KPA = (K @ np.hstack([np.eye(3), np.zeros((3, 1))])).astype(np.float32)
    KPB = (K @ np.hstack([R2, -t[:, None]])).astype(np.float32)
    for i in range(0, len(gold["A"]), 7):
        X = np.zeros(4)
        L.f("sm_hdlt")(H._p(_d(KPA.astype(np.float64))), H._p(_d(KPB.astype(np.float64))), H._p(_d(gold["A"][i])), H._p(_d(gold["B"][i])), H._p(X))
        ref = M.hdlt(KPA, KPB, gold["A"][i].astype(np.float64), gold["B"][i].astype(np.float64))
        assert np.allclose(X / X[3], ref / ref[3], rtol=1e-5, atol=1e-7)


def test_realroots_sorted_matches_oracle(L):
    rng = np.random.default_rng(4)
    for trial in range(100):
        a7 = np.poly(rng.uniform(-3, 3, 6))[::-1].copy() if trial % 2 else rng.normal(size=7)
        if trial % 5 == 0:
            a7[6] *= 1e-6                   # nearly a quintic, like the sextics of data.txt
        out = np.zeros(6)
        n = L.f("sm_realroots")(H._p(a7), H._p(out))
        ref = M.realroots(a7)
        assert n == len(ref)
        assert np.allclose(out[:n], ref, rtol=1e-8, atol=1e-10)


def test_sampson_oracle_golden_and_properties(gold):
    """oracle f_sampson (multiview.hpp:187-299): reproduces the committed fixture, is a fixed point
    of its own iteration up to float32 noise, and unit weights are a no-op.  (The reference's
    weighting -- W, not sqrt(W), of lines divided by their z -- does not lower the textbook Sampson
    cost on its own match list; it is restated as it is.)"""
    A, B = gold["A"], gold["B"]
    F = M.f_sampson(A, B)
    assert np.abs(F - gold["F_sampson"]).max() <= 1e-12 * np.abs(F).max()
    F101 = M.f_sampson(A, B, iters=101)
    assert np.abs(F101 - F).max() < 1e-4 * np.abs(F).max()
    assert np.abs(M.f_sampson(A, B, w=np.ones(len(A), np.float32)) - F).max() == 0.0

