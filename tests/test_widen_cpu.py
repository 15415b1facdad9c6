"""CPU checks of the oracles behind the widened rows (match selection, RGB-D plane fit): the pieces of them the
reference itself implements -- tpose::intriangle (utility.hpp:43-53) and triangulation::boundary(vec2)
(triangulation.hpp:132-140) -- against the reference's own headers compiled unmodified (oracle/_ref)."""
import ctypes as C
import os

import numpy as np
import pytest

import hostlib as H
from oracle import mview_oracle as M
from oracle import rgbd_oracle as R

pytestmark = pytest.mark.skipif(not os.path.exists(H.REF_SO), reason="oracle/_ref not built (needs /root/reference at build time)")


def test_rgbd_oracle_intriangle_is_the_reference_intriangle():
    L = H.ref()
    rng = np.random.default_rng(2)
    f = np.float32
    bad = 0
    for trial in range(300):
        a, b, c = (rng.uniform(-1, 1, 2).astype(f) for _ in range(3))
        if trial % 10 == 0:
            c = a.copy()                                  # zero-length side
        if trial % 10 == 1:
            c = (a + f(0.5) * (b - a)).astype(f)          # collinear
        P = rng.uniform(-1.2, 1.2, (40, 2)).astype(f)
        P[0], P[1], P[2] = a, (a + b) / f(2), (a + b + c) / f(3)     # a corner, an edge midpoint, the centroid
        got = R._intriangle(P, a, b, c)
        verts = np.ascontiguousarray(np.stack([a, b, c]), f)
        t4 = np.array([0, 1, 2, 0], np.int32)
        for i in range(len(P)):
            want = L.f("intriangle")(C.c_float(float(P[i, 0])), C.c_float(float(P[i, 1])), t4.ctypes.data_as(C.c_void_p),
                                     verts.ctypes.data_as(C.c_void_p), 3)
            bad += int(bool(want) != bool(got[i]))
    assert bad == 0


def test_select_matches_boundary_and_distance_follow_the_reference_types():
    L = H.ref()
    L.set_ratio(1.5)
    f = np.float32
    pts = np.array([[-1.5, 0.0], [1.5, 0.2], [0.1, -1.0], [0.3, 1.0], [0.0, 0.0], [1.4999999, 0.5], [-0.2, 0.99999994]], f)
    zero = np.zeros_like(pts)
    # every vertex is its own match with distance 0; only boundary / central-half tests decide
    X, Y, W, D = M.select_matches(pts, pts, pts, zero[:0], zero[:0], zero[:0], 1.5, max_dist=1.0)
    kept = []
    for p in pts:
        on_boundary = L.f("point_boundary")(C.c_float(float(p[0])), C.c_float(float(p[1])))
        central = abs(float(p[0])) <= 0.75 and abs(float(p[1])) <= 0.5
        kept.append((not on_boundary) and central)
    assert len(X) == sum(kept) and np.isinf(W).all() and not D.any()
    cx, cy = f(0.5) / f(1.5), f(-0.5) / f(1.5)
    want = np.stack([cx * pts[kept, 0], cy * pts[kept, 1]], 1).astype(f)
    assert np.array_equal(X, want) and np.array_equal(Y, want)


def test_rgbd_pixelmap_and_its_inverse():
    T = R.pixelmap(960 / 540)
    TB = R.inverse(T)
    p = np.array([[0.3, -0.4], [-1.7, 0.9]], np.float32)
    assert np.allclose(R.apply(T, p), [[0.3 * 270 + 480, -0.4 * 270 + 270], [-1.7 * 270 + 480, 0.9 * 270 + 270]], atol=1e-3)
    assert np.allclose(R.apply(TB, R.apply(T, p)), p, atol=1e-5)
