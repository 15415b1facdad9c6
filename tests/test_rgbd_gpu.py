"""software/rgbd's plane-fit meshing core (rgbd/main.cpp:244-349) on the device against its literal numpy
restatement (oracle/rgbd_oracle.py).  The scene is synthetic: a few tilted planes seen by a pinhole camera,
sampled on a jittered pixel grid with noise -- the reference's input (a RealSense capture) is not available."""
import numpy as np
import pytest

from oracle import rgbd_oracle as R

pytestmark = pytest.mark.gpu


def _scene(rng, n, intr, w=960, h=540):
    fx, fy, ppx, ppy = intr
    u = rng.uniform(0, w, n); v = rng.uniform(0, h, n)
    dx, dy = (u - ppx) / fx, (v - ppy) / fy
    # three planes z = z0 + a x + b y, picked by image region
    region = (u > w / 3).astype(int) + (u > 2 * w / 3).astype(int)
    z0 = np.array([2.0, 2.6, 3.1])[region]; a = np.array([0.2, -0.3, 0.0])[region]; b = np.array([0.1, 0.0, -0.25])[region]
    z = z0 / (1.0 - a * dx - b * dy) + rng.normal(0, 2e-3, n)
    P = np.stack([dx * z, dy * z, z, np.ones(n)], 1).astype(np.float32)
    nrm = np.stack([-a, -b, np.ones(n)], 1); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    N = np.concatenate([nrm + rng.normal(0, 0.02, nrm.shape), np.ones((n, 1))], 1).astype(np.float32)
    return P, N


@pytest.mark.parametrize("ntri,npoints", [(150, 20000), (1000, 60000)])
def test_planefit_matches_oracle(cuda_pkg, synth, ntri, npoints):
    rng = np.random.default_rng(3)
    ratio = 960 / 540
    intr = np.array([640.0, 642.0, 481.5, 268.25], np.float32)
    pts, tri, _ = synth.mesh_for(ntri, ratio)
    P, N = _scene(rng, npoints, intr)
    px = R.project(P, intr)
    T = R.pixelmap(ratio)
    TB = R.inverse(T)
    V, ring, cnt = cuda_pkg.rgbd_planefit(tri, pts, T, TB, px, P, N, intr)
    rV, rring, rcnt = R.planefit(tri, pts, T, px, P, N, intr)
    assert np.array_equal(cnt, rcnt) and cnt.sum() > 0.9 * npoints       # nearly every point lies strictly inside one triangle
    assert np.array_equal(ring, rring)
    ok = rring > 0
    assert ok.sum() > 0.9 * len(pts) and np.isnan(V[~ok]).all() == np.isnan(rV[~ok]).all()
    assert np.array_equal(V[ok].view(np.uint32), rV[ok].view(np.uint32)), np.abs(V[ok] - rV[ok]).max()
    # and the mesh is sensible: vertices inside the middle plane's region sit on that plane
    mid = ok & (np.abs(pts[:, 0]) < ratio / 4)
    z_expected = 2.6 / (1.0 + 0.3 * V[mid, 0] / V[mid, 2])
    assert np.abs(V[mid, 2] - z_expected).max() < 0.05


def test_planefit_without_points_gives_nan_vertices(cuda_pkg, synth):
    ratio = 960 / 540
    pts, tri, _ = synth.mesh_for(150, ratio)
    T = R.pixelmap(ratio)
    V, ring, cnt = cuda_pkg.rgbd_planefit(tri, pts, T, R.inverse(T), np.zeros((0, 2)), np.zeros((0, 4)), np.zeros((0, 4)), [600, 600, 480, 270])
    assert not cnt.any() and not ring.any() and np.isnan(V).all()       # 0 / 0, as coded (rgbd/main.cpp:338-339)
