"""Synthetic workloads: seeded images, view pairs and grid meshes.

Everything here is numpy on the host; it feeds both the CUDA path and the test
oracle with identical bytes.  No dataset is read (there is no network).
"""
import numpy as np


def make_image(w, h, seed=1000, nshapes=60, noise=2.0):
    """RGBA8 (h, w, 4): bilinear 4-corner gradient + random flat shapes + noise."""
    rng = np.random.default_rng(seed)
    corners = rng.integers(0, 256, size=(2, 2, 3)).astype(np.float32)
    u = np.linspace(0.0, 1.0, w, dtype=np.float32)[None, :, None]
    v = np.linspace(0.0, 1.0, h, dtype=np.float32)[:, None, None]
    img = (corners[0, 0] * (1 - u) * (1 - v) + corners[0, 1] * u * (1 - v)
           + corners[1, 0] * (1 - u) * v + corners[1, 1] * u * v)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    for _ in range(nshapes):
        col = rng.integers(0, 256, size=3).astype(np.float32)
        cx, cy = rng.uniform(0, w), rng.uniform(0, h)
        rad = rng.uniform(0.03, 0.18) * min(w, h)
        if rng.random() < 0.5:  # ellipse
            ax, ay = rad * rng.uniform(0.5, 1.5), rad * rng.uniform(0.5, 1.5)
            th = rng.uniform(0, np.pi)
            c, s = np.cos(th), np.sin(th)
            dx, dy = xx - cx, yy - cy
            m = ((dx * c + dy * s) / ax) ** 2 + ((-dx * s + dy * c) / ay) ** 2 <= 1.0
        else:  # convex polygon = intersection of half planes
            n = int(rng.integers(3, 7))
            ang = np.sort(rng.uniform(0, 2 * np.pi, n))
            m = np.ones((h, w), bool)
            for a in ang:
                nx, ny = np.cos(a), np.sin(a)
                m &= ((xx - cx) * nx + (yy - cy) * ny) <= rad * rng.uniform(0.6, 1.0)
        img[m] = col
    img = img + rng.normal(0.0, noise, size=img.shape).astype(np.float32)
    out = np.empty((h, w, 4), np.uint8)
    out[..., :3] = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    out[..., 3] = 255
    return out


def view_b_homography(w, h):
    """Pixel-space homography A -> B: rot 1.5 deg, scale 1.02, shift (+14,-9), persp 2e-5."""
    th = np.deg2rad(1.5)
    c, s = np.cos(th) * 1.02, np.sin(th) * 1.02
    cx, cy = 0.5 * w, 0.5 * h
    T0 = np.array([[1, 0, -cx], [0, 1, -cy], [0, 0, 1]], np.float64)
    R = np.array([[c, -s, 14.0], [s, c, -9.0], [2e-5, -2e-5, 1.0]], np.float64)
    T1 = np.array([[1, 0, cx], [0, 1, cy], [0, 0, 1]], np.float64)
    return T1 @ R @ T0


def warp_image(img, Hm):
    """View B = img resampled under homography Hm (bilinear, edge clamp)."""
    h, w = img.shape[:2]
    Hi = np.linalg.inv(Hm)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    d = Hi[2, 0] * xx + Hi[2, 1] * yy + Hi[2, 2]
    sx = (Hi[0, 0] * xx + Hi[0, 1] * yy + Hi[0, 2]) / d
    sy = (Hi[1, 0] * xx + Hi[1, 1] * yy + Hi[1, 2]) / d
    sx = np.clip(sx, 0, w - 1.001)
    sy = np.clip(sy, 0, h - 1.001)
    x0, y0 = np.floor(sx).astype(np.int64), np.floor(sy).astype(np.int64)
    fx, fy = (sx - x0)[..., None], (sy - y0)[..., None]
    f = img[..., :3].astype(np.float64)
    o = (f[y0, x0] * (1 - fx) * (1 - fy) + f[y0, x0 + 1] * fx * (1 - fy)
         + f[y0 + 1, x0] * (1 - fx) * fy + f[y0 + 1, x0 + 1] * fx * fy)
    out = np.empty_like(img)
    out[..., :3] = np.clip(np.rint(o), 0, 255).astype(np.uint8)
    out[..., 3] = 255
    return out


def make_pair(w, h, seed=1000):
    a = make_image(w, h, seed)
    return a, warp_image(a, view_b_homography(w, h))


GRIDS = {150: (15, 5), 500: (25, 10), 1000: (25, 20), 1500: (30, 25), 3000: (50, 30), 10000: (100, 50)}


def grid_mesh(nx, ny, ratio, seed=7, jitter=0.2):
    """Regular nx x ny x 2 triangle grid over [-ratio,ratio] x [-1,1].

    Points 0..3 are the domain corners in the reference's order
    (source/triangulation.hpp:44-47) because shift.cs pins indices < 4.
    Returns (points float32 [NP,2], triangles int32 [NT,4], halfedges int32 [3*NT]).
    """
    rng = np.random.default_rng(seed)
    gx = np.linspace(-ratio, ratio, nx + 1)
    gy = np.linspace(-1.0, 1.0, ny + 1)
    P = np.stack(np.meshgrid(gx, gy, indexing="ij"), -1)  # [nx+1, ny+1, 2]
    jit = rng.uniform(-jitter, jitter, size=P.shape) * np.array([2 * ratio / nx, 2.0 / ny])
    jit[0, :, :] = 0; jit[-1, :, :] = 0; jit[:, 0, :] = 0; jit[:, -1, :] = 0
    P = (P + jit).astype(np.float32)
    P[0, :, 0] = -ratio; P[-1, :, 0] = ratio; P[:, 0, 1] = -1.0; P[:, -1, 1] = 1.0
    corner_ij = [(0, 0), (0, ny), (nx, 0), (nx, ny)]
    order = list(corner_ij)
    for i in range(nx + 1):
        for j in range(ny + 1):
            if (i, j) not in corner_ij:
                order.append((i, j))
    index = {ij: n for n, ij in enumerate(order)}
    pts = np.array([P[i, j] for (i, j) in order], np.float32)
    tris = []
    for i in range(nx):
        for j in range(ny):
            a, b = index[(i, j)], index[(i, j + 1)]
            c, d = index[(i + 1, j)], index[(i + 1, j + 1)]
            tris.append((a, b, c, 0))
            tris.append((c, b, d, 0))
    tri = np.array(tris, np.int32)
    return pts, tri, halfedges_of(tri)


def halfedges_of(tri):
    """Twin table in the reference's convention: halfedge 3t+k is v_k -> v_{k+1}."""
    NT = tri.shape[0]
    he = -np.ones(3 * NT, np.int32)
    seen = {}
    for t in range(NT):
        for k in range(3):
            a, b = int(tri[t, k]), int(tri[t, (k + 1) % 3])
            if (b, a) in seen:
                o = seen.pop((b, a))
                he[3 * t + k] = o
                he[o] = 3 * t + k
            else:
                seen[(a, b)] = 3 * t + k
    return he


def mesh_for(ntri, ratio, seed=7):
    nx, ny = GRIDS[ntri]
    return grid_mesh(nx, ny, ratio, seed)
