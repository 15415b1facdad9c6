"""numpy restatement of the plane-fit meshing core of software/rgbd (rgbd/main.cpp:244-349 of the reference).

TEST INFRASTRUCTURE ONLY.  float32 throughout, one rounding per operation, sums in index order (the reference
iterates a std::set<size_t>).  PARITY STATUS: "parity unpinned" -- the reference program cannot be built
(librealsense, the unreleased piolib); piolib's project / unproject are restated as a pinhole model, the same
one the CUDA path uses (csrc/rgbd.cuh)."""
import numpy as np

f32 = np.float32


def mat3_colmajor(*rows9):
    """glm::mat3(a0..a8): the nine scalars fill the matrix column by column; returns m[col][row]."""
    return np.array(rows9, f32).reshape(3, 3)


def transpose(m):
    return m.T.copy()


def inverse(m):
    """glm's cofactor inverse of a mat3 (m[col][row]), operation for operation."""
    c0 = f32(m[1][1] * m[2][2]) - f32(m[2][1] * m[1][2])
    c1 = f32(m[0][1] * m[2][2]) - f32(m[2][1] * m[0][2])
    c2 = f32(m[0][1] * m[1][2]) - f32(m[1][1] * m[0][2])
    det = f32(f32(f32(m[0][0] * c0) - f32(m[1][0] * c1)) + f32(m[2][0] * c2))
    inv = f32(1.0) / det
    r = np.zeros((3, 3), f32)
    d = lambda a, b, c, e: f32(f32(a * b) - f32(c * e))
    r[0][0] = +d(m[1][1], m[2][2], m[2][1], m[1][2]) * inv
    r[1][0] = -d(m[1][0], m[2][2], m[2][0], m[1][2]) * inv
    r[2][0] = +d(m[1][0], m[2][1], m[2][0], m[1][1]) * inv
    r[0][1] = -d(m[0][1], m[2][2], m[2][1], m[0][2]) * inv
    r[1][1] = +d(m[0][0], m[2][2], m[2][0], m[0][2]) * inv
    r[2][1] = -d(m[0][0], m[2][1], m[2][0], m[0][1]) * inv
    r[0][2] = +d(m[0][1], m[1][2], m[1][1], m[0][2]) * inv
    r[1][2] = -d(m[0][0], m[1][2], m[1][0], m[0][2]) * inv
    r[2][2] = +d(m[0][0], m[1][1], m[1][0], m[0][1]) * inv
    return r


def pixelmap(ratio, width=960.0, height=540.0):
    """rgbd/main.cpp:94-98."""
    a = f32(f32(f32(1.0) / f32(ratio)) * f32(0.5)) * f32(width)
    return transpose(mat3_colmajor(a, 0, f32(0.5) * f32(width), 0, f32(f32(1.0) * f32(0.5)) * f32(height), f32(0.5) * f32(height), 0, 0, 1))


def apply(m, p):
    """(mat3 * vec3(p, 1)).xy for an (N, 2) array."""
    p = np.asarray(p, f32).reshape(-1, 2)
    x = (f32(m[0][0]) * p[:, 0] + f32(m[1][0]) * p[:, 1]).astype(f32) + f32(m[2][0]) * f32(1.0)
    y = (f32(m[0][1]) * p[:, 0] + f32(m[1][1]) * p[:, 1]).astype(f32) + f32(m[2][1]) * f32(1.0)
    return np.stack([x, y], 1).astype(f32)


def project(points, intr):
    fx, fy, ppx, ppy = [f32(v) for v in intr]
    p = np.asarray(points, f32)
    return np.stack([(fx * p[:, 0] / p[:, 2]).astype(f32) + ppx, (fy * p[:, 1] / p[:, 2]).astype(f32) + ppy], 1).astype(f32)


def _intriangle(P, a, b, c):
    """tpose::intriangle (utility.hpp:43-53) for an (N, 2) array of points against one triangle."""
    for u, v in ((a, b), (b, c), (c, a)):
        if f32(f32((u[0] - v[0]) * (u[0] - v[0])) + f32((u[1] - v[1]) * (u[1] - v[1]))) == 0:
            return np.zeros(len(P), bool)
    m01, m02, m11, m12, m21, m22 = a[0], a[1], b[0], b[1], c[0], c[1]
    c0 = f32(f32(m11 * m22) - f32(m21 * m12)); c1 = f32(f32(m01 * m22) - f32(m21 * m02)); c2 = f32(f32(m01 * m12) - f32(m11 * m02))
    det = f32(f32(c0 - c1) + c2)
    if abs(float(det)) < 1e-8:
        return np.zeros(len(P), bool)
    inv = f32(1.0) / det
    r00, r10, r20 = c0 * inv, -(m22 - m12) * inv, (m21 - m11) * inv
    r01, r11, r21 = -c1 * inv, (m22 - m02) * inv, -(m21 - m01) * inv
    r02, r12, r22 = c2 * inv, -(m12 - m02) * inv, (m11 - m01) * inv
    x, y = P[:, 0], P[:, 1]
    sx = ((r00 + (r10 * x).astype(f32)).astype(f32) + (r20 * y).astype(f32)).astype(f32)
    sy = ((r01 + (r11 * x).astype(f32)).astype(f32) + (r21 * y).astype(f32)).astype(f32)
    sz = ((r02 + (r12 * x).astype(f32)).astype(f32) + (r22 * y).astype(f32)).astype(f32)
    return ~((sx <= 0) | (sx >= 1) | (sy <= 0) | (sy >= 1) | (sz <= 0) | (sz >= 1))


def _seq(v):
    return np.cumsum(np.asarray(v, f32), dtype=f32)[-1]


def _unproject(u, intr, mean, n):
    fx, fy, ppx, ppy = [f32(v) for v in intr]
    dx, dy = f32(u[0] - ppx) / fx, f32(u[1] - ppy) / fy
    num = f32(f32(f32(n[0] * mean[0]) + f32(n[1] * mean[1])) + f32(n[2] * mean[2]))
    den = f32(f32(f32(n[0] * dx) + f32(n[1] * dy)) + n[2])
    s = num / den
    return np.array([s * dx, s * dy, s, 1.0], f32)


def planefit(tri, pts, T, pixels, points, normals, intr):
    """rgbd/main.cpp:244-339.  Returns (nvertices (NP, 4), nring (NP,), points gathered per triangle (NT,))."""
    tri = np.asarray(tri, np.int32); pts = np.asarray(pts, f32)
    pixels = np.asarray(pixels, f32).reshape(-1, 2); points = np.asarray(points, f32).reshape(-1, 4); normals = np.asarray(normals, f32).reshape(-1, 4)
    TB = inverse(T)
    back = apply(TB, pixels)
    NP = len(pts)
    nv = np.zeros((NP, 4), f32); ring = np.zeros(NP, f32); counts = np.zeros(len(tri), np.int32)
    for t in range(len(tri)):
        a, b, c = pts[tri[t, 0]], pts[tri[t, 1]], pts[tri[t, 2]]
        lo = apply(T, [np.minimum(np.minimum(a, b), c)])[0]; hi = apply(T, [np.maximum(np.maximum(a, b), c)])[0]
        box = ~((pixels[:, 0] < lo[0]) | (pixels[:, 0] > hi[0]) | (pixels[:, 1] < lo[1]) | (pixels[:, 1] > hi[1]))
        idx = np.flatnonzero(box)
        if len(idx):
            idx = idx[_intriangle(back[idx], a, b, c)]
        counts[t] = len(idx)
        if not len(idx):
            continue
        n = f32(len(idx))
        mean = np.array([_seq(points[idx, k]) for k in range(3)], f32) / n
        nrm = np.array([_seq(normals[idx, k]) for k in range(3)], f32) / n
        d = f32(f32(f32(nrm[0] * nrm[0]) + f32(nrm[1] * nrm[1])) + f32(nrm[2] * nrm[2]))
        nrm = (nrm * (f32(1.0) / np.sqrt(d))).astype(f32)
        corners = apply(T, np.stack([a, b, c]))
        for j in range(3):
            nv[tri[t, j]] = (nv[tri[t, j]] + _unproject(corners[j], intr, mean, nrm)).astype(f32)
            ring[tri[t, j]] += f32(1.0)
    with np.errstate(invalid="ignore", divide="ignore"):
        nv = (nv / ring[:, None]).astype(f32)
    return nv, ring, counts
