"""Exact-integer numpy model of the ALGORITHM the CUDA kernels use
(edge integrals over a row-prefix moment table + closed-form energy).

Test helper only: it lets the CPU suite prove, without a GPU, that the
span/edge-integral formulation and the moment closed form are identical to the
literal per-fragment oracle (oracle/step_oracle.c).  The product never imports it.
"""
import numpy as np


def moment_table(rgba):
    """Row-exclusive prefix sums P[r, c, :] over pixels [0, c) of row r.

    moments: 0 count, 1..3 sum r,g,b, 4 sum |rgb|^2, 5 count of odd (r+g+b)."""
    h, w = rgba.shape[:2]
    b = rgba[..., :3].astype(np.int64)
    m = np.zeros((h, w, 6), np.int64)
    m[..., 0] = 1
    m[..., 1:4] = b
    m[..., 4] = (b * b).sum(-1)
    m[..., 5] = b.sum(-1) & 1
    P = np.zeros((h, w + 1, 6), np.int64)
    P[:, 1:] = np.cumsum(m, axis=1)
    return P


def edge_rows(ya, yb, H):
    """Rows r whose centre 256r+128 lies in [min(ya,yb), max(ya,yb)), clipped to [0,H)."""
    lo, hi = min(ya, yb), max(ya, yb)
    r0 = -((-(lo - 128)) // 256)  # ceil((lo-128)/256)
    r1 = -((-(hi - 128)) // 256)  # first row with centre >= hi
    return max(r0, 0), min(r1, H)


def edge_cross(xa, ya, xb, yb, r, W):
    """Number of pixel centres of row r strictly left of the edge, clipped to [0,W]."""
    if ya > yb:
        xa, ya, xb, yb = xb, yb, xa, ya
    dy = yb - ya
    yc = 256 * r + 128
    num = xa * dy + (xb - xa) * (yc - ya) - 128 * dy
    den = 256 * dy
    c = -((-num) // den)  # ceil(num/den)
    return min(max(c, 0), W)


def edge_integral(P, xa, ya, xb, yb):
    H, W = P.shape[0], P.shape[1] - 1
    acc = np.zeros(6, np.int64)
    if ya == yb:
        return acc
    r0, r1 = edge_rows(ya, yb, H)
    for r in range(r0, r1):
        acc += P[r, edge_cross(xa, ya, xb, yb, r, W)]
    return acc


def triangle_moments(P, fx):
    """fx: int [3,2] snapped vertices (y down).  Returns the 6 moments of the covered set."""
    tot = np.zeros(6, np.int64)
    for j in range(3):
        (xa, ya), (xb, yb) = fx[j], fx[(j + 1) % 3]
        e = edge_integral(P, int(xa), int(ya), int(xb), int(yb))
        if yb > ya:
            tot += e
        elif yb < ya:
            tot -= e
    if tot[0] < 0:
        tot = -tot
    return tot


def energy_closed_form(mom, mean):
    """sum over covered pixels of trunc(0.5*|b - mean|^2), exact, from the six moments."""
    n, S, S2, nodd = int(mom[0]), [int(x) for x in mom[1:4]], int(mom[4]), int(mom[5])
    m = [int(x) for x in mean]
    mism = nodd if (sum(m) & 1) == 0 else n - nodd
    num = S2 - 2 * sum(a * b for a, b in zip(m, S)) + n * sum(a * a for a in m) - mism
    assert num % 2 == 0
    return num // 2


def wrap32(x):
    return int((int(x) + 2 ** 31) % 2 ** 32 - 2 ** 31)
