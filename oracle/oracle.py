"""ctypes front end of oracle/liboracle.so (step_oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Never by the product.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

TRIANGULATE, WARP = 0, 1


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "step_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.tpo_dp.restype = C.c_float
        _LIB.tpo_sum_energy.restype = C.c_float
    return _LIB


def _p(a, t=None):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def dp(NT, flavour):
    return float(lib().tpo_dp(C.c_int(NT), C.c_int(flavour)))


def variant_vertices(pts, tri, ratio, W, H, flavour):
    pts, tri = _f32(pts), _i32(tri)
    NT = tri.shape[0]
    out = np.zeros((13 * NT, 3, 2), np.int32)
    lib().tpo_variant_vertices(_p(pts), _p(tri), NT, C.c_float(ratio), W, H, flavour, _p(out))
    return out


def mode0(rgba, pts, tri, ratio, flavour, ca=None):
    rgba, pts, tri = _u8(rgba), _f32(pts), _i32(tri)
    H, W = rgba.shape[:2]
    NT = tri.shape[0]
    cn = np.zeros(13 * NT, np.int32)
    if ca is None:
        ca = np.zeros((13 * NT, 4), np.int32)
    else:
        ca = _i32(ca).copy()
    lib().tpo_mode0(_p(rgba), W, H, _p(pts), _p(tri), NT, C.c_float(ratio), flavour, _p(cn), _p(ca))
    return cn, ca


def mode1(rgba, pts, tri, ratio, flavour, cn, ca, NP=None, gr=None):
    """gr: the gradient SSBO as the previous frame left it (None = all zero, a fresh tpose::init).
    Mode 1 zeroes it only for vertices some triangle references (triangle.vs:93-98); a vertex no
    triangle uses keeps its stale entry, and shift.cs goes on applying it."""
    rgba, pts, tri = _u8(rgba), _f32(pts), _i32(tri)
    H, W = rgba.shape[:2]
    NT = tri.shape[0]
    NP = pts.shape[0] if NP is None else NP
    ten = np.zeros(13 * NT, np.int32)
    if gr is None:
        gr = np.zeros((NP, 2), np.int32)
    else:
        old = _i32(gr).reshape(-1, 2)
        gr = np.zeros((max(NP, old.shape[0]), 2), np.int32)
        gr[:old.shape[0]] = old
    lib().tpo_mode1(_p(rgba), W, H, _p(pts), _p(tri), NT, C.c_float(ratio), flavour,
                    _p(_i32(cn)), _p(_i32(ca)), _p(ten), _p(gr))
    return ten, gr


def gradient(tri, ten, gr):
    tri = _i32(tri)
    gr = _i32(gr).copy()
    lib().tpo_gradient(_p(tri), tri.shape[0], _p(_i32(ten)), _p(gr))
    return gr


def shift(pts, gr, ratio, flavour):
    pts = _f32(pts).copy()
    lib().tpo_shift(_p(pts), pts.shape[0], _p(_i32(gr)), C.c_float(ratio), flavour)
    return pts


def raster_ids(pts, tri, ratio, W, H):
    pts, tri = _f32(pts), _i32(tri)
    ids = np.zeros((H, W), np.int32)
    lib().tpo_raster_ids(_p(pts), _p(tri), tri.shape[0], C.c_float(ratio), W, H, _p(ids))
    return ids


def sum_energy(ten, NT):
    return float(lib().tpo_sum_energy(_p(_i32(ten)), NT))


def set_dp_law(which):
    """-1: dp follows the flavour (the shipped shaders); 0 / 1: force the triangulate / warp cooling law."""
    lib().tpo_set_dp_law(C.c_int(which))


def resample(rgba, rw, rh):
    """The raster image of the 1/1.5 window (see tpo_resample): (rh, rw, 4) uint8."""
    rgba = _u8(rgba)
    h, w = rgba.shape[:2]
    out = np.empty((rh, rw, 4), np.uint8)
    lib().tpo_resample(_p(rgba), w, h, _p(out), rw, rh)
    return out


def step(rgba, pts, tri, ratio, flavour, colors=None, gr=None):
    """One full evaluation + update at positions `pts`.

    Returns dict(cn, ca, ten, gr, pts) where cn/ca/ten are the 13*NT rows at
    the ENTRY positions and pts are the positions after the shift -- exactly
    what the reference reads back at software/triangulate/main.cpp:201-204 /
    software/warp/main.cpp:226-229 for that frame.
    """
    tri = _i32(tri)
    NT = tri.shape[0]
    if flavour == WARP:
        ca = np.zeros((13 * NT, 4), np.int32)
        col = _i32(colors).reshape(NT, 4)
        for k in range(13):  # source/triangulation.hpp:633-641
            ca[k * NT:(k + 1) * NT] = col
        cn, _ = mode0(rgba, pts, tri, ratio, WARP, ca)
    else:
        cn, ca = mode0(rgba, pts, tri, ratio, TRIANGULATE)
    ten, gr = mode1(rgba, pts, tri, ratio, flavour, cn, ca, gr=gr)
    gr = gradient(tri, ten, gr)
    new_pts = shift(pts, gr[:_f32(pts).shape[0]], ratio, flavour)
    return dict(cn=cn, ca=ca, ten=ten, gr=gr, pts=new_pts)


def wire_pixels(pts, tri, ratio, W, H):
    """Wireframe fragments under the convention of csrc/raster.cuh::k_wire (GL leaves line fragments to the
    driver): boolean (H, W) mask of the pixels the three sides of every triangle light up."""
    V = variant_vertices(pts, tri, ratio, W, H, TRIANGULATE)[:_i32(tri).shape[0]].astype(np.int64)     # base instances: snapped corners
    mask = np.zeros((H, W), bool)
    for t in range(V.shape[0]):
        for k in range(3):
            (ax, ay), (bx, by) = V[t, k], V[t, (k + 1) % 3]
            dx, dy = bx - ax, by - ay
            if dx == 0 and dy == 0:
                continue
            xmajor = abs(dx) >= abs(dy)
            ua, va, ub, vb = (ax, ay, bx, by) if xmajor else (ay, ax, by, bx)
            if ua > ub:
                ua, va, ub, vb = ub, vb, ua, va
            du, dv = ub - ua, vb - va
            ulim, vlim = (W, H) if xmajor else (H, W)
            c0, c1 = max((ua - 128 + 255) >> 8, 0), min((ub - 128 + 255) >> 8, ulim)
            for c in range(int(c0), int(c1)):
                m = 256 * c + 128 - ua
                r = (va + (dv * m) // du) >> 8
                if 0 <= r < vlim:
                    if xmajor:
                        mask[r, c] = True
                    else:
                        mask[c, r] = True
    return mask
