"""ctypes access to the two host-side libraries the topology / IO parity tests compare:

  Ref  -- oracle/_ref/libtpose_ref.so: the reference's own headers, unmodified (test authority)
  Ours -- t-pose_b200/libtpose_host.so: the product's include/tpose/* host layer

Both export the same functions under different prefixes (ref_ / tph_)."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libtpose_ref.so")
HOST_SO = os.path.join(ROOT, "t-pose_b200", "libtpose_host.so")


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class HostLib:
    def __init__(self, path, prefix):
        self.lib = C.CDLL(path, mode=C.RTLD_LOCAL)
        self.prefix = prefix
        for name, res in (("tri_new", C.c_void_p), ("get_ratio", C.c_float), ("tri_angle", C.c_float),
                          ("tri_hlength", C.c_float), ("geterr", C.c_float), ("gettoterr", C.c_float),
                          ("get_toterr", C.c_float), ("get_relerr", C.c_float), ("horner", C.c_double)):
            getattr(self.lib, prefix + name).restype = res

    def f(self, name):
        return getattr(self.lib, self.prefix + name)

    def set_ratio(self, r):
        self.f("set_ratio")(C.c_float(r))

    def get_ratio(self):
        return self.f("get_ratio")()


class Tri:
    """One triangulation object inside one of the libraries."""

    def __init__(self, L):
        self.L = L
        self.h = C.c_void_p(L.f("tri_new")())

    def close(self):
        if self.h:
            self.L.f("io_close")(self.h)
            self.L.f("tri_free")(self.h)
            self.h = None

    def counts(self):
        nt, np_ = C.c_int(), C.c_int()
        self.L.f("tri_counts")(self.h, C.byref(nt), C.byref(np_))
        return nt.value, np_.value

    def sizes(self):
        a = [C.c_int() for _ in range(4)]
        self.L.f("tri_sizes")(self.h, *[C.byref(x) for x in a])
        return tuple(x.value for x in a)

    def get(self):
        ntri, nhe, npts, nopts = self.sizes()
        tri = np.zeros((ntri, 4), np.int32)
        he = np.zeros(nhe, np.int32)
        col = np.zeros((ntri, 4), np.int32)
        pts = np.zeros((npts, 2), np.float32)
        opts = np.zeros((nopts, 2), np.float32)
        self.L.f("tri_get")(self.h, _p(tri), _p(he), _p(col), _p(pts), _p(opts))
        return dict(tri=tri, he=he, col=col, pts=pts, opts=opts, counts=self.counts())

    def set(self, tri, he, pts, opts=None, col=None):
        tri = np.ascontiguousarray(tri, np.int32); he = np.ascontiguousarray(he, np.int32)
        pts = np.ascontiguousarray(pts, np.float32)
        opts = None if opts is None else np.ascontiguousarray(opts, np.float32)
        col = None if col is None else np.ascontiguousarray(col, np.int32)
        self.L.f("tri_set")(self.h, tri.shape[0], _p(tri), _p(he), _p(col), pts.shape[0], _p(pts), _p(opts))

    def set_points(self, pts):
        pts = np.ascontiguousarray(pts, np.float32)
        self.L.f("tri_set_points")(self.h, _p(pts))

    def origin_from_points(self):
        self.L.f("tri_origin_from_points")(self.h)

    def angle(self, ha): return self.L.f("tri_angle")(self.h, ha)
    def hlength(self, ha): return self.L.f("tri_hlength")(self.h, ha)
    def boundary(self, t): return self.L.f("tri_boundary")(self.h, t)
    def eraset(self, t, adjust=True): return self.L.f("tri_eraset")(self.h, t, int(adjust))
    def erasep(self, p): return self.L.f("tri_erasep")(self.h, p)
    def prune(self, t): return self.L.f("tri_prune")(self.h, t)
    def flip(self, ha, minangle=None):
        if minangle is None:
            return self.L.f("tri_flip_default")(self.h, ha)
        return self.L.f("tri_flip")(self.h, ha, C.c_float(minangle))
    def collapse(self, ha): return self.L.f("tri_collapse")(self.h, ha)
    def split(self, t): return self.L.f("tri_split")(self.h, t)
    def optimize(self): return self.L.f("tri_optimize")(self.h)

    def warp(self, pts, reverse=False):
        pts = np.ascontiguousarray(pts, np.float32).copy()
        self.L.f("tri_reversewarp" if reverse else "tri_warp")(self.h, _p(pts), pts.shape[0])
        return pts

    def write(self, path): self.L.f("io_write")(self.h, path.encode())
    def read(self, path, dowarp=False): return self.L.f("io_read")(self.h, path.encode(), int(dowarp))
    def io_close(self): self.L.f("io_close")(self.h)

    def geterr(self): return self.L.f("geterr")(self.h)
    def gettoterr(self): return self.L.f("gettoterr")(self.h)
    def maxerrid(self): return self.L.f("maxerrid")(self.h)


_cache = {}


def ref():
    if "ref" not in _cache:
        _cache["ref"] = HostLib(REF_SO, "ref_")
        _cache["ref"].f("init")()
    return _cache["ref"]


def ours():
    if "ours" not in _cache:
        _cache["ours"] = HostLib(HOST_SO, "tph_")
    return _cache["ours"]


def same_state(a, b):
    """Byte-for-byte comparison of two Tri.get() dictionaries."""
    if a["counts"] != b["counts"]:
        return False, "counts"
    for k in ("tri", "he", "col"):
        if a[k].shape != b[k].shape or not np.array_equal(a[k], b[k]):
            return False, k
    for k in ("pts", "opts"):
        if a[k].shape != b[k].shape or not np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)):
            return False, k
    return True, ""
