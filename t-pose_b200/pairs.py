"""View pairs through the warp level loop + F + DLT in one process per GPU (BASELINE.json configs 3-5):
Python glue over libtpose_host.so's tph_pairs_* entry points (include/tpose/pairs).  Pairs are
independent problems; a PairSet batches all of its pairs into shared launches on one device, and
shard.py spreads PairSets over ranks.  Everything numeric happens in the CUDA library."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_SO = os.environ.get("TPOSE_HOST_LIB") or os.path.join(_HERE, "libtpose_host.so")   # override: development experiments
_host = None


def host():
    global _host
    if _host is None:
        if not os.path.exists(HOST_SO):
            raise RuntimeError(f"{HOST_SO} is not built; run __graft_entry__.build()")
        _host = C.CDLL(HOST_SO, mode=C.RTLD_LOCAL)
        _host.tph_pairs_new.restype = C.c_void_p
        _host.tph_tri_new.restype = C.c_void_p
    return _host


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def write_stack(path, levels, ratio):
    """Writes a .tri stack (source/io.hpp:170-220) with one record per (pts, tri, halfedges, colours) level,
    coarse to fine, through the product's own tpose::io::write."""
    L = host()
    if os.path.exists(path):
        os.remove(path)
    L.tph_set_ratio(C.c_float(ratio))
    h = C.c_void_p(L.tph_tri_new())
    for (pts, tri, he, col) in levels:
        pts = np.ascontiguousarray(pts, np.float32); tri = np.ascontiguousarray(tri, np.int32)
        he = np.ascontiguousarray(he, np.int32); col = np.ascontiguousarray(col, np.int32)
        L.tph_tri_set(h, tri.shape[0], _p(tri), _p(he), _p(col), pts.shape[0], _p(pts), _p(pts))
        L.tph_io_write(h, path.encode())
    L.tph_io_close(h)
    L.tph_tri_free(h)


def grid_stack(pkg, synth, path, img, ratio, sizes=(150, 500, 1500, 3000), device=0):
    """A synthetic stack: jittered grids of the given sizes, each triangle coloured with the mean colour the
    device's colour pass finds under it (what software/triangulate exports, main.cpp:221-224)."""
    levels = []
    ctx = pkg.Context(device=device, max_triangles=max(sizes), max_points=max(sizes))
    ctx.set_image(0, img)
    for n in sizes:
        pts, tri, he = synth.mesh_for(n, ratio)
        ctx.upload(pts, tri, ratio)
        ctx.colors(0, pkg.TRIANGULATE)
        _, cn, ca = ctx.retrieve_base()
        col = np.ones((tri.shape[0], 4), np.int32)
        col[:, :3] = ca[:, :3] // np.maximum(cn[:, None], 1)
        levels.append((pts, tri, he, col))
    ctx.close()
    write_stack(path, levels, ratio)
    return levels


class PairSet:
    """Independent view pairs on one device, advanced together."""

    def __init__(self, device=0):
        L = host()
        L.tph_init(C.c_int(device))
        self.device = device
        self._h = C.c_void_p(L.tph_pairs_new(C.c_int(device)))
        self.n = 0

    def add(self, img_a, img_b, tri_a, tri_b, raster=None, max_triangles=4096, max_points=4096):
        a = np.ascontiguousarray(img_a, np.uint8); b = np.ascontiguousarray(img_b, np.uint8)
        h, w = a.shape[:2]
        assert b.shape == a.shape and a.shape[2] == 4
        rw, rh = (w, h) if raster is None else raster
        for p in (tri_a + ".warp", tri_b + ".warp"):
            if os.path.exists(p):
                os.remove(p)
        i = host().tph_pairs_add(self._h, _p(a), _p(b), w, h, rw, rh, tri_a.encode(), tri_b.encode(), max_triangles, max_points)
        if i < 0:
            raise RuntimeError(f"tph_pairs_add -> {i}")
        self.n += 1
        return i

    def run(self, tol=1e-6, level_cap=-1, mview=True):
        """Warp level loop of every pair (one batched launch sequence), then F_8Point + triangulate(F,K,A,B) for all
        pairs in two launches.  Returns (frames, levels) per pair."""
        fr = (C.c_long * self.n)(); lv = (C.c_long * self.n)()
        host().tph_pairs_run(self._h, C.c_double(tol), C.c_long(level_cap), int(mview), fr, lv)
        return list(fr), list(lv)

    @staticmethod
    def timing():
        """Where the wall time of the last run() went: seconds in tpose_cuda_batch_run, in the host's level switches
        (reversewarp, stack I/O, uploads) and in F + DLT; batch_run calls, frames asked of / run by them."""
        t = (C.c_double * 6)()
        host().tph_pairs_timing(t)
        return dict(batch_run_s=t[0], level_switch_s=t[1], mview_s=t[2], runs=int(t[3]), frames_asked=int(t[4]), frames_run=int(t[5]))

    def result(self, i):
        """dict(ptsA, optsA, ptsB, optsB (finest level's vertex buffers), F (3,3), X (matches,4))."""
        sz = (C.c_int * 3)()
        host().tph_pairs_sizes(self._h, i, sz)
        npa, npb, nm = sz[0], sz[1], sz[2]
        out = dict(ptsA=np.zeros((npa, 2), np.float32), optsA=np.zeros((npa, 2), np.float32), ptsB=np.zeros((npb, 2), np.float32),
                   optsB=np.zeros((npb, 2), np.float32), F=np.zeros((3, 3), np.float32), X=np.zeros((nm, 4), np.float32))
        host().tph_pairs_result(self._h, i, _p(out["ptsA"]), _p(out["optsA"]), _p(out["ptsB"]), _p(out["optsB"]), _p(out["F"]),
                                _p(out["X"]) if nm else None)
        return out

    def packed(self, i):
        """Pair i's results as one float32 vector (what a rank contributes to the final gather)."""
        r = self.result(i)
        return np.concatenate([r[k].ravel() for k in ("ptsA", "optsA", "ptsB", "optsB", "F", "X")]).astype(np.float32)

    def close(self):
        if self._h:
            host().tph_pairs_free(self._h)
            self._h = None
