"""tpose_b200 -- Python front end (ctypes) of libtpose_cuda.so, the sm_100a implementation of
t-pose's optimisation step.  The directory name carries a hyphen, so load it with
`__graft_entry__.load_package()` (importlib) rather than `import`.

This module is glue only: every number comes from the CUDA library.  If the library is missing
or there is no CUDA device it raises -- there is no CPU fallback.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TPOSE_CUDA_LIB") or os.path.join(_HERE, "libtpose_cuda.so")   # override: development experiments

TRIANGULATE, WARP = 0, 1
IMAGE_A, IMAGE_B = 0, 1
TRIG_CONVERGED, TRIG_MAXSTEPS, TRIG_TOPOLOGY = 1, 2, 4
(BUF_POINTS, BUF_INDEX, BUF_COLACC, BUF_COLNUM, BUF_TENERGY, BUF_PENERGY, BUF_GRADIENT,
 BUF_NRING) = range(8)
_BUF_DTYPE = {0: (np.float32, 2), 1: (np.int32, 4), 2: (np.int32, 4), 3: (np.int32, 1), 4: (np.int32, 1),
              5: (np.int32, 1), 6: (np.int32, 2), 7: (np.int32, 1)}

_lib = None


class TposeCudaError(RuntimeError):
    pass


def build(verbose=False):
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", _HERE, "libtpose_cuda.so"], stdout=out, stderr=out)
    return LIB_PATH


EXPORTS = [
    "tpose_cuda_create", "tpose_cuda_destroy", "tpose_cuda_last_error", "tpose_cuda_version",
    "tpose_cuda_set_image", "tpose_cuda_set_image_device", "tpose_cuda_upload",
    "tpose_cuda_upload_points", "tpose_cuda_colors", "tpose_cuda_energy", "tpose_cuda_shift",
    "tpose_cuda_step", "tpose_cuda_frame", "tpose_cuda_run", "tpose_cuda_set_toterr", "tpose_cuda_get_toterr",
    "tpose_cuda_retrieve", "tpose_cuda_retrieve_base", "tpose_cuda_retrieve_gradient",
    "tpose_cuda_raster_ids", "tpose_cuda_render", "tpose_cuda_render_wire", "tpose_cuda_last_elapsed_ms", "tpose_cuda_last_kernel_ms", "tpose_cuda_last_setup_ms", "tpose_cuda_kernel_path", "tpose_cuda_set_kernel_timing", "tpose_cuda_set_step_scale", "tpose_cuda_sync",
    "tpose_cuda_launch_count", "tpose_cuda_host_alloc", "tpose_cuda_host_free", "tpose_cuda_flush_l2",
    "tpose_cuda_f8point", "tpose_cuda_fsampson", "tpose_cuda_rgbd_planefit", "tpose_cuda_select_matches", "tpose_cuda_carry", "tpose_cuda_geterr", "tpose_cuda_triangulate", "tpose_cuda_hdlt", "tpose_cuda_mview_last_error",
    "tpose_cuda_mview_last_ms",
    "tpose_cuda_watch", "tpose_cuda_watch_mask", "tpose_cuda_set_history", "tpose_cuda_eval_list", "tpose_cuda_history_frame", "tpose_cuda_history_restore_gradient", "tpose_cuda_read", "tpose_cuda_write", "tpose_cuda_batch_step", "tpose_cuda_batch_run",
]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TposeCudaError(f"{LIB_PATH} is not built; run __graft_entry__.build() "
                                 "(there is no CPU fallback)")
        _lib = C.CDLL(LIB_PATH)
        _lib.tpose_cuda_last_error.restype = C.c_char_p
        _lib.tpose_cuda_last_error.argtypes = [C.c_void_p]
        _lib.tpose_cuda_version.restype = C.c_char_p
        _lib.tpose_cuda_launch_count.restype = C.c_int64
        _lib.tpose_cuda_launch_count.argtypes = [C.c_void_p]
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Context:
    """One device + one stream; mirrors the call sequence of the reference's lambdas."""

    def __init__(self, device=0, max_triangles=40000, max_points=24000, stream=None):
        self._h = C.c_void_p()
        rc = lib().tpose_cuda_create(C.c_int(device), C.c_int(max_triangles), C.c_int(max_points),
                                     C.c_void_p(stream or 0), C.byref(self._h))
        if rc != 0:
            msg = lib().tpose_cuda_last_error(None).decode()
            self._h = None
            raise TposeCudaError(f"tpose_cuda_create -> {rc}: {msg}")
        self.NT = self.NP = 0
        self.shape = [None, None]

    def _ck(self, rc):
        if rc != 0:
            raise TposeCudaError(f"status {rc}: {lib().tpose_cuda_last_error(self._h).decode()}")

    def close(self):
        if self._h:
            lib().tpose_cuda_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- inputs
    def set_image(self, slot, rgba, raster=None):
        rgba = np.ascontiguousarray(rgba, dtype=np.uint8)
        h, w = rgba.shape[:2]
        assert rgba.shape[2] == 4
        rw, rh = (w, h) if raster is None else raster
        self._ck(lib().tpose_cuda_set_image(self._h, slot, _ptr(rgba), w, h, rw, rh))
        self.shape[slot] = (rh, rw)

    def set_image_device(self, slot, dev_ptr, w, h, raster=None):
        rw, rh = (w, h) if raster is None else raster
        self._ck(lib().tpose_cuda_set_image_device(self._h, slot, C.c_void_p(dev_ptr), w, h, rw, rh))
        self.shape[slot] = (rh, rw)

    def upload(self, pts, tri, ratio, colors=None):
        pts = np.ascontiguousarray(pts, dtype=np.float32)
        tri = np.ascontiguousarray(tri, dtype=np.int32)
        col = None if colors is None else np.ascontiguousarray(colors, dtype=np.int32)
        assert tri.ndim == 2 and tri.shape[1] == 4 and pts.shape[1] == 2
        self._ck(lib().tpose_cuda_upload(self._h, _ptr(pts), pts.shape[0], _ptr(tri), tri.shape[0],
                                         _ptr(col), C.c_float(ratio)))
        self.NP, self.NT = pts.shape[0], tri.shape[0]

    def upload_points(self, pts):
        pts = np.ascontiguousarray(pts, dtype=np.float32)
        self._ck(lib().tpose_cuda_upload_points(self._h, _ptr(pts), pts.shape[0]))

    # -- the lambdas
    def colors(self, slot, flavour):
        self._ck(lib().tpose_cuda_colors(self._h, slot, flavour))

    def energy(self, slot, flavour):
        self._ck(lib().tpose_cuda_energy(self._h, slot, flavour))

    def shift(self, flavour):
        self._ck(lib().tpose_cuda_shift(self._h, flavour))

    # -- fused
    def step(self, slot, flavour, nsteps=1):
        self._ck(lib().tpose_cuda_step(self._h, slot, flavour, nsteps))

    def frame(self, slot, flavour, pts_in=None, ten=None, cn=None, ca=None, pts_out=None):
        """One reference frame: optional pointbuf fill, step, read-backs into caller-owned arrays."""
        self._ck(lib().tpose_cuda_frame(self._h, slot, flavour, _ptr(pts_in), _ptr(ten), _ptr(cn), _ptr(ca), _ptr(pts_out)))

    def bind_frame(self, slot, flavour, pts_in=None, ten=None, cn=None, ca=None, pts_out=None):
        """frame() with its arguments converted once: returns a zero-argument callable for loops that reuse the
        same buffers every frame (the arrays must stay alive)."""
        fn, h = lib().tpose_cuda_frame, self._h
        args = (h, C.c_int(slot), C.c_int(flavour), _ptr(pts_in), _ptr(ten), _ptr(cn), _ptr(ca), _ptr(pts_out))
        keep = (pts_in, ten, cn, ca, pts_out)

        def call(_keep=keep):
            rc = fn(*args)
            if rc != 0:
                self._ck(rc)
        return call

    def run(self, slot, flavour, max_steps, rel_tol):
        n, f, e = C.c_int(), C.c_uint32(), C.c_float()
        self._ck(lib().tpose_cuda_run(self._h, slot, flavour, max_steps, C.c_double(rel_tol),
                                      C.byref(n), C.byref(f), C.byref(e)))
        return n.value, f.value, e.value

    def watch(self, tri=None, halfedges=None):
        """The host's current topology for the per-frame edit watch (None switches it off)."""
        if tri is None:
            self._ck(lib().tpose_cuda_watch(self._h, None, None, 0))
            return
        tri = np.ascontiguousarray(tri, dtype=np.int32)
        he = np.ascontiguousarray(halfedges, dtype=np.int32)
        self._ck(lib().tpose_cuda_watch(self._h, _ptr(tri), _ptr(he), tri.shape[0]))

    def watch_mask(self):
        """Bit t of the returned bool array: the device flagged triangle t on the frame the last run stopped on."""
        nw = (self.NT + 31) // 32
        m = np.zeros(nw, np.uint32)
        self._ck(lib().tpose_cuda_watch_mask(self._h, _ptr(m), nw))
        return np.unpackbits(m.view(np.uint8), bitorder="little")[:self.NT].astype(bool)

    def set_history(self, frames):
        self._ck(lib().tpose_cuda_set_history(self._h, C.c_int(frames)))

    def history_frame(self, frame):
        """Frame `frame` of the last run: dict(pts, mask (bool per triangle), ten (base energies), toterr)."""
        nw = (self.NT + 31) // 32
        pts = np.empty((self.NP, 2), np.float32); m = np.zeros(nw, np.uint32)
        ten = np.empty(self.NT, np.int32); tot = C.c_float()
        self._ck(lib().tpose_cuda_history_frame(self._h, C.c_int(frame), _ptr(pts), _ptr(m), nw, _ptr(ten), C.byref(tot)))
        return dict(pts=pts, mask=np.unpackbits(m.view(np.uint8), bitorder="little")[:self.NT].astype(bool), ten=ten, toterr=tot.value)

    def eval_list(self, slot, entries, reset_gradient=False):
        """entries (n, 4) int32 = (v0, v1, v2, k): (ten, cn, ca) of those instances at the current points."""
        en = np.ascontiguousarray(entries, np.int32).reshape(-1, 4)
        n = en.shape[0]
        ten = np.zeros(n, np.int32); cn = np.zeros(n, np.int32); ca = np.zeros((n, 4), np.int32)
        self._ck(lib().tpose_cuda_eval_list(self._h, slot, _ptr(en), n, int(reset_gradient), _ptr(ten), _ptr(cn), _ptr(ca)))
        return ten, cn, ca

    def history_restore_gradient(self, frame):
        self._ck(lib().tpose_cuda_history_restore_gradient(self._h, C.c_int(frame)))

    def read(self, which, first, count):
        dt, width = _BUF_DTYPE[which]
        out = np.empty((count, width) if width > 1 else (count,), dt)
        self._ck(lib().tpose_cuda_read(self._h, which, C.c_size_t(first), C.c_size_t(count), _ptr(out)))
        return out

    def write(self, which, first, data):
        dt, width = _BUF_DTYPE[which]
        data = np.ascontiguousarray(data, dtype=dt)
        self._ck(lib().tpose_cuda_write(self._h, which, C.c_size_t(first), C.c_size_t(data.shape[0]), _ptr(data)))

    def set_toterr(self, v):
        self._ck(lib().tpose_cuda_set_toterr(self._h, C.c_float(v)))

    def get_toterr(self):
        v = C.c_float()
        self._ck(lib().tpose_cuda_get_toterr(self._h, C.byref(v)))
        return v.value

    # -- read back
    def retrieve(self, ten=True, cn=True, ca=True, pts=True):
        nv = 13 * self.NT
        a_ten = np.empty(nv, np.int32) if ten else None
        a_cn = np.empty(nv, np.int32) if cn else None
        a_ca = np.empty((nv, 4), np.int32) if ca else None
        a_pts = np.empty((self.NP, 2), np.float32) if pts else None
        self._ck(lib().tpose_cuda_retrieve(self._h, _ptr(a_ten), _ptr(a_cn), _ptr(a_ca), _ptr(a_pts)))
        return dict(ten=a_ten, cn=a_cn, ca=a_ca, pts=a_pts)

    def retrieve_into(self, ten=None, cn=None, ca=None, pts=None):
        """Same read-back into caller-owned arrays (pinned ones copy without staging)."""
        self._ck(lib().tpose_cuda_retrieve(self._h, _ptr(ten), _ptr(cn), _ptr(ca), _ptr(pts)))

    def retrieve_base(self):
        ten = np.empty(self.NT, np.int32)
        cn = np.empty(self.NT, np.int32)
        ca = np.empty((self.NT, 4), np.int32)
        self._ck(lib().tpose_cuda_retrieve_base(self._h, _ptr(ten), _ptr(cn), _ptr(ca)))
        return ten, cn, ca

    def retrieve_gradient(self):
        gr = np.empty((self.NP, 2), np.int32)
        self._ck(lib().tpose_cuda_retrieve_gradient(self._h, _ptr(gr)))
        return gr

    def raster_ids(self, slot=0):
        h, w = self.shape[slot]
        ids = np.empty((h, w), np.int32)
        self._ck(lib().tpose_cuda_raster_ids(self._h, slot, _ptr(ids)))
        return ids

    def render(self, w, h, other_pts=None, s=0.0, lines=False):
        """software/view, headless: (h, w, 4) uint8, flat triangle colours on black; lines: the wireframe on top."""
        out = np.empty((h, w, 4), np.uint8)
        op = None if other_pts is None else np.ascontiguousarray(other_pts, np.float32)
        fn = lib().tpose_cuda_render_wire if lines else lib().tpose_cuda_render
        self._ck(fn(self._h, w, h, _ptr(op), C.c_float(s), _ptr(out)))
        return out

    def last_elapsed_ms(self):
        ms = C.c_float()
        self._ck(lib().tpose_cuda_last_elapsed_ms(self._h, C.byref(ms)))
        return ms.value

    def set_step_scale(self, scale):
        self._ck(lib().tpose_cuda_set_step_scale(self._h, C.c_float(scale)))

    def set_kernel_timing(self, on):
        self._ck(lib().tpose_cuda_set_kernel_timing(self._h, C.c_int(1 if on else 0)))

    def last_kernel_ms(self):
        """(k_edges ms, k_corners ms) of the most recent step(.., 1)."""
        a, b = C.c_float(), C.c_float()
        self._ck(lib().tpose_cuda_last_kernel_ms(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def last_setup_ms(self):
        a = C.c_float()
        self._ck(lib().tpose_cuda_last_setup_ms(self._h, C.byref(a)))
        return a.value

    def kernel_path(self):
        """'tiles' (TMA-staged image tiles) or 'table' (row-prefix table gathers, TPOSE_CUDA_PATH=table)."""
        return "tiles" if lib().tpose_cuda_kernel_path(self._h) == 1 else "table"

    def sync(self):
        self._ck(lib().tpose_cuda_sync(self._h))

    def launch_count(self):
        return int(lib().tpose_cuda_launch_count(self._h))

    def flush_l2(self, nbytes=256 << 20):
        self._ck(lib().tpose_cuda_flush_l2(self._h, C.c_size_t(nbytes)))


def _handles(ctxs):
    arr = (C.c_void_p * len(ctxs))(*[c._h for c in ctxs])
    return arr


def batch_step(ctxs, flavour, nsteps=1, slots=None):
    """nsteps frames of every context in one launch per frame (contexts on one device)."""
    sl = None if slots is None else (C.c_int * len(ctxs))(*slots)
    rc = lib().tpose_cuda_batch_step(_handles(ctxs), len(ctxs), sl, flavour, nsteps)
    if rc != 0:
        raise TposeCudaError(f"batch_step -> {rc}: {lib().tpose_cuda_last_error(ctxs[0]._h).decode()}")


def batch_run(ctxs, flavour, max_steps, rel_tol, slots=None):
    sl = None if slots is None else (C.c_int * len(ctxs))(*slots)
    done = (C.c_int * len(ctxs))()
    trig = (C.c_uint32 * len(ctxs))()
    rc = lib().tpose_cuda_batch_run(_handles(ctxs), len(ctxs), sl, flavour, max_steps, C.c_double(rel_tol), done, trig)
    if rc != 0:
        raise TposeCudaError(f"batch_run -> {rc}: {lib().tpose_cuda_last_error(ctxs[0]._h).decode()}")
    return list(done), list(trig)


def _mview_ck(rc, what):
    if rc != 0:
        lib().tpose_cuda_mview_last_error.restype = C.c_char_p
        raise TposeCudaError(f"{what} -> {rc}: {lib().tpose_cuda_mview_last_error().decode()}")


def _concat(lists):
    offs = np.zeros(len(lists) + 1, np.int32)
    offs[1:] = np.cumsum([len(x) for x in lists])
    cat = np.concatenate([np.asarray(x, np.float32).reshape(-1, 2) for x in lists]) if offs[-1] else np.zeros((0, 2), np.float32)
    return np.ascontiguousarray(cat, np.float32), offs


def f8point(A_lists, B_lists, device=0):
    """Batched mview::F_8Point: lists of (N_i, 2) float32 matches -> (nbatch, 3, 3) float32."""
    A, offs = _concat(A_lists)
    B, offs_b = _concat(B_lists)
    assert np.array_equal(offs, offs_b), "matched sets have different size"
    if A.shape[0] == 0:
        A = np.zeros((1, 2), np.float32); B = np.zeros((1, 2), np.float32)
    F = np.zeros((len(A_lists), 9), np.float32)
    _mview_ck(lib().tpose_cuda_f8point(device, _ptr(A), _ptr(B), _ptr(offs), len(A_lists), _ptr(F)), "f8point")
    return F.reshape(-1, 3, 3)


def fsampson(A_lists, B_lists, weights=None, iterations=100, device=0):
    """Batched mview::F_Sampson (weights: list of (N_i,) float32 or None) -> (nbatch, 3, 3) float32."""
    A, offs = _concat(A_lists)
    B, offs_b = _concat(B_lists)
    assert np.array_equal(offs, offs_b), "matched sets have different size"
    W = None
    if weights is not None:
        W = np.ascontiguousarray(np.concatenate([np.asarray(w, np.float32).ravel() for w in weights]))
        assert W.shape[0] == A.shape[0], "one weight per match"
    if A.shape[0] == 0:
        A = np.zeros((1, 2), np.float32); B = np.zeros((1, 2), np.float32); W = None
    F = np.zeros((len(A_lists), 9), np.float32)
    _mview_ck(lib().tpose_cuda_fsampson(device, _ptr(A), _ptr(B), _ptr(W), _ptr(offs), len(A_lists),
                                        C.c_int(iterations), _ptr(F)), "fsampson")
    return F.reshape(-1, 3, 3)


def select_matches(originA, pointsA, warp_originA, pointsB, originB, warp_originB, ratio, max_dist=0.0000002, device=0):
    """Two-way-inconsistency match selection (compute_fundamental_mat/model.h:1-82): (X, Y, weights, dist)."""
    f = lambda a: np.ascontiguousarray(np.asarray(a, np.float32).reshape(-1, 2))
    oA, pA, wA, pB, oB, wB = map(f, (originA, pointsA, warp_originA, pointsB, originB, warp_originB))
    npa, npb = len(pA), len(pB)
    n = max(npa + npb, 1)
    X = np.zeros((n, 2), np.float32); Y = np.zeros((n, 2), np.float32); W = np.zeros(n, np.float32); D = np.zeros(n, np.float32)
    kept = C.c_int()
    _mview_ck(lib().tpose_cuda_select_matches(device, _ptr(oA), _ptr(pA), _ptr(wA), npa, _ptr(pB), _ptr(oB), _ptr(wB), npb,
                                              C.c_float(ratio), C.c_double(max_dist), _ptr(X), _ptr(Y), _ptr(W), _ptr(D), C.byref(kept)),
              "select_matches")
    k = kept.value
    return X[:k], Y[:k], W[:k], D[:npa + npb]


def triangulate(F, K, A_lists, B_lists, correct=True, device=0):
    """Batched mview::triangulate(F,K,A,B): returns (list of (N_i,4) points, corrected A, corrected B)."""
    A, offs = _concat(A_lists)
    B, _ = _concat(B_lists)
    nb = len(A_lists)
    F = np.ascontiguousarray(np.asarray(F, np.float32).reshape(nb, 9))
    K = np.ascontiguousarray(np.broadcast_to(np.asarray(K, np.float32).reshape(-1, 9), (nb, 9)))
    X = np.zeros((A.shape[0], 4), np.float32)
    Ac, Bc = np.zeros_like(A), np.zeros_like(B)
    _mview_ck(lib().tpose_cuda_triangulate(device, _ptr(F), _ptr(K), _ptr(A), _ptr(B), _ptr(offs), nb, int(correct),
                                           _ptr(X), _ptr(Ac), _ptr(Bc)), "triangulate")
    sp = lambda a: [a[offs[i]:offs[i + 1]] for i in range(nb)]
    return sp(X), sp(Ac), sp(Bc)


def hdlt(PA, PB, A, B, device=0):
    PA = np.ascontiguousarray(PA, np.float32).reshape(12); PB = np.ascontiguousarray(PB, np.float32).reshape(12)
    A = np.ascontiguousarray(A, np.float32).reshape(-1, 2); B = np.ascontiguousarray(B, np.float32).reshape(-1, 2)
    X = np.zeros((A.shape[0], 4), np.float32)
    _mview_ck(lib().tpose_cuda_hdlt(device, _ptr(PA), _ptr(PB), _ptr(A), _ptr(B), A.shape[0], _ptr(X)), "hdlt")
    return X


def geterr(energies, toterr=1.0, rel_tol=1e-4, device=0, route=-1):
    """geterr over an int32 array on the device: (newerr, relerr, relerr < rel_tol).  route: -1 as a run chooses,
    0 the plain chain, 1 the block-parallel routine."""
    e = np.ascontiguousarray(energies, np.int32)
    tot, rel, below = C.c_float(toterr), C.c_float(), C.c_int()
    _mview_ck(lib().tpose_cuda_geterr(device, _ptr(e) if e.size else _ptr(np.zeros(1, np.int32)), int(e.size), C.byref(tot), C.byref(rel),
                                      C.byref(below), C.c_double(rel_tol), C.c_int(route)), "geterr")
    return np.float32(tot.value), np.float32(rel.value), bool(below.value)


def carry(tri, frm, to, pts, ratio, device=0):
    """triangulation::warp on the device: returns pts carried from embedding `frm` to embedding `to`."""
    tri = np.ascontiguousarray(tri, np.int32); frm = np.ascontiguousarray(frm, np.float32)
    to = np.ascontiguousarray(to, np.float32); out = np.ascontiguousarray(pts, np.float32).copy()
    assert frm.shape == to.shape
    _mview_ck(lib().tpose_cuda_carry(device, _ptr(tri), tri.shape[0], _ptr(frm), _ptr(to), frm.shape[0], _ptr(out),
                                     out.shape[0], C.c_float(ratio)), "carry")
    return out


def rgbd_planefit(tri, pts, T, TB, pixels, points, normals, intr, device=0):
    """software/rgbd's meshing core (rgbd/main.cpp:244-349): (vertices (NP,4), ring (NP,), points per triangle (NT,)).
    T / TB: 3x3 as m[col][row] (glm layout)."""
    tri = np.ascontiguousarray(tri, np.int32); pts = np.ascontiguousarray(pts, np.float32)
    T = np.ascontiguousarray(T, np.float32).reshape(9); TB = np.ascontiguousarray(TB, np.float32).reshape(9)
    px = np.ascontiguousarray(pixels, np.float32).reshape(-1, 2); X = np.ascontiguousarray(points, np.float32).reshape(-1, 4)
    N = np.ascontiguousarray(normals, np.float32).reshape(-1, 4); k = np.ascontiguousarray(intr, np.float32).reshape(4)
    V = np.zeros((pts.shape[0], 4), np.float32); ring = np.zeros(pts.shape[0], np.float32); cnt = np.zeros(tri.shape[0], np.int32)
    _mview_ck(lib().tpose_cuda_rgbd_planefit(device, _ptr(tri), tri.shape[0], _ptr(pts), pts.shape[0], _ptr(T), _ptr(TB),
                                             _ptr(px) if len(px) else None, _ptr(X) if len(px) else None, _ptr(N) if len(px) else None,
                                             len(px), _ptr(k), _ptr(V), _ptr(ring), _ptr(cnt)), "rgbd_planefit")
    return V, ring, cnt


def mview_last_ms():
    lib().tpose_cuda_mview_last_ms.restype = C.c_float
    return float(lib().tpose_cuda_mview_last_ms())


def pinned_empty(shape, dtype):
    """numpy array over page-locked host memory (cudaHostAlloc); keep the returned array alive."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = C.c_void_p()
    rc = lib().tpose_cuda_host_alloc(C.c_size_t(max(n, 1)), C.byref(p))
    if rc != 0:
        raise TposeCudaError("cudaHostAlloc failed")
    buf = (C.c_char * max(n, 1)).from_address(p.value)
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
