"""Frame-by-frame restatement of the two program loops of the reference, on the CPU.

TEST INFRASTRUCTURE ONLY (imported by tests/ and tools/, never by the product).

  triangulate_loop  -- software/triangulate/main.cpp:176-353
  warp_loop         -- software/warp/main.cpp:202-284

Every frame does what the reference's frame does, in its order, with no shortcut: the device work
is the per-fragment step oracle (oracle/step_oracle.c through oracle/oracle.py), the mesh and its
edits / error reductions / file IO are the reference's OWN code (oracle/_ref/libtpose_ref.so,
driven through the `Tri` wrapper the caller passes in -- tests/hostlib.py).  Only the glue that
lives in main.cpp itself (the flip set, the export rule, the A<->B switch) is restated here, with
the line it follows.
"""
import numpy as np

from . import oracle as O


class _Device:
    """The eight SSBOs as the reference's frame sees them."""

    def __init__(self, img):
        self.img = img
        self.pts = None
        self.tri = None
        self.cn = None
        self.ca = None
        self.ten = None
        self.gr = None          # pgradbuf persists across frames and uploads (triangle.vs:93-98 zeroes referenced vertices only)

    def upload(self, tr_state, colors13=None):      # source/triangulation.hpp:628-643
        self.pts = tr_state["pts"].copy()
        self.tri = tr_state["tri"].copy()
        if colors13 is not None:
            self.ca = colors13


def _flip_set(terr, halfedges, ntri):
    """triangulate/main.cpp:233-281.  std::set ordered by energy only (descending): a halfedge whose
    energy equals one already present is dropped; std::map hflipset iterates by halfedge index."""
    seen = {}
    for t in range(ntri):
        for k in range(3):
            o = halfedges[3 * t + k]
            if o >= 0:
                e = np.float32(np.int32(terr[t]) + np.int32(terr[o // 3]))
                key = float(e)
                if key not in seen:
                    seen[key] = 3 * t + k
    ordered = sorted(seen.items(), key=lambda kv: -kv[0])
    nflip, hflip = set(), {}
    for e, h in ordered:
        if h in nflip:
            continue
        o = halfedges[h]
        if o < 0 or o in nflip:
            continue
        hflip[h] = np.float32(e)
        ta, tb = h // 3, o // 3
        for k in range(3):
            nflip.add(3 * ta + k)
            nflip.add(3 * tb + k)
    return dict(sorted(hflip.items()))


def triangulate_loop(R, Tri, img, ratio, exportlist, trifile, max_frames, tol=1e-4, trace=None, updates=None):
    """R: the reference library (hostlib.ref()); Tri: hostlib.Tri.  exportlist is popped from the
    back like main.cpp:19-31.  Returns dict(frames, tr (final Tri.get()), finished)."""
    import ctypes as C
    exportlist = list(exportlist)
    R.set_ratio(ratio)
    R.f("set_toterr")(C.c_float(1.0))
    tr = Tri(R)
    dev = _Device(img)
    NPI = np.float32(3.14159265)

    def upload():
        dev.upload(tr.get())

    def computecolors():                                        # main.cpp:121-130
        dev.cn, dev.ca = O.mode0(img, dev.pts, dev.tri, ratio, O.TRIANGULATE)

    def doenergy():                                             # main.cpp:132-141
        dev.ten, dev.gr = O.mode1(img, dev.pts, dev.tri, ratio, O.TRIANGULATE, dev.cn, dev.ca, gr=dev.gr)

    def doshift():                                              # main.cpp:143-155
        dev.gr = O.gradient(dev.tri, dev.ten, dev.gr)
        dev.pts = O.shift(dev.pts, dev.gr[:len(dev.pts)], ratio, O.TRIANGULATE)

    def retrieve_terr():
        R.f("set_terr")(dev.ten.ctypes.data_as(C.c_void_p), len(dev.ten))
        return dev.ten

    upload()                                                    # main.cpp:106
    computecolors()                                             # main.cpp:176
    frames, finished = 0, False
    while max_frames < 0 or frames < max_frames:
        doenergy(); doshift()                                   # main.cpp:196-197
        frames += 1
        terr = retrieve_terr()                                  # main.cpp:201-204
        cn = dev.cn
        R.f("set_cn")(cn.ctypes.data_as(C.c_void_p), len(cn))
        tr.set_points(dev.pts)
        updated = False
        err = tr.geterr()
        if trace is not None:
            trace.append((frames, tr.counts()[0], float(err)))
        if float(err) < tol:                                    # main.cpp:210 (float vs double literal)
            if not exportlist:                                  # main.cpp:214-217
                finished = True
                break
            s = tr.get()
            NT = s["counts"][0]
            if NT >= exportlist[-1]:                            # main.cpp:219-231
                col = s["col"].copy()
                base = dev.ca[:NT].copy()
                col[:NT] = base
                nz = cn[:NT] > 0
                # ivec4 /= int: C integer division (truncation); sums are non-negative here
                col[:NT][nz] = (base[nz] // cn[:NT][nz, None]).astype(np.int32)
                tr.set(s["tri"], s["he"], s["pts"], s["pts"], col)
                tr.write(trifile)
                exportlist.pop()
                s = tr.get()
            hflip = _flip_set(terr, s["he"], len(s["tri"]))
            for h in hflip:                                     # main.cpp:285-288
                tr.flip(h, 0.0)
            upload(); computecolors(); doenergy()               # main.cpp:290-294
            terr = retrieve_terr()
            he = tr.get()["he"]
            for h, e in hflip.items():                          # main.cpp:296-301
                if np.float32(np.int32(terr[h // 3]) + np.int32(terr[he[h] // 3])) > e:
                    tr.flip(h, 0.0)
                    he = tr.get()["he"]
            upload(); computecolors(); doenergy()               # main.cpp:303-306
            terr = retrieve_terr()
            tta = tr.maxerrid()                                 # main.cpp:308-310
            if tta >= 0 and tr.split(tta):
                updated = True
        ta = 0                                                  # main.cpp:316-320
        while ta < tr.counts()[0]:
            if tr.boundary(ta) == 3 and tr.prune(ta):
                updated = True
            ta += 1
        lim = 0.8 * float(NPI)
        ta = 0                                                  # main.cpp:324-333
        while ta < tr.counts()[0]:
            for k in range(3):
                if float(tr.angle(3 * ta + k)) > lim:
                    tr.flip(3 * ta + k, 0.0)
            ta += 1
        ta = 0                                                  # main.cpp:337-347
        while ta < tr.sizes()[0]:
            ha = 3 * ta
            ml = tr.hlength(ha)
            if tr.hlength(ha + 1) < ml:
                ha += 1; ml = tr.hlength(ha)
            if tr.hlength(ha + 1) < ml:
                ha += 1; ml = tr.hlength(ha)
            if tr.collapse(ha):
                updated = True
            ta += 1
        if updated:                                             # main.cpp:348-351
            tot = tr.gettoterr()
            if updates is not None:
                updates.append((tr.counts()[0], float(tot), frames))   # what the program prints: "NT toterr"
            upload()
        computecolors()                                         # the pipeline, main.cpp:180-184
    out = dict(frames=frames, tr=tr.get(), finished=finished)
    tr.close()
    return out


def warp_loop(R, Tri, img_a, img_b, triA, triB, max_frames, tol=1e-6, level_cap=-1):
    """software/warp/main.cpp:112-284.  Writes triA+'.warp' / triB+'.warp'.
    level_cap > 0 (not in the reference; mirrors tpose::pairs::warp_batch): a level also ends after that many
    frames.  The returned dict carries the frames each level took."""
    import ctypes as C
    R.f("set_toterr")(C.c_float(1.0))                           # a fresh process: toterr starts at 1 (triangulation.hpp:648)
    trA, trB = Tri(R), Tri(R)
    trA.read(triA); trB.read(triB)                              # main.cpp:115-116
    warpA = True
    frames, levels, finished = 0, 0, False
    level_frames, per_level = 0, []
    state = {}

    def upload(tr):                                             # upload(tr): colours replicated 13x
        s = tr.get()
        NT = s["counts"][0]
        state["pts"], state["tri"] = s["pts"].copy(), s["tri"].copy()
        ca = np.zeros((13 * NT, 4), np.int32)
        for k in range(13):
            ca[k * NT:(k + 1) * NT] = s["col"][:NT]
        state["ca"] = ca

    tr = trA
    upload(tr)                                                  # main.cpp:126
    while max_frames < 0 or frames < max_frames:
        ratio = R.get_ratio()
        img = img_b if warpA else img_a                         # warp/shader/triangle.fs:49-50
        cn, _ = O.mode0(img, state["pts"], state["tri"], ratio, O.WARP, state["ca"])      # doreset
        ten, gr = O.mode1(img, state["pts"], state["tri"], ratio, O.WARP, cn, state["ca"], gr=state.get("gr"))  # doenergy
        gr = O.gradient(state["tri"], ten, gr)                  # doshift
        state["gr"] = gr
        state["pts"] = O.shift(state["pts"], gr[:len(state["pts"])], ratio, O.WARP)
        frames += 1
        level_frames += 1
        R.f("set_terr")(ten.ctypes.data_as(C.c_void_p), len(ten))
        tr.set_points(state["pts"])
        if float(tr.geterr()) < tol or (level_cap > 0 and level_frames >= level_cap):   # main.cpp:231
            if warpA:                                           # main.cpp:235-243
                sB = trB.get()
                trB.set_points(trA.warp(sB["opts"], reverse=True))
            else:
                sA = trA.get()
                trA.set_points(trB.warp(sA["opts"], reverse=True))
            warpA = not warpA
            tr = trA if warpA else trB
            trA.write(triA + ".warp"); trB.write(triB + ".warp")   # main.cpp:269-270
            levels += 1
            per_level.append(level_frames); level_frames = 0
            if not trA.read(triA, True) or not trB.read(triB, True):
                finished = True
                break
            upload(tr)
    out = dict(frames=frames, levels=levels, finished=finished, per_level=per_level, A=trA.get(), B=trB.get())
    trA.close(); trB.close()
    return out
