"""GPU parity tests: the CUDA path through the C ABI against the CPU oracle, bit for bit."""
import numpy as np
import pytest

from conftest import bits

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("kernel_path")]


def _ctx(pkg, nt=4096, np_=4096):
    return pkg.Context(device=0, max_triangles=nt, max_points=np_)


def _assert_step_equal(got, ref, keys):
    for k in keys:
        assert np.array_equal(got[k], ref[k]), f"{k} differs: {np.flatnonzero((got[k] != ref[k]).reshape(len(ref[k]), -1).any(1))[:8]}"
    assert np.array_equal(bits(got["pts"]), bits(ref["pts"])), "points differ"


@pytest.mark.parametrize("w,h,ntri", [(256, 160, 150), (300, 200, 150), (853, 480, 1000), (64, 48, 150)])
@pytest.mark.parametrize("group", [0, 1, 4, 32])
def test_fused_step_triangulate_matches_oracle(cuda_pkg, synth, O, w, h, ntri, group, monkeypatch):
    if group:
        monkeypatch.setenv("TPOSE_CUDA_GROUP", str(group))
    else:
        monkeypatch.delenv("TPOSE_CUDA_GROUP", raising=False)
    ratio = w / h
    img = synth.make_image(w, h, seed=1000)
    pts, tri, _ = synth.mesh_for(ntri, ratio)
    ctx = _ctx(cuda_pkg)
    ctx.set_image(0, img)
    ctx.upload(pts, tri, ratio)
    cur = pts
    for it in range(3):
        ctx.step(0, cuda_pkg.TRIANGULATE, 1)
        got = ctx.retrieve()
        ref = O.step(img, cur, tri, ratio, O.TRIANGULATE)
        _assert_step_equal(got, ref, ("cn", "ca", "ten"))
        assert np.array_equal(ctx.retrieve_gradient(), ref["gr"])
        cur = ref["pts"]
    ctx.close()


def test_dense_mesh_on_a_small_raster(cuda_pkg, synth, O):
    """10 000 triangles over 320x200 pixels: sub-pixel triangles, most instances cover no pixel centre at all, and
    on the tile path every 64x64 tile is crossed by ~2 000 edges -- far beyond its work-list capacity, so the
    tiles walk the edges in id-range chunks.  Still the oracle's bits."""
    w, h = 320, 200
    ratio = w / h
    img = synth.make_image(w, h, seed=77)
    pts, tri, _ = synth.mesh_for(10000, ratio)
    ctx = cuda_pkg.Context(device=0, max_triangles=tri.shape[0], max_points=pts.shape[0])
    ctx.set_image(0, img)
    ctx.upload(pts, tri, ratio)
    cur, gr = pts, None
    for it in range(2):
        ctx.step(0, cuda_pkg.TRIANGULATE, 1)
        ref = O.step(img, cur, tri, ratio, O.TRIANGULATE, gr=gr)
        _assert_step_equal(ctx.retrieve(), ref, ("cn", "ca", "ten"))
        assert np.array_equal(ctx.retrieve_gradient(), ref["gr"])
        cur, gr = ref["pts"], ref["gr"]
    ctx.close()


def test_fused_step_warp_matches_oracle(cuda_pkg, synth, O):
    w, h = 320, 180
    ratio = w / h
    a, b = synth.make_pair(w, h, seed=1001)
    pts, tri, _ = synth.mesh_for(150, ratio)
    NT = tri.shape[0]
    r0 = O.step(a, pts, tri, ratio, O.TRIANGULATE)
    colors = np.ones((NT, 4), np.int32)
    colors[:, :3] = r0["ca"][:NT, :3] // np.maximum(r0["cn"][:NT, None], 1)
    ctx = _ctx(cuda_pkg)
    ctx.set_image(0, a)
    ctx.set_image(1, b)
    ctx.upload(pts, tri, ratio, colors)
    cur = pts
    for it in range(4):
        ctx.step(1, cuda_pkg.WARP, 1)
        got = ctx.retrieve()
        ref = O.step(b, cur, tri, ratio, O.WARP, colors=colors)
        _assert_step_equal(got, ref, ("cn", "ten"))
        assert np.array_equal(got["ca"][:, :3], ref["ca"][:, :3])
        cur = ref["pts"]
    ctx.close()


def test_unfused_lambdas_match_oracle(cuda_pkg, synth, O):
    """computecolors / doenergy / doshift called one by one, as the reference's frame does."""
    w, h = 200, 120
    ratio = w / h
    img = synth.make_image(w, h, seed=5)
    pts, tri, _ = synth.mesh_for(150, ratio)
    ctx = _ctx(cuda_pkg)
    ctx.set_image(0, img)
    ctx.upload(pts, tri, ratio)
    ctx.colors(0, cuda_pkg.TRIANGULATE)
    cn, ca = O.mode0(img, pts, tri, ratio, O.TRIANGULATE)
    got = ctx.retrieve()
    assert np.array_equal(got["cn"], cn) and np.array_equal(got["ca"], ca)
    # move the points, keep the OLD colours: the energy pass must use the stored ca/cn
    moved = pts.copy()
    moved[4:] += np.float32(0.003)
    ctx.upload_points(moved)
    ctx.energy(0, cuda_pkg.TRIANGULATE)
    ten, gr = O.mode1(img, moved, tri, ratio, O.TRIANGULATE, cn, ca)
    assert np.array_equal(ctx.retrieve()["ten"], ten)
    ctx.shift(cuda_pkg.TRIANGULATE)
    gr = O.gradient(tri, ten, gr)
    assert np.array_equal(ctx.retrieve_gradient(), gr)
    assert np.array_equal(bits(ctx.retrieve()["pts"]), bits(O.shift(moved, gr, ratio, O.TRIANGULATE)))
    ctx.close()


def test_adversarial_vertices_on_pixel_centres(cuda_pkg, O):
    """Fill-rule ties: vertices exactly on pixel centres and corners, both windings, off-screen."""
    rng = np.random.default_rng(11)
    W, H = 64, 40
    ratio = W / H
    img = rng.integers(0, 256, size=(H, W, 4)).astype(np.uint8)
    ctx = _ctx(cuda_pkg, 64, 256)
    ctx.set_image(0, img)
    gr = None        # pgradbuf lives as long as the context: vertices no triangle references keep stale entries
    for trial in range(40):
        NT = int(rng.integers(1, 12))
        NP = 4 + 3 * NT
        k = rng.integers(-4, 2 * W + 4, NP) / 2.0
        l = rng.integers(-4, 2 * H + 4, NP) / 2.0
        pts = np.stack([((k / (W / 2)) - 1) * ratio, (l / (H / 2)) - 1], 1).astype(np.float32)
        tri = np.zeros((NT, 4), np.int32)
        tri[:, :3] = rng.integers(0, NP, size=(NT, 3))
        for fl in (cuda_pkg.TRIANGULATE, cuda_pkg.WARP):
            cols = rng.integers(0, 256, size=(NT, 4)).astype(np.int32)
            ctx.upload(pts, tri, ratio, cols)
            ctx.step(0, fl, 1)
            got = ctx.retrieve()
            ref = O.step(img, pts, tri, ratio, fl, colors=cols, gr=gr)
            gr = ref["gr"]
            _assert_step_equal(got, ref, ("cn", "ten"))
            assert np.array_equal(ctx.retrieve_gradient(), gr[:NP])
        ctx.upload_points(pts)   # the step moved the device points; the ID raster is of `pts`
        assert np.array_equal(ctx.raster_ids(0), O.raster_ids(pts, tri, ratio, W, H))
    ctx.close()


def test_raster_ids_match_oracle(cuda_pkg, synth, O):
    w, h = 640, 360
    ratio = w / h
    img = synth.make_image(w, h, seed=9)
    pts, tri, _ = synth.mesh_for(1000, ratio)
    ctx = _ctx(cuda_pkg)
    ctx.set_image(0, img)
    ctx.upload(pts, tri, ratio)
    ids = ctx.raster_ids(0)
    assert np.array_equal(ids, O.raster_ids(pts, tri, ratio, w, h))
    ctx.step(0, cuda_pkg.TRIANGULATE, 1)
    cn = ctx.retrieve()["cn"][: tri.shape[0]]
    assert np.array_equal(np.bincount(ids.ravel(), minlength=tri.shape[0]), cn)
    ctx.close()


def test_run_stops_on_convergence_like_geterr(cuda_pkg, synth, O):
    """tpose_cuda_run == stepping one frame at a time and applying geterr on the host."""
    w, h = 256, 160
    ratio = w / h
    img = synth.make_image(w, h, seed=21)
    pts, tri, _ = synth.mesh_for(150, ratio)
    NT = tri.shape[0]
    tol = 1e-3
    ctx = _ctx(cuda_pkg)
    ctx.set_image(0, img)
    ctx.upload(pts, tri, ratio)
    n, flags, err = ctx.run(0, cuda_pkg.TRIANGULATE, 60, tol)
    got = ctx.retrieve()
    # host replay with the oracle
    toterr = np.float32(1.0)
    cur = pts
    steps = 0
    for it in range(60):
        ref = O.step(img, cur, tri, ratio, O.TRIANGULATE)
        cur = ref["pts"]
        steps += 1
        newerr = np.float32(O.sum_energy(ref["ten"], NT))
        rel = np.float32((toterr - newerr) / toterr)
        toterr = newerr
        if abs(rel) < tol:
            break
    assert n == steps
    assert (flags == cuda_pkg.TRIG_CONVERGED) == (steps < 60 or abs(rel) < tol)
    assert np.array_equal(got["ten"], ref["ten"])
    assert np.array_equal(bits(got["pts"]), bits(cur))
    assert ctx.get_toterr() == float(toterr)
    ctx.close()


def test_run_with_quiet_watch_equals_plain_run(cuda_pkg, synth):
    """A watch topology on which no per-frame edit can apply must not change the run."""
    w, h = 256, 160
    ratio = w / h
    img = synth.make_image(w, h, seed=21)
    pts, tri, he = synth.mesh_for(150, ratio)
    out = []
    for watch in (False, True):
        ctx = _ctx(cuda_pkg)
        ctx.set_image(0, img)
        ctx.upload(pts, tri, ratio)
        if watch:
            ctx.watch(tri, he)
        n, flags, err = ctx.run(0, cuda_pkg.TRIANGULATE, 25, 1e-9)
        out.append((n, flags, ctx.retrieve()))
        ctx.close()
    assert out[0][0] == out[1][0] == 25 and out[0][1] == out[1][1] == cuda_pkg.TRIG_MAXSTEPS
    for k in ("cn", "ca", "ten"):
        assert np.array_equal(out[0][2][k], out[1][2][k])
    assert np.array_equal(bits(out[0][2]["pts"]), bits(out[1][2]["pts"]))


def test_watch_stops_after_the_frame_where_an_edit_applies(cuda_pkg, synth, O):
    """Short edge (<= 0.01, triangulation.hpp:338): the run must stop right after the first frame,
    holding that frame's energies and the shifted points -- what the reference has in hand when its
    per-frame collapse loop (triangulate/main.cpp:336-346) fires."""
    w, h = 256, 160
    ratio = w / h
    img = synth.make_image(w, h, seed=22)
    pts, tri, he = synth.mesh_for(150, ratio)
    pts = pts.copy()
    a, b = int(tri[40, 0]), int(tri[40, 1])
    a, b = (a, b) if a >= 4 else (b, a)
    pts[a] = pts[b] + np.float32([0.004, 0.0])
    ctx = _ctx(cuda_pkg)
    ctx.set_image(0, img)
    ctx.upload(pts, tri, ratio)
    ctx.watch(tri, he)
    n, flags, err = ctx.run(0, cuda_pkg.TRIANGULATE, 10, 1e-12)
    assert flags & cuda_pkg.TRIG_TOPOLOGY
    ref = O.step(img, pts, tri, ratio, O.TRIANGULATE)
    ln = np.linalg.norm(ref["pts"][a] - ref["pts"][b])
    assert n == 1 and ln <= 0.0100005
    got = ctx.retrieve()
    _assert_step_equal(got, ref, ("cn", "ca", "ten"))
    assert ctx.get_toterr() == O.sum_energy(ref["ten"], tri.shape[0])
    ctx.close()


def test_read_write_buffers_like_tinyengine_buffers(cuda_pkg, synth, O):
    w, h = 128, 96
    ratio = w / h
    img = synth.make_image(w, h, seed=2)
    pts, tri, _ = synth.mesh_for(150, ratio)
    NT, NP = tri.shape[0], pts.shape[0]
    ctx = _ctx(cuda_pkg)
    ctx.set_image(0, img)
    ctx.upload(pts, tri, ratio)
    ctx.step(0, cuda_pkg.TRIANGULATE, 1)
    ref = O.step(img, pts, tri, ratio, O.TRIANGULATE)
    assert np.array_equal(ctx.read(cuda_pkg.BUF_TENERGY, 0, 13 * NT), ref["ten"])
    assert np.array_equal(ctx.read(cuda_pkg.BUF_COLNUM, NT, 5), ref["cn"][NT:NT + 5])
    assert np.array_equal(ctx.read(cuda_pkg.BUF_COLACC, 0, NT), ref["ca"][:NT])
    assert np.array_equal(ctx.read(cuda_pkg.BUF_GRADIENT, 0, NP), ref["gr"])
    assert np.array_equal(bits(ctx.read(cuda_pkg.BUF_POINTS, 0, NP)), bits(ref["pts"]))
    assert np.array_equal(ctx.read(cuda_pkg.BUF_INDEX, 0, NT), tri)
    assert not ctx.read(cuda_pkg.BUF_PENERGY, 0, 13 * NT).any()
    ctx.write(cuda_pkg.BUF_POINTS, 0, pts)
    ctx.step(0, cuda_pkg.TRIANGULATE, 1)
    assert np.array_equal(ctx.read(cuda_pkg.BUF_TENERGY, 0, 13 * NT), ref["ten"])
    ctx.close()


def test_batch_equals_individual_contexts(cuda_pkg, synth, O):
    """Independent jobs sharing launches: ragged NT, different images; results identical to solo runs."""
    cases = [(256, 160, 150, 31), (320, 200, 1000, 32), (200, 120, 150, 33)]
    solo, ctxs = [], []
    for (w, h, ntri, seed) in cases:
        ratio = w / h
        img = synth.make_image(w, h, seed=seed)
        pts, tri, _ = synth.mesh_for(ntri, ratio)
        c = _ctx(cuda_pkg)
        c.set_image(0, img); c.upload(pts, tri, ratio)
        c.step(0, cuda_pkg.TRIANGULATE, 4)
        solo.append(c.retrieve()); c.close()
        c = _ctx(cuda_pkg)
        c.set_image(0, img); c.upload(pts, tri, ratio)
        ctxs.append(c)
    # contexts need not have done the same number of steps before they join a batch
    ctxs[1].step(0, cuda_pkg.TRIANGULATE, 1)
    cuda_pkg.batch_step(ctxs, cuda_pkg.TRIANGULATE, 3)
    cuda_pkg.batch_step([ctxs[0], ctxs[2]], cuda_pkg.TRIANGULATE, 1)
    for c, s in zip(ctxs, solo):
        got = c.retrieve()
        for k in ("cn", "ca", "ten"):
            assert np.array_equal(got[k], s[k])
        assert np.array_equal(bits(got["pts"]), bits(s["pts"]))
    # batch_run: each job stops on its own convergence, same step counts as solo runs
    tol = 2e-3
    for c, (w, h, ntri, seed) in zip(ctxs, cases):
        pts, tri, _ = synth.mesh_for(ntri, w / h)
        c.upload(pts, tri, w / h); c.set_toterr(1.0)
    done, trig = cuda_pkg.batch_run(ctxs, cuda_pkg.TRIANGULATE, 40, tol)
    for i, (w, h, ntri, seed) in enumerate(cases):
        ratio = w / h
        img = synth.make_image(w, h, seed=seed)
        pts, tri, _ = synth.mesh_for(ntri, ratio)
        c = _ctx(cuda_pkg)
        c.set_image(0, img); c.upload(pts, tri, ratio)
        n, fl, _ = c.run(0, cuda_pkg.TRIANGULATE, 40, tol)
        assert (n, fl) == (done[i], trig[i])
        a, b = c.retrieve(), ctxs[i].retrieve()
        assert np.array_equal(a["ten"], b["ten"]) and np.array_equal(bits(a["pts"]), bits(b["pts"]))
        c.close()
    for c in ctxs:
        c.close()


def test_headless_view_matches_the_id_raster(cuda_pkg, synth, O):
    """tpose_cuda_render (software/view without a window): flat triangle colours where the oracle's
    triangle-ID raster says a triangle is drawn, black elsewhere; s interpolates towards the other
    embedding like view/shader/triangle.vs:44-46."""
    w, h = 300, 200
    ratio = w / h
    pts, tri, _ = synth.mesh_for(150, ratio)
    rng = np.random.default_rng(3)
    colors = np.zeros((tri.shape[0], 4), np.int32)
    colors[:, :3] = rng.integers(0, 256, (tri.shape[0], 3))
    colors[3, :3] = (300, -5, 128)                                   # clamped like an RGBA8 target
    colors[:, 3] = 1
    c = _ctx(cuda_pkg)
    c.upload(pts, tri, ratio, colors)
    other = pts.copy()
    other[4:] += rng.normal(0, 0.01, other[4:].shape).astype(np.float32)
    for s in (0.0, 0.25, 1.0):
        got = c.render(w, h, other, s)
        f32 = np.float32
        mixed = (pts * (f32(1.0) - f32(s))).astype(f32) + (other * f32(s)).astype(f32)
        ids = O.raster_ids(mixed.astype(f32), tri, ratio, w, h)
        want = np.zeros((h, w, 4), np.uint8); want[..., 3] = 255
        drawn = ids >= 0
        want[drawn, :3] = np.clip(colors[ids[drawn], :3], 0, 255).astype(np.uint8)
        assert np.array_equal(got, want), (s, int((got != want).any(axis=2).sum()))
    assert np.array_equal(c.render(w, h), c.render(w, h, other, 0.0))
    # the viewer's wireframe (`n` key): white one-pixel outlines on top of the same picture
    plain, wired = c.render(w, h), c.render(w, h, lines=True)
    wm = O.wire_pixels(pts, tri, ratio, w, h)
    assert wm.sum() > 2 * (w + h)
    assert (wired[wm] == 255).all() and np.array_equal(wired[~wm], plain[~wm])
    c.close()


def test_far_offscreen_vertices_take_the_wide_walk(cuda_pkg, O):
    """Vertices thousands of rows outside the window (snap clamp at +-32768 px): edge heights beyond
    2^22 sub-pixels send k_edges down its 64-bit remainder path; results still equal the oracle."""
    rng = np.random.default_rng(21)
    W, H = 96, 64
    ratio = W / H
    img = rng.integers(0, 256, size=(H, W, 4)).astype(np.uint8)
    ctx = _ctx(cuda_pkg, 64, 256)
    ctx.set_image(0, img)
    gr = None
    for trial in range(12):
        NT = 6
        NP = 4 + 3 * NT
        pts = np.stack([rng.uniform(-ratio, ratio, NP), rng.uniform(-1, 1, NP)], 1).astype(np.float32)
        far = rng.choice(NP, 6, replace=False)
        pts[far, 1] = rng.choice([-1.0, 1.0], 6) * rng.uniform(300, 1200, 6)      # up to 38 000 rows away (clamped)
        pts[far[:2], 0] = rng.choice([-1.0, 1.0], 2) * rng.uniform(200, 900, 2)
        tri = np.zeros((NT, 4), np.int32)
        tri[:, :3] = rng.integers(0, NP, size=(NT, 3))
        for fl in (cuda_pkg.TRIANGULATE, cuda_pkg.WARP):
            cols = rng.integers(0, 256, size=(NT, 4)).astype(np.int32)
            ctx.upload(pts, tri, ratio, cols)
            ctx.step(0, fl, 1)
            got = ctx.retrieve()
            ref = O.step(img, pts, tri, ratio, fl, colors=cols, gr=gr)
            gr = ref["gr"]
            _assert_step_equal(got, ref, ("cn", "ten"))
    ctx.close()


def test_maximum_width_image_takes_the_wide_sums(cuda_pkg, O):
    """4095 x 2060 (the table's width limit; 255*W*H >= 2^31 switches k_edges to 64-bit colour sums),
    saturated pixels, two triangles: counts, colour sums and (wrapping) energies equal the oracle."""
    W, H = 4095, 2060
    ratio = W / H
    rng = np.random.default_rng(4)
    img = np.full((H, W, 4), 255, np.uint8)
    img[::7, ::5, :3] = rng.integers(0, 256, size=img[::7, ::5, :3].shape)
    pts = np.array([[-ratio, -1], [-ratio, 1], [ratio, -1], [ratio, 1], [0.1, 0.05]], np.float32)
    tri = np.array([[0, 1, 4, 0], [1, 3, 4, 0], [3, 2, 4, 0], [2, 0, 4, 0]], np.int32)
    ctx = _ctx(cuda_pkg, 64, 64)
    ctx.set_image(0, img)
    ctx.upload(pts, tri, ratio)
    ctx.step(0, cuda_pkg.TRIANGULATE, 1)
    got = ctx.retrieve()
    ref = O.step(img, pts, tri, ratio, O.TRIANGULATE)
    _assert_step_equal(got, ref, ("cn", "ca", "ten"))
    assert int(got["cn"][:4].sum()) == W * H
    ctx.close()


def test_run_history_equals_stepping_frame_by_frame(cuda_pkg, synth, O):
    """tpose_cuda_set_history: every frame of a run -- points left, base energies evaluated, toterr
    after its geterr, watch mask at its points -- equals what single-frame runs of a twin context
    produce, and the recording does not change the run itself."""
    w, h = 240, 160
    ratio = w / h
    img = synth.make_image(w, h, seed=41)
    pts, tri, he = synth.mesh_for(150, ratio)
    pts = pts.copy()
    pts[60] = pts[61] + np.float32([0.02, 0.0])          # a short edge that the steps pull shorter
    a = _ctx(cuda_pkg); b = _ctx(cuda_pkg)
    for c in (a, b):
        c.set_image(0, img); c.upload(pts, tri, ratio); c.watch(tri, he)
    a.set_history(12)
    n, flags, _ = a.run(0, cuda_pkg.TRIANGULATE, 12, 1e-12)
    assert n >= 1
    for f in range(n):
        nb, fb, _ = b.run(0, cuda_pkg.TRIANGULATE, 1, 1e-12)          # one frame, looked at like the reference does
        assert nb == 1
        hf = a.history_frame(f)
        rb = b.retrieve()
        assert np.array_equal(bits(hf["pts"]), bits(rb["pts"])), f
        assert np.array_equal(hf["ten"], rb["ten"][:tri.shape[0]]), f
        assert hf["toterr"] == b.get_toterr(), f
        if fb & cuda_pkg.TRIG_TOPOLOGY:
            assert np.array_equal(hf["mask"], b.watch_mask()), f
        else:
            assert not hf["mask"].any(), f
    # with the history on, the 0.8*pi flips alone do not stop a run; prune / collapse candidates do
    if flags & cuda_pkg.TRIG_TOPOLOGY:
        assert a.history_frame(n - 1)["mask"].any()
    ra, rb = a.retrieve(), b.retrieve()
    assert np.array_equal(bits(ra["pts"]), bits(rb["pts"])) and np.array_equal(ra["ten"], rb["ten"])
    with pytest.raises(cuda_pkg.TposeCudaError):
        a.run(0, cuda_pkg.TRIANGULATE, 13, 1e-12)         # beyond the history capacity
    with pytest.raises(cuda_pkg.TposeCudaError):
        a.history_frame(n)                                # not recorded
    a.set_history(0)
    a.run(0, cuda_pkg.TRIANGULATE, 20, 1e-12)             # off again: any length
    a.close(); b.close()


@pytest.mark.parametrize("pinned", [False, True])
def test_frame_call_equals_upload_step_retrieve(cuda_pkg, synth, O, pinned):
    """tpose_cuda_frame = pointbuf fill + fused step + the frame's read-backs, in one call.  With page-locked
    caller buffers (pinned) the frame is one graph that moves the data with kernels over mapped memory; with
    pageable ones it goes through the copy engine.  Same bytes either way."""
    w, h = 256, 160
    ratio = w / h
    img = synth.make_image(w, h, seed=9)
    pts, tri, _ = synth.mesh_for(150, ratio)
    nv = 13 * tri.shape[0]
    a, b = _ctx(cuda_pkg), _ctx(cuda_pkg)
    for c in (a, b):
        c.set_image(0, img); c.upload(pts, tri, ratio)
    alloc = cuda_pkg.pinned_empty if pinned else (lambda shape, dt: np.empty(shape, dt))
    cur = alloc(pts.shape, np.float32); cur[...] = pts
    ten = alloc((nv,), np.int32); cn = alloc((nv,), np.int32); ca = alloc((nv, 4), np.int32); out = alloc(pts.shape, np.float32)
    for i in range(3):
        a.frame(0, cuda_pkg.TRIANGULATE, pts_in=cur if i else None, ten=ten, cn=cn, ca=ca, pts_out=out)
        if i:
            b.upload_points(cur)
        b.step(0, cuda_pkg.TRIANGULATE, 1)
        r = b.retrieve()
        assert np.array_equal(ten, r["ten"]) and np.array_equal(cn, r["cn"]) and np.array_equal(ca, r["ca"])
        assert np.array_equal(bits(out), bits(r["pts"]))
        cur[...] = out
        cur[10] += np.float32(0.001)                      # the host edits a point between frames
    ref = O.step(img, pts, tri, ratio, O.TRIANGULATE)
    a.upload(pts, tri, ratio)
    a.frame(0, cuda_pkg.TRIANGULATE, ten=ten, pts_out=out)
    assert np.array_equal(ten, ref["ten"]) and np.array_equal(bits(out), bits(ref["pts"]))
    a.close(); b.close()


def test_eval_list_equals_rows_of_the_full_evaluation(cuda_pkg, synth, O, kernel_path):
    """tpose_cuda_eval_list: instances evaluated on their own (the device half of the flip set,
    triangulate/main.cpp:287-306) give the rows the full 13*NT evaluation holds for them -- also for
    triangles that are not part of the uploaded mesh (a flipped pair)."""
    if kernel_path == "tiles":
        pytest.skip("eval_list walks the row-prefix table")
    w, h = 320, 200
    ratio = w / h
    img = synth.make_image(w, h, seed=14)
    pts, tri, he = synth.mesh_for(1000, ratio)
    NT = tri.shape[0]
    ctx = _ctx(cuda_pkg)
    ctx.set_image(0, img)
    ctx.upload(pts, tri, ratio)
    ref = O.step(img, pts, tri, ratio, O.TRIANGULATE)
    rng = np.random.default_rng(0)
    en = np.concatenate([tri[:, :3], np.zeros((NT, 1), np.int32)], 1)               # every base row
    extra = [(int(t), int(k)) for t, k in zip(rng.integers(0, NT, 40), rng.integers(1, 13, 40))]
    en = np.concatenate([en, np.array([[*tri[t, :3], k] for t, k in extra], np.int32)])
    ten, cn, ca = ctx.eval_list(0, en)
    assert np.array_equal(ten[:NT], ref["ten"][:NT]) and np.array_equal(cn[:NT], ref["cn"][:NT])
    assert np.array_equal(ca[:NT, :3], ref["ca"][:NT, :3])
    for i, (t, k) in enumerate(extra):
        g = k * NT + t
        assert ten[NT + i] == ref["ten"][g] and cn[NT + i] == ref["cn"][g] and np.array_equal(ca[NT + i, :3], ref["ca"][g, :3])
    # a flipped pair: triangles (a, b, c) + (c, b, d) -> (a, b, d) + (a, d, c); compare with a mesh that holds them
    t0 = 2 * 37
    a, b, c = tri[t0, :3]; d = tri[t0 + 1, 2]
    tri2 = tri.copy(); tri2[t0, :3] = (a, b, d); tri2[t0 + 1, :3] = (a, d, c)
    ref2 = O.step(img, pts, tri2, ratio, O.TRIANGULATE)
    ten2, cn2, _ = ctx.eval_list(0, np.array([[a, b, d, 0], [a, d, c, 0]], np.int32))
    assert np.array_equal(ten2, ref2["ten"][[t0, t0 + 1]]) and np.array_equal(cn2, ref2["cn"][[t0, t0 + 1]])
    with pytest.raises(cuda_pkg.TposeCudaError):
        ctx.eval_list(0, np.array([[0, 1, 99999, 0]], np.int32))
    ctx.close()


def test_run_results_reach_the_host_without_copy_engine_operations(cuda_pkg, synth):
    """A run's verdict block and recorded frames are pushed into the context's page-locked buffers by one kernel
    behind the chain (k_run_out): 2 * (max_steps + 1) evaluation launches + 1, with and without the history."""
    w, h = 160, 120
    ratio = w / h
    img = synth.make_image(w, h, seed=2)
    pts, tri, he = synth.mesh_for(150, ratio)
    c = _ctx(cuda_pkg)
    c.set_image(0, img); c.upload(pts, tri, ratio)
    per_eval = 3 if c.kernel_path() == "tiles" else 2
    for hist in (0, 8):
        c.set_history(hist)
        c.run(0, cuda_pkg.TRIANGULATE, 5, 1e-30)
        n0 = c.launch_count()
        done, _, _ = c.run(0, cuda_pkg.TRIANGULATE, 5, 1e-30)
        assert done == 5
        assert c.launch_count() - n0 == per_eval * 6 + 1
    c.close()


def test_geterr_on_the_device_equals_the_serial_float_sum(cuda_pkg):
    """tpose_cuda_geterr = geterr_block (csrc/eval.cuh), the routine every frame of a run ends with: the float32 sum of
    the energies in index order (source/triangulation.hpp:653-674), computed from per-chunk summaries that depend on
    one bit of the running sum.  Bit-equal to the literal serial loop on plain, adversarial (ties in every binade,
    powers of two, zeros, negative energies, int -> float rounding of the addends) and large inputs."""
    from geterr_model import serial_f32, cases
    rng = np.random.default_rng(23)
    inputs = list(cases(rng, 150)) + [np.abs(rng.normal(1.5e9 / n, 0.5e9 / n, n)).astype(np.int32) for n in (150, 3000, 10000, 40000)]
    inputs += [np.zeros(0, np.int32), np.array([7], np.int32), np.array([2 ** 24, 1, 1, 1], np.int32), np.full(5000, 1, np.int32)]
    for x in inputs:
        want = serial_f32(x)
        tot = np.float32(rng.uniform(1.0, 3e9))
        wrel = np.float32(abs(np.float32((tot - want) / tot)))
        for route in (1, 0, -1):                      # the block routine, the plain chain, what a run picks
            got, rel, below = cuda_pkg.geterr(x, toterr=float(tot), rel_tol=1e-4, route=route)
            assert got.view(np.uint32) == want.view(np.uint32), (len(x), route, got, want)
            assert rel.view(np.uint32) == wrel.view(np.uint32) and below == (float(wrel) < 1e-4)


def test_run_on_a_large_mesh_takes_the_block_geterr_and_matches_the_serial_sum(cuda_pkg, synth):
    """Above 3 072 triangles the check a run's frames end with is geterr_block (csrc/eval.cuh) inside k_edges' service block:
    after every run of one frame, toterr and |relerr| equal the serial float32 sum over the base energies the device holds
    (source/triangulation.hpp:653-674), bit for bit; and a run of several frames equals those single-frame runs."""
    from geterr_model import serial_f32
    w, h = 1280, 720
    ratio = w / h
    img = synth.make_image(w, h, seed=8)
    pts, tri, _ = synth.mesh_for(10000, ratio)
    NT = tri.shape[0]
    a = cuda_pkg.Context(device=0, max_triangles=NT, max_points=pts.shape[0])
    b = cuda_pkg.Context(device=0, max_triangles=NT, max_points=pts.shape[0])
    for c in (a, b):
        c.set_image(0, img); c.upload(pts, tri, ratio)
    toterr = np.float32(1.0)
    for frame in range(4):
        n, flags, err = a.run(0, cuda_pkg.TRIANGULATE, 1, 1e-30)
        assert n == 1
        want = serial_f32(a.retrieve()["ten"][:NT])
        rel = np.float32(abs(np.float32((toterr - want) / toterr)))
        assert np.float32(a.get_toterr()).view(np.uint32) == want.view(np.uint32)
        assert np.float32(err).view(np.uint32) == rel.view(np.uint32)
        toterr = want
    n, flags, err = b.run(0, cuda_pkg.TRIANGULATE, 4, 1e-30)
    assert n == 4 and np.float32(b.get_toterr()).view(np.uint32) == toterr.view(np.uint32)
    assert np.float32(err).view(np.uint32) == rel.view(np.uint32)
    assert np.array_equal(a.retrieve()["ten"], b.retrieve()["ten"])
    a.close(); b.close()
