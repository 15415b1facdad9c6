"""GPU parity tests: the CUDA path through the C ABI against the CPU oracle, bit for bit."""
import numpy as np
import pytest

from conftest import bits

pytestmark = pytest.mark.gpu


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
            ref = O.step(img, pts, tri, ratio, fl, colors=cols)
            _assert_step_equal(got, ref, ("cn", "ten"))
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


def test_full_size_properties_1080p(cuda_pkg, synth):
    """BASELINE size (1920x1080, 3000 triangles): size-independent properties instead of the oracle."""
    w, h = 1920, 1080
    ratio = w / h
    img = synth.make_image(w, h, seed=1000)
    pts, tri, _ = synth.mesh_for(3000, ratio)
    NT = tri.shape[0]
    ctx = _ctx(cuda_pkg)
    ctx.set_image(0, img)
    ctx.upload(pts, tri, ratio)
    ctx.step(0, cuda_pkg.TRIANGULATE, 1)
    got = ctx.retrieve()
    # the base triangles tile the window: counts and colour sums are checksums of the image
    assert int(got["cn"][:NT].sum()) == w * h
    assert np.array_equal(got["ca"][:NT, :3].astype(np.int64).sum(0), img[..., :3].reshape(-1, 3).astype(np.int64).sum(0))
    # determinism / idempotence of evaluation: same positions -> same rows
    ctx.upload(pts, tri, ratio)
    ctx.step(0, cuda_pkg.TRIANGULATE, 1)
    again = ctx.retrieve()
    for k in ("cn", "ca", "ten"):
        assert np.array_equal(got[k], again[k])
    assert np.array_equal(bits(got["pts"]), bits(again["pts"]))
    # corners never move (shift.cs:20)
    assert np.array_equal(bits(got["pts"][:4]), bits(pts[:4]))
    # ids raster agrees with the counts
    ctx.upload(pts, tri, ratio)
    ids = ctx.raster_ids(0)
    assert np.array_equal(np.bincount(ids.ravel(), minlength=NT), got["cn"][:NT])
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
