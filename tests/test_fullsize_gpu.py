"""GPU parity at the sizes BASELINE.json names and bench.py times: the CUDA path through the C ABI
against the per-fragment CPU oracle, bit for bit.

  * 1920x1080, 1 000 / 3 000 / 10 000 triangles, triangulate and warp flavour, three consecutive steps,
    default lane policy and a forced one (what BENCH times is the 3 000-triangle triangulate case);
  * a 64-image batch_step at 1080p / 3 000 (bench.py's `batched` workload), two sampled jobs checked;
  * config 1 (configs[0]): centre 512x512 crop of the reference's resource/fruit.png, 150 triangles,
    50 steps, every step compared; also at the reference's 1/1.5 window (341x341);
  * the 1/1.5 window at 1200x675 -> 800x450 (software/triangulate/main.cpp:53).

Raster parity is parity with oracle/step_oracle.c -- the conventions a GL driver would decide are
written down there (and soft-pinned to the reference's traces by tests/test_kat_cpu.py).
"""
import functools
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import bits

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("kernel_path")]
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@functools.lru_cache(maxsize=8)
def _pair_1080p(seed):
    import importlib
    synth = importlib.import_module("tpose_b200.synth")
    return synth.make_pair(1920, 1080, seed=seed)


def _equal_step(got, ref, keys, gr=None):
    for k in keys:
        assert np.array_equal(got[k], ref[k]), f"{k} differs at rows {np.flatnonzero((got[k] != ref[k]).reshape(len(ref[k]), -1).any(1))[:8]}"
    assert np.array_equal(bits(got["pts"]), bits(ref["pts"])), "points differ"
    if gr is not None:
        assert np.array_equal(gr, ref["gr"][:len(gr)]), "gradient differs"


@pytest.mark.parametrize("flavour", ["triangulate", "warp"])
@pytest.mark.parametrize("ntri,group", [(1000, 0), (3000, 0), (3000, 8), (10000, 0), (10000, 2)])
def test_1080p_steps_match_oracle(cuda_pkg, synth, O, ntri, group, flavour, monkeypatch):
    if group:
        monkeypatch.setenv("TPOSE_CUDA_GROUP", str(group))
    else:
        monkeypatch.delenv("TPOSE_CUDA_GROUP", raising=False)
    w, h = 1920, 1080
    ratio = w / h
    a, b = _pair_1080p(1000)
    pts, tri, _ = synth.mesh_for(ntri, ratio)
    NT = tri.shape[0]
    ctx = cuda_pkg.Context(device=0, max_triangles=NT, max_points=pts.shape[0])
    ctx.set_image(0, a)
    colors = None
    if flavour == "warp":
        # colours of A's mesh over image A (what triangulate exports, main.cpp:221-224), then A's mesh is
        # scored against image B with dp = 0.05/(1+9*NT/1000): 0.64 px at 3 000 triangles
        r0 = O.step(a, pts, tri, ratio, O.TRIANGULATE)
        colors = np.ones((NT, 4), np.int32)
        colors[:, :3] = r0["ca"][:NT, :3] // np.maximum(r0["cn"][:NT, None], 1)
        ctx.set_image(1, b)
    ctx.upload(pts, tri, ratio, colors)
    cur, gr = pts, None
    for it in range(3):
        if flavour == "warp":
            ctx.step(1, cuda_pkg.WARP, 1)
            ref = O.step(b, cur, tri, ratio, O.WARP, colors=colors, gr=gr)
            got = ctx.retrieve()
            _equal_step(got, ref, ("cn", "ten"), ctx.retrieve_gradient())
            assert np.array_equal(got["ca"][:, :3], ref["ca"][:, :3])
        else:
            ctx.step(0, cuda_pkg.TRIANGULATE, 1)
            ref = O.step(a, cur, tri, ratio, O.TRIANGULATE, gr=gr)
            _equal_step(ctx.retrieve(), ref, ("cn", "ca", "ten"), ctx.retrieve_gradient())
        cur, gr = ref["pts"], ref["gr"]
    if flavour == "triangulate":
        assert int(ref["cn"][:NT].sum()) == w * h       # the base triangles tile the window
    ctx.close()


def test_batch_of_64_images_1080p_sampled_jobs_match_oracle(cuda_pkg, synth, O):
    """bench.py's `batched` workload: 64 contexts, 4 distinct images, one launch set per step."""
    w, h, ntri, B = 1920, 1080, 3000, 64
    ratio = w / h
    pool = [synth.make_image(w, h, seed=s) for s in (1000, 2001, 2002, 2003)]
    pts, tri, _ = synth.mesh_for(ntri, ratio)
    ctxs = []
    for i in range(B):
        c = cuda_pkg.Context(device=0, max_triangles=tri.shape[0], max_points=pts.shape[0])
        c.set_image(0, pool[i % 4])
        c.upload(pts, tri, ratio)
        ctxs.append(c)
    cuda_pkg.batch_step(ctxs, cuda_pkg.TRIANGULATE, 2)
    # a non-lead context is read first: its stream must have been ordered behind the batch
    for job in (37, 0, 63):
        img = pool[job % 4]
        r1 = O.step(img, pts, tri, ratio, O.TRIANGULATE)
        r2 = O.step(img, r1["pts"], tri, ratio, O.TRIANGULATE, gr=r1["gr"])
        _equal_step(ctxs[job].retrieve(), r2, ("cn", "ca", "ten"), ctxs[job].retrieve_gradient())
    for c in ctxs:
        c.close()


def _fruit():
    from PIL import Image
    return np.array(Image.open(os.path.join(GOLD, "fruit_512.png")).convert("RGBA"))


@pytest.mark.parametrize("key,raster", [("native_512x512", None), ("window_341x341", (341, 341))])
def test_config1_fruit_150_triangles_50_steps(cuda_pkg, synth, O, key, raster):
    """configs[0]: every one of the 50 steps bit-equal to the oracle (ten, cn, ca, gr, points), and the
    digest chain over the CUDA results equals the one committed from the oracle run in the build
    container (tests/golden/make_config1.py)."""
    img = _fruit()
    ratio = 1.0
    pts, tri, _ = synth.mesh_for(150, ratio)
    ctx = cuda_pkg.Context(device=0, max_triangles=256, max_points=256)
    ctx.set_image(0, img, raster=raster)
    ctx.upload(pts, tri, ratio)
    oimg = img if raster is None else O.resample(img, raster[0], raster[1])
    want = json.load(open(os.path.join(GOLD, "config1_digests.json")))[key]
    hsh = hashlib.sha256()
    cur, gr = pts, None
    for it in range(1, 51):
        ctx.step(0, cuda_pkg.TRIANGULATE, 1)
        got = ctx.retrieve()
        ggr = ctx.retrieve_gradient()
        ref = O.step(oimg, cur, tri, ratio, O.TRIANGULATE, gr=gr)
        _equal_step(got, ref, ("cn", "ca", "ten"), ggr)
        for k in ("ten", "cn", "ca"):
            hsh.update(np.ascontiguousarray(got[k]).tobytes())
        hsh.update(np.ascontiguousarray(ggr).tobytes())
        hsh.update(np.ascontiguousarray(got["pts"]).tobytes())
        if str(it) in want:
            assert hsh.hexdigest() == want[str(it)], f"digest after step {it}"
        cur, gr = ref["pts"], ref["gr"]
    ctx.close()


@pytest.mark.parametrize("flavour", ["triangulate", "warp"])
def test_reference_window_raster_truncated_filtered_texel(cuda_pkg, synth, O, flavour):
    """raster = (w/1.5, h/1.5) = 800x450 of a 1200x675 image (Tiny::window, triangulate/main.cpp:53; the size
    of the reference's own resource/imageA.png).  What is compared: the CUDA resampler + step against the
    oracle's resampler + step, both with the written-down float32 bilinear formula and the filtered texel
    TRUNCATED to a byte in both modes (the reference's mode 1 keeps the fraction, triangle.fs:40-42 --
    stated divergence)."""
    w, h = 1200, 675
    rw, rh = int(w / 1.5), int(h / 1.5)
    ratio = w / h
    a, b = synth.make_pair(w, h, seed=77)
    ra, rb = O.resample(a, rw, rh), O.resample(b, rw, rh)
    pts, tri, _ = synth.mesh_for(1000, ratio)
    NT = tri.shape[0]
    ctx = cuda_pkg.Context(device=0, max_triangles=NT, max_points=pts.shape[0])
    ctx.set_image(0, a, raster=(rw, rh))
    ctx.set_image(1, b, raster=(rw, rh))
    colors = None
    if flavour == "warp":
        r0 = O.step(ra, pts, tri, ratio, O.TRIANGULATE)
        colors = np.ones((NT, 4), np.int32)
        colors[:, :3] = r0["ca"][:NT, :3] // np.maximum(r0["cn"][:NT, None], 1)
    ctx.upload(pts, tri, ratio, colors)
    cur, gr = pts, None
    for it in range(3):
        if flavour == "warp":
            ctx.step(1, cuda_pkg.WARP, 1)
            ref = O.step(rb, cur, tri, ratio, O.WARP, colors=colors, gr=gr)
            _equal_step(ctx.retrieve(), ref, ("cn", "ten"), ctx.retrieve_gradient())
        else:
            ctx.step(0, cuda_pkg.TRIANGULATE, 1)
            ref = O.step(ra, cur, tri, ratio, O.TRIANGULATE, gr=gr)
            _equal_step(ctx.retrieve(), ref, ("cn", "ca", "ten"), ctx.retrieve_gradient())
        cur, gr = ref["pts"], ref["gr"]
    ctx.close()


def test_unreferenced_vertex_keeps_its_stale_gradient(cuda_pkg, synth, O):
    """triangle.vs:93-98 zeroes gr only for vertices a triangle references.  After a triangle is taken
    away (prune -> eraset, "vertices stay") its lone vertex keeps the gradient of the last frame that
    referenced it, and shift.cs:16-47 keeps applying it -- on the device as in the oracle."""
    w, h = 256, 160
    ratio = w / h
    img = synth.make_image(w, h, seed=5)
    pts, tri, _ = synth.mesh_for(150, ratio)
    ctx = cuda_pkg.Context(device=0, max_triangles=256, max_points=256)
    ctx.set_image(0, img)
    ctx.upload(pts, tri, ratio)
    ctx.step(0, cuda_pkg.TRIANGULATE, 1)
    r1 = O.step(img, pts, tri, ratio, O.TRIANGULATE)
    # drop every triangle that touches vertex v: v is now unreferenced
    v = int(tri[77, 1])
    assert v >= 4 and np.any(r1["gr"][v] != 0)
    keep = ~(tri[:, :3] == v).any(1)
    tri2 = np.ascontiguousarray(tri[keep])
    ctx.upload(r1["pts"], tri2, ratio)
    cur, gr = r1["pts"], r1["gr"]
    for it in range(3):
        ctx.step(0, cuda_pkg.TRIANGULATE, 1)
        ref = O.step(img, cur, tri2, ratio, O.TRIANGULATE, gr=gr)
        assert np.array_equal(ref["gr"][v], r1["gr"][v])                 # stale, not zero
        _equal_step(ctx.retrieve(), ref, ("cn", "ca", "ten"), ctx.retrieve_gradient())
        cur, gr = ref["pts"], ref["gr"]
    assert not np.array_equal(bits(cur[v]), bits(r1["pts"][v]))          # and it kept moving
    # the literal lambdas: doenergy resets referenced vertices only, doshift applies the rest
    ctx.colors(0, cuda_pkg.TRIANGULATE)
    ctx.energy(0, cuda_pkg.TRIANGULATE)
    ctx.shift(cuda_pkg.TRIANGULATE)
    ref = O.step(img, cur, tri2, ratio, O.TRIANGULATE, gr=gr)
    assert np.array_equal(ctx.retrieve_gradient(), ref["gr"][:len(cur)])
    assert np.array_equal(bits(ctx.retrieve()["pts"]), bits(ref["pts"]))
    ctx.close()
