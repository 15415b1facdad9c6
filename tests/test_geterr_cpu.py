"""The algorithm of geterr_block (csrc/eval.cuh) -- the in-order float32 sum computed from per-chunk summaries that
depend on one bit of the running sum -- against the literal serial loop, on the CPU (tests/geterr_model.py)."""
import numpy as np

from geterr_model import serial_f32, block_f32, cases


def test_block_sum_equals_the_serial_float_sum_bit_for_bit():
    rng = np.random.default_rng(11)
    stats = dict(group=0, chunk=0, chain=0)
    for x in cases(rng, 400):
        a, b = serial_f32(x), block_f32(x, stats=stats)
        assert a.view(np.uint32) == b.view(np.uint32), (len(x), a, b)
    assert stats["group"] > 0 and stats["chunk"] > 0 and stats["chain"] > 0         # every route was taken


def test_a_frame_like_input_takes_the_fast_routes():
    rng = np.random.default_rng(3)
    for n in (150, 1256, 3000, 10000):
        x = np.abs(rng.normal(1.5e9 / n, 0.5e9 / n, n)).astype(np.int32)
        stats = dict(group=0, chunk=0, chain=0)
        a, b = serial_f32(x), block_f32(x, stats=stats)
        assert a.view(np.uint32) == b.view(np.uint32)
        assert stats["chain"] <= 16            # a handful of chunks around the powers of two, the rest through summaries


def test_small_blocks_and_edge_cases():
    for x in ([], [0], [5], [2 ** 31 - 1], [2 ** 24, 1, 1, 1], [2 ** 24 - 1, 1, 1], [1] * 5000, [-5, 7, 9]):
        for nthreads, group in ((256, 16), (32, 4), (1, 1)):
            a, b = serial_f32(x), block_f32(x, nthreads=nthreads, group=group)
            assert a.view(np.uint32) == b.view(np.uint32), (x[:4], nthreads)
