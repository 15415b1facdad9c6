"""The raster/step oracle against the only numbers the reference itself holds for this path, and
against its own committed fixtures (CPU only).

1. Soft pin.  tests/opt_min_energy_edgeflip/strat/*.txt in the reference are traces "NT  sum-of-energies"
   printed at every topology change of that test program.  Their input image was not recorded; SURVEY.md
   section 4 infers resource/imageB.png (flip_set_maxenergy) and resource/imageA.png (flip_delaunay) at the
   800x450 window.  The first line of each trace is reproduced by the oracle within +-0.5 %.  This cannot
   make raster parity green (the GL driver's fill rule and filter precision stay unknown) but it ties
   orientation, coverage, energy definition and the 13*NT row layout to something the reference produced.
2. Config 1 (BASELINE.json configs[0]): centre 512x512 crop of resource/fruit.png,
   frozen 15x5x2 grid, 50 steps -- the oracle's digest chain equals the committed one.
"""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
sys.path.insert(0, GOLD)
import make_config1 as G  # noqa: E402


def test_committed_kat_values_are_within_half_a_percent_of_the_reference_traces():
    kat = json.load(open(os.path.join(GOLD, "kat_traces.json")))
    assert kat["flip_set_maxenergy.txt"]["reference_line"] == [4, 1537340672.0]       # strat/flip_set_maxenergy.txt:1
    assert kat["flip_delaunay.txt"]["reference_line"] == [4, 1577750016.0]            # strat/flip_delaunay.txt:1
    for name in ("flip_set_maxenergy.txt", "flip_delaunay.txt"):
        k = kat[name]
        assert abs(k["oracle"] / k["reference_line"][1] - 1.0) < 5e-3, (name, k)


@pytest.mark.skipif(not os.path.isdir(os.path.join(G.REF, "resource")), reason="needs /root/reference (build container only)")
@pytest.mark.skipif(not os.path.exists(os.path.join(G.ROOT, "oracle", "_ref", "libtpose_ref.so")), reason="oracle/_ref not built")
@pytest.mark.parametrize("image,strategy,trace", [("imageB.png", "set", "flip_set_maxenergy.txt"), ("imageA.png", "none", "flip_delaunay.txt")])
def test_oracle_reproduces_first_trace_line_from_the_reference_images(O, image, strategy, trace):
    """Image inferred, not recorded by the reference: see the module docstring."""
    from PIL import Image
    import hostlib as H
    rgba = np.array(Image.open(os.path.join(G.REF, "resource", image)).convert("RGBA"))
    first = open(os.path.join(G.REF, "tests", "opt_min_energy_edgeflip", "strat", trace)).readline().split()
    assert int(first[0]) == 4
    val = G.kat_first_split(O, H, rgba, strategy)
    assert abs(val / float(first[1]) - 1.0) < 5e-3, (val, first)
    assert val == json.load(open(os.path.join(GOLD, "kat_traces.json")))[trace]["oracle"]


def _fruit():
    from PIL import Image
    return np.array(Image.open(os.path.join(GOLD, "fruit_512.png")).convert("RGBA"))


@pytest.mark.skipif(not os.path.isdir(os.path.join(G.REF, "resource")), reason="needs /root/reference (build container only)")
def test_fruit_fixture_is_the_centre_crop_of_the_reference_image():
    assert np.array_equal(_fruit(), G.fruit_crop())


@pytest.mark.parametrize("key,raster", [("native_512x512", None), ("window_341x341", (341, 341))])
def test_config1_oracle_digests(O, synth, key, raster):
    want = json.load(open(os.path.join(GOLD, "config1_digests.json")))[key]
    assert G.config1_digests(O, synth, _fruit(), raster) == want
