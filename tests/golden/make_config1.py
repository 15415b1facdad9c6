"""Regenerates the fixtures of BASELINE.json config 1 and of the soft pin to the reference's traces.

Run in the build container (needs /root/reference and PIL); the outputs travel to the GPU box.

  fruit_512.png        centre 512x512 crop (x0 = 249, y0 = 81) of
                       /root/reference/resource/fruit.png (1011x674 RGBA) -- no resampling
  config1_digests.json SHA-256 of the oracle's (ten, cn, ca, gr, points) after steps 1, 2, 10, 25, 50 of
                       the frozen 15x5x2 = 150-triangle grid, at the native 512x512 raster and at the
                       reference's 1/1.5 window (341x341)
  kat_traces.json      Sum of terr[0..4) after the first split of the two-triangle mesh, as the
                       reference's tests/opt_min_energy_edgeflip program prints it, computed by the
                       oracle on resource/imageA.png / imageB.png at the 800x450 window
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference"
CROP_X0, CROP_Y0, CROP = 249, 81, 512
DIGEST_STEPS = (1, 2, 10, 25, 50)


def fruit_crop():
    from PIL import Image
    a = np.array(Image.open(os.path.join(REF, "resource", "fruit.png")).convert("RGBA"))
    assert a.shape == (674, 1011, 4)
    return np.ascontiguousarray(a[CROP_Y0:CROP_Y0 + CROP, CROP_X0:CROP_X0 + CROP])


def config1_digests(O, synth, img, raster=None):
    """The frozen-topology half of config 1 (1a): 50 steps; digest chain over every step's outputs."""
    ratio = 1.0
    if raster is not None:
        img = O.resample(img, raster[0], raster[1])
    pts, tri, _ = synth.mesh_for(150, ratio)
    h = hashlib.sha256()
    out, cur, gr = {}, pts, None
    for it in range(1, max(DIGEST_STEPS) + 1):
        r = O.step(img, cur, tri, ratio, O.TRIANGULATE, gr=gr)
        for k in ("ten", "cn", "ca", "gr", "pts"):
            h.update(np.ascontiguousarray(r[k]).tobytes())
        cur, gr = r["pts"], r["gr"]
        if it in DIGEST_STEPS:
            out[str(it)] = h.hexdigest()
    return out


def kat_first_split(O, H, rgba, strategy):
    """tests/opt_min_energy_edgeflip/main.cpp:168-196 up to its first printed line.  The corners
    never move (shift.cs:20), so the two-triangle mesh is static: geterr is 0 on the second frame,
    the strategy header runs, the split sets `updated`, and the program prints NT = 4 and gettoterr
    = float sum of terr[0..4) -- energies of the two-triangle evaluation read at the new NT:
    the two base rows and the first moved variant of both triangles.
      strategy 'set'  : strat/flip_set_maxenergy.h -- flip the diagonal, keep it if the pair's energy
                        did not rise (:96-113), re-evaluate, split
      strategy 'none' : strat/flip_delaunay.h      -- split only
    The shader copy of that test cools dp like the warp shader (its triangle.vs:70)."""
    h, w = rgba.shape[:2]
    rw, rh = int(w / 1.5), int(h / 1.5)                  # Tiny::window(IMG->w/1.5f, IMG->h/1.5f), main.cpp:36
    ratio = float(np.float32(w) / np.float32(h))
    img = O.resample(rgba, rw, rh)
    R = H.ref()
    R.set_ratio(ratio)
    tr = H.Tri(R)
    O.set_dp_law(O.WARP)
    try:
        def evaluate():
            s = tr.get()
            cn, ca = O.mode0(img, s["pts"], s["tri"], ratio, O.TRIANGULATE)
            ten, _ = O.mode1(img, s["pts"], s["tri"], ratio, O.TRIANGULATE, cn, ca)
            return ten
        terr = evaluate()
        if strategy == "set":
            f32 = np.float32
            before = f32(f32(terr[0]) + f32(terr[1]))
            ha = 1                                        # the only interior half-edge pair is (1, 3); std::set keeps the first
            tr.flip(ha, 0.0)
            terr = evaluate()
            he = tr.get()["he"]
            if f32(f32(terr[ha // 3]) + f32(terr[he[ha] // 3])) > before:
                tr.flip(ha, 0.0)
            terr = evaluate()
        tot = np.float32(0)
        for i in range(4):
            tot = np.float32(tot + np.float32(terr[i]))
        return float(tot)
    finally:
        O.set_dp_law(-1)
        tr.close()


def main():
    from PIL import Image
    import importlib
    import __graft_entry__ as entry
    import hostlib as H
    entry.load_package()
    synth = importlib.import_module("tpose_b200.synth")
    O = entry.load_oracle()
    crop = fruit_crop()
    Image.fromarray(crop).save(os.path.join(HERE, "fruit_512.png"), optimize=True)
    dig = {"native_512x512": config1_digests(O, synth, crop),
           "window_341x341": config1_digests(O, synth, crop, (int(512 / 1.5), int(512 / 1.5))),
           "what": "sha256 chain over (ten, cn, ca, gr, pts) of every step up to the keyed step; oracle/step_oracle.c"}
    json.dump(dig, open(os.path.join(HERE, "config1_digests.json"), "w"), indent=1)
    kat = {}
    for name, strat, trace in (("imageB.png", "set", "flip_set_maxenergy.txt"), ("imageA.png", "none", "flip_delaunay.txt")):
        rgba = np.array(Image.open(os.path.join(REF, "resource", name)).convert("RGBA"))
        first = open(os.path.join(REF, "tests", "opt_min_energy_edgeflip", "strat", trace)).readline().split()
        val = kat_first_split(O, H, rgba, strat)
        kat[trace] = {"image": name, "strategy": strat, "reference_line": [int(first[0]), float(first[1])], "oracle": val,
                      "rel_diff": val / float(first[1]) - 1.0}
    kat["note"] = ("the traces do not record their input image; imageB / imageA are inferred (SURVEY.md section 4). "
                   "Agreement to a fraction of a percent pins orientation, coverage, the energy definition and the "
                   "13*NT row layout softly; it cannot pin the fill rule or the filter's weight precision.")
    json.dump(kat, open(os.path.join(HERE, "kat_traces.json"), "w"), indent=1)
    print(json.dumps(kat, indent=1))
    print(json.dumps(dig, indent=1))


if __name__ == "__main__":
    main()
