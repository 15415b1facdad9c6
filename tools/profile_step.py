"""Runs N fused steps of one configuration (for ncu captures)."""
import os, sys, argparse
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
import importlib
import numpy as np
ap = argparse.ArgumentParser()
ap.add_argument("--ntri", type=int, default=3000)
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--steps", type=int, default=30)
ap.add_argument("--flavour", type=int, default=0)
a = ap.parse_args()
pkg = entry.load_package()
synth = importlib.import_module("tpose_b200.synth")
ratio = a.width / a.height
img = synth.make_image(a.width, a.height, seed=1000)
pts, tri, _ = synth.mesh_for(a.ntri, ratio)
ctx = pkg.Context(0, 16384, 16384)
ctx.set_image(0, img)
col = np.full((tri.shape[0], 4), 128, np.int32)
ctx.upload(pts, tri, ratio, col if a.flavour else None)
ctx.step(0, a.flavour, a.steps)
ctx.sync()
print("ms/step", ctx.last_elapsed_ms() / a.steps)
