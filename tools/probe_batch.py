"""Device timing of batched steps: B independent 1080p images advance in lock step, one launch
pair per step for all of them (development aid; bench.py reports the same numbers)."""
import os, sys, json, argparse
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
import importlib
import numpy as np

ap = argparse.ArgumentParser()
ap.add_argument("--batches", default="1,8,32,64,128")
ap.add_argument("--ntri", type=int, default=3000)
ap.add_argument("--steps", type=int, default=64)
ap.add_argument("--groups", default="0")
a = ap.parse_args()
pkg = entry.load_package()
synth = importlib.import_module("tpose_b200.synth")
w, h = 1920, 1080
ratio = w / h
pts, tri, _ = synth.mesh_for(a.ntri, ratio)
NT, NP = tri.shape[0], pts.shape[0]
balg = 4 * w * h + 16 * NT + 16 * NP
imgs = [synth.make_image(w, h, seed=1000 + i) for i in range(4)]
out = []
for g in [int(x) for x in a.groups.split(",")]:
    os.environ["TPOSE_CUDA_GROUP"] = str(g)
    for B in [int(x) for x in a.batches.split(",")]:
        ctxs = [pkg.Context(0, max(NT, 16), max(NP, 16)) for _ in range(B)]
        for i, c in enumerate(ctxs):
            c.set_image(0, imgs[i % len(imgs)])
            c.upload(pts, tri, ratio)
        pkg.batch_step(ctxs, pkg.TRIANGULATE, 8)
        ctxs[0].sync()
        pkg.batch_step(ctxs, pkg.TRIANGULATE, a.steps)
        ctxs[0].sync()
        ms = ctxs[0].last_elapsed_ms()
        us = 1000 * ms / a.steps / B
        rec = dict(batch=B, ntri=NT, group=g, us_per_image_step=us, image_steps_per_s=1e6 / us,
                   algorithmic_GBps=balg / us / 1e3)
        print(json.dumps(rec), flush=True)
        out.append(rec)
        for c in ctxs:
            c.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe_batch.json"), "w"), indent=1)
