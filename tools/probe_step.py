"""Quick device timing of the fused step for a few lane-group sizes (development aid)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
import importlib
import numpy as np

pkg = entry.load_package()
synth = importlib.import_module("tpose_b200.synth")
w, h = 1920, 1080
ratio = w / h
img = synth.make_image(w, h, seed=1000)
out = []
for ntri in (1000, 3000, 10000, 150):
    pts, tri, _ = synth.mesh_for(ntri, ratio)
    for g in (0, 1, 2, 4, 8, 16, 32):
        os.environ["TPOSE_CUDA_GROUP"] = str(g)
        ctx = pkg.Context(0, 16384, 16384)
        t0 = time.perf_counter(); ctx.set_image(0, img); ctx.sync(); t_img = time.perf_counter() - t0
        t0 = time.perf_counter(); ctx.set_image(0, img); ctx.sync(); t_img2 = time.perf_counter() - t0
        ctx.upload(pts, tri, ratio)
        ctx.step(0, 0, 50); ctx.sync()
        ctx.upload(pts, tri, ratio)
        ctx.step(0, 0, 500); ctx.sync()
        ms = ctx.last_elapsed_ms()
        rec = dict(ntri=ntri, group=g, us_per_step=1000 * ms / 500, steps_per_s=500 / (ms / 1000), set_image_ms=t_img2 * 1000)
        print(json.dumps(rec), flush=True)
        out.append(rec)
        ctx.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe_step.json"), "w"), indent=1)
