"""Device timing of one image-step on both kernel paths (development aid; bench.py reports the product's):
  tiles  k_setup + k_tiles (TMA-staged image tiles) + k_corners
  table  k_edges (row-prefix table gathers) + k_corners        (TPOSE_CUDA_PATH=table)
per mesh size: cold (L2 flushed before every step, per-kernel event times), warm (graph replay), batched.

  python tools/probe_paths.py [--ntri 1000,3000,10000] [--batch 64] [--out gpurun_out/probe_paths.json]
"""
import argparse, importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
import numpy as np

ap = argparse.ArgumentParser()
ap.add_argument("--ntri", default="1000,3000,10000")
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--paths", default="tiles,table")
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "probe_paths.json"))
a = ap.parse_args()
pkg = entry.load_package()
synth = importlib.import_module("tpose_b200.synth")
w, h = 1920, 1080
ratio = w / h
imgs = [synth.make_image(w, h, seed=s) for s in (1000, 2001, 2002, 2003)]
out = []
for path in a.paths.split(","):
    os.environ["TPOSE_CUDA_PATH"] = path
    for ntri in [int(x) for x in a.ntri.split(",")]:
        pts, tri, _ = synth.mesh_for(ntri, ratio)
        NT, NP = tri.shape[0], pts.shape[0]
        balg = 4 * w * h + 16 * NT + 16 * NP
        ctx = pkg.Context(0, max(NT, 16), max(NP, 16))
        assert ctx.kernel_path() == path
        ctx.set_image(0, imgs[0]); ctx.upload(pts, tri, ratio)
        for _ in range(5):
            ctx.flush_l2(); ctx.step(0, pkg.TRIANGULATE, 1)
        ctx.sync()
        ctx.set_kernel_timing(True)
        K = 100
        t_setup = t_main = t_corn = t_step = 0.0
        for _ in range(K):
            ctx.flush_l2(); ctx.step(0, pkg.TRIANGULATE, 1)
            m, c = ctx.last_kernel_ms(); t_main += m; t_corn += c
            t_setup += ctx.last_setup_ms(); t_step += ctx.last_elapsed_ms()
        ctx.set_kernel_timing(False)
        cold_graph = 0.0
        for _ in range(K):
            ctx.flush_l2(); ctx.step(0, pkg.TRIANGULATE, 1); cold_graph += ctx.last_elapsed_ms()
        ctx.upload(pts, tri, ratio); ctx.step(0, pkg.TRIANGULATE, 512); ctx.sync()
        ctx.upload(pts, tri, ratio); ctx.step(0, pkg.TRIANGULATE, 512); ctx.sync()
        warm = ctx.last_elapsed_ms() / 512
        rec = dict(path=path, ntri=NT, cold_us=dict(setup=1e3 * t_setup / K, main=1e3 * t_main / K, corners=1e3 * t_corn / K, step=1e3 * t_step / K,
                                                    step_graph=1e3 * cold_graph / K),
                   warm_us=1e3 * warm, frac_main_cold=balg / (t_main / K * 1e-3) / 1e9 / 6547.5)
        B = a.batch
        if B > 1:
            ctxs = [ctx] + [pkg.Context(0, max(NT, 16), max(NP, 16)) for _ in range(B - 1)]
            for i, c in enumerate(ctxs):
                c.set_image(0, imgs[i % 4]); c.upload(pts, tri, ratio)
            pkg.batch_step(ctxs, pkg.TRIANGULATE, 16); ctx.sync()
            pkg.batch_step(ctxs, pkg.TRIANGULATE, 64); ctx.sync()
            us = 1e3 * ctx.last_elapsed_ms() / 64 / B
            rec["batched_us_per_image_step"] = us
            rec["batched_frac_step"] = balg / (us * 1e-6) / 1e9 / 6547.5
            for c in ctxs[1:]:
                c.close()
        print(json.dumps(rec), flush=True)
        out.append(rec)
        ctx.close()
os.makedirs(os.path.dirname(a.out), exist_ok=True)
json.dump(out, open(a.out, "w"), indent=1)
