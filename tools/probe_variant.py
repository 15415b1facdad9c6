"""One-line timing of the 1080p/3000 step for the library named by TPOSE_CUDA_LIB: warm (graph), cold (L2 flushed), batch 64."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
import importlib
pkg = entry.load_package()
synth = importlib.import_module("tpose_b200.synth")
w, h = 1920, 1080
ratio = w / h
img = synth.make_image(w, h, seed=1000)
pts, tri, _ = synth.mesh_for(3000, ratio)
ctx = pkg.Context(0, 4096, 4096)
ctx.set_image(0, img); ctx.upload(pts, tri, ratio)
ctx.step(0, 0, 512); ctx.sync(); ctx.upload(pts, tri, ratio)
ctx.step(0, 0, 512); ctx.sync()
warm = 1000 * ctx.last_elapsed_ms() / 512
cold = 0.0
for i in range(120):
    ctx.flush_l2(256 << 20); ctx.step(0, 0, 1)
    if i >= 20: cold += ctx.last_elapsed_ms()
cold = 1000 * cold / 100
ctxs = [ctx] + [pkg.Context(0, 4096, 4096) for _ in range(63)]
for c in ctxs[1:]:
    c.set_image(0, img); c.upload(pts, tri, ratio)
pkg.batch_step(ctxs, 0, 64); ctx.sync(); pkg.batch_step(ctxs, 0, 64); ctx.sync()
b64 = 1000 * ctx.last_elapsed_ms() / 64 / 64
print(json.dumps(dict(lib=os.environ.get("TPOSE_CUDA_LIB", "default"), pdl=os.environ.get("TPOSE_CUDA_NO_PDL", "0") != "1",
                      warm_us=round(warm, 2), cold_us=round(cold, 2), batch64_us_per_image=round(b64, 2))))
