"""Kernel-internal timeline (library built with -DTPOSE_TIMELINE, named by TPOSE_CUDA_LIB)."""
import os, sys, ctypes as C
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
for ntri in (3000, 150):
    pts, tri, _ = synth.mesh_for(ntri, ratio)
    ctx = pkg.Context(0, 4096, 4096)
    ctx.set_image(0, img); ctx.upload(pts, tri, ratio)
    ctx.step(0, 0, 64); ctx.sync()
    for mode in ("direct", "graph4"):
        buf = (C.c_ulonglong * (2 * 6 * 8192))()
        pkg.lib().tpose_cuda_debug_timeline(None, 1)
        if mode == "direct":
            ctx.step(0, 0, 1)
        else:
            ctx.step(0, 0, 4)     # stamps of the 4th step overwrite the earlier ones (t0 is then off: read differences)
        ctx.sync()
        pkg.lib().tpose_cuda_debug_timeline(buf, 0)
        a = np.frombuffer(buf, dtype=np.uint64).reshape(2, 6, 8192).astype(np.float64)
        t0 = a[0, 0][a[0, 0] > 0].min()
        names = [["entry", "J loaded", "pts loaded", "init done", "walk done", "exit"], ["entry", "J loaded", "base done", "variants done", "exit"]]
        print(f"--- NT={ntri} {mode}: us relative to the first k_edges warp (last step if several); first / median / p95 / last over warps")
        for k in range(2):
            for p, nm in enumerate(names[k]):
                v = a[k, p][a[k, p] > 0]
                if v.size == 0: continue
                v = (v - t0) / 1e3
                print(f"  {'k_edges  ' if k == 0 else 'k_corners'} {nm:14s} n={v.size:5d}  {v.min():7.2f} {np.median(v):7.2f} {np.percentile(v, 95):7.2f} {v.max():7.2f}")
    ctx.close()
