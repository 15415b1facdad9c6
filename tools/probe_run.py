"""Device time per frame of tpose_cuda_run against plain steps (development aid): where a run's frame goes --
the previous frame's geterr + topology watch in k_edges' service block, the history records."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
import importlib
import numpy as np
pkg = entry.load_package()
synth = importlib.import_module("tpose_b200.synth")
w, h, ntri = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
ratio = w / h
img = synth.make_image(w, h, seed=1000)
pts, tri, he = synth.mesh_for(ntri, ratio)
for _ in (0,):
    c = pkg.Context(0, tri.shape[0], pts.shape[0])
    c.set_image(0, img); c.upload(pts, tri, ratio)
    c.step(0, pkg.TRIANGULATE, 64); c.sync()
    c.step(0, pkg.TRIANGULATE, 64); c.sync()
    out = {"us_per_step": 1000 * c.last_elapsed_ms() / 64}
    for name, watch, hist in (("run", False, 0), ("run+watch", True, 0), ("run+watch+history", True, 32)):
        c.watch(tri, he) if watch else c.watch(None)
        c.set_history(hist)
        for n in (32, 8):
            c.run(0, pkg.TRIANGULATE, n, 1e-30)
            done, flags, _ = c.run(0, pkg.TRIANGULATE, n, 1e-30)
            out[f"{name} n={n}"] = (done, round(1000 * c.last_elapsed_ms() / max(done, 1), 2))
    print(json.dumps(out))
    c.close()
