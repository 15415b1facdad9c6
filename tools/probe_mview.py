"""Device time of the batched multiview kernels for 64 match lists of 2*NP matches (development aid)."""
import importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
import numpy as np
sys.argv = [sys.argv[0]]
import bench
pkg = entry.load_package(); synth = importlib.import_module("tpose_b200.synth")
w, h = 1920, 1080; ratio = w / h
pts, tri, _ = synth.mesh_for(3000, ratio)
lists = [bench.synthetic_matches(synth, pts, w, h, ratio, seed=500 + i) for i in range(64)]
A, B = [a for a, _ in lists], [b for _, b in lists]
out = {}
pkg.f8point(A, B); pkg.f8point(A, B); out["f8point_ms"] = pkg.mview_last_ms()
pkg.fsampson(A, B); F = pkg.fsampson(A, B); out["fsampson_100_iterations_ms"] = pkg.mview_last_ms()
K = np.array([[960 / 673.101, 0, 488.421 / 960], [0, 960 / 673.328, 268.8 / 960], [0, 0, 1]], np.float32)
pkg.triangulate(F, K, A, B); out["triangulate_ms"] = pkg.mview_last_ms()
out["lists"], out["matches_per_list"] = 64, len(A[0])
print(json.dumps(out))
