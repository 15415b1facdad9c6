"""Wall and device time of tpose_cuda_carry (triangulation::warp / reversewarp on the device) -- development aid."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
import importlib
import numpy as np
pkg = entry.load_package()
synth = importlib.import_module("tpose_b200.synth")
ratio = 1920 / 1080
for ntri, nfine in ((150, 500), (500, 1500), (1500, 3000), (3000, 3000), (10000, 10000)):
    pts, tri, _ = synth.mesh_for(ntri, ratio)
    rng = np.random.default_rng(3)
    moved = (pts + rng.normal(0, 0.002, pts.shape)).astype(np.float32)
    q = synth.mesh_for(nfine, ratio, seed=11)[0]
    pkg.carry(tri, pts, moved, q, ratio)
    t0 = time.perf_counter()
    for _ in range(20):
        out = pkg.carry(tri, pts, moved, q, ratio)
    wall = (time.perf_counter() - t0) / 20
    print(json.dumps(dict(ntri=ntri, npts=len(q), wall_us=round(1e6 * wall, 1), device_ms=pkg.mview_last_ms() if hasattr(pkg, "mview_last_ms") else None)))
