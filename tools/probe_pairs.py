"""Development aid: frames to convergence per level of the warp loop at a given raster, and pairs/s of a PairSet."""
import argparse, importlib, json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
import numpy as np
ap = argparse.ArgumentParser()
ap.add_argument("--w", type=int, default=1280); ap.add_argument("--h", type=int, default=720)
ap.add_argument("--pairs", type=int, default=2); ap.add_argument("--cap", type=int, default=-1)
ap.add_argument("--sizes", default="150,500,1500,3000")
a = ap.parse_args()
pkg = entry.load_package(); synth = importlib.import_module("tpose_b200.synth"); P = importlib.import_module("tpose_b200.pairs")
w, h = a.w, a.h; ratio = float(np.float32(w) / np.float32(h))
sizes = tuple(int(x) for x in a.sizes.split(","))
work = tempfile.mkdtemp()
A, B = synth.make_pair(w, h, seed=1000)
ta, tb = os.path.join(work, "a.tri"), os.path.join(work, "b.tri")
P.grid_stack(pkg, synth, ta, A, ratio, sizes); P.grid_stack(pkg, synth, tb, B, ratio, sizes)
ps = P.PairSet(0)
import shutil
for i in range(a.pairs):
    pa, pb = os.path.join(work, f"a{i}.tri"), os.path.join(work, f"b{i}.tri")
    shutil.copy(ta, pa); shutil.copy(tb, pb)
    ps.add(A, B, pa, pb, max_triangles=max(sizes), max_points=max(sizes))
t0 = time.perf_counter()
fr, lv = ps.run(level_cap=a.cap)
dt = time.perf_counter() - t0
r = ps.result(0)
print(json.dumps(dict(raster=f"{w}x{h}", pairs=a.pairs, cap=a.cap, frames=fr, levels=lv, wall_s=dt, pairs_per_s=a.pairs / dt,
                      moved=float(np.abs(r["ptsA"] - r["optsA"]).max()), F=r["F"].tolist())))
