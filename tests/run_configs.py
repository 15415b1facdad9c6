"""BASELINE.json configs 2-4 at full size on the GPU box, through the ported programs and the C ABI:
wall times and size-independent checks (record sizes, level counts, half-edge consistency, F / 3-D
points against the numpy oracle).  Writes gpurun_out/configs.json.

  python tests/run_configs.py [--quick]      (--quick: 640x360, 500 triangles -- a smoke run)
"""
import argparse, json, os, struct, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))     # (this file lives in tests/ because it uses the oracle as its checker)
import numpy as np
import __graft_entry__ as entry
import importlib

ap = argparse.ArgumentParser()
ap.add_argument("--quick", action="store_true")
ap.add_argument("--frames", type=int, default=400000, help="frame cap of each full-size continuation (the levels reached are reported)")
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out"))
a = ap.parse_args()
pkg = entry.load_package()
synth = importlib.import_module("tpose_b200.synth")
BIN = os.path.join(ROOT, "t-pose_b200", "bin")
os.makedirs(a.out, exist_ok=True)
work = os.path.join(a.out, "configs_work")
os.makedirs(work, exist_ok=True)


def ppm(path, img):
    h, w = img.shape[:2]
    with open(path, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (w, h)); f.write(np.ascontiguousarray(img[..., :3]).tobytes())


def records(path):
    """Parses a .tri stack (source/io.hpp:67-220): list of dicts with ratio, tri, he, col, pts, opts."""
    out, buf, o = [], open(path, "rb").read(), 0
    while o < len(buf):
        ratio, nt = struct.unpack_from("<fi", buf, o); o += 8
        t = np.frombuffer(buf, np.int32, 9 * nt, o).reshape(nt, 9); o += 36 * nt
        (npnt,) = struct.unpack_from("<i", buf, o); o += 4
        p = np.frombuffer(buf, np.float32, 4 * npnt, o).reshape(npnt, 4); o += 16 * npnt
        out.append(dict(ratio=ratio, tri=t[:, 0:3], he=t[:, 3:6], col=t[:, 6:9], pts=p[:, 0:2], opts=p[:, 2:4]))
    assert o == len(buf), "trailing bytes in .tri"
    return out


def check_stack(recs, exportlist=None):
    for r in recs:
        nt = len(r["tri"]); he = r["he"].reshape(-1)
        assert r["tri"].min() >= 0 and r["tri"].max() < len(r["pts"])
        twins = he[he >= 0]
        assert np.array_equal(he[twins], np.nonzero(he >= 0)[0]), "half-edge twins are not mutual"
        # twin edges run between the same two vertices in opposite directions
        hidx = np.nonzero(he >= 0)[0]
        va = r["tri"].reshape(-1)[hidx]; vb = r["tri"][hidx // 3, (hidx % 3 + 1) % 3]
        ta = r["tri"].reshape(-1)[he[hidx]]; tb = r["tri"][he[hidx] // 3, (he[hidx] % 3 + 1) % 3]
        assert np.array_equal(va, tb) and np.array_equal(vb, ta), "twin half-edges disagree on their vertices"
        assert 2 * len(r["pts"]) - nt - 2 >= 0      # Euler: NT = 2 NP - b - 2
    nts = [len(r["tri"]) for r in recs]
    assert nts == sorted(nts), "levels are not coarse to fine"
    return nts


def run(cmd, timeout):
    t0 = time.perf_counter()
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    dt = time.perf_counter() - t0
    if p.returncode != 0:
        raise RuntimeError(" ".join(cmd) + "\n" + p.stdout[-2000:] + p.stderr[-2000:])
    # the program's summary: the "frames ..." line and the "runs ..." line under it
    return " | ".join(p.stdout.strip().splitlines()[-2:]), dt


res = {"quick": a.quick}
W2, H2, MAX2 = (640, 360, 1500) if a.quick else (1920, 1080, 3000)
W3, H3 = (480, 270) if a.quick else (1280, 720)

def shrink(img, f):
    """box-filtered copy, f x smaller (exact integer factor)"""
    h, w = img.shape[:2]
    return np.rint(img.reshape(h // f, f, w // f, f, 4).astype(np.float32).mean(axis=(1, 3))).astype(np.uint8)


def grow(tag, img, boot_factor, upto):
    """The reference schedule on a reduced copy (its step constant does not settle on large rasters
    with a handful of triangles), continued on the full image up to `upto`
    triangles.  Returns (path of the combined stack, timings)."""
    small = shrink(img, boot_factor)
    ps, pf = os.path.join(work, tag + "_boot.ppm"), os.path.join(work, tag + ".ppm")
    ppm(ps, small); ppm(pf, img)
    for p in (ps, pf):
        if os.path.exists(p + ".tri"): os.remove(p + ".tri")
    l1, t1 = run([os.path.join(BIN, "triangulate"), "-i", ps, "-native", "-frames", "3000000"], 1200)
    l2, t2 = run([os.path.join(BIN, "triangulate"), "-i", pf, "-native", "-max", str(upto), "-t", ps + ".tri", "-frames", str(a.frames)], 1800)
    stack = os.path.join(work, tag + "_stack.tri")
    with open(stack, "wb") as f:
        f.write(open(ps + ".tri", "rb").read())
        if os.path.exists(pf + ".tri"): f.write(open(pf + ".tri", "rb").read())
    return stack, dict(boot_raster=f"{small.shape[1]}x{small.shape[0]}", boot_line=l1, boot_wall_s=t1, full_line=l2, full_wall_s=t2)


# ---- config 2: single-image triangulate, full topology schedule --------------------------------
img = synth.make_image(W2, H2, seed=1000)
stack2, tim = grow("c2", img, 2 if a.quick else 3, MAX2)
recs = records(stack2)
nts = check_stack(recs)
res["config2"] = {"workload": f"{W2}x{H2} triangulate: reference schedule on a {tim['boot_raster']} copy to 1000 triangles, continued at {W2}x{H2} "
                              f"(native raster) to {MAX2} triangles", **tim, "levels": len(recs), "triangles_per_level": nts,
                  "bytes": os.path.getsize(stack2),
                  "bytes_expected": int(sum(12 + 36 * len(r["tri"]) + 16 * len(r["pts"]) for r in recs))}
assert res["config2"]["bytes"] == res["config2"]["bytes_expected"]
print(json.dumps(res["config2"]), flush=True)

# ---- config 3: two-view coarse-to-fine warp -------------------------------------------------------
A, B = synth.make_pair(W3, H3, seed=1000)
sa, ta_ = grow("c3a", A, 2 if not a.quick else 1, MAX2)
sb, tb_ = grow("c3b", B, 2 if not a.quick else 1, MAX2)
pa, pb = os.path.join(work, "c3a.ppm"), os.path.join(work, "c3b.ppm")
for p in (sa, sb):
    if os.path.exists(p + ".warp"): os.remove(p + ".warp")
line, dt = run([os.path.join(BIN, "warp"), "-ia", pa, "-ib", pb, "-ta", sa, "-tb", sb, "-native"], 3000)
ra, rb = records(sa), records(sb)
wa, wb = records(sa + ".warp"), records(sb + ".warp")
check_stack(wa); check_stack(wb)
for s, d in ((ra, wa), (rb, wb)):
    assert len(s) == len(d)
    for x, y in zip(s, d):      # topology frozen by the warp (warp/main.cpp never edits it)
        assert np.array_equal(x["tri"], y["tri"]) and np.array_equal(x["he"], y["he"])
moved = float(np.abs(wa[-1]["pts"] - wa[-1]["opts"]).max())
res["config3"] = {"workload": f"{W3}x{H3} pair: both views grown to {MAX2} triangles (as in config 2), then the warp level loop over all levels, native raster",
                  "triangulate_a": ta_, "triangulate_b": tb_, "warp_program_line": line, "warp_wall_s": dt, "levels": len(wa),
                  "max_vertex_displacement_finest": moved}
print(json.dumps(res["config3"]), flush=True)

# ---- config 4: warp + F + DLT at the finest level ---------------------------------------------------
from oracle import mview_oracle as M
fa, fb = wa[-1], wb[-1]
ratio = fa["ratio"]
T = lambda P: np.stack([np.float32(0.5) * P[:, 0] / np.float32(ratio), np.float32(-0.5) * P[:, 1] / np.float32(ratio)], 1).astype(np.float32)
# compute_fundamental_mat/main.cpp:137-167 with every triangle selected
used_a = np.unique(fa["tri"]); used_b = np.unique(fb["tri"])
X = np.concatenate([T(fa["opts"][used_a]), T(fb["pts"][used_b])])
Y = np.concatenate([T(fa["pts"][used_a]), T(fb["opts"][used_b])])
t0 = time.perf_counter()
F8 = pkg.f8point([X], [Y])[0]
ms_f8 = pkg.mview_last_ms()
Fs = pkg.fsampson([X], [Y])[0]
ms_fs = pkg.mview_last_ms()
K = M.camera()
X3, Ac, Bc = pkg.triangulate([F8], K, [X], [Y], correct=True)
ms_tr = pkg.mview_last_ms()
dt = time.perf_counter() - t0
refF = M.f_8point(X, Y)
errF = float(np.abs(F8 - refF).max() / np.abs(refF).max())
sub = np.random.default_rng(0).choice(len(X), size=min(200, len(X)), replace=False)
refX = M.triangulate(F8, K, X[sub], Y[sub])
errX = float(np.abs(X3[0][sub] - refX).max() / np.abs(refX).max())
res["config4"] = {"workload": f"matches = all vertices of both finest warped meshes ({len(X)} matches), F_8Point, F_Sampson, triangulate(F,K,A,B)",
                  "kernel_ms": {"f8point": ms_f8, "fsampson": ms_fs, "triangulate": ms_tr}, "wall_s": dt,
                  "F_rel_err_vs_oracle": errF, "X_rel_err_vs_oracle_200_matches": errX,
                  "sampson_vs_8point_rel_diff": float(np.abs(Fs - F8).max() / np.abs(F8).max())}
assert errF < 1e-4 and errX < 1e-4, (errF, errX)
print(json.dumps(res["config4"]), flush=True)
json.dump(res, open(os.path.join(a.out, "configs.json"), "w"), indent=1)
