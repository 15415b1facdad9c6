"""BASELINE config 5 in small: P independent image pairs, sharded over the ranks of one box (one
process per GPU, no data-path collective), each pair = triangulate both views -> coarse-to-fine warp
-> matches -> F_8Point -> triangulate(F,K,A,B); one NCCL gather of the per-pair results to rank 0
(t-pose_b200/shard.py).  Rank 0 then recomputes a pair that another rank owned and checks that the
bytes agree (sharding must not change results).  Writes gpurun_out/pairs_n<world>.json.

  python tools/run_pairs.py                                             # one GPU
  python -m torch.distributed.run --nproc-per-node 2 ... tools/run_pairs.py   # sharded
The raster is 480x270 (the reference's step constant does not settle on large rasters).
"""
import argparse, importlib, json, os, struct, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as entry

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=8)
ap.add_argument("--width", type=int, default=480)
ap.add_argument("--height", type=int, default=270)
a = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", rank))
dist = None
if world > 1:
    import torch, torch.distributed as dist
    torch.cuda.set_device(local)
    saved = os.dup(1); os.dup2(2, 1)
    try:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier(); torch.cuda.synchronize()
    finally:
        sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)
pkg = entry.load_package()
synth = importlib.import_module("tpose_b200.synth")
shard = importlib.import_module("tpose_b200.shard")

def camera():
    """mview::Camera() (multiview.hpp:35-52): 1/fx on the diagonal, as coded; float32."""
    f32 = np.float32
    px, py = f32(488.421 / 960.0), f32(268.8 / 960.0)
    fx, fy = f32(673.101 / 960.0), f32(673.328 / 960.0)
    return np.array([[f32(1.0) / fx, 0, px], [0, f32(1.0) / fy, py], [0, 0, 1]], f32)

BIN = os.path.join(ROOT, "t-pose_b200", "bin")
work = os.path.join(ROOT, "gpurun_out", f"pairs_work_r{rank}")
os.makedirs(work, exist_ok=True)


def ppm(path, img):
    h, w = img.shape[:2]
    with open(path, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (w, h)); f.write(np.ascontiguousarray(img[..., :3]).tobytes())


def finest(path):
    buf, o, last = open(path, "rb").read(), 0, None
    while o < len(buf):
        ratio, nt = struct.unpack_from("<fi", buf, o); o += 8
        t = np.frombuffer(buf, np.int32, 9 * nt, o).reshape(nt, 9); o += 36 * nt
        (npnt,) = struct.unpack_from("<i", buf, o); o += 4
        p = np.frombuffer(buf, np.float32, 4 * npnt, o).reshape(npnt, 4); o += 16 * npnt
        last = dict(ratio=ratio, tri=t[:, 0:3], pts=p[:, 0:2], opts=p[:, 2:4])
    return last


def run(cmd):
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    if p.returncode != 0:
        raise RuntimeError(" ".join(cmd) + "\n" + p.stdout[-1500:] + p.stderr[-1500:])


def solve_pair(u, tag):
    """Everything for pair u on this rank's GPU; returns the flat float32 result vector."""
    A, B = synth.make_pair(a.width, a.height, seed=1000 + u)
    pa, pb = os.path.join(work, f"{tag}{u}a.ppm"), os.path.join(work, f"{tag}{u}b.ppm")
    ppm(pa, A); ppm(pb, B)
    dev = ["-device", str(local)]
    for p in (pa, pb):
        for ext in (".tri", ".tri.warp"):
            if os.path.exists(p + ext): os.remove(p + ext)
        run([os.path.join(BIN, "triangulate"), "-i", p, "-native", "-frames", "400000"] + dev)
    run([os.path.join(BIN, "warp"), "-ia", pa, "-ib", pb, "-ta", pa + ".tri", "-tb", pb + ".tri", "-native"] + dev)
    fa, fb = finest(pa + ".tri.warp"), finest(pb + ".tri.warp")
    r = np.float32(fa["ratio"])
    T = lambda P: np.stack([np.float32(0.5) * P[:, 0] / r, np.float32(-0.5) * P[:, 1] / r], 1).astype(np.float32)
    ua, ub = np.unique(fa["tri"]), np.unique(fb["tri"])
    X = np.concatenate([T(fa["opts"][ua]), T(fb["pts"][ub])]); Y = np.concatenate([T(fa["pts"][ua]), T(fb["opts"][ub])])
    F = pkg.f8point([X], [Y], device=local)[0]
    X3, _, _ = pkg.triangulate([F], camera(), [X], [Y], correct=True, device=local)
    return np.concatenate([fa["pts"].ravel(), fb["pts"].ravel(), F.ravel(), X3[0].ravel()]).astype(np.float32)


lo, hi = shard.shard_range(a.pairs, world, rank)
t0 = time.perf_counter()
mine = {u: solve_pair(u, "p") for u in range(lo, hi)}
t_local = time.perf_counter() - t0
t_max = shard.max_over_ranks(t_local, dist, device=f"cuda:{local}" if dist else "cpu")
t1 = time.perf_counter()
everything = shard.gather_units(mine, a.pairs, dist, device=f"cuda:{local}" if dist else "cpu")
if dist:
    torch.cuda.synchronize()
t_gather = time.perf_counter() - t1
if rank == 0:
    probe = a.pairs - 1                                    # owned by the last rank
    again = solve_pair(probe, "check")
    same = again.shape == everything[probe].shape and np.array_equal(again.view(np.uint32), everything[probe].view(np.uint32))
    out = {"pairs": a.pairs, "raster": f"{a.width}x{a.height}", "n_gpus": world, "pairs_per_s": a.pairs / t_max,
           "slowest_rank_s": t_max, "gather_s": t_gather, "gather_bytes": int(sum(v.nbytes for v in everything.values())),
           "pair_recomputed_on_rank0": probe, "recomputed_identical": bool(same),
           "floats_per_pair": [int(everything[u].size) for u in range(a.pairs)]}
    assert same, "a pair computed on another rank differs from rank 0's own computation"
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"pairs_n{world}.json"), "w"), indent=1)
    print(json.dumps(out))
if dist:
    dist.barrier(); dist.destroy_process_group()
