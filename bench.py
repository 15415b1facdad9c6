#!/usr/bin/env python
"""bench.py -- triangulation-optimisation steps/s at 1920x1080 / 3000 triangles (BASELINE.json).

A "step" is one frame of software/triangulate at frozen topology: colours (mode 0) + energy
(mode 1) + gradient + shift for all 13*NT variants, leaving cn/ca/ten/points valid in the
reference's SSBO layout.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path (one JSON line)
  python bench.py --impl reference [--gpus N] [--steps K] ...    the reference algorithm on the host CPU

N>1 is launched by torch.distributed.run, one rank per GPU; every rank optimises its own image
(independent problems, no data-path collective -- weak scaling) and the vertex buffers are
gathered to rank 0 over NCCL at the end of the timed region.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

L2_FLUSH_BYTES = 256 << 20
FALLBACK_HBM_GBS = 6650.0


def b_alg(w, h, nt, np_):
    """Algorithmic bytes of one image-step: every texel once + indices + points r/w."""
    return 4 * w * h + 16 * nt + 16 * np_


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback 6.65 TB/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def dist_setup(n):
    """Returns (rank, world, local_rank, dist-or-None)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 0, 1, 0, None
    import torch
    import torch.distributed as dist
    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    # stdout carries exactly one JSON line.  NCCL writes its version banner (and any NCCL_DEBUG output)
    # to file descriptor 1 while the communicator comes up, so fd 1 points at stderr until the first
    # collective has run.
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
        torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
    return rank, world, local, dist


def cpu_reference_steps_per_s(O, img, pts, tri, ratio, budget_s, min_steps=1, stride=None):
    """Times the oracle (literal 13*NT-variant, two-pass restatement of the shaders, OpenMP over
    variants on all host cores).  Returns (steps/s, description)."""
    t0 = time.perf_counter()
    O.step(img, pts, tri, ratio, O.TRIANGULATE)
    t_first = time.perf_counter() - t0
    n = max(min_steps, int(budget_s / max(t_first, 1e-3)))
    n = min(n, 200)
    cur = pts
    t0 = time.perf_counter()
    for _ in range(n):
        cur = O.step(img, cur, tri, ratio, O.TRIANGULATE)["pts"]
    dt = time.perf_counter() - t0
    return n / dt, f"{n} full steps (all 13*NT variants, both passes) of the same image+mesh, {dt:.1f} s"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun pins OMP_NUM_THREADS=1 for its workers; the CPU arm is meant to use every host core
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count())
    pkg = entry.load_package()
    synth = importlib.import_module("tpose_b200.synth")
    O = entry.load_oracle()
    w, h, nt = args.width, args.height, args.ntri
    ratio = w / h
    img = synth.make_image(w, h, seed=1000)
    pts, tri, _ = synth.mesh_for(nt, ratio)
    cores = os.cpu_count()
    # warm-up + timed steps, bounded so the whole run stays within a few minutes
    t0 = time.perf_counter()
    O.step(img, pts, tri, ratio, O.TRIANGULATE)
    t_one = time.perf_counter() - t0
    budget = 150.0
    k = max(1, min(args.steps, int(budget / max(t_one, 1e-3))))
    wu = max(1, min(args.warmup, 3))
    cur = pts
    for _ in range(wu):
        cur = O.step(img, cur, tri, ratio, O.TRIANGULATE)["pts"]
    t0 = time.perf_counter()
    for _ in range(k):
        cur = O.step(img, cur, tri, ratio, O.TRIANGULATE)["pts"]
    dt = time.perf_counter() - t0
    val = k / dt
    sample = f"{k} full steps of the {w}x{h}/{nt}-triangle workload (requested {args.steps})"
    line = {
        "impl": "reference", "metric": "triangulation-opt steps/sec @1080p/3k-tri", "value": val, "unit": "steps/s",
        "n_gpus": args.gpus, "steps": k, "warmup": wu, "ms_per_step": 1000.0 * dt / k, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "i32", "data": "synthetic",
        "config": {"workload": f"single {w}x{h} image, {nt} triangles, frozen topology, triangulate flavour",
                   "note": "the reference's GL path cannot run here (no GL/TinyEngine); this is the CPU port "
                           "(oracle/step_oracle.c) of its algorithm on the host cores"},
        "cpu_baseline": {"value": val, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    rank, world, local, dist = dist_setup(args.gpus)
    pkg = entry.load_package()
    synth = importlib.import_module("tpose_b200.synth")
    w, h, nt = args.width, args.height, args.ntri
    ratio = w / h
    K, W = args.steps, max(args.warmup, 3)
    # every rank owns one independent image (seed 1000 + rank): weak scaling
    img = synth.make_image(w, h, seed=1000 + rank)
    pts, tri, _ = synth.mesh_for(nt, ratio)
    NT, NP = tri.shape[0], pts.shape[0]
    ctx = pkg.Context(device=local, max_triangles=max(NT, 16), max_points=max(NP, 16))
    ctx.set_image(0, img)          # image resident in HBM before any timed region
    ctx.upload(pts, tri, ratio)

    def barrier():
        ctx.sync()
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident loop, L2 flushed before every step (rule: flush or exceed L2) ------
    for _ in range(W):
        ctx.flush_l2(L2_FLUSH_BYTES)
        ctx.step(0, pkg.TRIANGULATE, 1)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count()
    barrier()
    t_wall0 = time.perf_counter()
    ms_cold = 0.0
    for _ in range(K):
        ctx.flush_l2(L2_FLUSH_BYTES)          # not inside the event bracket of the step
        ctx.step(0, pkg.TRIANGULATE, 1)       # ev0 . k_edges . k_corners . ev1 on the context's stream
        ms_cold += ctx.last_elapsed_ms()
    barrier()
    launches = ctx.launch_count() - launches0
    ms_cold = max_over_ranks(ms_cold)
    # the dominant kernel alone (an event between the two kernels; not part of the timed K steps)
    ctx.set_kernel_timing(True)
    ms_edges, Kk = 0.0, min(K, 100)
    for _ in range(Kk):
        ctx.flush_l2(L2_FLUSH_BYTES)
        ctx.step(0, pkg.TRIANGULATE, 1)
        ms_edges += ctx.last_kernel_ms()[0]
    ctx.set_kernel_timing(False)
    # ---- same K steps without the flush: the optimiser's steady state (table stays in L2).
    # One untimed pass first, so that the graphs the K steps replay are already instantiated.
    ctx.upload(pts, tri, ratio)
    ctx.step(0, pkg.TRIANGULATE, K)
    ctx.upload(pts, tri, ratio)
    barrier()
    ctx.step(0, pkg.TRIANGULATE, K)
    barrier()
    ms_warm = max_over_ranks(ctx.last_elapsed_ms())
    # ---- batch of independent images in lock step (tables exceed L2: the HBM-bound regime) ----
    B = args.batch
    batched = None
    if B > 1:
        bctx = [ctx] + [pkg.Context(device=local, max_triangles=max(NT, 16), max_points=max(NP, 16)) for _ in range(B - 1)]
        pool = [img] + [synth.make_image(w, h, seed=2000 + 10 * rank + i) for i in range(3)]   # 4 distinct images, B slots
        for i, c in enumerate(bctx):
            if i:
                c.set_image(0, pool[i % len(pool)])
            c.upload(pts, tri, ratio)
        Kb = max(16, min(K, 64))
        pkg.batch_step(bctx, pkg.TRIANGULATE, Kb)      # warm-up + graph instantiation
        barrier()
        pkg.batch_step(bctx, pkg.TRIANGULATE, Kb)
        barrier()
        ms_b = max_over_ranks(ctx.last_elapsed_ms())
        batched = (Kb, ms_b)
        for c in bctx[1:]:
            c.close()
        ctx.upload(pts, tri, ratio)
    if dist is not None:      # final vertex-buffer gather (the only collective of the job)
        import torch
        mine = torch.from_numpy(ctx.retrieve(ten=False, cn=False, ca=False)["pts"]).to(f"cuda:{local}")
        bucket = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
        dist.gather(mine, bucket, dst=0)
        torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end through the C ABI with host buffers --------------------------------------
    # one job = set_image from pinned host memory (H2D + table build) + K frames, each frame
    # filling pointbuf from the host and reading back what triangulate/main.cpp:201-204 reads
    # back every frame (tenergybuf, tcolnumbuf, pointbuf; penergybuf is dead in the reference)
    himg = pkg.pinned_empty(img.shape, np.uint8); himg[...] = img
    hpts = pkg.pinned_empty(pts.shape, np.float32); hpts[...] = pts
    hten = pkg.pinned_empty((13 * NT,), np.int32)
    hcn = pkg.pinned_empty((13 * NT,), np.int32)
    Ke = min(K, 200)
    ctx.upload(pts, tri, ratio)
    barrier()
    t0 = time.perf_counter()
    ctx.set_image(0, himg)
    for _ in range(Ke):
        # pointbuf->fill(points) + step + read-backs, returns after the copies have landed
        ctx.frame(0, pkg.TRIANGULATE, pts_in=hpts, ten=hten, cn=hcn, pts_out=hpts)
    ctx.sync()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    h2d = (img.nbytes / Ke) + pts.nbytes
    d2h = 13 * NT * (4 + 4) + NP * 8

    if rank == 0:
        peak, peak_src = peaks()
        balg = b_alg(w, h, NT, NP)
        cold_s = ms_cold / 1000.0
        warm_s = ms_warm / 1000.0
        edges_s = ms_edges / 1000.0 / Kk         # mean duration of one k_edges launch (cold)
        achieved = balg / edges_s / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("k_edges_dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": "triangulation-opt steps/sec @1080p/3k-tri", "value": world * K / cold_s, "unit": "steps/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": 1000.0 * cold_s / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "i32/i64 (integer moments; float32 vertex update)",
            "data": "synthetic",
            "config": {"workload": f"single {w}x{h} image per GPU, {NT} triangles ({NP} points), frozen topology, "
                                   "triangulate flavour (BASELINE.json configs[1])",
                       "raster": "native", "l2": f"flushed before every timed step ({L2_FLUSH_BYTES >> 20} MiB write)",
                       "parallelism": f"replicas x{world}: one independent image per GPU"},
            "value_l2_warm": world * K / warm_s,
            "value_l2_warm_note": "same K steps back to back (CUDA-graph replay), no flush: the optimiser's steady "
                                  "state (one image's 33 MB table stays L2-resident between its own steps)",
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "k_edges", "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": balg,
                         "kernel_us": 1e6 * edges_s, "step_us": 1e6 * cold_s / K,
                         "note": "one launch = one image-step of ONE image (latency-bound: 41 k work items cannot "
                                 "fill 148 SMs); see `batched` for the HBM-bound regime",
                         "frac_step_l2_warm": (balg / (warm_s / K) / 1e9) / peak},
            "e2e": {"value": world * Ke / e2e_s, "unit": "steps/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "steps": Ke,
                    "what": "set_image from pinned host memory once per job, then per frame tpose_cuda_frame: pointbuf fill (H2D) + "
                            "step + read-back of tenergybuf, tcolnumbuf and pointbuf (13*NT ints each + NP vec2), one sync"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "wall_s_timed_region": time.perf_counter() - t_wall0,
        }
        if batched is not None:
            Kb, ms_b = batched
            per = ms_b / 1000.0 / Kb / B         # seconds per image-step
            btraffic = None
            if os.path.exists(tp):
                try:
                    btraffic = json.load(open(tp)).get("batched_dram_bytes_per_image_step")
                except Exception:
                    btraffic = None
            line["batched"] = {
                "workload": f"{B} independent {w}x{h} images per GPU in lock step, {NT} triangles each, one launch pair "
                            f"per step for all of them ({B * 33.2:.0f} MB of tables: larger than L2)",
                "value": world * 1.0 / per, "unit": "image-steps/s", "steps": Kb, "us_per_image_step": 1e6 * per,
                "roofline": {"bound": "hbm", "achieved": balg / per / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": balg / per / 1e9 / peak, "traffic": btraffic,
                             "kernel": "k_edges + k_corners (whole step)", "algorithmic_bytes_per_image_step": balg}}
        if world == 1 and not args.no_cpu_baseline:
            O = entry.load_oracle()
            v, sample = cpu_reference_steps_per_s(O, img, pts, tri, ratio, budget_s=args.cpu_budget)
            line["cpu_baseline"] = {"value": v, "unit": "steps/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": sample}
        print(json.dumps(line), flush=True)
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--ntri", type=int, default=3000)
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batch", type=int, default=64, help="images in the batched (HBM-bound) measurement; 0/1 = skip")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
