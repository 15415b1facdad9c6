"""Turns what a GPU visit left in gpurun_out/ into the tracked summaries under profiles/:
   <tag>_bench_n1.json, <tag>_bench_reference_n1.json, <tag>_launches_bench.csv, <tag>_launch_shares.json,
   <tag>_k_edges_ncu_full.csv, <tag>_k_corners_ncu_full.csv, <tag>_batched_traffic.csv, traffic.json
Run here (no GPU needed): python tools/summarize_profiles.py r01c"""
import csv, json, os, shutil, subprocess, sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1]

KEEP = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__waves_per_multiprocessor", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]


def copy(src, dst):
    if os.path.exists(os.path.join(G, src)):
        shutil.copy(os.path.join(G, src), os.path.join(P, dst))


def ncu_raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def full_summary(rep, dst):
    if not os.path.exists(rep):
        return None
    hdr, units, rows = ncu_raw(rep)
    idx = [hdr.index(k) for k in KEEP if k in hdr]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx]); w.writerow([units[i] for i in idx])
        for r in rows:
            w.writerow([r[i] for i in idx])
    return hdr, units, rows


copy("bench.json", f"{tag}_bench_n1.json")
copy("bench_ref.json", f"{tag}_bench_reference_n1.json")
copy("launches.csv", f"{tag}_launches_bench.csv")
copy("batched_traffic.csv", f"{tag}_batched_traffic.csv")
copy("probe_step.json", f"{tag}_probe_group_sweep.json")
copy("probe_batch.json", f"{tag}_probe_batch.json")

# share of each kernel in the bench run's launch list
lp = os.path.join(G, "launches.csv")
if os.path.exists(lp):
    rows = [r for r in csv.reader(open(lp)) if len(r) > 5]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        t = tot[r[ik].split("(")[0]]
        t[0] += 1; t[1] += float(r[iv].replace(",", "")) / 1000.0
    s = sum(v[1] for v in tot.values())
    json.dump([{"kernel": k, "launches": v[0], "total_us": v[1], "mean_us": v[1] / v[0], "share": v[1] / s}
               for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])],
              open(os.path.join(P, f"{tag}_launch_shares.json"), "w"), indent=1)

traffic = {}
e = full_summary(os.path.join(G, "prof_edges.ncu-rep"), os.path.join(P, f"{tag}_k_edges_ncu_full.csv"))
full_summary(os.path.join(G, "prof_corners.ncu-rep"), os.path.join(P, f"{tag}_k_corners_ncu_full.csv"))
if e:
    hdr, units, rows = e
    def col(name):
        i = hdr.index(name)
        scale = {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Gbyte": 1e9}[units[i]]
        return sum(float(r[i].replace(",", "")) * scale for r in rows) / len(rows)
    traffic["k_edges_dram_bytes_per_launch"] = col("dram__bytes_read.sum") + col("dram__bytes_write.sum")
    traffic["k_edges_source"] = f"profiles/{tag}_k_edges_ncu_full.csv (ncu --set full, single 1080p image, 3000 triangles, cold caches)"
bp = os.path.join(G, "batched_traffic.csv")
if os.path.exists(bp):
    rows = [r for r in csv.reader(open(bp)) if len(r) > 5]
    hdr = rows[0]
    ik, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    per = defaultdict(float); n = defaultdict(set)
    for r in rows[1:]:
        if r[im].startswith("dram__bytes"):
            k = r[ik].split("(")[0].split("<")[0].replace("void ", "")
            per[k] += float(r[iv].replace(",", "")); n[k].add(r[iid])
    batch = 64
    traffic["batched_dram_bytes_per_image_step"] = sum(per[k] / len(n[k]) for k in per) / batch
    traffic["batched_source"] = f"profiles/{tag}_batched_traffic.csv (64 images per launch; k_edges + k_corners read+write, per image)"
if traffic:
    json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
print(json.dumps(traffic, indent=1))
