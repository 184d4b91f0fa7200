"""Turn gpurun_out/ ncu outputs into the small tracked summaries under profiles/.

    python scripts/summarize_profiles.py r01
"""
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
ncu_only = len(sys.argv) > 2 and sys.argv[2] == "ncu_only"   # on the GPU box: only the ncu report, written next to it
if len(sys.argv) > 3:
    OUT = sys.argv[3]
os.makedirs(OUT, exist_ok=True)


def to_us(v, unit):
    v = float(v.replace(",", ""))
    return {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(unit, v)


def launches():
    path = os.path.join(SRC, "launches.csv")
    if not os.path.isfile(path):
        return
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    tot = collections.defaultdict(lambda: [0, 0.0])
    ours = []
    n = 0
    for row in csv.DictReader(lines):
        us = to_us(row["Metric Value"], row["Metric Unit"])
        full = row["Kernel Name"]
        name = re.sub(r"\(.*", "", full)[:110]
        tot[name][0] += 1
        tot[name][1] += us
        n += 1
        if ("grafp" in full or "knn_" in full or "mr_aggregate" in full or "bn_" in full or "ntxent" in full or "peak_extract" in full
                or "conv1x1" in full or "taps_" in full):
            ours.append((row["ID"], name, row.get("Grid Size", ""), row.get("Block Size", ""), f"{us:.2f}"))
    total = sum(v[1] for v in tot.values())
    with open(os.path.join(OUT, f"{tag}_launches_summary.csv"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 2 --warmup 1 --graph off --cudnn-benchmark off --no-cpu-baseline --no-gpu-eager\n")
        f.write(f"# {n} launches, {total/1e3:.1f} ms of kernel time (cold-cache, serialised: compare shares, not absolutes)\n")
        f.write("share_pct,total_ms,launches,kernel\n")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{100*v[1]/total:.2f},{v[1]/1e3:.3f},{v[0]},\"{k}\"\n")
    with open(os.path.join(OUT, f"{tag}_launches_ours.csv"), "w") as f:
        f.write("id,kernel,grid,block,duration_us\n")
        for r in ours:
            f.write(",".join(f'"{c}"' if "," in c or "<" in c else c for c in r) + "\n")
    print("launch list:", n, "launches")


KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__cluster_dim_x",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
]


def ncu_report(rep, out_name):
    path = os.path.join(SRC, rep)
    if not os.path.isfile(path):
        return
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        return
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(os.path.join(OUT, out_name), "w") as f:
        f.write(f"# from {rep}: ncu --set full --clock-control none --import-source on (one row per captured launch)\n")
        for d in data:
            f.write(f"\n== {d[idx['Kernel Name']]}\n")
            for k in KEYS:
                if k in idx:
                    f.write(f"{k:75s} {d[idx[k]]} {units[idx[k]]}\n")
            for h in hdr:
                if "warp_issue_stalled" in h and h.endswith("_per_warp_active.pct"):
                    try:
                        if float(d[idx[h]]) >= 5:
                            f.write(f"stall {h.split('issue_stalled_')[1]:69s} {d[idx[h]]} %\n")
                    except ValueError:
                        pass
    print("wrote", out_name, len(data), "launches")


if not ncu_only:
    launches()
ncu_report("prof_ops.ncu-rep", f"{tag}_ncu_hot_kernels.txt")
