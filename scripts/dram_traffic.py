"""profiles/dram_traffic.json from an `ncu --set full` capture of scripts/ncu_ops.py: per-launch DRAM bytes
(dram__bytes_read.sum + dram__bytes_write.sum) of every hot-path op at the stage-0 shape of the bench workload
(B = 512 segments, N = 1024, C = 64; BatchNorm on the (B, N, 2C) aggregation output).  bench.py copies these into
`roofline.traffic`.    python scripts/dram_traffic.py <report.ncu-rep> [out.json]"""
import csv, json, subprocess, sys

rep = sys.argv[1]
out_path = sys.argv[2] if len(sys.argv) > 2 else "profiles/dram_traffic.json"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units = rows[0], rows[1]
kn, rd, wr = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
ops = {"knn_fwd": ("knn_normalize", "knn_stream_kernel"), "mr_aggregate_fwd": ("mr_aggregate_fwd",),
       "mr_aggregate_bwd": ("mr_aggregate_bwd",), "bn_train_fwd": ("bn_fwd", "bn_stats", "bn_apply"),
       "bn_train_bwd": ("bn_bwd",), "conv1x1_bn_stats_fwd": ("conv1x1_stats_kernel",)}
result = {}
seen = {k: set() for k in ops}
for row in rows[2:]:
    name = row[kn]
    for op, pats in ops.items():
        for p in pats:
            if p in name and p not in seen[op]:       # first launch of each kernel = the stage-0 shape
                seen[op].add(p)
                b = float(row[rd].replace(",", "")) * scale.get(units[rd], 1.0) + float(row[wr].replace(",", "")) * scale.get(units[wr], 1.0)
                result[op] = result.get(op, 0.0) + b
                break
result = {k: int(v) for k, v in result.items()}
result["_note"] = ("per-launch DRAM bytes (read + write) from ncu --set full, first launch of each kernel in scripts/ncu_ops.py: "
                   "B=512, N=1024, C=64 (K1 = normalise + Gram/top-k launches; K5 on the (B, N, 128) tensor; convolution 64 -> 256)")
json.dump(result, open(out_path, "w"), indent=1)
print(json.dumps(result, indent=1))
