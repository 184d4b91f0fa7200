"""One line per captured launch: python scripts/ncu_summary.py <rep> [regex]"""
import csv, re, subprocess, sys
rep = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "us"), ("launch__grid_size", "grid"), ("launch__block_size", "blk"),
        ("launch__registers_per_thread", "regs"), ("dram__bytes_read.sum", "rdMB"), ("dram__bytes_write.sum", "wrMB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"),
        ("lts__t_sector_hit_rate.pct", "l2hit%"), ("smsp__inst_executed.sum", "inst")]
idx = [(h.index(k), n) for k, n in want if k in h]
units = rows[1]
print(" | ".join(n for _, n in idx))
for row in rows[2:]:
    if pat and not pat.search(row[h.index("Kernel Name")]):
        continue
    cells = []
    for i, n in idx:
        v = row[i]
        if n == "kernel":
            v = re.sub(r"\(.*", "", v).replace("void ", "").replace("grafp::", "")[:48]
        else:
            try:
                f = float(v.replace(",", ""))
                if n in ("rdMB", "wrMB"):
                    u = units[i]
                    f = f * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
                if n == "us":
                    f = f * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(units[i], 1.0)
                v = f"{f:.1f}" if n != "inst" else f"{f/1e6:.1f}M"
            except ValueError:
                pass
        cells.append(v)
    print(" | ".join(cells))
