"""Per-kernel warp-stall breakdown (pc sampling) from an ncu report: python scripts/ncu_stalls.py <rep> [regex]"""
import csv, re, subprocess, sys
rep = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
kn = h.index("Kernel Name")
cols = [i for i, n in enumerate(h) if n.startswith("smsp__pcsamp_warps_issue_stalled_") and not n.endswith("_not_issued")]
for row in rows[2:]:
    if pat and not pat.search(row[kn]):
        continue
    vals = []
    for i in cols:
        try:
            vals.append((float(row[i].replace(",", "")), h[i].replace("smsp__pcsamp_warps_issue_stalled_", "")))
        except ValueError:
            pass
    tot = sum(v for v, _ in vals) or 1.0
    top = sorted(vals, reverse=True)[:7]
    print(row[kn][:60], "|", ", ".join(f"{n} {100*v/tot:.0f}%" for v, n in top))
