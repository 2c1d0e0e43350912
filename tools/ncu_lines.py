"""Per-source-line instruction counts from an .ncu-rep captured with --import-source on (needs -lineinfo).
usage: python tools/ncu_lines.py rep.ncu-rep [top_n]  ->  file:line  warp-instructions  share  samples  source"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None
hdr = None
acc = {}
total = 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    # rows with an address are SASS rows attributed to the last source line; source rows carry aggregated counts
    if r[2] not in ("-", ""):
        continue
    ie = hdr.index("Instructions Executed")
    sm = hdr.index("# Samples")
    try:
        n = int(r[ie] or 0)
        smp = int(r[sm] or 0)
    except ValueError:
        continue
    if n:
        acc[(cur_file, line)] = (n, smp, r[1].strip()[:110])
        total += n
print("total warp-instructions (source rows):", total)
for (f, line), (n, smp, src) in sorted(acc.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f}:{line:<5d} {n:>12d} {100.0 * n / total:5.1f}%  smp {smp:>6d}  {src}")

# optional: sums over line ranges of one file:  ... rep top file:lo-hi[,lo-hi...]
if len(sys.argv) > 3:
    f, spec = sys.argv[3].split(":")
    for part in spec.split(","):
        lo, hi = (int(x) for x in part.split("-"))
        tot = sum(n for (ff, line), (n, _, _) in acc.items() if ff == f and lo <= line <= hi)
        print(f"{f}:{lo}-{hi}  {tot}  ({100.0 * tot / total:.1f}% of source rows)")
