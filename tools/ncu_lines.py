"""Aggregate an ncu report's warp-stall samples per CUDA source line.
usage: python tools/ncu_lines.py report.ncu-rep [kernel-regex] [top]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else None
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
if kre:
    cmd += ["--kernel-name", "regex:" + kre]
txt = subprocess.run(cmd, capture_output=True, text=True).stdout
agg = collections.defaultdict(lambda: [0, 0, collections.Counter(), ""])
fname, hdr = None, None
srcs = {}
for row in csv.reader(io.StringIO(txt)):
    if not row:
        continue
    if row[0] == "File Path":
        fname = row[1].split("/")[-1]; hdr = None; continue
    if row[0] == "Function Name":
        continue
    if row[0] == "Line No":
        hdr = row
        si = hdr.index("# Samples"); ie = hdr.index("Instructions Executed")
        stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(row) < len(hdr):
        continue
    line = row[0]
    if row[1].strip():
        srcs[(fname, line)] = row[1].strip()
    try:
        s = int(row[si]); e = int(row[ie])
    except ValueError:
        continue
    a = agg[(fname, line)]
    a[0] += s; a[1] += e
    for i, h in stall_cols:
        try:
            a[2][h[6:]] += int(row[i])
        except ValueError:
            pass
tot = sum(a[0] for a in agg.values())
print("total samples", tot)
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = ",".join("%s:%d" % (k, v) for k, v in a[2].most_common(3))
    print("%5.1f%% %7d ex=%8d %-14s L%-4s %-40s | %s" % (100.0 * a[0] / max(tot, 1), a[0], a[1], f, l, st, srcs.get((f, l), "")[:90]))
