#!/usr/bin/env python
"""Summarise an ncu report's source page per CUDA source line (dev tool):
python tools/ncu_lines.py report.ncu-rep [kernel-substring] [top]"""
import csv
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname = None
agg = defaultdict(lambda: [0, 0, ""])
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] in ("File Name", "File Path"):
        fname = r[1].split("/")[-1]
        hdr = None
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[0] == "":
        continue  # SASS row belonging to the previous source line (already aggregated there)
    try:
        ln = int(r[0])
    except ValueError:
        continue
    def num(v):
        try:
            return int(v)
        except ValueError:
            return 0
    si = hdr.index("# Samples") if "# Samples" in hdr else None
    ii = hdr.index("Instructions Executed") if "Instructions Executed" in hdr else None
    if si is None or ii is None:
        continue
    a = agg[(fname, ln)]
    a[0] += num(r[si])
    a[1] += num(r[ii])
    a[2] = r[1].strip()
ts = sum(a[0] for a in agg.values()) or 1
ti = sum(a[1] for a in agg.values()) or 1
print("total samples %d, instructions %d" % (ts, ti))
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% smp %5.1f%% inst  %s:%d  %s" % (100 * a[0] / ts, 100 * a[1] / ti, f, ln, a[2][:90]))
