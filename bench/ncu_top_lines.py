#!/usr/bin/env python3
"""Aggregate an ncu `--page source --csv --print-source cuda,sass` dump by CUDA source line.
usage: ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > src.csv; ncu_top_lines.py src.csv [N]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
agg = collections.defaultdict(lambda: [0, 0, 0.0, ""])
fname = "?"
hdr = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Name":
        fname = r[1].split("/")[-1]
        continue
    if "# Samples" in r:
        hdr = r
        iS = hdr.index("# Samples"); iI = hdr.index("Instructions Executed"); iT = hdr.index("Thread Instructions Executed")
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    try:
        ln = int(r[0])
    except Exception:
        continue
    try:
        s = int(r[iS] or 0); ins = int(r[iI] or 0); ti = float(r[iT] or 0)
    except Exception:
        continue
    a = agg[(fname, ln)]
    a[0] += s; a[1] += ins; a[2] += ti
    if r[1].strip():
        a[3] = r[1].strip()
tot = sum(a[0] for a in agg.values()) or 1
toti = sum(a[1] for a in agg.values()) or 1
print("total samples", tot, "warp instructions", toti)
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:N]:
    thr = a[2] / a[1] if a[1] else 0
    print("%5.1f%% smp %5.1f%% inst thr/inst %4.1f  %s:%d  %s" % (100.0 * a[0] / tot, 100.0 * a[1] / toti, thr, f, ln, a[3][:90]))
