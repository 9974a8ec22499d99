#!/usr/bin/env python
"""Per-source-line summary of an Nsight Compute report (needs -lineinfo + --import-source on):
    ncu -i rep.ncu-rep --page source --print-source cuda,sass --csv > x.csv ; python scripts/ncu_lines.py x.csv [top]
Prints, per CUDA source line, the share of executed warp instructions, the average active lanes and the share of stall samples."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
fname, hdr, cur, acc = None, None, None, {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None:
        continue
    if r[0]:  # a source line starts a group of SASS rows
        cur = (fname, r[0], r[1].strip()[:100])
        acc.setdefault(cur, [0.0, 0.0, 0.0])
        continue
    d = dict(zip(hdr[2:], r[2:]))
    try:
        ie, te, sm = float(d["Instructions Executed"]), float(d["Thread Instructions Executed"]), float(d["# Samples"])
    except (ValueError, KeyError):
        continue
    a = acc[cur]
    a[0] += ie
    a[1] += te
    a[2] += sm
tot = sum(a[0] for a in acc.values()) or 1
tots = sum(a[2] for a in acc.values()) or 1
print("warp instructions %d, avg active lanes %.1f, samples %d" % (tot, sum(a[1] for a in acc.values()) / tot, tots))
for k, a in sorted(acc.items(), key=lambda kv: -kv[1][2])[:top]:
    if a[0] == 0 and a[2] == 0:
        continue
    print("%5.1f%% samples %5.1f%% instr  lanes %4.1f  %s:%s  %s" % (100 * a[2] / tots, 100 * a[0] / tot, a[1] / a[0] if a[0] else 0, k[0], k[1], k[2]))
