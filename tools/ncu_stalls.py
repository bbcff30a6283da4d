#!/usr/bin/env python
"""Per-CUDA-source-line stall-reason breakdown of an .ncu-rep captured with --import-source on.

    python tools/ncu_stalls.py rep.ncu-rep <source file name> [first line] [last line] [topN]

Prints, for the lines of that file with the most stall samples, the samples per stall reason (every column of the source
page whose name starts with "stall_") and the instructions executed.
"""
import csv, io, subprocess, sys
from collections import defaultdict

rep, fname = sys.argv[1], sys.argv[2]
lo = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4]) if len(sys.argv) > 4 else 10 ** 9
top = int(sys.argv[5]) if len(sys.argv) > 5 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = None; fpath = ""; cur = None
acc = defaultdict(lambda: defaultdict(float)); src = {}
cols = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fpath = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No":
        hdr = r
        cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") or h in ("Instructions Executed", "Warp Stall Sampling (All Samples)")]
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[0] != "":
        cur = (fpath, int(r[0])); src[cur] = r[1]
        continue
    if cur is None or cur[0] != fname or not (lo <= cur[1] <= hi):
        continue
    for i in cols:
        try:
            acc[cur][hdr[i]] += float(r[i])
        except ValueError:
            pass
tot = defaultdict(float)
for k in acc:
    for h, v in acc[k].items():
        tot[h] += v
print("file %s lines %d-%d: totals" % (fname, lo, hi))
for h, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v > 0:
        print("   %-50s %12.0f" % (h, v))
key = "Warp Stall Sampling (All Samples)"
for k in sorted(acc, key=lambda k: -acc[k].get(key, 0))[:top]:
    a = acc[k]
    reasons = ", ".join("%s %.0f" % (h.replace("stall_", ""), v) for h, v in sorted(a.items(), key=lambda kv: -kv[1])
                        if h.startswith("stall_") and v > 0.02 * max(a.get(key, 1), 1))
    print("%s:%d  inst %.0f samples %.0f | %s | %s" % (k[0], k[1], a.get("Instructions Executed", 0), a.get(key, 0), reasons, src[k].strip()[:90]))
