#!/usr/bin/env python
"""Per-CUDA-source-line instruction and stall-sample totals of an .ncu-rep captured with
--import-source on (`--page source --print-source cuda,sass`).

    python tools/ncu_lines.py rep.ncu-rep [topN]
"""
import csv, io, subprocess, sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
inst = defaultdict(float); samp = defaultdict(float); src = {}
hdr = None; fpath = ""; cur = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fpath = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No":
        hdr = r
        iI = hdr.index("Instructions Executed"); iS = hdr.index("Warp Stall Sampling (All Samples)")
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[0] != "":
        cur = (fpath, int(r[0])); src[cur] = r[1]
        continue
    if cur is None:
        continue
    try:
        inst[cur] += float(r[iI]); samp[cur] += float(r[iS])
    except ValueError:
        pass
ti = sum(inst.values()) or 1; ts = sum(samp.values()) or 1
print("total warp instructions %.0f, stall samples %.0f" % (ti, ts))
print("%6s %6s  %s" % ("inst%", "samp%", "line"))
for k in sorted(inst, key=lambda k: -inst[k])[:top]:
    print("%6.2f %6.2f  %s:%d  %s" % (100 * inst[k] / ti, 100 * samp[k] / ts, k[0], k[1], src[k].strip()[:110]))
