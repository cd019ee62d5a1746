#!/usr/bin/env python3
"""Top-sampled SASS instructions of a kernel with a few lines of context (where do warps actually wait?).
Usage: python tools/ncu_sass_hot.py rep.ncu-rep <kernel regex> [N]"""
import csv, subprocess, sys, io
rep, pat = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 12
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k", "regex:" + pat], capture_output=True, text=True).stdout
hdr, data = None, []
for r in csv.reader(io.StringIO(out)):
    if r and r[0] == "Address":
        if hdr is not None and data: break          # first function only
        hdr = r; continue
    if hdr and len(r) == len(hdr): data.append(r)
si, src = 2, 1
tot = sum(int(r[si]) for r in data if r[si].isdigit()) or 1
for i in sorted(range(len(data)), key=lambda i: -int(data[i][si]))[:N]:
    print("---- #%d  %.1f%% of samples" % (i, 100 * int(data[i][si]) / tot))
    for j in range(max(0, i - 4), min(len(data), i + 2)):
        print("    %6d %-90s %s" % (j, data[j][src][:90], data[j][si]))
