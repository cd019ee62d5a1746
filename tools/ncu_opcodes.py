#!/usr/bin/env python3
"""Executed warp-instructions per SASS opcode for the kernels matching a regex (all functions of the match, i.e. the
kernel plus its non-inlined device functions).  Usage: python tools/ncu_opcodes.py rep.ncu-rep <kernel regex> [N]"""
import csv, subprocess, sys, io, collections, re
rep, pat = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k", "regex:" + pat], capture_output=True, text=True).stdout
hdr, ops = None, collections.Counter()
for r in csv.reader(io.StringIO(out)):
    if r and r[0] == "Address": hdr = r; continue
    if hdr and len(r) == len(hdr):
        ie = hdr.index("Instructions Executed")
        try: n = int(r[ie])
        except ValueError: continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[1])
        if m: ops[m.group(2)] += n
tot = sum(ops.values()) or 1
print("total warp-instructions", tot)
for op, n in ops.most_common(N): print("%-12s %12d  %5.1f%%" % (op, n, 100 * n / tot))
