#!/usr/bin/env python3
"""Per-source-line instruction / stall histogram of one kernel from an .ncu-rep (needs -lineinfo, --import-source on).
Usage: python tools/ncu_lines.py rep.ncu-rep <kernel-substring> [N] [--by samples|inst] [--stalls]"""
import csv, io, subprocess, sys, collections
rep, pat = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3].isdigit() else 40
by = "samples" if "--by" in sys.argv and sys.argv[sys.argv.index("--by") + 1] == "samples" else "inst"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, hdr, data, hdrs = None, None, collections.defaultdict(list), {}
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0] in ("Kernel Name", "Function Name"): cur = r[1]; continue
    if r[0] == "Line No": hdr = r; hdrs[cur] = r; continue      # the stall columns differ from kernel to kernel
    if hdr and r[0].isdigit() and len(r) > 8 and r[2] == '-': data[cur].append(r)
for k, v in data.items():
    if k is None or pat not in k: continue
    hdr = hdrs.get(k, hdr)
    ti = sum(int(r[7]) for r in v) or 1; ts = sum(int(r[4]) for r in v) or 1
    print("##", k, "warp-inst", ti, "samples", ts)
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = collections.Counter()
    for r in v:
        for i in stall_cols:
            try: tot[hdr[i]] += int(r[i])
            except ValueError: pass
    print("   stall totals:", ", ".join(f"{a}={100*b/ts:.1f}%" for a, b in tot.most_common(8)))
    v.sort(key=lambda r: -(int(r[7]) if by == "inst" else int(r[4])))
    for r in v[:N]:
        extra = ""
        if "--stalls" in sys.argv:
            st = sorted(((int(r[i]) if r[i].isdigit() else 0, hdr[i][6:]) for i in stall_cols), reverse=True)[:3]
            extra = " | " + " ".join(f"{b}:{a}" for a, b in st if a)
        print("%5.1f%% inst %5.1f%% smp  L%-5s %s%s" % (100 * int(r[7]) / ti, 100 * int(r[4]) / ts, r[0], r[1].strip()[:100], extra))
