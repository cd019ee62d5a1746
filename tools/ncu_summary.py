#!/usr/bin/env python3
"""Summarise an .ncu-rep (ncu --set full) into a small JSON/markdown: per kernel duration, DRAM bytes,
occupancy, pipe utilisation, issue activity and the top stall lines by source line.
Usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [--top 25]"""
import csv
import io
import json
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.sum.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for i, h in enumerate(hdr):
            if h in WANT:
                d[h] = (r[i], units[i])
        res.append(d)
    return res


def stalls(rep, top):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    res, cur_file, cur_fn, hdr = {}, None, None, None
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]; continue
        if r[0] == "Function Name":
            cur_fn = r[1]; continue
        if r[0] == "Kernel Name":
            cur_fn = r[1]; continue
        if r[0] == "Line No":
            hdr = r; continue
        if hdr and len(r) > 8 and r[2] == "-" and r[0].isdigit():
            try:
                res.setdefault(cur_fn, []).append((int(r[4]), int(r[7]), cur_file, int(r[0]), r[1].strip()[:100]))
            except ValueError:
                pass
    outd = {}
    for fn, lst in res.items():
        tot = sum(a[0] for a in lst) or 1
        toti = sum(a[1] for a in lst) or 1
        lst.sort(reverse=True)
        outd[fn] = {"total_samples": tot, "total_warp_inst": toti,
                    "top": [{"samples_pct": round(100 * a[0] / tot, 1), "inst_pct": round(100 * a[1] / toti, 1), "where": f"{a[2]}:{a[3]}", "src": a[4]} for a in lst[:top]]}
    return outd


if __name__ == "__main__":
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
    print(json.dumps({"kernels": raw(rep), "stalls": stalls(rep, top)}, indent=1))
