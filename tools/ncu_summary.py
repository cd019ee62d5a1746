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


def traffic(rep):
    """{short kernel name: {...}} for bench.py (profiles/ncu_kernels.json); averages over the captured launches."""
    acc = {}
    for d in raw(rep):
        name = d["kernel"].split("(")[0].split("<")[0].replace("void ", "").replace("srla::", "").strip()
        def val(k):
            v, u = d.get(k, ("0", ""))
            f = float(v.replace(",", ""))
            return f * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(u, 1.0)
        e = acc.setdefault(name, {"n": 0, "dram": 0.0, "t": 0.0, "fp64": 0.0, "issue": 0.0, "lsu_smem": 0.0, "alu": 0.0, "fma": 0.0, "warps": 0.0})
        e["n"] += 1; e["dram"] += val("dram__bytes_read.sum") + val("dram__bytes_write.sum"); e["t"] += val("gpu__time_duration.sum")
        e["fp64"] += val("sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active"); e["issue"] += val("smsp__issue_active.avg.pct_of_peak_sustained_active")
        e["lsu_smem"] += val("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed")
        e["alu"] += val("sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active"); e["fma"] += val("sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active")
        e["warps"] += val("sm__warps_active.avg.pct_of_peak_sustained_active")
    out = {}
    for k, e in acc.items():
        n = e["n"]
        out[k] = {"dram_bytes_per_launch": int(e["dram"] / n), "ncu_duration_ms": round(e["t"] / n * 1e3, 4), "launches_captured": n,
                  "issue_slots_active_pct": round(e["issue"] / n, 1), "fp64_pipe_pct": round(e["fp64"] / n, 1), "alu_pipe_pct": round(e["alu"] / n, 1),
                  "fma_pipe_pct": round(e["fma"] / n, 1), "shared_mem_wavefronts_pct": round(e["lsu_smem"] / n, 1), "warps_active_pct": round(e["warps"] / n, 1)}
    return out


if __name__ == "__main__":
    rep = sys.argv[1]
    if "--traffic" in sys.argv:
        print(json.dumps(traffic(rep), indent=1)); sys.exit(0)
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
    print(json.dumps({"kernels": raw(rep), "stalls": stalls(rep, top)}, indent=1))
