#!/usr/bin/env python3
"""Issue-rate floors of every kernel in an .ncu-rep (ncu --set full --import-source on): executed warp instructions per
SASS opcode, grouped by the pipe that executes them, times the issue cost measured on this pool's B200 with
tools/pipe_rates.cu (lanes/clk/SM: DADD, DMUL, DFMA, IMAD, IDP, SHF, VIADDMNMX ~63 = one warp instruction per 2 clocks per
scheduler; IADD3, LOP3, MOV, ISETP, SEL, PRMT ~117 = one per clock; I2F.F64 / F2I.F64 ~16 = one per 8 clocks).

  floor_ms(pipe) = sum(count x clocks per warp instruction) / (SMs x 4 schedulers x SM clock)
  issue_floor_ms = the largest of {every instruction at one issue slot, FP64 pipe, FMA pipe, ALU pipe}

Usage: python tools/ncu_floors.py rep.ncu-rep [--sms 148] [--mhz 1965] > profiles/rN_kernel_floors.json
"""
import collections, csv, io, json, re, subprocess, sys

rep = sys.argv[1]
sms = int(sys.argv[sys.argv.index("--sms") + 1]) if "--sms" in sys.argv else 148
mhz = float(sys.argv[sys.argv.index("--mhz") + 1]) if "--mhz" in sys.argv else 1965.0

FP64 = {"DADD": 2, "DMUL": 2, "DFMA": 2, "DSETP": 2, "DMNMX": 2}
FMA = {"IMAD": 2, "IDP": 2, "FFMA": 1, "FMUL": 1, "FADD": 1, "HFMA2": 1, "IMUL": 2}
ALU_HALF = {"SHF": 2, "VIADDMNMX": 2, "VIMNMX": 2, "VIADD": 1, "LEA": 1, "POPC": 2, "FLO": 2, "BREV": 2, "REDUX": 2}
CONV = {"I2F": 8, "F2I": 8, "F2F": 8, "I2I": 2, "MUFU": 8}
LSU = ("LD", "ST", "ATOM", "RED", "LDS", "STS", "LDG", "STG", "LDL", "STL", "LDSM", "UBLKCP", "LDGSTS", "SHFL", "MATCH", "VOTE")

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
kern, hdr = None, None
ops = collections.defaultdict(collections.Counter)
launches = collections.Counter()
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "Kernel Name":
        kern = r[1].split("(")[0].split("<")[0].replace("void ", "").replace("srla::", "").strip(); launches[kern] += 1; continue
    if r[0] == "Address":
        hdr = r; continue
    if hdr and len(r) == len(hdr) and kern:
        try:
            n = int(r[hdr.index("Instructions Executed")])
        except ValueError:
            continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[1])
        if m and n:
            ops[kern][m.group(2)] += n

res = {}
per_ms = sms * 4 * mhz * 1e3                      # scheduler-clocks per millisecond
for k, c in ops.items():
    n = launches[k]
    tot = sum(c.values()) / n
    cyc = {"issue": tot, "fp64": 0.0, "fma": 0.0, "alu": 0.0, "lsu": 0.0}
    grp = collections.Counter()
    for op, cnt in c.items():
        cnt = cnt / n
        if op in FP64:
            cyc["fp64"] += cnt * FP64[op]; grp["fp64"] += cnt
        elif op in CONV:
            cyc["fp64" if op in ("I2F", "F2I", "F2F") else "alu"] += cnt * CONV[op]; grp["convert"] += cnt
        elif op in FMA:
            cyc["fma"] += cnt * FMA[op]; grp["fma_pipe(imad/idp)"] += cnt
        elif op in ALU_HALF:
            cyc["alu"] += cnt * ALU_HALF[op]; grp["alu"] += cnt
        elif op.startswith(LSU):
            cyc["lsu"] += cnt; grp["lsu"] += cnt
        elif op.startswith(("BRA", "BAR", "BSSY", "BSYNC", "EXIT", "CALL", "RET", "WARPSYNC", "NANOSLEEP", "SYNCS", "ERRBAR", "MEMBAR", "FENCE")):
            grp["control"] += cnt
        elif op.startswith(("U", "R2UR", "S2UR", "CS2R", "S2R")):
            grp["uniform/special"] += cnt
        else:
            cyc["alu"] += cnt; grp["alu"] += cnt
    floors = {p: v / per_ms for p, v in cyc.items()}
    res[k] = {"launches_captured": n, "warp_instructions": int(tot),
              "instruction_groups": {g: int(v) for g, v in grp.most_common()},
              "top_opcodes": {op: int(v / n) for op, v in c.most_common(14)},
              "floor_ms": {p: round(v, 4) for p, v in floors.items()},
              "issue_floor_ms": round(max(floors.values()), 4), "binding": max(floors, key=floors.get)}
print(json.dumps({"sms": sms, "sm_mhz": mhz, "method": __doc__.split("Usage")[0].strip(), "kernels": res}, indent=1))
