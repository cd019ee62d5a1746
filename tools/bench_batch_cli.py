#!/usr/bin/env python3
"""End-to-end timing of the many-file front end (SURVEY 8f N1, config 5 shape): N stereo 16-bit WAV files of 30 s
in a RAM disk -> srla_b200_batch -> .srl files, next to the reference CLI (`srla -e`, one process per file as a
user would run it) timed on a few of the same files and compared byte for byte.
Usage: python tools/bench_batch_cli.py [--files 128] [--dir /dev/shm/srla_cli_bench]"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import struct
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def write_wav(path: str, pcm: np.ndarray, rate: int = 48000) -> None:
    data = np.ascontiguousarray(pcm.T).astype("<i2").tobytes()
    nch = pcm.shape[0]
    fmt = struct.pack("<HHIIHH", 1, nch, rate, rate * nch * 2, nch * 2, 16)
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVE" + b"fmt " + struct.pack("<I", 16) + fmt + b"data" + struct.pack("<I", len(data)))
        f.write(data)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--files", type=int, default=128)
    ap.add_argument("--dir", default="/dev/shm/srla_cli_bench")
    ap.add_argument("--ref-files", type=int, default=2)
    args = ap.parse_args()
    from srla_b200.workload import make_blocks_workload
    tool = os.path.join(ROOT, "srla_b200", "srla_b200_batch")
    ref_cli = os.path.join(ROOT, "oracle", "_ref", "srla_ref")
    shutil.rmtree(args.dir, ignore_errors=True)
    os.makedirs(args.dir)
    frames = 1_440_000
    base = make_blocks_workload((frames + 4095) // 4096 + 1, 4096, 2, 16, seed=79, num_templates=2, template_blocks=88)
    names = []
    for k in range(args.files):
        p = os.path.join(args.dir, f"f{k:04d}.wav")
        write_wav(p, np.roll(base, 4099 * k, axis=1)[:, :frames])
        names.append(p)
    opts = ["-m", "4", "-B", "4096", "-V", "0"]
    out_dir = os.path.join(args.dir, "out")
    best = None
    for _ in range(3):
        shutil.rmtree(out_dir, ignore_errors=True)
        t0 = time.perf_counter()
        r = subprocess.run([tool] + opts + ["--timing", "-o", out_dir] + names, capture_output=True, text=True)
        dt = time.perf_counter() - t0
        assert r.returncode == 0, r.stderr[-2000:]
        best = dt if best is None else min(best, dt)
    samples = args.files * frames * 2
    line = {"files": args.files, "Msamples": samples / 1e6, "batch_cli_wall_s": best, "batch_cli_Msamples_per_s": samples / best / 1e6,
            "batch_cli_summary": r.stdout.strip().splitlines()[-1], "batch_cli_stages": r.stderr.strip().splitlines()[-1]}
    if os.path.exists(ref_cli):
        t = 0.0
        same = True
        for p in names[:args.ref_files]:
            want = p.replace(".wav", ".ref.srl")
            t0 = time.perf_counter()
            subprocess.run([ref_cli, "-e"] + opts + [p, want], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            t += time.perf_counter() - t0
            same = same and open(want, "rb").read() == open(os.path.join(out_dir, os.path.basename(p).replace(".wav", ".srl")), "rb").read()
        line["reference_cli_Msamples_per_s_one_process"] = args.ref_files * frames * 2 / t / 1e6
        line["identical_to_reference_cli"] = same
    print(json.dumps(line), flush=True)
    shutil.rmtree(args.dir, ignore_errors=True)


if __name__ == "__main__":
    main()
