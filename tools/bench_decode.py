#!/usr/bin/env python3
"""Decode throughput on the config-2 stream: 10 000 stereo 16-bit blocks of 4096 samples, mode 4.
The stream is written by our encoder, decoded by SRLADecoder_DecodeWhole (host bytes in, host planar int32 out:
H2D of the stream, one kernel launch, D2H of 328 MB of PCM), checked against the original samples, and timed next
to the reference decoder on one host thread over a slice.
Usage: python tools/bench_decode.py [--blocks 10000]"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=10000)
    args = ap.parse_args()
    from helpers import have_ref, ref_decode
    from srla_b200 import decoder as D
    from srla_b200 import encoder as E
    from srla_b200.workload import make_blocks_workload
    pcm = np.ascontiguousarray(make_blocks_workload(args.blocks, 4096, 2, 16, seed=1234).astype(np.int32))
    with E.Encoder(max_channels=2, max_block=4096) as enc:
        assert enc.set_parameter(2, 16, 48000, 4096, 4096, 4096, 0, 4) == E.OK
        out, offs = enc.encode_streams_host([np.ascontiguousarray(pcm.astype(np.int16))])
    stream = out[:offs[1]].tobytes()
    samples = pcm.size
    with D.Decoder() as dec:
        best, kernel = None, None
        got = np.zeros_like(pcm)                     # the caller's buffers exist before the call, as in the C API
        for it in range(4):
            got[:] = -1
            t0 = time.perf_counter()
            dec.decode_whole(stream, got)
            dt = time.perf_counter() - t0
            if it:
                best = dt if best is None else min(best, dt)
                kernel = dec.kernel_ms() if kernel is None else min(kernel, dec.kernel_ms())
    line = {"workload": f"{args.blocks} stereo 16-bit blocks x 4096, mode 4", "Msamples": samples / 1e6, "stream_bytes": len(stream),
            "identical_to_source": bool(np.array_equal(got, pcm)), "decode_kernel_ms": kernel,
            "kernel_Msamples_per_s": samples / (kernel * 1e-3) / 1e6, "e2e_ms": best * 1e3, "e2e_Msamples_per_s": samples / best / 1e6,
            "e2e_note": "SRLADecoder_DecodeWhole: pageable stream in, pageable planar int32 out (D2H of 4 bytes per sample)"}
    if have_ref():
        cut = 4096 * 200
        with E.Encoder(max_channels=2, max_block=4096) as enc:
            assert enc.set_parameter(2, 16, 48000, 4096, 4096, 4096, 0, 4) == E.OK
            part = enc.encode_whole(pcm[:, :cut])
        t0 = time.perf_counter()
        back = ref_decode(part)
        dt = time.perf_counter() - t0
        line["reference_1thread_Msamples_per_s"] = 2 * cut / dt / 1e6
        line["reference_agrees_on_slice"] = bool(np.array_equal(back, pcm[:, :cut]))
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
