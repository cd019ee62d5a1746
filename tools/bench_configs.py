#!/usr/bin/env python3
"""Throughput of the other BASELINE.json configs (they are parity-test cases, not bench.py lines; this table is
extra evidence that the widened paths run at scale).  Every configuration is encoded through the batch C ABI with
pinned host buffers (H2D + kernels + D2H timed, best of 3 after a warm-up), a slice of it is encoded by the
compiled reference (oracle/_ref) on one host thread and compared byte for byte, and the single-thread reference
rate is reported beside it.

  config 3: 48 kHz / 24-bit stereo, block 8192, mode 4, LTP order 3
  config 4: 16-bit stereo, variable blocks -V 2 -L 4 (min 1024, max 4096, look-ahead 16384), mode 4
  config 2 + SVR: the config-2 signal with the SVR coefficient refinement switched on (3 iterations)
  config 5: one GPU's shard of the 1024-file batch: 128 stereo 16-bit files of 30 s (1 440 000 frames = 351 blocks
            of 4096 + a 2304-frame tail each), mode 4, one SRLAB200_EncodeStreamsHost call
Usage: python tools/bench_configs.py [--files 128] [--seconds 60]
"""
from __future__ import annotations

import argparse
import ctypes as C
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
    ap.add_argument("--files", type=int, default=128)
    ap.add_argument("--seconds", type=int, default=900, help="length of the config 3 / 4 streams (900 s = 86.4 M channel-samples)")
    ap.add_argument("--only", type=str, default="", help="run only the configurations whose name starts with this (e.g. 3)")
    args = ap.parse_args()
    import torch
    from helpers import have_ref, ref_encode
    from srla_b200 import encoder as E
    from srla_b200.workload import make_blocks_workload

    def pinned(a: np.ndarray) -> np.ndarray:
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t.numpy()

    def payload(a: np.ndarray, bits: int) -> np.ndarray:
        """planar [channels, frames] -> pinned bytes of the WAV data chunk that holds them"""
        inter = np.ascontiguousarray(a.T)
        if bits == 16:
            return pinned(inter.astype("<i2").view(np.uint8).reshape(-1))
        return pinned(np.ascontiguousarray(inter.astype("<i4").view(np.uint8).reshape(-1, 4)[:, :3]).reshape(-1))

    def run(name, streams, bits, min_block, max_block, lookahead, ltp, ref_frames, svr=0):
        if args.only and not name.startswith(args.only):
            return
        with E.Encoder(max_channels=2, max_block=max_block, min_block=min_block, lookahead=lookahead) as enc:
            assert enc.set_parameter(2, bits, 48000, min_block, max_block, lookahead, ltp, 4, svr) == E.OK
            cap = sum(enc.max_encoded_size(s.shape[1]) for s in streams)
            out = pinned(np.empty(cap, dtype=np.uint8))
            best = None
            for it in range(4):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                out, offs = enc.encode_streams_host(streams, out)
                dt = time.perf_counter() - t0
                if it:
                    best = dt if best is None else min(best, dt)
            st = enc.stats()
            samples = sum(s.size for s in streams)
            # per-kernel times: the call above overlaps the kernels of its groups on three lanes, so the time between two of its
            # events is not one kernel's; a handle with ONE lane runs the groups one after the other (the copies still overlap)
            os.environ["SRLA_B200_LANES"] = "1"
            try:
                with E.Encoder(max_channels=2, max_block=max_block, min_block=min_block, lookahead=lookahead) as e1:
                    assert e1.set_parameter(2, bits, 48000, min_block, max_block, lookahead, ltp, 4, svr) == E.OK
                    for _ in range(2):
                        e1.encode_streams_host(streams, out)
                    st1 = e1.stats()
            finally:
                del os.environ["SRLA_B200_LANES"]
            # the same streams as WAV data-chunk payloads (interleaved; 3-byte samples for 24 bits): SRLAB200_EncodeInterleavedHost
            raws = [payload(s, bits) for s in streams]
            wav_best = None
            for it in range(4):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                out2, offs2 = enc.encode_interleaved_host(raws, out)
                dt = time.perf_counter() - t0
                if it:
                    wav_best = dt if wav_best is None else min(wav_best, dt)
            assert list(offs2) == list(offs)
            line = {"config": name, "streams": len(streams), "Msamples": samples / 1e6, "ms": best * 1e3,
                    "e2e_Msamples_per_s": samples / best / 1e6,
                    "wav_payload_e2e_Msamples_per_s": samples / wav_best / 1e6, "blocks": int(st.num_blocks), "analysed_blocks": int(st.num_analysed),
                    "compression": offs[-1] / float(st.bytes_in),
                    "device_kernel_ms": {"front": round(st1.ms_front, 3), "lpc": round(st1.ms_lpc, 3), "residual": round(st1.ms_residual, 3),
                                         "decide+scan+emit": round(st1.ms_emit, 3), "all_analysis_passes": round(st1.ms_analyse, 3),
                                         "how": "one lane (SRLA_B200_LANES=1): the groups' kernels do not overlap"},
                    "device_Msamples_per_s": samples / max(1e-9, (st1.ms_analyse + st1.ms_emit) * 1e-3) / 1e6}
            if have_ref():
                sl = np.ascontiguousarray(streams[0][:, :ref_frames].astype(np.int32))
                t0 = time.perf_counter()
                want = ref_encode(sl, bps=bits, max_block=max_block, min_block=min_block, lookahead=lookahead, ltp=ltp, preset=4, svr=svr)
                dt = time.perf_counter() - t0
                with E.Encoder(max_channels=2, max_block=max_block, min_block=min_block, lookahead=lookahead) as e2:
                    assert e2.set_parameter(2, bits, 48000, min_block, max_block, lookahead, ltp, 4, svr) == E.OK
                    got = e2.encode_whole(sl) if min_block == max_block else E.encode(sl, bps=bits, max_block=max_block, min_block=min_block, lookahead=lookahead, ltp=ltp, preset=4)
                line["reference_1thread_Msamples_per_s"] = sl.size / dt / 1e6
                line["identical_to_reference_on_slice"] = bool(got == want)
            print(json.dumps(line), flush=True)

    n = 48000 * args.seconds
    wide = make_blocks_workload((n + 8191) // 8192, 8192, 2, 24, seed=77, num_templates=2, template_blocks=24)[:, :n]
    run("3: 24-bit stereo, block 8192, mode 4, LTP 3", [pinned(wide.astype(np.int32))], 24, 8192, 8192, 8192, 3, 8192 * 12)
    narrow = make_blocks_workload((n + 4095) // 4096, 4096, 2, 16, seed=78, num_templates=2, template_blocks=48)[:, :n]
    run("4: 16-bit stereo, -V 2 -L 4 (1024..4096, look-ahead 16384), mode 4", [pinned(narrow.astype(np.int16))], 16, 1024, 4096, 16384, 0, 16384 * 6)
    short = narrow[:, :48000 * min(args.seconds, 60)]                     # the SVR refinement is a slow optional mode: one minute
    run("2 + SVR: 16-bit stereo, block 4096, mode 4, --svr-filter-learning-iteration 3", [pinned(short.astype(np.int16))], 16, 4096, 4096, 4096, 0, 4096 * 4, svr=3)
    big = make_blocks_workload((n // 4 + 65534) // 65535, 65535, 2, 16, seed=80, num_templates=2, template_blocks=4)[:, :n // 4]
    run("blocks of 65535 samples (odd: serial-stream mode), 16-bit stereo, mode 4", [pinned(big.astype(np.int16))], 16, 65535, 65535, 65535, 0, 65535 * 2)
    run("blocks of 32768 samples, 16-bit stereo, mode 4", [pinned(big.astype(np.int16))], 16, 32768, 32768, 32768, 0, 32768 * 2)
    frames = 1_440_000
    base = make_blocks_workload((frames + 4095) // 4096 + 1, 4096, 2, 16, seed=79, num_templates=2, template_blocks=88)
    files = [pinned(np.roll(base, 4099 * k, axis=1)[:, :frames].astype(np.int16)) for k in range(args.files)]
    run(f"5 (one GPU's shard): {args.files} stereo 16-bit files x 30 s, block 4096, mode 4", files, 16, 4096, 4096, 4096, 0, 4096 * 40)


if __name__ == "__main__":
    main()
