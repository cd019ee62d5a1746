"""Decoder-only probe for compute-sanitizer (memcheck / racecheck): the lockstep walk of decode_parse_kernel (predicated
loads, line prefetches, the general reader for long codes) and the unrolled synthesis rounds of decode_blocks_kernel on
well-formed and on damaged streams -- with the checksum test off, so that the walk really runs through the damage.
Usage: compute-sanitizer --tool memcheck python tools/sanitizer_probe_decode.py"""
import os
import sys

sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
from srla_b200 import decoder as D
from srla_b200 import encoder as E
from srla_b200.synth import synth_stereo

rng = np.random.default_rng(5)
n = 1024
parts = []
for k in range(24):
    kind = k % 4
    if kind == 0:
        seg = synth_stereo(n, seed=100 + k)
    elif kind == 1:
        seg = np.zeros((2, n), dtype=np.int32); seg[:, rng.integers(0, n, size=3)] = rng.integers(-30000, 30000, size=3)
    elif kind == 2:
        seg = rng.integers(-32768, 32768, size=(2, n)).astype(np.int32)
    else:
        seg = (synth_stereo(n, seed=300 + k) >> 5).astype(np.int32)
    parts.append(seg)
parts.append(synth_stereo(333, seed=9))
pcm = np.ascontiguousarray(np.concatenate(parts, axis=1).astype(np.int32))
stream = E.encode(pcm, preset=4, max_block=n)
big = synth_stereo(4096 * 3 + 77, seed=2)
stream_big = E.encode(big, preset=4, max_block=4096)
for lanes in ("0", "32"):
    os.environ["SRLA_B200_DECODE_LANES"] = lanes
    for pipe in ("0", "1"):
        os.environ["SRLA_B200_DECODE_PIPELINE"] = pipe
        with D.Decoder() as dec:
            print("lanes", lanes, "pipeline", pipe, "decode", np.array_equal(dec.decode_whole(stream), pcm), np.array_equal(dec.decode_whole(stream_big), big))
        with D.Decoder(check_checksum=False) as dec:
            codes = []
            for at in range(40, len(stream), max(1, len(stream) // 40)):
                bad = bytearray(stream); bad[at] ^= 0xFF; bad[min(at + 1, len(bad) - 1)] ^= 0x55
                codes.append(dec.decode_whole_rc(bytes(bad), 2, pcm.shape[1])[0])
            print("  damaged, unchecked ->", sorted(set(codes)))
print("done")
