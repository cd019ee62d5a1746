#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libsrla_ref.so, built by
`make -C oracle ref` from /root/reference).  Run in the build container only; the fixtures are
committed because /root/reference does not exist on the GPU box.

Each fixture holds the PCM input (so no libm/numpy version can change it), the encode parameters
and the byte stream the reference produced (its handle created on a zeroed work area, i.e. what
a fresh `srla` CLI process emits).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from helpers import ref_encode, ref_decode, reference_test_signals  # noqa: E402
from srla_b200.synth import synth_stereo  # noqa: E402


def save(name, pcm, **kw):
    if os.path.exists(os.path.join(HERE, name + ".npz")) and "--force" not in sys.argv:
        return                                     # fixtures are deterministic; --force regenerates all of them
    pcm = np.ascontiguousarray(pcm, dtype=np.int32)
    srl = ref_encode(pcm, **kw)
    assert np.array_equal(ref_decode(srl), pcm), name
    store = pcm.astype(np.int16) if kw.get("bps", 16) <= 16 else pcm
    keys = sorted(kw)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), pcm=store,
                        srl=np.frombuffer(srl, dtype=np.uint8),
                        param_names=np.array(keys), param_values=np.array([kw[k] for k in keys], dtype=np.int64))
    print(f"{name}: {pcm.shape} -> {len(srl)} bytes")


def main():
    s16 = synth_stereo(4096 * 3 + 2304, seed=1234)
    save("config1_mono_m0", s16[:1, :4096], preset=0, max_block=4096)
    save("stereo16_m4_b4096", s16, preset=4, max_block=4096)
    save("stereo16_m2_b4096", s16[:, :9000], preset=2, max_block=4096)
    save("stereo16_m6_b4096", s16[:, :8192], preset=6, max_block=4096)
    s24 = synth_stereo(8192 * 2 + 1000, seed=77, bits=24)
    save("stereo24_m4_b8192_ltp3", s24, bps=24, preset=4, max_block=8192, ltp=3)
    save("stereo16_m4_v2_l4", synth_stereo(20000, seed=99), preset=4, max_block=4096, min_block=1024, lookahead=16384)
    save("mono16_m4_shifted", (synth_stereo(10000, seed=5)[:1] >> 3) << 3, preset=4, max_block=4096)
    three = synth_stereo(6000, seed=11, channels=3)
    save("three_ch_m3", three, preset=3, max_block=2048)
    # the reference's own round-trip matrix generators (test/srla_encode_decode/main.cpp:51-208), its
    # block configuration (min 512 / max 1024 / look-ahead 2048, preset 0), with and without LTP
    for name, sig in reference_test_signals(n=8500, bps=16, nch=2).items():
        for ltp in (0, 3):
            save(f"refgen_{name}_ltp{ltp}", sig, preset=0, min_block=512, max_block=1024, lookahead=2048, ltp=ltp, rate=44100)
    for name, sig in reference_test_signals(n=4000, bps=24, nch=1, seed=3).items():
        save(f"refgen24_{name}", sig, bps=24, preset=4, max_block=1024, ltp=3, rate=44100)
    # stale-scratch corners of the reference (lpc.c:260-264, 371-373): odd stream / tail lengths, where the Welch
    # window's middle sample is what the previous call left behind, and LTP on a tail shorter than 263 samples
    for n in (1, 3, 4095, 8969, 9001, 9193):
        for ltp in (0, 3):
            save(f"odd{n}_m4_b4096_ltp{ltp}", synth_stereo(n, seed=n), preset=4, max_block=4096, ltp=ltp)
    save("odd7001_mono_m3_b2048", synth_stereo(7001, seed=71)[:1], preset=3, max_block=2048)
    save("odd5001_three_ch_m2_b2048_ltp3", synth_stereo(5001, seed=72, channels=3), preset=2, max_block=2048, ltp=3)
    save("odd16999_24bit_m4_b8192_ltp3", synth_stereo(16999, seed=73, bits=24), bps=24, preset=4, max_block=8192, ltp=3)
    save("ltp_short_tail_4196_m4_b4096", synth_stereo(4196, seed=74), preset=4, max_block=4096, ltp=3)
    save("ltp_short_tail_odd_8391_m4_b4096", synth_stereo(8391, seed=75), preset=4, max_block=4096, ltp=3)
    save("odd9001_m4_v2_l4", synth_stereo(9001, seed=76), preset=4, max_block=4096, min_block=1024, lookahead=16384)
    # the reference CLI's defaults (-m 4 -B 4096 -V 1 -L 4) on odd frame counts: the clipped segments at the end follow one another,
    # a stream end inside the first segment of a chunk, a silent stretch in front of the end, LTP
    save("odd50001_m4_v1_l4", synth_stereo(50001, seed=77), preset=4, max_block=4096, min_block=2048, lookahead=16384)
    save("odd16385_m4_v1_l4", synth_stereo(16385, seed=78), preset=4, max_block=4096, min_block=2048, lookahead=16384)
    quiet = synth_stereo(16384 + 3001, seed=81); quiet[:, 16384 - 100:16384 + 2048] = 0
    save("odd19385_silence_m4_v1_l4", quiet, preset=4, max_block=4096, min_block=2048, lookahead=16384)
    save("odd21385_m3_v2_l2_ltp3", synth_stereo(16384 + 5001, seed=82), preset=3, max_block=4096, min_block=1024, lookahead=8192, ltp=3)
    # blocks beyond 8192 samples (the reference CLI admits -B up to 65535; this implementation up to 16384)
    save("stereo16_m4_b16384", synth_stereo(16384 * 2 + 5001, seed=79), preset=4, max_block=16384)
    save("stereo24_m3_b16384_ltp3", synth_stereo(16384 + 9000, seed=80, bits=24), bps=24, preset=3, max_block=16384, ltp=3)
    save("mono16_m5_b12000", synth_stereo(12000 * 2 + 100, seed=81)[:1], preset=5, max_block=12000)
    # ... and beyond the shared-memory capacity (front_big_kernel: transform in global memory), up to the format's limit;
    # an ODD block size chains every block's analysis to the one before it (front_big_kernel's serial-stream mode)
    save("stereo16_m4_b32768", synth_stereo(32768 * 2 + 12345, seed=82), preset=4, max_block=32768)
    save("stereo16_m4_b65535", synth_stereo(65535 * 2 + 4001, seed=83), preset=4, max_block=65535)
    save("stereo24_m2_b20000_ltp3", synth_stereo(20000 * 2 + 777, seed=84, bits=24), bps=24, preset=2, max_block=20000, ltp=3)
    save("stereo16_m4_b4095", synth_stereo(4095 * 4 + 100, seed=85), preset=4, max_block=4095)
    save("mono16_m3_b1001_silence_inside", np.concatenate([synth_stereo(2002, seed=86)[:1], np.zeros((1, 1001), dtype=np.int32),
                                                           synth_stereo(1500, seed=87)[:1]], axis=1), preset=3, max_block=1001)
    save("odd_after_silence_m4_b4096", np.concatenate([synth_stereo(4096, seed=77), np.zeros((2, 4096), dtype=np.int32),
                                                       synth_stereo(1001, seed=78)], axis=1), preset=4, max_block=4096)


if __name__ == "__main__":
    main()
