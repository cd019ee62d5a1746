import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
from srla_b200 import encoder as E
from srla_b200.synth import synth_stereo
from helpers import oracle_encode
pcm = synth_stereo(4096 * 3 + 2304, seed=1234)
got = E.encode(pcm, preset=4, max_block=4096); want = oracle_encode(pcm, preset=4, max_block=4096); print("16-bit", got == want)
pcm24 = synth_stereo(8192 * 2 + 1000, seed=5, bits=24)
got = E.encode(pcm24, bps=24, preset=4, max_block=8192, ltp=3); want = oracle_encode(pcm24, bps=24, preset=4, max_block=8192, ltp=3); print("24-bit ltp", got == want)
v = synth_stereo(16384 + 3000, seed=6)
kw = dict(preset=2, max_block=4096, min_block=1024, lookahead=16384)
got = E.encode(v, **kw); want = oracle_encode(v, **kw); print("variable", got == want)
# ---- round-1 additions: WAV ingest, SVR refinement, decoder ----
from srla_b200 import decoder as D
pay = np.ascontiguousarray(pcm.T).astype("<i2").view(np.uint8).reshape(-1)
with E.Encoder(max_channels=2, max_block=4096) as enc:
    assert enc.set_parameter(2, 16, 48000, 4096, 4096, 4096, 0, 4) == E.OK
    o, offs = enc.encode_interleaved_host([pay, pay[: 4 * 5001]])
    print("wav ingest", o[:offs[1]].tobytes() == oracle_encode(pcm, preset=4, max_block=4096))
    assert enc.set_parameter(2, 16, 48000, 4096, 4096, 4096, 0, 3, 2) == E.OK
    svr = enc.encode_whole(pcm[:, :4096 + 600])
with D.Decoder() as dec:
    print("decode", np.array_equal(dec.decode_whole(got), v), np.array_equal(dec.decode_whole(svr), pcm[:, :4096 + 600]))
    s24 = E.encode(pcm24, bps=24, preset=4, max_block=8192, ltp=3)
    print("decode 24-bit ltp", np.array_equal(dec.decode_whole(s24), pcm24))
    bad = bytearray(s24); bad[200] ^= 1
    print("corrupt ->", dec.decode_whole_rc(bytes(bad), 2, pcm24.shape[1])[0])
with D.Decoder(check_checksum=False) as dec:
    print("corrupt, unchecked ->", dec.decode_whole_rc(bytes(bad), 2, pcm24.shape[1])[0])
# ---- round-2 additions: odd tails (front_tail_kernel), residual16_kernel's bulk copies incl. the plain-load fallback for a tail that
# is not a multiple of 8, blocks beyond the shared-memory capacity and an odd block size (front_big / residual_big / emit_big),
# a block of 16384 samples (front_kernel<512>), a hostile stream with a zero-sample compressed block ----
odd = synth_stereo(4096 * 2 + 1235, seed=21)
print("odd tail", E.encode(odd, preset=4, max_block=4096) == oracle_encode(odd, preset=4, max_block=4096))
print("odd tail ltp", E.encode(odd[:, :4096 + 101], preset=3, max_block=4096, ltp=3) == oracle_encode(odd[:, :4096 + 101], preset=3, max_block=4096, ltp=3))
with E.Encoder(max_channels=2, max_block=4096) as enc:
    assert enc.set_parameter(2, 16, 48000, 4096, 4096, 4096, 0, 4) == E.OK
    many = [synth_stereo(n, seed=30 + i).astype(np.int16) for i, n in enumerate((4096 * 3, 4096 + 7, 9001, 100))]
    o, offs = enc.encode_streams_host(many)
    print("batch with ragged tails", all(o[offs[i]:offs[i + 1]].tobytes() == oracle_encode(m.astype(np.int32), preset=4, max_block=4096) for i, m in enumerate(many)))
big = synth_stereo(20000 * 2 + 333, seed=22)
print("block 20000", E.encode(big, preset=3, max_block=20000) == oracle_encode(big, preset=3, max_block=20000))
print("block 16384", E.encode(big, preset=3, max_block=16384) == oracle_encode(big, preset=3, max_block=16384))
print("odd block size 4095", E.encode(big[:, :13000], preset=2, max_block=4095) == oracle_encode(big[:, :13000], preset=2, max_block=4095))
w24 = synth_stereo(17000 * 2 + 55, seed=23, bits=24)
print("block 17000 24-bit ltp", E.encode(w24, bps=24, preset=2, max_block=17000, ltp=3) == oracle_encode(w24, bps=24, preset=2, max_block=17000, ltp=3))
with D.Decoder() as dec:
    s = bytearray(E.encode(odd, preset=4, max_block=4096))
    s[30 + 9] = 0; s[30 + 10] = 0                      # first block announces zero samples
    print("zero-sample block ->", dec.decode_whole_rc(bytes(s), 2, odd.shape[1])[0])
# ---- second half of round 2: front16_kernel (rows by cp.async.bulk, plain-load fallback, offset shift, mono), the lane split of a
# device-resident-sized call, variable blocks on odd lengths (front_tail_kernel's call chains) ----
sh = (synth_stereo(4096 * 2 + 37, seed=41) // 4) * 4                         # common trailing zeros: offset shift 2, ragged tail
print("front16 shifted + ragged", E.encode(sh, preset=4, max_block=4096) == oracle_encode(sh, preset=4, max_block=4096))
mono = synth_stereo(4096 * 2 + 100, seed=42)[:1]
print("front16 mono", E.encode(mono, preset=4, max_block=4096) == oracle_encode(mono, preset=4, max_block=4096))
for n, v in ((16384 + 2049, 1), (9001, 2), (777, 1)):
    x = synth_stereo(n, seed=43 + n)
    kw = dict(preset=4, max_block=4096, min_block=4096 >> v, lookahead=16384)
    print("variable blocks, odd length", n, v, E.encode(x, **kw) == oracle_encode(x, **kw))
long_ = synth_stereo(4096 * 2100, seed=44)                                    # >= 2048 blocks: three groups on three lanes
with E.Encoder(max_channels=2, max_block=4096) as enc:
    assert enc.set_parameter(2, 16, 48000, 4096, 4096, 4096, 0, 4) == E.OK
    o, offs = enc.encode_streams_host([long_.astype(np.int16)])
    import hashlib
    print("lane split, 2100 blocks", len(o[:offs[1]]), hashlib.sha1(o[:offs[1]].tobytes()).hexdigest()[:12])
