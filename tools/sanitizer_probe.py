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
