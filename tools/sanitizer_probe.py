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
