import os, sys, time, ctypes as C
import numpy as np
sys.path.insert(0, os.getcwd())
from srla_b200 import encoder as E
from srla_b200.synth import synth_stereo
pcm = np.tile(synth_stereo(4096 * 8, seed=29), (1, 1300))[:, :4096 * 10000]
for it in range(3):
    ts = []
    t0 = time.perf_counter()
    out = E.encode(pcm, preset=4, max_block=4096, callback=lambda n, p, ptr, size: ts.append(time.perf_counter() - t0))
    t1 = time.perf_counter() - t0
    ts = np.array(ts)
    print("call %.1f ms; callbacks %d; first at %.2f ms, median at %.2f ms, last at %.2f ms" % (t1 * 1e3, len(ts), ts[0] * 1e3, np.median(ts) * 1e3, ts[-1] * 1e3))
