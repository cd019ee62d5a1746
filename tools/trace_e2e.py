import os, sys, ctypes as C, time
import numpy as np
sys.path.insert(0, os.getcwd())
from srla_b200 import encoder as E
from srla_b200.workload import make_blocks_workload
pcm = make_blocks_workload(10000, 4096, 2, 16)
pcm32 = np.ascontiguousarray(pcm.astype(np.int32))
enc = E.Encoder(max_channels=2, max_block=4096)
assert enc.set_parameter(2, 16, 48000, 4096, 4096, 4096, 0, 4) == E.OK
lib = enc.lib
rows = (C.POINTER(C.c_int32) * 2)()
for ch in range(2): rows[ch] = C.cast(pcm32[ch].ctypes.data, C.POINTER(C.c_int32))
cap = enc.max_encoded_size(pcm32.shape[1])
out = np.empty(cap, dtype=np.uint8); size = C.c_uint32(0)
lib.SRLAEncoder_EncodeWhole.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_int32)), C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.c_void_p]
for it in range(6):
    t0 = time.perf_counter()
    rc = lib.SRLAEncoder_EncodeWhole(enc.handle, rows, pcm32.shape[1], out.ctypes.data, cap, C.byref(size), None)
    dt = time.perf_counter() - t0
    print("call", it, rc, size.value, "%.3f ms  %.1f Msamples/s" % (dt * 1e3, pcm32.size / dt / 1e6), flush=True)
