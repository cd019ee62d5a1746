"""Host-side mirror of the reference encoder interface (include/srla_encoder.h) on top of the C ABI
of ``libsrla_b200.so`` (include/srla_b200.h).

The reference is a C library; its "operator interface" for the encode path is the handle API
``SRLAEncoder_Create / SetEncodeParameter / EncodeWhole / EncodeBlock / ComputeBlockSize /
EncodeOptimalPartitionedBlock / Destroy`` (tools/srla_codec/srla_codec.c:91-155 is the production
caller).  This module binds exactly those symbols with ``ctypes`` -- plain pointers and sizes, no
torch types -- and adds thin numpy conveniences used by the tests and by bench.py.

There is no fallback: if the shared library is missing or no CUDA device is usable, loading /
``Encoder()`` raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsrla_b200.so")

# result codes, include/srla.h:29-38
OK, INVALID_ARGUMENT, INVALID_FORMAT, INSUFFICIENT_BUFFER, INSUFFICIENT_DATA, PARAMETER_NOT_SET, DATA_CORRUPTION, NG = range(8)
RESULT_NAMES = ["OK", "INVALID_ARGUMENT", "INVALID_FORMAT", "INSUFFICIENT_BUFFER", "INSUFFICIENT_DATA",
                "PARAMETER_NOT_SET", "DETECT_DATA_CORRUPTION", "NG"]
PRESET_MAX_ORDER = [0, 8, 16, 32, 64, 128, 255]


class SRLAHeader(C.Structure):                      # include/srla.h:41-51
    _fields_ = [("format_version", C.c_uint32), ("codec_version", C.c_uint32),
                ("num_channels", C.c_uint16), ("num_samples", C.c_uint32),
                ("sampling_rate", C.c_uint32), ("bits_per_sample", C.c_uint16),
                ("offset_lshift", C.c_uint8), ("max_num_samples_per_block", C.c_uint32),
                ("preset", C.c_uint8)]


class SRLAEncodeParameter(C.Structure):             # include/srla_encoder.h:8-18
    _fields_ = [("num_channels", C.c_uint16), ("bits_per_sample", C.c_uint16),
                ("sampling_rate", C.c_uint32), ("min_num_samples_per_block", C.c_uint32),
                ("max_num_samples_per_block", C.c_uint32), ("num_lookahead_samples", C.c_uint32),
                ("ltp_order", C.c_uint32), ("num_svr_filter_learning_iteration", C.c_uint32),
                ("preset", C.c_uint8)]


class SRLAEncoderConfig(C.Structure):               # include/srla_encoder.h:21-27
    _fields_ = [("max_num_channels", C.c_uint32), ("min_num_samples_per_block", C.c_uint32),
                ("max_num_samples_per_block", C.c_uint32), ("max_num_lookahead_samples", C.c_uint32),
                ("max_num_parameters", C.c_uint32)]


class SRLAB200Stream(C.Structure):
    _fields_ = [("pcm", C.c_void_p), ("channel_stride", C.c_uint64), ("num_samples", C.c_uint32),
                ("sample_bytes", C.c_uint32)]


class SRLAB200Frames(C.Structure):
    _fields_ = [("frames", C.c_void_p), ("num_samples", C.c_uint32)]


class SRLAB200Stats(C.Structure):
    _fields_ = [("num_blocks", C.c_uint64), ("num_analysed", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("bytes_in", C.c_uint64), ("bytes_out", C.c_uint64),
                ("ms_analyse", C.c_float), ("ms_emit", C.c_float), ("ms_total_device", C.c_float),
                ("order_histogram", C.c_uint32 * 256), ("method_histogram", C.c_uint32 * 4),
                ("type_histogram", C.c_uint32 * 3),
                ("ms_front", C.c_float), ("ms_lpc", C.c_float), ("ms_residual", C.c_float)]


class SRLAB200ChannelResult(C.Structure):
    _fields_ = [("pre_coef", C.c_int32), ("pre_prev", C.c_int32), ("order", C.c_uint32), ("rshift", C.c_uint32),
                ("use_sum", C.c_uint32), ("coef", C.c_int32 * 255), ("ltp_period", C.c_uint32),
                ("ltp_coef", C.c_int32 * 3), ("code_type", C.c_uint32), ("porder", C.c_uint32),
                ("residual_bits", C.c_uint32), ("total_bits", C.c_uint32),
                ("autocorr", C.c_double * 256), ("error_vars", C.c_double * 256), ("lpc_double", C.c_double * 255)]


CALLBACK = C.CFUNCTYPE(None, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint8), C.c_uint32)

EXPORTED_SYMBOLS = [
    "SRLAEncoder_EncodeHeader", "SRLAEncoder_CalculateWorkSize", "SRLAEncoder_Create", "SRLAEncoder_Destroy",
    "SRLAEncoder_SetEncodeParameter", "SRLAEncoder_ComputeBlockSize", "SRLAEncoder_EncodeBlock",
    "SRLAEncoder_EncodeOptimalPartitionedBlock", "SRLAEncoder_EncodeWhole",
    "SRLAB200_EncodeStreamsDevice", "SRLAB200_EncodeStreamsHost", "SRLAB200_EncodeInterleavedHost",
    "SRLAB200_AllocPinned", "SRLAB200_FreePinned", "SRLAB200_MaxEncodedSize", "SRLAB200_GetStats",
    "SRLAB200_SetDevice", "SRLAB200_SetStream", "SRLAB200_Version", "SRLAB200_TestAnalyseChannel",
    "SRLADecoder_DecodeHeader", "SRLADecoder_CalculateWorkSize", "SRLADecoder_Create", "SRLADecoder_Destroy",
    "SRLADecoder_SetHeader", "SRLADecoder_DecodeBlock", "SRLADecoder_DecodeWhole", "SRLAB200_DecoderKernelMs",
    "SRLAB200_TestNarrow", "SRLAB200_GetDeviceCount", "SRLAB200_GetDevicePciBusId",
]

_lib: Optional[C.CDLL] = None


def load_library() -> C.CDLL:
    """Load libsrla_b200.so and declare every prototype of include/srla_b200.h.  Raises when the
    library has not been built (``python -c 'import __graft_entry__ as g; g.build()'``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with __graft_entry__.build(); "
                           "the SRLA B200 encode path has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    PP = C.POINTER(C.POINTER(C.c_int32))
    lib.SRLAEncoder_EncodeHeader.argtypes = [C.POINTER(SRLAHeader), C.c_void_p, C.c_uint32]
    lib.SRLAEncoder_EncodeHeader.restype = C.c_int
    lib.SRLAEncoder_CalculateWorkSize.argtypes = [C.POINTER(SRLAEncoderConfig)]
    lib.SRLAEncoder_CalculateWorkSize.restype = C.c_int32
    lib.SRLAEncoder_Create.argtypes = [C.POINTER(SRLAEncoderConfig), C.c_void_p, C.c_int32]
    lib.SRLAEncoder_Create.restype = C.c_void_p
    lib.SRLAEncoder_Destroy.argtypes = [C.c_void_p]
    lib.SRLAEncoder_Destroy.restype = None
    lib.SRLAEncoder_SetEncodeParameter.argtypes = [C.c_void_p, C.POINTER(SRLAEncodeParameter)]
    lib.SRLAEncoder_SetEncodeParameter.restype = C.c_int
    lib.SRLAEncoder_ComputeBlockSize.argtypes = [C.c_void_p, PP, C.c_uint32, C.POINTER(C.c_uint32)]
    lib.SRLAEncoder_ComputeBlockSize.restype = C.c_int
    for name in ("SRLAEncoder_EncodeBlock", "SRLAEncoder_EncodeOptimalPartitionedBlock"):
        f = getattr(lib, name)
        f.argtypes = [C.c_void_p, PP, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
        f.restype = C.c_int
    lib.SRLAEncoder_EncodeWhole.argtypes = [C.c_void_p, PP, C.c_uint32, C.c_void_p, C.c_uint32,
                                            C.POINTER(C.c_uint32), C.c_void_p]
    lib.SRLAEncoder_EncodeWhole.restype = C.c_int
    for name in ("SRLAB200_EncodeStreamsDevice", "SRLAB200_EncodeStreamsHost"):
        f = getattr(lib, name)
        f.argtypes = [C.c_void_p, C.POINTER(SRLAB200Stream), C.c_uint32, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        f.restype = C.c_int
    lib.SRLAB200_EncodeInterleavedHost.argtypes = [C.c_void_p, C.POINTER(SRLAB200Frames), C.c_uint32, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
    lib.SRLAB200_EncodeInterleavedHost.restype = C.c_int
    lib.SRLAB200_AllocPinned.argtypes = [C.c_size_t]
    lib.SRLAB200_AllocPinned.restype = C.c_void_p
    lib.SRLAB200_FreePinned.argtypes = [C.c_void_p]
    lib.SRLAB200_FreePinned.restype = None
    lib.SRLAB200_TestNarrow.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    lib.SRLAB200_TestNarrow.restype = C.c_uint32
    lib.SRLAB200_MaxEncodedSize.argtypes = [C.c_void_p, C.c_uint32]
    lib.SRLAB200_MaxEncodedSize.restype = C.c_uint64
    lib.SRLAB200_GetStats.argtypes = [C.c_void_p, C.POINTER(SRLAB200Stats)]
    lib.SRLAB200_GetStats.restype = C.c_int
    lib.SRLAB200_SetDevice.argtypes = [C.c_int]
    lib.SRLAB200_SetDevice.restype = C.c_int
    lib.SRLAB200_SetStream.argtypes = [C.c_void_p, C.c_void_p]
    lib.SRLAB200_SetStream.restype = C.c_int
    lib.SRLAB200_Version.argtypes = []
    lib.SRLAB200_Version.restype = C.c_char_p
    lib.SRLAB200_TestAnalyseChannel.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.POINTER(SRLAB200ChannelResult)]
    lib.SRLAB200_TestAnalyseChannel.restype = C.c_int
    _lib = lib
    return lib


class SRLAError(RuntimeError):
    def __init__(self, what: str, code: int):
        super().__init__(f"{what} -> {RESULT_NAMES[code] if 0 <= code < 8 else code}")
        self.code = code


def _planar_ptrs(pcm: np.ndarray):
    assert pcm.dtype == np.int32 and pcm.flags.c_contiguous and pcm.ndim == 2
    rows = (C.POINTER(C.c_int32) * pcm.shape[0])()
    for c in range(pcm.shape[0]):
        rows[c] = C.cast(pcm[c].ctypes.data, C.POINTER(C.c_int32))
    return rows


class Encoder:
    """One encoder handle, driven the way tools/srla_codec does (srla_codec.c:91-134):
    Create(config) -> SetEncodeParameter(parameter) -> EncodeWhole(...) -> Destroy."""

    def __init__(self, max_channels: int = 8, max_block: int = 4096, min_block: Optional[int] = None,
                 lookahead: Optional[int] = None, max_params: int = 255, device: Optional[int] = None):
        self.lib = load_library()
        min_block = max_block if min_block is None else min_block
        lookahead = (4 * max_block if min_block != max_block else max_block) if lookahead is None else lookahead
        if device is not None:
            rc = self.lib.SRLAB200_SetDevice(device)
            if rc != OK:
                raise SRLAError(f"SRLAB200_SetDevice({device})", rc)
        self.config = SRLAEncoderConfig(max_channels, min_block, max_block, lookahead, max_params)
        self.handle = self.lib.SRLAEncoder_Create(C.byref(self.config), None, 0)
        if not self.handle:
            raise RuntimeError("SRLAEncoder_Create failed (invalid config, or no usable CUDA device: there is no CPU fallback)")
        self.param: Optional[SRLAEncodeParameter] = None

    def close(self) -> None:
        if getattr(self, "handle", None):
            self.lib.SRLAEncoder_Destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- reference API -------------------------------------------------------------------------
    def set_parameter(self, num_channels: int, bits_per_sample: int = 16, sampling_rate: int = 48000,
                      min_block: Optional[int] = None, max_block: Optional[int] = None, lookahead: Optional[int] = None,
                      ltp_order: int = 0, preset: int = 4, svr_iterations: int = 0) -> int:
        max_block = self.config.max_num_samples_per_block if max_block is None else max_block
        min_block = (self.config.min_num_samples_per_block if self.config.min_num_samples_per_block != self.config.max_num_samples_per_block
                     else max_block) if min_block is None else min_block
        lookahead = (self.config.max_num_lookahead_samples if min_block != max_block else max_block) if lookahead is None else lookahead
        p = SRLAEncodeParameter(num_channels, bits_per_sample, sampling_rate, min_block, max_block, lookahead,
                                ltp_order, svr_iterations, preset)
        rc = self.lib.SRLAEncoder_SetEncodeParameter(self.handle, C.byref(p))
        if rc == OK:
            self.param = p
        return rc

    def encode_whole(self, pcm: np.ndarray, callback=None) -> bytes:
        """SRLAEncoder_EncodeWhole on planar int32 host PCM [channels, samples]."""
        pcm = np.ascontiguousarray(pcm, dtype=np.int32)
        nch, n = pcm.shape
        cap = 2 * (nch * n * 4) + 4096
        out = np.zeros(cap, dtype=np.uint8)
        size = C.c_uint32(0)
        cb = CALLBACK(callback) if callback is not None else None
        rc = self.lib.SRLAEncoder_EncodeWhole(self.handle, _planar_ptrs(pcm), n, out.ctypes.data, cap, C.byref(size),
                                              C.cast(cb, C.c_void_p) if cb is not None else None)
        if rc != OK:
            raise SRLAError("SRLAEncoder_EncodeWhole", rc)
        return out[:size.value].tobytes()

    def encode_block(self, pcm: np.ndarray, optimal_partition: bool = False) -> bytes:
        pcm = np.ascontiguousarray(pcm, dtype=np.int32)
        nch, n = pcm.shape
        cap = 2 * (nch * n * 4) + 4096
        out = np.zeros(cap, dtype=np.uint8)
        size = C.c_uint32(0)
        fn = self.lib.SRLAEncoder_EncodeOptimalPartitionedBlock if optimal_partition else self.lib.SRLAEncoder_EncodeBlock
        rc = fn(self.handle, _planar_ptrs(pcm), n, out.ctypes.data, cap, C.byref(size))
        if rc != OK:
            raise SRLAError(fn.__name__, rc)
        return out[:size.value].tobytes()

    def compute_block_size(self, pcm: np.ndarray) -> int:
        pcm = np.ascontiguousarray(pcm, dtype=np.int32)
        size = C.c_uint32(0)
        rc = self.lib.SRLAEncoder_ComputeBlockSize(self.handle, _planar_ptrs(pcm), pcm.shape[1], C.byref(size))
        if rc != OK:
            raise SRLAError("SRLAEncoder_ComputeBlockSize", rc)
        return size.value

    # -- batch extension -------------------------------------------------------------------------
    def max_encoded_size(self, num_samples: int) -> int:
        return int(self.lib.SRLAB200_MaxEncodedSize(self.handle, num_samples))

    def encode_streams_host(self, streams: Sequence[np.ndarray], out: Optional[np.ndarray] = None):
        """streams: planar arrays [channels, samples] of dtype int16 or int32 (host, ideally pinned).
        Returns (out_buffer, offsets[num_streams + 1])."""
        descs = (SRLAB200Stream * len(streams))()
        cap = 0
        for i, s in enumerate(streams):
            assert s.ndim == 2 and s.flags.c_contiguous and s.dtype in (np.int16, np.int32)
            descs[i] = SRLAB200Stream(s.ctypes.data, s.shape[1], s.shape[1], s.dtype.itemsize)
            cap += self.max_encoded_size(s.shape[1])
        if out is None:
            out = np.empty(cap, dtype=np.uint8)
        offsets = (C.c_uint64 * (len(streams) + 1))()
        rc = self.lib.SRLAB200_EncodeStreamsHost(self.handle, descs, len(streams), out.ctypes.data, out.size, offsets)
        if rc != OK:
            raise SRLAError("SRLAB200_EncodeStreamsHost", rc)
        return out, list(offsets)

    def encode_interleaved_host(self, payloads: Sequence[np.ndarray], out: Optional[np.ndarray] = None):
        """payloads: one uint8 array per stream holding the payload of a WAV `data` chunk (interleaved little-endian
        frames of the handle's channel count and bits_per_sample / 8 bytes per sample).
        Returns (out_buffer, offsets[num_streams + 1])."""
        frame = self.param.num_channels * (self.param.bits_per_sample // 8)
        items = (SRLAB200Frames * len(payloads))()
        cap = 0
        for i, b in enumerate(payloads):
            assert b.ndim == 1 and b.dtype == np.uint8 and b.flags.c_contiguous and b.size % frame == 0
            items[i] = SRLAB200Frames(b.ctypes.data, b.size // frame)
            cap += self.max_encoded_size(b.size // frame)
        if out is None:
            out = np.empty(cap, dtype=np.uint8)
        offsets = (C.c_uint64 * (len(payloads) + 1))()
        rc = self.lib.SRLAB200_EncodeInterleavedHost(self.handle, items, len(payloads), out.ctypes.data, out.size, offsets)
        if rc != OK:
            raise SRLAError("SRLAB200_EncodeInterleavedHost", rc)
        return out, list(offsets)

    def encode_streams_device(self, descs, num_streams: int, d_out_ptr: int, capacity: int):
        """descs: ctypes array of SRLAB200Stream whose pcm fields are DEVICE pointers."""
        offsets = (C.c_uint64 * (num_streams + 1))()
        rc = self.lib.SRLAB200_EncodeStreamsDevice(self.handle, descs, num_streams, d_out_ptr, capacity, offsets)
        if rc != OK:
            raise SRLAError("SRLAB200_EncodeStreamsDevice", rc)
        return list(offsets)

    def set_stream(self, cuda_stream_ptr: int) -> None:
        rc = self.lib.SRLAB200_SetStream(self.handle, cuda_stream_ptr)
        if rc != OK:
            raise SRLAError("SRLAB200_SetStream", rc)

    def stats(self) -> SRLAB200Stats:
        st = SRLAB200Stats()
        self.lib.SRLAB200_GetStats(self.handle, C.byref(st))
        return st

    def analyse_channel(self, sig: np.ndarray):
        sig = np.ascontiguousarray(sig, dtype=np.int32)
        res = np.zeros_like(sig)
        out = SRLAB200ChannelResult()
        rc = self.lib.SRLAB200_TestAnalyseChannel(self.handle, sig.ctypes.data, len(sig), res.ctypes.data, C.byref(out))
        if rc != OK:
            raise SRLAError("SRLAB200_TestAnalyseChannel", rc)
        return out, res


def encode(pcm: np.ndarray, bps: int = 16, rate: int = 48000, max_block: int = 4096, min_block: Optional[int] = None,
           lookahead: Optional[int] = None, ltp: int = 0, preset: int = 4, callback=None) -> bytes:
    """One-shot convenience with the keyword names the test helpers use for the reference/oracle."""
    pcm = np.ascontiguousarray(pcm, dtype=np.int32)
    min_block = max_block if min_block is None else min_block
    lookahead = (4 * max_block if min_block != max_block else max_block) if lookahead is None else lookahead
    with Encoder(max_channels=8, max_block=max_block, min_block=min_block, lookahead=lookahead) as enc:
        rc = enc.set_parameter(pcm.shape[0], bps, rate, min_block, max_block, lookahead, ltp, preset)
        if rc != OK:
            raise SRLAError("SRLAEncoder_SetEncodeParameter", rc)
        return enc.encode_whole(pcm, callback)
