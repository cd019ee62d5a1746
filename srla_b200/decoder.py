"""ctypes mirror of the decoder half of include/srla_b200.h (the reference's include/srla_decoder.h:8-56)."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from .encoder import OK, SRLAError, SRLAHeader, load_library


class SRLADecoderConfig(C.Structure):               # include/srla_decoder.h:8-12
    _fields_ = [("max_num_channels", C.c_uint32), ("max_num_parameters", C.c_uint32), ("check_checksum", C.c_uint8)]


def _bind(lib: C.CDLL) -> C.CDLL:
    if getattr(lib, "_srla_decoder_bound", False):
        return lib
    PP = C.POINTER(C.POINTER(C.c_int32))
    lib.SRLADecoder_DecodeHeader.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(SRLAHeader)]
    lib.SRLADecoder_DecodeHeader.restype = C.c_int
    lib.SRLADecoder_CalculateWorkSize.argtypes = [C.POINTER(SRLADecoderConfig)]
    lib.SRLADecoder_CalculateWorkSize.restype = C.c_int32
    lib.SRLADecoder_Create.argtypes = [C.POINTER(SRLADecoderConfig), C.c_void_p, C.c_int32]
    lib.SRLADecoder_Create.restype = C.c_void_p
    lib.SRLADecoder_Destroy.argtypes = [C.c_void_p]
    lib.SRLADecoder_Destroy.restype = None
    lib.SRLADecoder_SetHeader.argtypes = [C.c_void_p, C.POINTER(SRLAHeader)]
    lib.SRLADecoder_SetHeader.restype = C.c_int
    lib.SRLADecoder_DecodeBlock.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, PP, C.c_uint32, C.c_uint32,
                                            C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    lib.SRLADecoder_DecodeBlock.restype = C.c_int
    lib.SRLADecoder_DecodeWhole.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, PP, C.c_uint32, C.c_uint32]
    lib.SRLADecoder_DecodeWhole.restype = C.c_int
    lib.SRLAB200_DecoderKernelMs.argtypes = [C.c_void_p]
    lib.SRLAB200_DecoderKernelMs.restype = C.c_float
    lib._srla_decoder_bound = True
    return lib


def _rows(buf: np.ndarray):
    rows = (C.POINTER(C.c_int32) * buf.shape[0])()
    for c in range(buf.shape[0]):
        rows[c] = C.cast(buf[c].ctypes.data, C.POINTER(C.c_int32))
    return rows


class Decoder:
    """SRLADecoder_Create / DecodeWhole / DecodeBlock / Destroy.  Raises when no CUDA device is usable."""

    def __init__(self, max_channels: int = 8, max_parameters: int = 255, check_checksum: bool = True):
        self.lib = _bind(load_library())
        self.config = SRLADecoderConfig(max_channels, max_parameters, 1 if check_checksum else 0)
        self.handle: Optional[int] = self.lib.SRLADecoder_Create(C.byref(self.config), None, 0)
        if not self.handle:
            raise RuntimeError("SRLADecoder_Create failed (no CUDA device? the decode path has no CPU fallback)")

    def close(self) -> None:
        if self.handle:
            self.lib.SRLADecoder_Destroy(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def decode_header(self, stream: bytes) -> SRLAHeader:
        h = SRLAHeader()
        rc = self.lib.SRLADecoder_DecodeHeader(stream, len(stream), C.byref(h))
        if rc != OK:
            raise SRLAError("SRLADecoder_DecodeHeader", rc)
        return h

    def decode_whole(self, stream: bytes, out: Optional[np.ndarray] = None) -> np.ndarray:
        """-> int32 [channels, samples] (into `out` when given: a C-contiguous int32 array of that shape)"""
        h = self.decode_header(stream)
        if out is None:
            out = np.zeros((h.num_channels, h.num_samples), dtype=np.int32)
        assert out.dtype == np.int32 and out.flags.c_contiguous and out.shape == (h.num_channels, h.num_samples)
        rc = self.lib.SRLADecoder_DecodeWhole(self.handle, stream, len(stream), _rows(out), h.num_channels, h.num_samples)
        if rc != OK:
            raise SRLAError("SRLADecoder_DecodeWhole", rc)
        return out

    def decode_whole_rc(self, stream: bytes, channels: int, samples: int):
        """result code and buffer, for the error-path tests"""
        out = np.zeros((max(channels, 1), max(samples, 1)), dtype=np.int32)
        rc = self.lib.SRLADecoder_DecodeWhole(self.handle, stream, len(stream), _rows(out), channels, samples)
        return rc, out

    def set_header(self, header: SRLAHeader) -> int:
        return self.lib.SRLADecoder_SetHeader(self.handle, C.byref(header))

    def decode_block(self, data: bytes, channels: int, capacity: int):
        out = np.zeros((channels, max(capacity, 1)), dtype=np.int32)
        size, n = C.c_uint32(0), C.c_uint32(0)
        rc = self.lib.SRLADecoder_DecodeBlock(self.handle, data, len(data), _rows(out), channels, capacity, C.byref(size), C.byref(n))
        return rc, out[:, :n.value], size.value, n.value

    def kernel_ms(self) -> float:
        return float(self.lib.SRLAB200_DecoderKernelMs(self.handle))


def decode(stream: bytes, check_checksum: bool = True) -> np.ndarray:
    with Decoder(check_checksum=check_checksum) as dec:
        return dec.decode_whole(stream)
