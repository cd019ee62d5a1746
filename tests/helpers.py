"""ctypes access to the two checkers used by the tests:

* oracle/_ref/libsrla_ref.so -- the UNMODIFIED reference compiled from /root/reference by oracle/Makefile
  (exports the reference's own SRLAEncoder_* / SRLADecoder_* API),
* oracle/liboracle.so        -- our CPU restatement (oracle/srla_oracle.c).

Test infrastructure only.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libsrla_ref.so")
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")

PRESET_MAX_ORDER = [0, 8, 16, 32, 64, 128, 255]


# ----------------------------------------------------------------------------- reference ABI
class SRLAHeader(C.Structure):
    _fields_ = [("format_version", C.c_uint32), ("codec_version", C.c_uint32),
                ("num_channels", C.c_uint16), ("num_samples", C.c_uint32),
                ("sampling_rate", C.c_uint32), ("bits_per_sample", C.c_uint16),
                ("offset_lshift", C.c_uint8), ("max_num_samples_per_block", C.c_uint32),
                ("preset", C.c_uint8)]


class SRLAEncodeParameter(C.Structure):
    _fields_ = [("num_channels", C.c_uint16), ("bits_per_sample", C.c_uint16),
                ("sampling_rate", C.c_uint32), ("min_num_samples_per_block", C.c_uint32),
                ("max_num_samples_per_block", C.c_uint32), ("num_lookahead_samples", C.c_uint32),
                ("ltp_order", C.c_uint32), ("num_svr_filter_learning_iteration", C.c_uint32),
                ("preset", C.c_uint8)]


class SRLAEncoderConfig(C.Structure):
    _fields_ = [("max_num_channels", C.c_uint32), ("min_num_samples_per_block", C.c_uint32),
                ("max_num_samples_per_block", C.c_uint32), ("max_num_lookahead_samples", C.c_uint32),
                ("max_num_parameters", C.c_uint32)]


class SRLADecoderConfig(C.Structure):
    _fields_ = [("max_num_channels", C.c_uint32), ("max_num_parameters", C.c_uint32),
                ("check_checksum", C.c_uint8)]


def bind_encoder_api(lib: C.CDLL) -> C.CDLL:
    """Declare the SRLAEncoder_* prototypes (identical for the reference and for libsrla_b200)."""
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
    return lib


_ref = None


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def ref_lib() -> C.CDLL:
    global _ref
    if _ref is None:
        lib = bind_encoder_api(C.CDLL(REF_SO))
        PP = C.POINTER(C.POINTER(C.c_int32))
        lib.SRLADecoder_DecodeHeader.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(SRLAHeader)]
        lib.SRLADecoder_DecodeHeader.restype = C.c_int
        lib.SRLADecoder_Create.argtypes = [C.POINTER(SRLADecoderConfig), C.c_void_p, C.c_int32]
        lib.SRLADecoder_Create.restype = C.c_void_p
        lib.SRLADecoder_Destroy.argtypes = [C.c_void_p]
        lib.SRLADecoder_DecodeWhole.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, PP, C.c_uint32, C.c_uint32]
        lib.SRLADecoder_DecodeWhole.restype = C.c_int
        lib.SRLAUtility_CalculateFletcher16CheckSum.argtypes = [C.c_void_p, C.c_size_t]
        lib.SRLAUtility_CalculateFletcher16CheckSum.restype = C.c_uint16
        _ref = lib
    return _ref


def planar_ptrs(pcm: np.ndarray):
    """pcm: C-contiguous int32 [ch, n] -> (int32** array, keepalive)."""
    assert pcm.dtype == np.int32 and pcm.flags.c_contiguous and pcm.ndim == 2
    rows = (C.POINTER(C.c_int32) * pcm.shape[0])()
    for c in range(pcm.shape[0]):
        rows[c] = C.cast(pcm[c].ctypes.data, C.POINTER(C.c_int32))
    return rows


def make_param(nch, bps, rate, min_block, max_block, lookahead, ltp, preset, svr=0):
    p = SRLAEncodeParameter()
    p.num_channels, p.bits_per_sample, p.sampling_rate = nch, bps, rate
    p.min_num_samples_per_block, p.max_num_samples_per_block = min_block, max_block
    p.num_lookahead_samples, p.ltp_order = lookahead, ltp
    p.num_svr_filter_learning_iteration, p.preset = svr, preset
    return p


def api_encode_whole(lib: C.CDLL, pcm: np.ndarray, bps=16, rate=48000, max_block=4096, min_block=None,
                     lookahead=None, ltp=0, preset=4, max_params=255, max_channels=8, svr=0) -> bytes:
    """Drive Create / SetEncodeParameter / EncodeWhole / Destroy the way tools/srla_codec does
    (srla_codec.c:91-134) on any library exporting the SRLAEncoder_* API."""
    pcm = np.ascontiguousarray(pcm, dtype=np.int32)
    nch, n = pcm.shape
    min_block = max_block if min_block is None else min_block
    lookahead = (4 * max_block if min_block != max_block else max_block) if lookahead is None else lookahead
    cfg = SRLAEncoderConfig(max_channels, min_block, max_block, lookahead, max_params)
    # Caller-provided, ZEROED work area: the reference reads two autocorrelation lags it never
    # writes during the LTP pitch search (lpc.c:1497-1513); a fresh CLI process sees zero pages
    # there, a recycled heap does not.  Zeroing makes the reference deterministic and equal to
    # what the `srla` CLI produces.
    work_size = lib.SRLAEncoder_CalculateWorkSize(C.byref(cfg))
    assert work_size > 0, "SRLAEncoder_CalculateWorkSize failed"
    work = np.zeros(work_size + 64, dtype=np.uint8)
    enc = lib.SRLAEncoder_Create(C.byref(cfg), work.ctypes.data, work_size)
    assert enc, "SRLAEncoder_Create failed"
    try:
        prm = make_param(nch, bps, rate, min_block, max_block, lookahead, ltp, preset, svr)
        rc = lib.SRLAEncoder_SetEncodeParameter(enc, C.byref(prm))
        assert rc == 0, f"SetEncodeParameter -> {rc}"
        cap = 2 * (nch * n * 4) + 4096
        out = np.zeros(cap, dtype=np.uint8)
        size = C.c_uint32(0)
        rows = planar_ptrs(pcm)
        rc = lib.SRLAEncoder_EncodeWhole(enc, rows, n, out.ctypes.data, cap, C.byref(size), None)
        assert rc == 0, f"EncodeWhole -> {rc}"
        return out[:size.value].tobytes()
    finally:
        lib.SRLAEncoder_Destroy(enc)


def ref_encode(pcm, **kw) -> bytes:
    return api_encode_whole(ref_lib(), pcm, **kw)


def ref_decode(stream: bytes) -> np.ndarray:
    """Decode a whole .srl stream with the reference decoder -> int32 [ch, n]."""
    lib = ref_lib()
    buf = np.frombuffer(stream, dtype=np.uint8).copy()
    hdr = SRLAHeader()
    rc = lib.SRLADecoder_DecodeHeader(buf.ctypes.data, len(buf), C.byref(hdr))
    assert rc == 0, f"DecodeHeader -> {rc}"
    cfg = SRLADecoderConfig(8, 255, 1)
    dec = lib.SRLADecoder_Create(C.byref(cfg), None, 0)
    assert dec
    try:
        out = np.zeros((hdr.num_channels, hdr.num_samples), dtype=np.int32)
        rows = planar_ptrs(out)
        rc = lib.SRLADecoder_DecodeWhole(dec, buf.ctypes.data, len(buf), rows, hdr.num_channels, hdr.num_samples)
        assert rc == 0, f"DecodeWhole -> {rc}"
        return out
    finally:
        lib.SRLADecoder_Destroy(dec)


# ----------------------------------------------------------------------------- oracle (our restatement)
class SoParams(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("num_channels", "bits_per_sample", "sampling_rate", "min_block",
                                          "max_block", "lookahead", "ltp_order", "preset", "offset_lshift")]


class SoChannel(C.Structure):
    _fields_ = [("pre_coef", C.c_int32), ("pre_prev", C.c_int32), ("order", C.c_uint32), ("rshift", C.c_uint32),
                ("use_sum", C.c_uint32), ("coef", C.c_int32 * 255), ("ltp_period", C.c_uint32),
                ("ltp_coef", C.c_int32 * 3), ("code_type", C.c_uint32), ("porder", C.c_uint32),
                ("residual_bits", C.c_uint32), ("total_bits", C.c_uint32),
                ("autocorr", C.c_double * 256), ("error_vars", C.c_double * 256), ("lpc_double", C.c_double * 255)]


_oracle = None


def build_oracle() -> None:
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])


def oracle_lib() -> C.CDLL:
    global _oracle
    if _oracle is None:
        src = os.path.join(ROOT, "oracle", "srla_oracle.c")
        if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
            build_oracle()
        lib = C.CDLL(ORACLE_SO)
        lib.so_fletcher16.argtypes = [C.c_void_p, C.c_size_t]
        lib.so_fletcher16.restype = C.c_uint16
        lib.so_huffman_codes.argtypes = [C.c_int, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        lib.so_huffman_codes.restype = None
        lib.so_encode_header.argtypes = [C.POINTER(SoParams), C.c_uint32, C.c_void_p, C.c_uint32]
        lib.so_real_fft.argtypes = [C.c_int, C.c_int, C.c_void_p]
        lib.so_real_fft.restype = None
        lib.so_autocorr.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
        lib.so_autocorr.restype = None
        lib.so_rice_search.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        lib.so_rice_search.restype = C.c_uint32
        lib.so_analyse_channel.argtypes = [C.POINTER(SoParams), C.c_void_p, C.c_uint32, C.c_void_p, C.POINTER(SoChannel)]
        lib.so_reset_state.argtypes = []
        lib.so_reset_state.restype = None
        lib.so_encode_whole_flat.argtypes = [C.POINTER(SoParams), C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32,
                                             C.POINTER(C.c_uint32)]
        lib.so_set_svr_iterations.argtypes = [C.c_uint32]
        lib.so_set_svr_iterations.restype = None
        lib.so_decode_header.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(SoParams), C.POINTER(C.c_uint32)]
        lib.so_decode_whole_flat.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_uint32, C.c_uint32]
        _oracle = lib
    return _oracle


def so_params(nch, bps=16, rate=48000, max_block=4096, min_block=None, lookahead=None, ltp=0, preset=4, lshift=0):
    min_block = max_block if min_block is None else min_block
    lookahead = (4 * max_block if min_block != max_block else max_block) if lookahead is None else lookahead
    return SoParams(nch, bps, rate, min_block, max_block, lookahead, ltp, preset, lshift)


def oracle_encode(pcm, bps=16, rate=48000, max_block=4096, min_block=None, lookahead=None, ltp=0, preset=4, svr=0) -> bytes:
    lib = oracle_lib()
    lib.so_set_svr_iterations(svr)
    pcm = np.ascontiguousarray(pcm, dtype=np.int32)
    nch, n = pcm.shape
    prm = so_params(nch, bps, rate, max_block, min_block, lookahead, ltp, preset)
    cap = 2 * (nch * n * 4) + 4096
    out = np.zeros(cap, dtype=np.uint8)
    size = C.c_uint32(0)
    try:
        rc = lib.so_encode_whole_flat(C.byref(prm), pcm.ctypes.data, n, out.ctypes.data, cap, C.byref(size))
    finally:
        lib.so_set_svr_iterations(0)
    assert rc == 0, f"so_encode_whole -> {rc}"
    return out[:size.value].tobytes()


def oracle_decode_rc(stream: bytes, channels: int, samples: int, check: int = 1):
    """the restatement's DecodeWhole: (result code, int32 [channels, samples])"""
    lib = oracle_lib()
    buf = np.frombuffer(stream, dtype=np.uint8).copy() if len(stream) else np.zeros(1, dtype=np.uint8)
    out = np.zeros((max(channels, 1), max(samples, 1)), dtype=np.int32)
    rc = lib.so_decode_whole_flat(buf.ctypes.data, len(stream), check, out.ctypes.data, channels, max(samples, 1) if samples else 0)
    return rc, out


def oracle_decode(stream: bytes) -> np.ndarray:
    lib = oracle_lib()
    buf = np.frombuffer(stream, dtype=np.uint8).copy()
    prm, n = SoParams(), C.c_uint32(0)
    rc = lib.so_decode_header(buf.ctypes.data, len(buf), C.byref(prm), C.byref(n))
    assert rc == 0, f"so_decode_header -> {rc}"
    rc, out = oracle_decode_rc(stream, prm.num_channels, n.value)
    assert rc == 0, f"so_decode_whole -> {rc}"
    return out


def oracle_analyse(x: np.ndarray, bps=16, preset=4, ltp=0):
    """Analyse one candidate channel; returns (SoChannel, pre-emphasised signal, residual)."""
    lib = oracle_lib()
    sig = np.ascontiguousarray(x, dtype=np.int32).copy()
    res = np.zeros_like(sig)
    ch = SoChannel()
    prm = so_params(1, bps, preset=preset, ltp=ltp)
    lib.so_reset_state()                                 # a freshly created handle (odd lengths read the calculator's scratch)
    rc = lib.so_analyse_channel(C.byref(prm), sig.ctypes.data, len(sig), res.ctypes.data, C.byref(ch))
    assert rc == 0
    return ch, sig, res


# ----------------------------------------------------------------------------- signals
def walk_blocks(stream: bytes):
    """Yield (offset, size_field, type, nsmpl) for each block of a .srl stream (block layout:
    0xFFFF, u32 size, u16 checksum, u8 type, u16 nsmpl, payload)."""
    pos = 30
    while pos < len(stream):
        assert stream[pos] == 0xFF and stream[pos + 1] == 0xFF, f"lost sync at {pos}"
        size = int.from_bytes(stream[pos + 2:pos + 6], "big")
        yield pos, size, stream[pos + 8], int.from_bytes(stream[pos + 9:pos + 11], "big")
        pos += 6 + size


def reference_test_signals(n=8500, bps=16, nch=2, seed=0):
    """The generator families of the reference's round-trip matrix
    (test/srla_encode_decode/main.cpp:51-208), re-expressed with numpy."""
    rng = np.random.default_rng(seed)
    full = (1 << (bps - 1))
    t = np.arange(n)
    out = {}
    out["silence"] = np.zeros((nch, n))
    out["sine440"] = np.tile(np.sin(2 * np.pi * 440.0 * t / 44100.0), (nch, 1)) * (full - 1)
    flip = np.tile(np.sin(2 * np.pi * 440.0 * t / 44100.0), (nch, 1)) * (full - 1)
    flip[1::2] *= -1
    out["sine_flipped"] = flip
    out["white"] = rng.uniform(-1, 1, (nch, n)) * (full - 1)
    out["chirp"] = np.tile(np.sin(2 * np.pi * (t / n) * t / 8.0), (nch, 1)) * (full - 1)
    out["pos_const"] = np.full((nch, n), full - 1.0)
    out["neg_const"] = np.full((nch, n), -float(full))
    out["nyquist"] = np.tile(np.where(t % 2 == 0, 1.0, -1.0), (nch, 1)) * (full - 1)
    out["gauss"] = np.clip(rng.standard_normal((nch, n)) * 0.2, -1, 1) * (full - 1)
    imp = np.zeros((nch, n)); imp[:, ::100] = 1
    out["mini_impulse"] = imp
    return {k: np.round(v).astype(np.int32) for k, v in out.items()}


# ----------------------------------------------------------------------------- golden fixtures
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz"))


def load_golden(name):
    """-> (pcm int32 [ch, n], kwargs for *_encode, reference .srl bytes)"""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    kw = {str(k): int(v) for k, v in zip(z["param_names"], z["param_values"])}
    return z["pcm"].astype(np.int32), kw, z["srl"].tobytes()
