"""CPU tests that PIN the oracle (oracle/srla_oracle.c):
 * against the reference's own known-answer vectors,
 * against the committed reference-generated fixtures (tests/golden/, made by make_golden.py),
 * against the compiled unmodified reference (oracle/_ref/libsrla_ref.so) on seeded inputs, when it
   is present (build container; it is prebuilt and travels to the GPU box).
"""
import ctypes as C

import numpy as np
import pytest

from helpers import (SoParams, golden_names, have_ref, load_golden, oracle_encode, oracle_lib, ref_decode,
                     ref_encode, ref_lib, walk_blocks)
from srla_b200.synth import synth_stereo


# ---- reference KATs -------------------------------------------------------------------------
@pytest.mark.parametrize("text,expect", [(b"abcde", 0xC8F0), (b"abcdef", 0x2057), (b"abcdefgh", 0x0627)])
def test_fletcher16_kat(text, expect):
    """test/srla_internal/main.cpp:27-29"""
    buf = np.frombuffer(text, dtype=np.uint8).copy()
    assert oracle_lib().so_fletcher16(buf.ctypes.data, len(buf)) == expect


def _huff(counts):
    counts = np.asarray(counts, dtype=np.uint32)
    codes = np.zeros(len(counts), dtype=np.uint32)
    lens = np.zeros(len(counts), dtype=np.uint8)
    oracle_lib().so_huffman_codes(2, counts.ctypes.data, len(counts), codes.ctypes.data, lens.ctypes.data)
    return list(zip(codes.tolist(), lens.tolist()))


def test_huffman_kat():
    """test/static_huffman/main.cpp:36-110"""
    assert _huff([4, 3, 2, 1]) == [(0x0, 1), (0x2, 2), (0x7, 3), (0x6, 3)]
    assert _huff([5, 3, 2, 1, 1]) == [(0x0, 1), (0x2, 2), (0x6, 3), (0xE, 4), (0xF, 4)]
    for counts, total in (([8, 4, 4, 4, 2, 2], 60), ([50, 20, 10, 8, 5, 4, 2, 1], 220)):
        assert sum(c * l for c, (_, l) in zip(counts, _huff(counts))) == total


def test_header_signature_and_layout():
    """test/srla_encoder/srla_encoder_test.cpp:52-66 + field layout of srla_encoder.c:85-165"""
    prm = SoParams(2, 16, 48000, 4096, 4096, 4096, 0, 4, 3)
    out = np.zeros(30, dtype=np.uint8)
    assert oracle_lib().so_encode_header(C.byref(prm), 12345, out.ctypes.data, 30) == 0
    b = out.tobytes()
    assert b[:4] == b"1249"
    assert int.from_bytes(b[4:8], "big") == 10 and int.from_bytes(b[8:12], "big") == 18
    assert int.from_bytes(b[12:14], "big") == 2 and int.from_bytes(b[14:18], "big") == 12345
    assert int.from_bytes(b[18:22], "big") == 48000 and int.from_bytes(b[22:24], "big") == 16
    assert b[24] == 3 and int.from_bytes(b[25:29], "big") == 4096 and b[29] == 4
    assert oracle_lib().so_encode_header(C.byref(prm), 12345, out.ctypes.data, 29) == 3   # INSUFFICIENT_BUFFER


def test_fft_matches_dft():
    """test/fft/main.cpp:40- (real FFT vs naive DFT, 1e-8); here also at the sizes the path uses"""
    rng = np.random.default_rng(0)
    for n in (32, 4096):
        x = rng.standard_normal(n)
        y = x.copy()
        oracle_lib().so_real_fft(n, -1, y.ctypes.data)
        ref = np.fft.rfft(x)
        assert abs(y[0] - ref[0].real) < 1e-8 and abs(y[1] - ref[n // 2].real) < 1e-8
        got = y[2::2] + 1j * y[3::2]
        assert np.max(np.abs(got - ref[1:n // 2])) < 1e-8 * n
        oracle_lib().so_real_fft(n, 1, y.ctypes.data)
        assert np.max(np.abs(y * (2.0 / n) - x)) < 1e-10


def test_autocorr_is_scaled_circular():
    """FFT autocorrelation == (N/n) x circular autocorrelation of the Welch-windowed signal"""
    rng = np.random.default_rng(1)
    for n in (4096, 2304):
        x = rng.standard_normal(n)
        r = np.zeros(65)
        oracle_lib().so_autocorr(x.ctypes.data, n, r.ctypes.data, 64)
        N = 1 << (n - 1).bit_length()
        i = np.arange(n)
        w = 4.0 / (n - 1) ** 2 * np.minimum(i, n - 1 - i) * (n - 1 - np.minimum(i, n - 1 - i))
        xw = np.zeros(N)
        xw[:n] = x * w
        direct = np.array([np.dot(xw, np.roll(xw, -k)) for k in range(65)]) * (N / n)
        assert np.max(np.abs(r - direct)) < 1e-10 * abs(direct[0])


# ---- golden fixtures (reference-generated) ----------------------------------------------------
@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_golden(name):
    pcm, kw, srl = load_golden(name)
    assert oracle_encode(pcm, **kw) == srl


# ---- live comparison with the compiled reference ----------------------------------------------
needs_ref = pytest.mark.skipif(not have_ref(), reason="oracle/_ref/libsrla_ref.so not built")


@needs_ref
@pytest.mark.parametrize("seed,kw", [
    (1, dict(preset=4, max_block=4096)),
    (2, dict(preset=0, max_block=4096)),
    (3, dict(preset=5, max_block=2048)),
    (4, dict(preset=4, max_block=4096, ltp=3)),
    (5, dict(preset=4, max_block=4096, min_block=2048, lookahead=8192)),
    (6, dict(preset=1, max_block=1024, min_block=512, lookahead=2048, ltp=3)),
])
def test_oracle_matches_reference_16bit(seed, kw):
    pcm = synth_stereo(30000, seed=seed)
    ref = ref_encode(pcm, **kw)
    assert oracle_encode(pcm, **kw) == ref
    assert np.array_equal(ref_decode(ref), pcm)


@needs_ref
@pytest.mark.parametrize("kw", [dict(preset=4, max_block=4096), dict(preset=2, max_block=1024), dict(preset=4, max_block=4096, min_block=1024),
                                dict(preset=4, max_block=4095), dict(preset=3, max_block=1001), dict(preset=4, max_block=3000, min_block=750)])
@pytest.mark.parametrize("ltp", [0, 3])
def test_oracle_reproduces_the_stale_scratch_corners(kw, ltp):
    """odd block lengths keep the previous call's inverse transform in the Welch window's middle sample (lpc.c:260-264)
    and LTP on blocks shorter than 263 samples copies lags from beyond the transform (lpc.c:371-373): the restatement
    keeps the calculator's scratch from call to call like the reference does (a handle on zeroed memory = the CLI)"""
    for n in (1, 3, 65, 67, 100, 263, 265, 511, 4095, 4097, 8969, 9001, 9193):
        pcm = synth_stereo(n, seed=n)
        assert oracle_encode(pcm, ltp=ltp, **kw) == ref_encode(pcm, ltp=ltp, **kw), n


@needs_ref
@pytest.mark.parametrize("bits,nch", [(8, 2), (24, 2), (24, 1), (16, 5)])
def test_oracle_matches_reference_widths_and_channels(bits, nch):
    pcm = synth_stereo(20000, seed=40 + bits + nch, bits=bits, channels=nch)
    kw = dict(bps=bits, preset=4, max_block=4096, ltp=3 if bits == 24 else 0)
    assert oracle_encode(pcm, **kw) == ref_encode(pcm, **kw)


@needs_ref
def test_oracle_matches_reference_loud_24bit_8192():
    """loud 24-bit / 8192: the pre-emphasis double sums exceed 2^53 and round"""
    pcm = synth_stereo(8192 * 3, seed=9, bits=24)
    pcm = np.clip(pcm.astype(np.int64) * 3 // 2, -(1 << 23), (1 << 23) - 1).astype(np.int32)
    kw = dict(bps=24, preset=4, max_block=8192)
    assert oracle_encode(pcm, **kw) == ref_encode(pcm, **kw)


@needs_ref
def test_reference_fletcher_of_blocks():
    """every block of a reference stream carries the checksum our Fletcher-16 computes"""
    srl = ref_encode(synth_stereo(10000, seed=8), preset=4)
    for pos, size, _type, _n in walk_blocks(srl):
        body = np.frombuffer(srl[pos + 8:pos + 6 + size], dtype=np.uint8).copy()
        assert oracle_lib().so_fletcher16(body.ctypes.data, len(body)) == int.from_bytes(srl[pos + 6:pos + 8], "big")
        assert ref_lib().SRLAUtility_CalculateFletcher16CheckSum(body.ctypes.data, len(body)) == int.from_bytes(srl[pos + 6:pos + 8], "big")


# ---- decoder restatement (checks the GPU decoder when oracle/_ref is not at hand) ------------------------------------

@pytest.mark.parametrize("name", golden_names())
def test_oracle_decoder_on_the_reference_streams(name):
    """every committed stream the REFERENCE encoder wrote decodes to its PCM"""
    from helpers import oracle_decode
    pcm, _kw, srl = load_golden(name)
    assert np.array_equal(oracle_decode(srl), pcm)


def test_oracle_decoder_agrees_with_the_reference_decoder_on_malformed_streams():
    from helpers import oracle_decode_rc, ref_decode
    if not have_ref():
        pytest.skip("oracle/_ref not present")
    import ctypes as C
    from helpers import SRLADecoderConfig, planar_ptrs, ref_lib
    rng = np.random.default_rng(3)
    pcm = (rng.standard_normal((2, 4096 * 3 + 500)) * 3000).astype(np.int32)
    good = ref_encode(pcm, preset=3, max_block=4096)
    n = pcm.shape[1]
    assert np.array_equal(ref_decode(good), pcm)
    starts, at = [], 30
    while at < len(good):
        starts.append(at)
        at += 6 + int.from_bytes(good[at + 2:at + 6], "big")

    def ref_rc(stream, channels, samples, check=1):
        lib = ref_lib()
        buf = np.frombuffer(stream, dtype=np.uint8).copy()
        cfg = SRLADecoderConfig(8, 255, check)
        dec = lib.SRLADecoder_Create(C.byref(cfg), None, 0)
        out = np.zeros((max(channels, 1), max(samples, 1)), dtype=np.int32)
        try:
            return lib.SRLADecoder_DecodeWhole(dec, buf.ctypes.data, len(buf), planar_ptrs(out), channels, samples), out
        finally:
            lib.SRLADecoder_Destroy(dec)

    cases = {"good": good, "truncated block": good[:starts[2] + 100], "truncated header": good[:20]}
    b = bytearray(good); b[starts[1] + 40] ^= 0x10; cases["flipped bit"] = bytes(b)
    b = bytearray(good); b[starts[2]] = 0x7F; cases["bad sync"] = bytes(b)
    b = bytearray(good); b[0] = ord("X"); cases["bad signature"] = bytes(b)
    b = bytearray(good); b[7] = 99; cases["bad format version"] = bytes(b)
    b = bytearray(good); b[29] = 7; cases["bad preset"] = bytes(b)
    b = bytearray(good); b[starts[0] + 8] = 3; cases["bad block type"] = bytes(b)
    for name, stream in cases.items():
        for check in (1, 0):
            if name == "flipped bit" and check == 0:
                continue                                  # both decoders then run over damaged codes: output unspecified
            want, wout = ref_rc(stream, 2, n, check)
            got, gout = oracle_decode_rc(stream, 2, n, check)
            assert got == want, (name, check, got, want)
            if want == 0:
                assert np.array_equal(gout, wout)
    for ch, smp in ((1, n), (2, n - 1)):
        assert oracle_decode_rc(good, ch, smp)[0] == ref_rc(good, ch, smp)[0] == 3


@pytest.mark.parametrize("kw", [dict(preset=0, max_block=1024), dict(preset=2, max_block=2048, ltp=3), dict(preset=4, max_block=4096, bps=24, ltp=1),
                                dict(preset=5, max_block=4096), dict(preset=3, max_block=4096, min_block=1024, lookahead=8192)])
def test_oracle_round_trip(kw):
    from helpers import oracle_decode
    rng = np.random.default_rng(11)
    bits = kw.get("bps", 16)
    t = np.arange(4096 * 2 + 777)
    base = np.sin(t / 17.0) * 0.4 + np.sin(t / 3.1) * 0.1 + rng.standard_normal(t.size) * 0.01
    pcm = np.stack([base, np.roll(base, 5) * 0.9]) * (1 << (bits - 1)) * 0.9
    pcm = pcm.astype(np.int32)
    assert np.array_equal(oracle_decode(oracle_encode(pcm, **kw)), pcm)


@pytest.mark.parametrize("preset,bits,ltp,svr", [(4, 16, 0, 1), (4, 16, 0, 3), (3, 16, 0, 8), (2, 24, 3, 2), (5, 16, 0, 2), (1, 8, 0, 4)])
def test_oracle_svr_matches_reference(preset, bits, ltp, svr):
    """the SVR coefficient refinement of the restatement (lpc.c:1036-1136) against the compiled reference"""
    if not have_ref():
        pytest.skip("oracle/_ref not present")
    pcm = synth_stereo(4096 * 2 + 1500, seed=40 + preset)
    if bits == 24:
        pcm = np.clip(pcm.astype(np.int64) * 180 + 3, -(1 << 23), (1 << 23) - 1).astype(np.int32)
    if bits == 8:
        pcm = (pcm >> 8).astype(np.int32)
    kw = dict(bps=bits, preset=preset, max_block=4096, ltp=ltp, svr=svr)
    got, want = oracle_encode(pcm, **kw), ref_encode(pcm, **kw)
    assert got == want


def test_oracle_svr_on_the_reference_test_signals():
    """silence (singular covariance), constants, impulses, noise, Nyquist -- the cases the GPU test uses"""
    if not have_ref():
        pytest.skip("oracle/_ref not present")
    from helpers import reference_test_signals
    for name, pcm in sorted(reference_test_signals(n=4096 + 700, bps=16, nch=2, seed=5).items()):
        pcm = np.ascontiguousarray(pcm, dtype=np.int32)
        kw = dict(preset=3, max_block=4096, svr=3)
        assert oracle_encode(pcm, **kw) == ref_encode(pcm, **kw), name
