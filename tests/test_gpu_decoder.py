"""GPU decoder against the reference decoder: every stream the encoder tests produce must decode to
the original PCM, sample for sample, and malformed input must be refused with the reference's result codes."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import (ROOT, SRLADecoderConfig, SRLAHeader, have_ref, oracle_encode, planar_ptrs, ref_decode, ref_encode, ref_lib,
                     reference_test_signals)
from srla_b200 import decoder as D
from srla_b200 import encoder as E
from srla_b200.synth import synth_stereo

pytestmark = pytest.mark.gpu


def _enc(pcm, **kw):
    return ref_encode(pcm, **kw) if have_ref() else oracle_encode(pcm, **kw)


def _ref_rc(stream: bytes, channels: int, samples: int, check: int = 1):
    """result code (and buffer) of the reference's SRLADecoder_DecodeWhole on the same input; of the oracle's decoder
    restatement (pinned against the reference on exactly these cases by tests/test_oracle.py) when oracle/_ref is absent"""
    if not have_ref():
        from helpers import oracle_decode_rc
        return oracle_decode_rc(stream, channels, samples, check)
    lib = ref_lib()
    buf = np.frombuffer(stream, dtype=np.uint8).copy()
    cfg = SRLADecoderConfig(8, 255, check)
    dec = lib.SRLADecoder_Create(C.byref(cfg), None, 0)
    out = np.zeros((max(channels, 1), max(samples, 1)), dtype=np.int32)
    try:
        return lib.SRLADecoder_DecodeWhole(dec, buf.ctypes.data, len(buf), planar_ptrs(out), channels, samples), out
    finally:
        lib.SRLADecoder_Destroy(dec)


@pytest.mark.parametrize("preset,bits,nch,ltp,block", [(0, 16, 2, 0, 4096), (1, 16, 1, 0, 1024), (2, 16, 2, 3, 2048), (3, 8, 2, 0, 4096),
                                                        (4, 16, 2, 0, 4096), (4, 24, 2, 3, 8192), (5, 16, 2, 0, 4096), (6, 16, 2, 1, 4096),
                                                        (4, 16, 3, 0, 4096), (2, 24, 8, 0, 512)])
def test_decodes_what_the_encoders_write(preset, bits, nch, ltp, block):
    n = block * 3 + block // 2 + 6
    base = synth_stereo(n, seed=100 + preset + nch)
    rows = [base[c % 2] if c < 2 else np.roll(base[c % 2], 17 * c) // (c + 1) for c in range(nch)]
    pcm = np.ascontiguousarray(np.stack(rows), dtype=np.int32)
    if bits == 24:
        pcm = np.clip(pcm.astype(np.int64) * 190 + 5, -(1 << 23), (1 << 23) - 1).astype(np.int32)
    if bits == 8:
        pcm = (pcm >> 8).astype(np.int32)
    stream = _enc(pcm, bps=bits, preset=preset, max_block=block, ltp=ltp)
    with D.Decoder() as dec:
        got = dec.decode_whole(stream)
        assert dec.kernel_ms() > 0.0
    assert got.shape == pcm.shape
    assert np.array_equal(got, pcm), np.argwhere(got != pcm)[:5]
    if have_ref():
        assert np.array_equal(ref_decode(stream), got)


def test_decodes_reference_test_signals_shifted_and_variable_blocks():
    """silence / constants / impulses / noise (SILENT and RAW blocks, all-zero channels, plain Rice), a stream whose
    samples are multiples of 16 (offset shift), variable block sizes, and a stream written by OUR encoder with SVR"""
    sigs = reference_test_signals(n=4096 * 2 + 700, bps=16, nch=2, seed=9)
    with D.Decoder() as dec:
        for name, pcm in sorted(sigs.items()):
            pcm = np.ascontiguousarray(pcm, dtype=np.int32)
            stream = _enc(pcm, preset=4, max_block=4096)
            assert np.array_equal(dec.decode_whole(stream), pcm), name
        pcm = (synth_stereo(4096 * 2 + 100, seed=3) // 16 * 16).astype(np.int32)
        stream = _enc(pcm, preset=3, max_block=4096)
        assert stream[24] == 4 and np.array_equal(dec.decode_whole(stream), pcm)
        pcm = synth_stereo(16384 * 2 + 3000, seed=4)
        stream = _enc(pcm, preset=4, max_block=4096, min_block=1024, lookahead=16384)
        assert np.array_equal(dec.decode_whole(stream), pcm)
        with E.Encoder(max_channels=2, max_block=4096) as enc:
            assert enc.set_parameter(2, 16, 48000, 4096, 4096, 4096, 0, 4, 2) == E.OK
            pcm = synth_stereo(4096 * 2, seed=5)
            assert np.array_equal(dec.decode_whole(enc.encode_whole(pcm)), pcm)


def test_many_blocks_simple_and_pipelined_paths():
    """640 blocks: the first long stream of a handle takes the simple path (one copy in, two launches, one copy out),
    later ones the pipelined path (groups of blocks staged through page-locked memory by host threads); a bad block in
    the middle stops the delivery where the reference stops it, on both paths"""
    pcm = np.concatenate([np.roll(synth_stereo(4096 * 40, seed=6), 31 * k, axis=1) for k in range(16)], axis=1)
    stream = E.encode(pcm, preset=4, max_block=4096)
    n = pcm.shape[1]
    at, starts = 30, []
    while at < len(stream):
        starts.append(at)
        at += 6 + int.from_bytes(stream[at + 2:at + 6], "big")
    bad = bytearray(stream); bad[starts[333] + 50] ^= 0x40; bad = bytes(bad)
    with D.Decoder() as dec:
        for round_ in range(3):                                   # round 0: simple path, then pipelined
            assert np.array_equal(dec.decode_whole(stream), pcm), round_
            rc, out = dec.decode_whole_rc(bad, 2, n)
            assert rc == E.DATA_CORRUPTION, (round_, rc)
            assert np.array_equal(out[:, :333 * 4096], pcm[:, :333 * 4096]), round_
            assert not out[:, 334 * 4096:].any(), round_           # nothing behind the bad block is delivered
    if have_ref():
        rc, out = _ref_rc(bad, 2, n)
        assert rc == E.DATA_CORRUPTION and np.array_equal(out[:, :333 * 4096], pcm[:, :333 * 4096]) and not out[:, 334 * 4096:].any()


def test_malformed_streams_get_the_reference_result_codes():
    pcm = synth_stereo(4096 * 3 + 500, seed=8)
    good = _enc(pcm, preset=3, max_block=4096)
    n = pcm.shape[1]
    blocks = []
    at = 30
    while at < len(good):
        size = int.from_bytes(good[at + 2:at + 6], "big")
        blocks.append(at)
        at += 6 + size
    cases = {}
    b = bytearray(good); b[blocks[1] + 40] ^= 0x10; cases["flipped bit in block 1"] = bytes(b)
    b = bytearray(good); b[blocks[2]] = 0x7F; cases["bad sync code in block 2"] = bytes(b)
    cases["truncated inside block 2"] = good[:blocks[2] + 100]
    cases["truncated inside the header"] = good[:20]
    b = bytearray(good); b[0] = ord("X"); cases["bad signature"] = bytes(b)
    b = bytearray(good); b[7] = 99; cases["bad format version"] = bytes(b)
    b = bytearray(good); b[blocks[0] + 8] = 3; cases["bad block type (checksum no longer matches)"] = bytes(b)
    with D.Decoder() as dec:
        for name, stream in cases.items():
            want, wout = _ref_rc(stream, 2, n)
            got, gout = dec.decode_whole_rc(stream, 2, n)
            assert got == want != 0, (name, got, want)
            if "block 2" in name or "block 1" in name:
                k = 4096 if "block 1" in name else 8192
                assert np.array_equal(gout[:, :k], pcm[:, :k]) and np.array_equal(wout[:, :k], pcm[:, :k]), name    # earlier blocks are delivered
        # capacity errors
        for ch, smp in ((1, n), (2, n - 1)):
            assert dec.decode_whole_rc(good, ch, smp)[0] == _ref_rc(good, ch, smp)[0] == E.INSUFFICIENT_BUFFER
    # without the checksum test a flipped residual bit goes unnoticed by both decoders; a bad block type does not
    with D.Decoder(check_checksum=False) as dec:
        stream = cases["bad block type (checksum no longer matches)"]
        assert dec.decode_whole_rc(stream, 2, n)[0] == _ref_rc(stream, 2, n, check=0)[0] == E.INVALID_FORMAT


def test_decode_block_and_header_api():
    pcm = synth_stereo(4096 + 1000, seed=12)
    stream = _enc(pcm, preset=2, max_block=4096)
    with D.Decoder() as dec:
        h = dec.decode_header(stream)
        assert (h.num_channels, h.num_samples, h.bits_per_sample, h.max_num_samples_per_block, h.preset) == (2, 5096, 16, 4096, 2)
        rc, _, _, _ = dec.decode_block(stream[30:], 2, 4096)
        assert rc == E.PARAMETER_NOT_SET
        assert dec.set_header(h) == E.OK
        rc, blk, size, n = dec.decode_block(stream[30:], 2, 4096)
        assert rc == E.OK and n == 4096 and np.array_equal(blk, pcm[:, :4096])
        rc, blk2, size2, n2 = dec.decode_block(stream[30 + size:], 2, 4096)
        assert rc == E.OK and n2 == 1000 and np.array_equal(blk2, pcm[:, 4096:]) and 30 + size + size2 == len(stream)
        assert dec.decode_block(stream[30:], 1, 4096)[0] == E.INSUFFICIENT_BUFFER
        assert dec.decode_block(stream[30:], 2, 4095)[0] == E.INSUFFICIENT_BUFFER
        assert dec.decode_block(stream[30:30 + size - 1], 2, 4096)[0] == E.INSUFFICIENT_DATA
        bad = E.SRLAHeader.from_buffer_copy(bytes(h)); bad.preset = 7
        assert dec.set_header(bad) == E.INVALID_FORMAT
    with D.Decoder(max_channels=1) as small:
        assert small.set_header(h) == E.INSUFFICIENT_BUFFER
    with D.Decoder(max_parameters=8) as small:
        assert small.set_header(h) == E.INSUFFICIENT_BUFFER


def test_reference_cli_with_encoder_and_decoder_from_libsrla_b200(tmp_path):
    """the unmodified reference CLI linked against libsrla_b200.so for BOTH directions (oracle/Makefile `dropin_full`):
    encode -> decode returns the WAV, and it decodes the reference encoder's file as well"""
    import wave
    ref_cli = os.path.join(ROOT, "oracle", "_ref", "srla_ref")
    our_cli = os.path.join(ROOT, "oracle", "_ref", "srla_b200_cli_full")
    if not (os.path.exists(ref_cli) and os.path.exists(our_cli)):
        pytest.skip("oracle/_ref CLIs not built")
    pcm = synth_stereo(48000 + 776, seed=77)
    wav = tmp_path / "in.wav"
    with wave.open(str(wav), "wb") as w:
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(48000)
        w.writeframes(pcm.T.astype("<i2").tobytes())
    subprocess.run([ref_cli, "-e", "-m", "4", str(wav), str(tmp_path / "ref.srl")], check=True, stdout=subprocess.DEVNULL)
    subprocess.run([our_cli, "-e", "-m", "4", str(wav), str(tmp_path / "our.srl")], check=True, stdout=subprocess.DEVNULL)
    assert (tmp_path / "ref.srl").read_bytes() == (tmp_path / "our.srl").read_bytes()
    for src in ("ref.srl", "our.srl"):
        back = tmp_path / f"back_{src}.wav"
        subprocess.run([our_cli, "-d", str(tmp_path / src), str(back)], check=True, stdout=subprocess.DEVNULL)
        with wave.open(str(back), "rb") as w:
            got = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").reshape(-1, 2).T
        assert np.array_equal(got, pcm)


@pytest.mark.parametrize("name", __import__("helpers").golden_names())
def test_decodes_the_committed_reference_streams(name):
    """tests/golden/*.npz hold streams written by the REFERENCE encoder here (make_golden.py) next to their PCM: the GPU
    decoder must return that PCM (no oracle/_ref needed on the GPU box)"""
    from helpers import load_golden
    pcm, _kw, srl = load_golden(name)
    with D.Decoder() as dec:
        got = dec.decode_whole(srl)
    assert got.shape == pcm.shape and np.array_equal(got, pcm)


def test_hostile_blocks_are_refused_without_touching_unwritten_records():
    """a compressed block that announces ZERO samples (valid checksum, so only a hostile or broken writer produces it) used
    to make the synthesis kernel read side records nobody had written; a walk that runs past the end of its block is
    corruption even when the next block's bytes happen to parse"""
    pcm = synth_stereo(4096 * 3 + 100, seed=3)
    good = E.encode(pcm, preset=4, max_block=4096)
    starts, at = [], 30
    while at < len(good):
        starts.append(at)
        at += 6 + int.from_bytes(good[at + 2:at + 6], "big")

    def fletcher(b):
        lo = hi = 0
        for v in b:
            lo = (lo + v) % 255
            hi = (hi + lo) % 255
        return (hi << 8) | lo

    for which in (0, 1, len(starts) - 1):
        s = bytearray(good)
        p0 = starts[which]
        size = int.from_bytes(s[p0 + 2:p0 + 6], "big")
        s[p0 + 9] = 0; s[p0 + 10] = 0
        s[p0 + 6:p0 + 8] = fletcher(s[p0 + 8:p0 + 6 + size]).to_bytes(2, "big")        # keep the checksum valid
        with D.Decoder() as dec:
            for _ in range(2):                                                         # the handle stays usable afterwards
                rc, _out = dec.decode_whole_rc(bytes(s), 2, pcm.shape[1])
                assert rc == E.INVALID_FORMAT, (which, rc)
            assert np.array_equal(dec.decode_whole(good), pcm)
        with D.Decoder(check_checksum=False) as dec:
            rc, _out = dec.decode_whole_rc(bytes(s), 2, pcm.shape[1])
            assert rc == E.INVALID_FORMAT
    # a block whose size field is cut short by 12 bytes (checksum test off): its walk needs bits beyond its end
    s = bytearray(good)
    p0 = starts[1]
    size = int.from_bytes(s[p0 + 2:p0 + 6], "big")
    cut = bytes(s[:p0 + 2]) + (size - 12).to_bytes(4, "big") + bytes(s[p0 + 6:p0 + 6 + size - 12]) + bytes(s[p0 + 6 + size:])
    with D.Decoder(check_checksum=False) as dec:
        rc, _out = dec.decode_whole_rc(cut, 2, pcm.shape[1])
        assert rc == E.DATA_CORRUPTION, rc


@pytest.mark.parametrize("lanes", ["1", "5", "32", "0"])
def test_lockstep_walk_of_unlike_blocks(lanes, monkeypatch):
    """The lanes of a warp of decode_parse_kernel walk different blocks in lockstep: a stream whose neighbouring blocks
    differ in everything the walk branches on -- music, silence (SILENT blocks), full-scale noise (RAW blocks), sparse
    clicks on silence (codes longer than 32 bits: the general reader; tiny Rice parameters), a constant (all-zero
    residual channels), a short odd tail -- decodes to the source with 1, 5 and 32 blocks per warp and with the
    launch's own choice, on the one-launch path and on the pipelined one"""
    rng = np.random.default_rng(77)
    n = 1024
    parts = []
    for k in range(96):
        kind = k % 6
        if kind == 0:
            seg = synth_stereo(n, seed=100 + k)
        elif kind == 1:
            seg = np.zeros((2, n), dtype=np.int32)
        elif kind == 2:
            seg = rng.integers(-32768, 32768, size=(2, n)).astype(np.int32)
        elif kind == 3:
            seg = np.zeros((2, n), dtype=np.int32)
            seg[:, rng.integers(0, n, size=3)] = rng.integers(-30000, 30000, size=3)
        elif kind == 4:
            seg = np.full((2, n), 1234, dtype=np.int32); seg[1] = synth_stereo(n, seed=k)[1]
        else:
            seg = (synth_stereo(n, seed=300 + k) >> 6).astype(np.int32)
        parts.append(seg)
    parts.append(synth_stereo(333, seed=9))
    pcm = np.ascontiguousarray(np.concatenate(parts, axis=1).astype(np.int32))
    stream = E.encode(pcm, preset=4, max_block=n)
    monkeypatch.setenv("SRLA_B200_DECODE_LANES", lanes)
    for pipe in ("0", "1"):
        monkeypatch.setenv("SRLA_B200_DECODE_PIPELINE", pipe)
        with D.Decoder() as dec:
            assert np.array_equal(dec.decode_whole(stream), pcm), (lanes, pipe)
    if have_ref() and lanes == "0":
        assert np.array_equal(ref_decode(stream), pcm)
