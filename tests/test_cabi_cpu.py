"""CPU-side checks of the drop-in boundary (no GPU compute):
 * libsrla_b200.so loads and exports every symbol include/srla_b200.h declares,
 * host-only entry points behave like the reference (header bytes, config validation),
 * without a CUDA device the product path FAILS LOUDLY (no CPU fallback),
 * the sharding helper used by bench.py / multi-GPU runs partitions streams without overlap
   (world_size-2 gloo test).
"""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from helpers import ROOT, SoParams, have_ref, oracle_lib, ref_lib
from srla_b200 import encoder as E

LIB = os.path.join(ROOT, "srla_b200", "libsrla_b200.so")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        import __graft_entry__ as g
        g.build()
    return E.load_library()


def test_every_declared_symbol_is_exported(lib):
    header = open(os.path.join(ROOT, "include", "srla_b200.h")).read()
    declared = set(re.findall(r"\b(SRLAEncoder_[A-Za-z]+|SRLADecoder_[A-Za-z]+|SRLAB200_[A-Za-z]+)\s*\(", header))
    declared.discard("SRLAEncoder_EncodeBlockCallback")
    assert declared == set(E.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.SRLAB200_Version()


def test_library_embeds_sm100a_code_only():
    out = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out)


def test_encode_header_matches_oracle_and_reference(lib):
    """srla_encoder.c:85-165; signature bytes test/srla_encoder/srla_encoder_test.cpp:63-66"""
    h = E.SRLAHeader(10, 18, 2, 12345, 48000, 16, 3, 4096, 4)
    out = np.zeros(30, dtype=np.uint8)
    assert lib.SRLAEncoder_EncodeHeader(C.byref(h), out.ctypes.data, 30) == E.OK
    assert out[:4].tobytes() == b"1249"
    prm = SoParams(2, 16, 48000, 4096, 4096, 4096, 0, 4, 3)
    want = np.zeros(30, dtype=np.uint8)
    assert oracle_lib().so_encode_header(C.byref(prm), 12345, want.ctypes.data, 30) == 0
    assert out.tobytes() == want.tobytes()
    if have_ref():
        from helpers import SRLAHeader as RefHeader
        rh = RefHeader(10, 18, 2, 12345, 48000, 16, 3, 4096, 4)
        ref = np.zeros(30, dtype=np.uint8)
        assert ref_lib().SRLAEncoder_EncodeHeader(C.byref(rh), ref.ctypes.data, 30) == 0
        assert out.tobytes() == ref.tobytes()
    # error behaviour (srla_encoder_test.cpp:69-113)
    assert lib.SRLAEncoder_EncodeHeader(None, out.ctypes.data, 30) == E.INVALID_ARGUMENT
    assert lib.SRLAEncoder_EncodeHeader(C.byref(h), None, 30) == E.INVALID_ARGUMENT
    assert lib.SRLAEncoder_EncodeHeader(C.byref(h), out.ctypes.data, 29) == E.INSUFFICIENT_BUFFER
    for field, bad in (("num_channels", 0), ("num_samples", 0), ("sampling_rate", 0), ("bits_per_sample", 0),
                       ("offset_lshift", 32), ("max_num_samples_per_block", 0), ("preset", 7)):
        hb = E.SRLAHeader(10, 18, 2, 12345, 48000, 16, 3, 4096, 4)
        setattr(hb, field, bad)
        assert lib.SRLAEncoder_EncodeHeader(C.byref(hb), out.ctypes.data, 30) == E.INVALID_FORMAT, field


def test_work_size_validation_matches_reference(lib):
    """srla_encoder.c:468-496 (test/srla_encoder/srla_encoder_test.cpp:118-170)"""
    good = E.SRLAEncoderConfig(8, 1024, 4096, 16384, 255)
    assert lib.SRLAEncoder_CalculateWorkSize(C.byref(good)) > 0
    assert lib.SRLAEncoder_CalculateWorkSize(None) == -1
    cases = [(0, 1024, 4096, 16384, 255), (8, 0, 4096, 16384, 255), (8, 1024, 0, 16384, 255), (8, 1024, 4096, 0, 255),
             (8, 1024, 4096, 16384, 5000), (8, 8192, 4096, 16384, 255), (8, 1024, 4096, 2048, 255)]
    for c in cases:
        cfg = E.SRLAEncoderConfig(*c)
        assert lib.SRLAEncoder_CalculateWorkSize(C.byref(cfg)) == -1, c
        if have_ref():
            from helpers import SRLAEncoderConfig as RefCfg
            assert ref_lib().SRLAEncoder_CalculateWorkSize(C.byref(RefCfg(*c))) == -1, c


def test_no_cpu_fallback_without_a_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    cfg = E.SRLAEncoderConfig(8, 4096, 4096, 4096, 255)
    assert not lib.SRLAEncoder_Create(C.byref(cfg), None, 0)
    with pytest.raises(RuntimeError):
        E.Encoder()
    with pytest.raises(Exception):
        E.encode(np.zeros((2, 4096), dtype=np.int32))
    # NULL handle behaves like the reference: INVALID_ARGUMENT, Destroy(NULL) is a no-op
    size = C.c_uint32(0)
    assert lib.SRLAEncoder_EncodeWhole(None, None, 0, None, 0, C.byref(size), None) == E.INVALID_ARGUMENT
    assert lib.SRLAEncoder_SetEncodeParameter(None, None) == E.INVALID_ARGUMENT
    lib.SRLAEncoder_Destroy(None)


def test_decoder_host_entry_points_match_the_reference(lib):
    """SRLADecoder_DecodeHeader / CalculateWorkSize / Create argument handling (host only, no GPU needed)"""
    from srla_b200 import decoder as D
    D._bind(lib)
    hdr = E.SRLAHeader(10, 18, 2, 123456, 44100, 24, 3, 8192, 5)
    raw = (C.c_uint8 * 30)()
    assert lib.SRLAEncoder_EncodeHeader(C.byref(hdr), raw, 30) == E.OK
    data = bytes(raw)
    back = E.SRLAHeader()
    assert lib.SRLADecoder_DecodeHeader(data, 30, C.byref(back)) == E.OK
    assert bytes(back)[:0] == b"" and [getattr(back, f) for f, _ in E.SRLAHeader._fields_] == [getattr(hdr, f) for f, _ in E.SRLAHeader._fields_]
    assert lib.SRLADecoder_DecodeHeader(data, 29, C.byref(back)) == E.INSUFFICIENT_DATA
    assert lib.SRLADecoder_DecodeHeader(b"X" + data[1:], 30, C.byref(back)) == E.INVALID_FORMAT
    assert lib.SRLADecoder_DecodeHeader(None, 30, C.byref(back)) == E.INVALID_ARGUMENT
    assert lib.SRLADecoder_CalculateWorkSize(None) == -1
    assert lib.SRLADecoder_CalculateWorkSize(C.byref(D.SRLADecoderConfig(0, 255, 1))) == -1
    assert lib.SRLADecoder_CalculateWorkSize(C.byref(D.SRLADecoderConfig(8, 255, 1))) > 0
    if have_ref():
        from helpers import SRLADecoderConfig as RefCfg, SRLAHeader as RefHdr
        ref = ref_lib()
        rb = RefHdr()
        buf = np.frombuffer(data, dtype=np.uint8).copy()
        assert ref.SRLADecoder_DecodeHeader(buf.ctypes.data, 30, C.byref(rb)) == E.OK
        assert [getattr(rb, f) for f, _ in RefHdr._fields_] == [getattr(back, f) for f, _ in E.SRLAHeader._fields_]
        assert ref.SRLADecoder_DecodeHeader(buf.ctypes.data, 29, C.byref(rb)) == E.INSUFFICIENT_DATA
        assert ref.SRLADecoder_CalculateWorkSize(C.byref(RefCfg(0, 255, 1))) == -1
    assert lib.SRLADecoder_SetHeader(None, C.byref(hdr)) == E.INVALID_ARGUMENT
    size, n = C.c_uint32(0), C.c_uint32(0)
    assert lib.SRLADecoder_DecodeBlock(None, data, 30, None, 2, 0, C.byref(size), C.byref(n)) == E.INVALID_ARGUMENT
    assert lib.SRLADecoder_DecodeWhole(None, data, 30, None, 2, 0) == E.INVALID_ARGUMENT
    lib.SRLADecoder_Destroy(None)
    import torch
    if not torch.cuda.is_available():
        assert not lib.SRLADecoder_Create(C.byref(D.SRLADecoderConfig(8, 255, 1)), None, 0)       # no CPU fallback
        with pytest.raises(RuntimeError):
            D.Decoder()


def test_feeder_narrowing_on_the_host(lib):
    """the int32 -> int16 narrowing of the EncodeWhole feeder (AVX2 + streaming stores on this CPU, scalar otherwise):
    exact for every alignment and length, and every out-of-range sample is reported"""
    rng = np.random.default_rng(5)
    for n in (0, 1, 7, 15, 16, 17, 31, 33, 1000, 65536, 65537):
        src = rng.integers(-32768, 32768, size=n + 1, dtype=np.int32)
        for off in (0, 1, 3, 8, 15):
            dst = np.full(n + off + 40, 77, dtype=np.int16)
            assert lib.SRLAB200_TestNarrow(src.ctypes.data, dst[off:].ctypes.data, n) == 0
            assert np.array_equal(dst[off:off + n], src[:n].astype(np.int16)) and dst[off + n] == 77 and (off == 0 or dst[off - 1] == 77)
    src = rng.integers(-32768, 32768, size=4099, dtype=np.int32)
    dst = np.zeros(4099, dtype=np.int16)
    for pos in (0, 5, 16, 2048, 4095, 4098):
        for bad in (32768, -32769, 1 << 20, -(1 << 31)):
            keep = src[pos]; src[pos] = bad
            assert lib.SRLAB200_TestNarrow(src.ctypes.data, dst.ctypes.data, 4099) != 0, (pos, bad)
            src[pos] = keep
    assert lib.SRLAB200_TestNarrow(src.ctypes.data, dst.ctypes.data, 4099) == 0


def test_batch_cli_options_follow_the_reference_cli(tmp_path):
    """srla_b200_batch takes the reference CLI's encode options with its range checks (srla_codec.c:311-403); on a
    box without a GPU it fails loudly instead of falling back"""
    import struct
    tool = os.path.join(ROOT, "srla_b200", "srla_b200_batch")
    if not os.path.exists(tool):
        import __graft_entry__ as g
        g.build()
    run = lambda *a: subprocess.run([tool, *a], capture_output=True, text=True)
    assert run("--help").returncode == 0
    assert "sm_100a" in run("-v").stdout
    r = run("-o", str(tmp_path)); assert r.returncode == 1 and "input file must be specified" in r.stderr
    r = run("x.wav"); assert r.returncode == 1 and "output directory must be specified" in r.stderr
    for args, msg in ((("-m", "7"), "encode preset number is out of range"), (("-m", "4x"), "irregular character found in 4x at x"),
                      (("-B", "65536"), "number of block samples is out of range"), (("-L", "0"), "lookahead factor is out of range"),
                      (("-V", "13"), "number of variable block divisions is too large"), (("-P", "2"), "must be odd"),
                      (("-P", "5"), "long term prediction order is too large")):
        r = run(*args, "-o", str(tmp_path), "x.wav")
        assert r.returncode == 1 and msg in r.stderr, (args, r.stderr)
    # header rules of the reference reader (wav.c:136-281): fmt first, size 16 or 40, linear PCM
    bad = tmp_path / "float.wav"
    fmt = struct.pack("<HHIIHH", 3, 1, 48000, 192000, 4, 32)
    bad.write_bytes(b"RIFF" + struct.pack("<I", 36) + b"WAVE" + b"fmt " + struct.pack("<I", 16) + fmt + b"data" + struct.pack("<I", 0))
    # more header shapes the reference reader refuses (wav.c:150-166, 242-245): each is reported by name
    def riff(chunks):
        return b"RIFF" + struct.pack("<I", 4 + len(chunks)) + b"WAVE" + chunks
    pcm16 = struct.pack("<HHIIHH", 1, 2, 44100, 176400, 4, 16)
    shapes = {
        "fmt18.wav": riff(b"fmt " + struct.pack("<I", 18) + pcm16 + b"\0\0" + b"data" + struct.pack("<I", 8) + bytes(8)),
        "listfirst.wav": riff(b"LIST" + struct.pack("<I", 4) + b"INFO" + b"fmt " + struct.pack("<I", 16) + pcm16 + b"data" + struct.pack("<I", 8) + bytes(8)),
        "bits32.wav": riff(b"fmt " + struct.pack("<I", 16) + struct.pack("<HHIIHH", 1, 2, 44100, 352800, 8, 32) + b"data" + struct.pack("<I", 16) + bytes(16)),
        "nodata.wav": riff(b"fmt " + struct.pack("<I", 16) + pcm16),
        "short.wav": riff(b"fmt " + struct.pack("<I", 16) + pcm16 + b"data" + struct.pack("<I", 4000) + bytes(8)),
        "ch9.wav": riff(b"fmt " + struct.pack("<I", 16) + struct.pack("<HHIIHH", 1, 9, 44100, 793800, 18, 16) + b"data" + struct.pack("<I", 18) + bytes(18)),
    }
    for name, blob in shapes.items():
        (tmp_path / name).write_bytes(blob)
    r = run("-o", str(tmp_path / "none"), *[str(tmp_path / n) for n in shapes])
    assert r.returncode == 1
    for name in shapes:
        assert f"Failed to open {tmp_path / name}" in r.stderr, (name, r.stderr)
    import torch
    if not torch.cuda.is_available():
        good = tmp_path / "ok.wav"
        fmt = struct.pack("<HHIIHH", 1, 1, 48000, 96000, 2, 16)
        good.write_bytes(b"RIFF" + struct.pack("<I", 36 + 64) + b"WAVE" + b"fmt " + struct.pack("<I", 16) + fmt + b"data" + struct.pack("<I", 64) + bytes(64))
        r = run("-o", str(tmp_path / "out"), str(good), str(bad))
        assert r.returncode == 1 and "Failed to create encoder handle" in r.stderr and "float.wav" in r.stderr, r.stderr
        assert not (tmp_path / "out" / "ok.srl").exists()


def test_product_code_never_touches_the_oracle():
    """the oracle is test infrastructure: nothing under srla_b200/ or include/ may reference it"""
    for base, _dirs, files in os.walk(os.path.join(ROOT, "srla_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(base, f), errors="ignore").read()
                assert "liboracle" not in text and "srla_oracle" not in text and "libsrla_ref" not in text, f


def test_shard_streams_world2_gloo(tmp_path):
    """multi-GPU sharding is by whole streams, contiguous ranges, no exchange (DESIGN.md section e)"""
    script = tmp_path / "shard.py"
    script.write_text(
        "import os, sys, torch, torch.distributed as dist\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "from srla_b200.sharding import shard_range\n"
        "dist.init_process_group('gloo')\n"
        "r, w = dist.get_rank(), dist.get_world_size()\n"
        "lo, hi = shard_range(1024 + 3, r, w)\n"
        "t = torch.zeros(1027, dtype=torch.int64); t[lo:hi] = 1\n"
        "dist.all_reduce(t)\n"
        "assert bool((t == 1).all()), 'ranges overlap or leave holes'\n"
        "sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(w)]\n"
        "dist.all_gather(sizes, torch.tensor([hi - lo]))\n"
        "assert max(int(s) for s in sizes) - min(int(s) for s in sizes) <= 1\n"
        "dist.destroy_process_group()\n")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29617", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]


def test_host_placement_helpers(tmp_path):
    """srla_b200/sharding.py: sysfs cpulist parsing, the CPUs next to a GPU, and how ranks sharing a node split them"""
    from srla_b200.sharding import gpu_locality, parse_cpulist, split_cpus
    assert parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert parse_cpulist("") == []
    dev = tmp_path / "0000:1b:00.0"
    dev.mkdir()
    (dev / "numa_node").write_text("1\n")
    (dev / "local_cpulist").write_text("32-63\n")
    assert gpu_locality("0000:1B:00.0", str(tmp_path)) == (1, list(range(32, 64)))
    assert gpu_locality("0000:ff:00.0", str(tmp_path)) == (None, [])
    (dev / "numa_node").write_text("-1\n")                       # a host that does not report the node
    assert gpu_locality("0000:1b:00.0", str(tmp_path))[0] is None
    cpus = list(range(32))
    parts = [split_cpus(cpus, 8, i) for i in range(8)]
    assert sorted(c for p in parts for c in p) == cpus and all(len(p) == 4 for p in parts)
    assert split_cpus(cpus, 1, 0) == cpus
    assert all(len(split_cpus([0, 1, 2], 8, i)) == 1 for i in range(8))      # more ranks than CPUs: one each, wrapped


def test_pinned_ranks_world2_gloo(tmp_path):
    """two ranks whose GPUs sit on the same node take disjoint halves of its CPUs (world_size-2, gloo)"""
    fake = tmp_path / "sys"
    for bdf in ("0000:1b:00.0", "0000:1c:00.0"):
        (fake / bdf).mkdir(parents=True)
        (fake / bdf / "numa_node").write_text("0\n")
        (fake / bdf / "local_cpulist").write_text(",".join(str(c) for c in sorted(os.sched_getaffinity(0))) + "\n")
    script = tmp_path / "pin.py"
    script.write_text(
        "import os, sys, torch, torch.distributed as dist\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "import srla_b200.sharding as S\n"
        f"real = S.gpu_locality; S.gpu_locality = lambda bdf, root=None: real(bdf, {str(fake)!r})\n"
        "dist.init_process_group('gloo')\n"
        "r, w = dist.get_rank(), dist.get_world_size()\n"
        "before = sorted(os.sched_getaffinity(0))\n"
        "info = S.pin_rank_to_gpu_cpus(['0000:1b:00.0', '0000:1c:00.0'], r, w)\n"
        "mine = sorted(os.sched_getaffinity(0))\n"
        "t = torch.zeros(4096, dtype=torch.int64); t[mine] = 1\n"
        "dist.all_reduce(t)\n"
        "assert info['pinned'] and info['ranks_sharing_them'] == 2, info\n"
        "if len(before) >= 2:\n"
        "    assert int(t.max()) == 1 and int(t.sum()) == len(before), 'halves overlap or leave CPUs unused'\n"
        "dist.destroy_process_group()\n")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29618", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
