#!/usr/bin/env python3
"""bench.py -- SRLA encode throughput on B200 (BASELINE.json: "encode Msamples/sec at mode 4 / block 4096").

A "step" is one pass of the hot path over the whole workload (configs[1]: 10 000 stereo 16-bit blocks of
4096 samples, mode 4, fixed blocks).  Per rank:

  value  device-resident: PCM already in HBM, SRLAB200_EncodeStreamsDevice, output left in HBM
  e2e    the same workload through the reference's own entry point, SRLAEncoder_EncodeWhole (planar int32 PCM and
         the output buffer in pageable HOST memory; narrowing, H2D, kernels, D2H inside the timed region)
  e2e_batch_api   the same through SRLAB200_EncodeStreamsHost with pinned host int16 PCM in, pinned bytes out

`--impl reference` times the UNMODIFIED reference (oracle/_ref/libsrla_ref.so, built from /root/reference
by oracle/Makefile; falls back to our C port oracle/liboracle.so) on the host cores, one handle per thread.
Multi-GPU: the path shards by independent streams/blocks -- every rank encodes its own copy of the workload
(weak scaling), no data-path collective; torch.distributed carries only the barrier and the max-over-ranks
timing.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BLOCK = 4096
PRESET = 4
RATE = 48000
CHANNELS = 2
BITS = 16


# ------------------------------------------------------------------------------------------ reference arm
class _Param(C.Structure):
    _fields_ = [("num_channels", C.c_uint16), ("bits_per_sample", C.c_uint16), ("sampling_rate", C.c_uint32),
                ("min_num_samples_per_block", C.c_uint32), ("max_num_samples_per_block", C.c_uint32),
                ("num_lookahead_samples", C.c_uint32), ("ltp_order", C.c_uint32),
                ("num_svr_filter_learning_iteration", C.c_uint32), ("preset", C.c_uint8)]


class _Config(C.Structure):
    _fields_ = [("max_num_channels", C.c_uint32), ("min_num_samples_per_block", C.c_uint32),
                ("max_num_samples_per_block", C.c_uint32), ("max_num_lookahead_samples", C.c_uint32),
                ("max_num_parameters", C.c_uint32)]


class _SoParams(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("num_channels", "bits_per_sample", "sampling_rate", "min_block", "max_block",
                                          "lookahead", "ltp_order", "preset", "offset_lshift")]


class CpuEncoder:
    """The reference's own CPU implementation of the path (kind 'reference'), or our C port of it
    (kind 'port') when oracle/_ref was not built.  Used ONLY as the measured CPU baseline."""

    def __init__(self):
        ref = os.path.join(ROOT, "oracle", "_ref", "libsrla_ref.so")
        port = os.path.join(ROOT, "oracle", "liboracle.so")
        if os.path.exists(ref):
            self.kind, self.lib = "reference", C.CDLL(ref)
            PP = C.POINTER(C.POINTER(C.c_int32))
            self.lib.SRLAEncoder_Create.restype = C.c_void_p
            self.lib.SRLAEncoder_Create.argtypes = [C.POINTER(_Config), C.c_void_p, C.c_int32]
            self.lib.SRLAEncoder_SetEncodeParameter.argtypes = [C.c_void_p, C.POINTER(_Param)]
            self.lib.SRLAEncoder_EncodeWhole.argtypes = [C.c_void_p, PP, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.c_void_p]
            self.lib.SRLAEncoder_Destroy.argtypes = [C.c_void_p]
        else:
            if not os.path.exists(port):
                subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])
            self.kind, self.lib = "port", C.CDLL(port)
            self.lib.so_encode_whole_flat.argtypes = [C.POINTER(_SoParams), C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]

    def make_handle(self):
        if self.kind != "reference":
            return None
        cfg = _Config(8, BLOCK, BLOCK, BLOCK, 255)
        h = self.lib.SRLAEncoder_Create(C.byref(cfg), None, 0)      # serial: Create rebuilds a static table
        prm = _Param(CHANNELS, BITS, RATE, BLOCK, BLOCK, BLOCK, 0, 0, PRESET)
        assert h and self.lib.SRLAEncoder_SetEncodeParameter(h, C.byref(prm)) == 0
        return h

    def encode(self, handle, pcm32: np.ndarray, out: np.ndarray) -> int:
        nch, n = pcm32.shape
        size = C.c_uint32(0)
        if self.kind == "reference":
            rows = (C.POINTER(C.c_int32) * nch)()
            for c in range(nch):
                rows[c] = C.cast(pcm32[c].ctypes.data, C.POINTER(C.c_int32))
            rc = self.lib.SRLAEncoder_EncodeWhole(handle, rows, n, out.ctypes.data, out.size, C.byref(size), None)
        else:
            prm = _SoParams(nch, BITS, RATE, BLOCK, BLOCK, BLOCK, 0, PRESET, 0)
            rc = self.lib.so_encode_whole_flat(C.byref(prm), pcm32.ctypes.data, n, out.ctypes.data, out.size, C.byref(size))
        assert rc == 0, rc
        return size.value

    def destroy(self, handle):
        if handle:
            self.lib.SRLAEncoder_Destroy(handle)


def cpu_throughput(pcm16: np.ndarray, blocks_per_thread: int, threads: int, repeats: int = 1):
    """Each of `threads` host threads encodes its own slice of `blocks_per_thread` blocks of the workload
    (ctypes releases the GIL).  Returns (channel-samples per second, kind, first slice's bytes)."""
    enc = CpuEncoder()
    total_blocks = pcm16.shape[1] // BLOCK
    slices, handles, outs = [], [], []
    for t in range(threads):
        b0 = (t * max(1, total_blocks // threads)) % max(1, total_blocks - blocks_per_thread + 1)
        slices.append(np.ascontiguousarray(pcm16[:, b0 * BLOCK:(b0 + blocks_per_thread) * BLOCK].astype(np.int32)))
        handles.append(enc.make_handle())
        outs.append(np.zeros(slices[-1].size * 4 + 4096, dtype=np.uint8))
    sizes = [0] * threads

    def work(t):
        for _ in range(repeats):
            sizes[t] = enc.encode(handles[t], slices[t], outs[t])

    ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    t0 = time.perf_counter()
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    dt = time.perf_counter() - t0
    for h in handles:
        enc.destroy(h)
    samples = sum(s.size for s in slices) * repeats
    return samples / dt, enc.kind, outs[0][:sizes[0]].tobytes(), slices[0]


def run_reference_arm(args, rank: int) -> None:
    if rank != 0:
        return
    from srla_b200.workload import make_blocks_workload
    threads = os.cpu_count() or 1
    pcm = make_blocks_workload(args.blocks, BLOCK, CHANNELS, BITS)          # the SAME signal the GPU arm encodes
    # calibrate so that the whole --steps/--warmup run stays within a few minutes
    rate1, kind, _, _ = cpu_throughput(pcm, 32, 1)
    budget_s = 8.0
    bpt = int(max(16, min(pcm.shape[1] // BLOCK, rate1 * budget_s / (BLOCK * CHANNELS))))
    times = []
    for step in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        rate, kind, _, _ = cpu_throughput(pcm, bpt, threads)
        if step >= args.warmup:
            times.append((time.perf_counter() - t0, rate))
    rate = float(np.mean([r for _, r in times]))
    ms = 1e3 * float(np.mean([t for t, _ in times]))
    sample = (f"{threads} host threads x {bpt} blocks of {BLOCK} stereo 16-bit frames each per step "
              f"(evenly spaced slices of the same {args.blocks}-block config-2 signal the GPU arm encodes)")
    line = {"impl": "reference", "metric": "encode Msamples/s (mode 4, block 4096, stereo 16-bit; 1 sample = 1 channel-sample)",
            "value": rate / 1e6, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32+f64",
            "data": "synthetic", "config": workload_config(args.blocks),
            "cpu_baseline": {"value": rate / 1e6, "unit": "Msamples/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": rate / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ helpers
def workload_config(blocks: int):
    return {"workload": f"configs[1]: {blocks} stereo 16-bit blocks of {BLOCK} samples, mode 4 (-m 4 -B 4096 -V 0), 48 kHz, "
                        "fixed blocks, LTP off",
            "blocks": blocks, "block_samples": BLOCK, "channels": CHANNELS, "bits": BITS, "preset": PRESET,
            "l2_policy": "inputs larger than L2 (163.84 MB int16 PCM + 655 MB residual scratch per step vs 126 MB L2)"}


class ClockSampler:
    """nvidia-smi sampling in the background DURING the timed region (B200_PROFILING.md clocks line)."""
    QUERY = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.3)          # let the sampler come up before the timed region starts
        except Exception:
            self.proc = None

    def stop(self):
        rows = []
        if self.proc is not None:
            time.sleep(0.1)
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
            except Exception:
                self.proc.kill()
                out = ""
            rows = [[x.strip() for x in ln.split(",")] for ln in out.splitlines() if ln.strip()]
        sm = [float(r[0]) for r in rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            for i, nm in enumerate(names):
                if len(r) > 3 + i and r[3 + i].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(rows)}


def load_traffic():
    """Per-kernel figures of the committed `ncu --set full` capture (profiles/ncu_kernels.json, written by
    tools/ncu_summary.py --traffic from the same bench command): dram bytes per launch, pipe utilisation."""
    path = os.path.join(ROOT, "profiles", "ncu_kernels.json")
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:
        return None


def load_floors():
    """Issue-rate floors per kernel (profiles/r2_kernel_floors.json, written by tools/ncu_floors.py from the committed
    `ncu --set full` capture: executed warp instructions per SASS opcode x the issue cost measured with tools/pipe_rates.cu)."""
    path = os.path.join(ROOT, "profiles", "r2_kernel_floors.json")
    try:
        with open(path) as f:
            return json.load(f).get("kernels")
    except Exception:
        return None


# ------------------------------------------------------------------------------------------ config 5
def run_config5(args, rank, world, local_rank, enc, lib, barrier, torch, dist):
    """BASELINE configs[4]: `--files` synthetic 48 kHz stereo 16-bit WAV payloads of 30 s (1 440 000 frames = 351 blocks of
    4096 + a 2304-frame tail), mode 4, split over the ranks by srla_b200.sharding.shard_range -- total work is fixed, so
    the throughput over N is a STRONG-scaling curve.  Per rank and step the shard goes through
    SRLAB200_EncodeInterleavedHost (page-locked WAV data-chunk bytes in, page-locked .srl bytes out: H2D, de-interleave,
    encode, D2H inside the timed region); the same shard is also timed device-resident, and a copy-only pass (the shard's
    bytes host->device, its encoded bytes device->host, all ranks at once) gives the ceiling the host side allows."""
    from srla_b200 import encoder as E
    from srla_b200.sharding import shard_range
    from srla_b200.synth import synth_stereo
    frames = 30 * RATE
    lo, hi = shard_range(args.files, rank, world)
    mine = hi - lo
    if mine == 0:
        lo, hi, mine = 0, 1, 1                       # more ranks than files: this rank repeats file 0 (not counted)
        counted = 0
    else:
        counted = mine
    # file k = rotation + integer gain of one synthetic 30 s signal (seed 1234): deterministic, every file different
    base = synth_stereo(frames, seed=1234).astype(np.int16)
    inter = np.ascontiguousarray(base.T)                                   # [frames, 2] = WAV data-chunk order
    pitch = (frames * CHANNELS * 2 + 255) // 256 * 256                     # bytes per file in the staging buffers
    h_raw = torch.empty(mine * pitch, dtype=torch.uint8).pin_memory()
    raw_np = h_raw.numpy()
    for i in range(mine):
        k = lo + i
        seg = np.roll(inter, 7919 * k, axis=0).astype(np.int32) * (16 - (k % 7)) // 16
        raw_np[i * pitch:i * pitch + frames * 4] = seg.astype("<i2").reshape(-1).view(np.uint8)
    items = (E.SRLAB200Frames * mine)()
    for i in range(mine):
        items[i] = E.SRLAB200Frames(h_raw.data_ptr() + i * pitch, frames)
    cap = mine * enc.max_encoded_size(frames)
    h_out = torch.empty(cap, dtype=torch.uint8).pin_memory()
    offs = (C.c_uint64 * (mine + 1))()

    def step_e2e():
        rc = lib.SRLAB200_EncodeInterleavedHost(enc.handle, items, mine, h_out.data_ptr(), cap, offs)
        assert rc == E.OK, rc

    def wall(fn, steps):
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
        barrier()
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    steps = 3
    step_e2e(); step_e2e()                                                  # warm-up: buffers grow, tiling cached
    ms_e2e = wall(step_e2e, steps)
    out_bytes = int(offs[mine])
    total_samples = args.files * frames * CHANNELS                           # the whole batch, all ranks
    # device resident: the same shard as planar int16 in HBM, output left in HBM
    d_raw = h_raw.cuda()
    stride = (frames + 15) // 16 * 16
    d_planar = torch.zeros((mine, CHANNELS, stride), dtype=torch.int16, device="cuda")
    for i in range(mine):
        d_planar[i, :, :frames] = d_raw[i * pitch:i * pitch + frames * 4].view(torch.int16).view(frames, CHANNELS).t()
    descs = (E.SRLAB200Stream * mine)()
    for i in range(mine):
        descs[i] = E.SRLAB200Stream(d_planar[i].data_ptr(), stride, frames, 2)
    d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")

    def step_dev():
        rc = lib.SRLAB200_EncodeStreamsDevice(enc.handle, descs, mine, d_out.data_ptr(), cap, offs)
        assert rc == E.OK, rc

    step_dev()
    ms_dev = wall(step_dev, steps)
    same = int(offs[mine]) == out_bytes
    # copy-only ceiling: what the host side allows when every rank moves its shard at once
    side = torch.cuda.Stream()

    def step_copy():
        # both directions at once, as the encode pipeline runs them (two copy engines)
        d_raw.copy_(h_raw, non_blocking=True)
        with torch.cuda.stream(side):
            h_out[:out_bytes].copy_(d_out[:out_bytes], non_blocking=True)

    step_copy()
    ms_copy = wall(step_copy, steps)
    rate = lambda ms: total_samples * steps / (ms * 1e-3) / 1e6
    res = {"workload": f"configs[4]: {args.files} synthetic 48 kHz stereo 16-bit files of 30 s (351 blocks of 4096 + a 2304-frame tail), "
                       f"mode 4, sharded {world} ways by contiguous file ranges (strong scaling: total work fixed)",
           "files": args.files, "files_per_rank": counted, "frames_per_file": frames, "scaling": "strong",
           "e2e": {"value": rate(ms_e2e), "unit": "Msamples/s", "ms_per_step": ms_e2e / steps,
                   "h2d_bytes_per_step_per_rank": mine * frames * 4, "d2h_bytes_per_step_per_rank": out_bytes,
                   "api": "SRLAB200_EncodeInterleavedHost per rank on its shard (page-locked WAV payloads in, page-locked .srl bytes out); "
                          "wall clock, max over ranks"},
           "device": {"value": rate(ms_dev), "unit": "Msamples/s", "ms_per_step": ms_dev / steps,
                      "api": "SRLAB200_EncodeStreamsDevice per rank on its shard (planar int16 in HBM)", "same_size_as_e2e": bool(same)},
           "copy_ceiling": {"value": rate(ms_copy), "unit": "Msamples/s", "ms_per_step": ms_copy / steps,
                            "h2d_gbs_per_rank": mine * frames * 4 * steps / (ms_copy * 1e-3) / 1e9,
                            "what": "the shard's WAV bytes host->device and its encoded bytes device->host from the same page-locked buffers, "
                                    "both directions concurrently, no kernels, all ranks at once: the end-to-end ceiling the host/PCIe side sets"},
           "e2e_fraction_of_copy_ceiling": ms_copy / ms_e2e}
    del d_raw, d_planar, d_out, h_raw, h_out
    torch.cuda.empty_cache()
    return res


# ------------------------------------------------------------------------------------------ our arm
def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--blocks", type=int, default=10000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--files", type=int, default=1024, help="config 5: number of 30 s stereo files in the sharded batch (0: skip)")
    ap.add_argument("--no-decode", action="store_true", help="skip the decoder figures (N=1 only)")
    ap.add_argument("--no-pin", action="store_true", help="do not restrict the rank to the CPUs next to its GPU")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    from srla_b200 import encoder as E
    from srla_b200.workload import make_blocks_workload

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the SRLA B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    # host placement: this rank's page-locked staging memory and feeder threads stay next to its GPU, and the ranks that
    # share a memory node split its CPUs (srla_b200/sharding.py).  Before any pinned allocation and before the library
    # starts its thread pool.
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    placement = None
    if not args.no_pin:
        from srla_b200.sharding import pin_rank_to_gpu_cpus
        try:
            ids = []
            for i in range(torch.cuda.device_count()):
                pr = torch.cuda.get_device_properties(i)
                ids.append("%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id))
            placement = pin_rank_to_gpu_cpus(ids, local_rank, min(local_world, len(ids)))
        except Exception as exc:                                     # placement is an optimisation, never a failure
            placement = {"pinned": False, "error": str(exc)}
    # Everything libraries print on the way (NCCL's version banner goes to stdout) is sent to stderr: stdout carries
    # exactly ONE line, the JSON below.
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # ---- workload: pinned host int16 PCM (one stream of blocks*4096 frames) and its device copy ----
    pcm_np = make_blocks_workload(args.blocks, BLOCK, CHANNELS, BITS)
    nsamp = pcm_np.shape[1]
    stride = (nsamp + 15) // 16 * 16
    h_pcm = torch.empty((CHANNELS, stride), dtype=torch.int16).pin_memory()
    h_pcm.zero_()
    h_pcm[:, :nsamp] = torch.from_numpy(pcm_np)
    d_pcm = h_pcm.cuda(non_blocking=False)

    E.load_library().SRLAB200_SetDevice(local_rank)
    enc = E.Encoder(max_channels=CHANNELS, max_block=BLOCK, device=local_rank)
    assert enc.set_parameter(CHANNELS, BITS, RATE, BLOCK, BLOCK, BLOCK, 0, PRESET) == E.OK
    stream = torch.cuda.current_stream()
    enc.set_stream(stream.cuda_stream)
    cap = enc.max_encoded_size(nsamp)
    d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    h_out = torch.empty(cap, dtype=torch.uint8).pin_memory()

    dev_desc = (E.SRLAB200Stream * 1)(E.SRLAB200Stream(d_pcm.data_ptr(), stride, nsamp, 2))
    host_desc = (E.SRLAB200Stream * 1)(E.SRLAB200Stream(h_pcm.data_ptr(), stride, nsamp, 2))
    offs = (C.c_uint64 * 2)()
    lib = enc.lib

    def step_device():
        rc = lib.SRLAB200_EncodeStreamsDevice(enc.handle, dev_desc, 1, d_out.data_ptr(), cap, offs)
        assert rc == E.OK, rc

    def step_host():
        rc = lib.SRLAB200_EncodeStreamsHost(enc.handle, host_desc, 1, h_out.data_ptr(), cap, offs)
        assert rc == E.OK, rc

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- device-resident throughput ----
    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local_rank)
    sampler.start()
    an_ms, em_ms, launches = [], [], 0
    k_ms = {"front_kernel": [], "lpc_kernels(levinson+select)": [], "residual_kernel": [], "emit_kernel(+decide+scan)": []}
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
        st = enc.stats()
        an_ms.append(st.ms_analyse); em_ms.append(st.ms_emit); launches += int(st.kernel_launches)
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop()
    # ---- per-kernel durations: the default device-resident call runs as groups on several lanes whose kernels overlap, so
    # the time between two of its events is not one kernel's.  A second handle created with SRLA_B200_SPLIT_DEVICE=0 runs
    # the same call as ONE batch on one stream; its CUDA-event times are the kernels' own (and what ncu's launch list shows).
    prev_split = os.environ.get("SRLA_B200_SPLIT_DEVICE")
    os.environ["SRLA_B200_SPLIT_DEVICE"] = "0"
    enc_serial = E.Encoder(max_channels=CHANNELS, max_block=BLOCK, device=local_rank)
    if prev_split is None:
        del os.environ["SRLA_B200_SPLIT_DEVICE"]
    else:
        os.environ["SRLA_B200_SPLIT_DEVICE"] = prev_split
    assert enc_serial.set_parameter(CHANNELS, BITS, RATE, BLOCK, BLOCK, BLOCK, 0, PRESET) == E.OK
    enc_serial.set_stream(stream.cuda_stream)
    serial_steps = max(3, min(args.steps, 10))
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(args.warmup + serial_steps):
        if i == args.warmup:
            torch.cuda.synchronize()
            s0.record(stream)
        rc = lib.SRLAB200_EncodeStreamsDevice(enc_serial.handle, dev_desc, 1, d_out.data_ptr(), cap, offs)
        assert rc == E.OK, rc
        if i >= args.warmup:
            sst = enc_serial.stats()
            k_ms["front_kernel"].append(sst.ms_front); k_ms["lpc_kernels(levinson+select)"].append(sst.ms_lpc)
            k_ms["residual_kernel"].append(sst.ms_residual); k_ms["emit_kernel(+decide+scan)"].append(sst.ms_emit)
    s1.record(stream)
    torch.cuda.synchronize()
    serial_ms = [s0.elapsed_time(s1) / serial_steps]
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    st = enc.stats()
    bytes_out = int(st.bytes_out)
    samples_per_step = nsamp * CHANNELS
    value = world * samples_per_step * args.steps / (ms_total * 1e-3) / 1e6

    # ---- end to end through the batch extension of the C ABI: pinned int16 host PCM in, pinned host bytes out ----
    for _ in range(max(1, args.warmup // 2)):
        step_host()
    e2e_steps = max(2, args.steps // 2)
    ms_batch = timed(step_host, e2e_steps)
    batch_value = world * samples_per_step * e2e_steps / (ms_batch * 1e-3) / 1e6
    batch_api = {"value": batch_value, "unit": "Msamples/s", "h2d_bytes_per_step": CHANNELS * stride * 2, "d2h_bytes_per_step": int(offs[1]),
                 "ms_per_step": ms_batch / e2e_steps,
                 "api": "SRLAB200_EncodeStreamsHost (pinned int16 host PCM in, pinned host bytes out; groups of blocks pipelined over "
                        "copy streams and compute lanes)"}

    # ---- end to end through the REFERENCE'S OWN entry point, every rank: SRLAEncoder_EncodeWhole(int32_t *const *input, ...)
    # with planar int32 PCM and the output buffer in ordinary (pageable) host memory, as tools/srla_codec calls it.  Wall
    # clock around the calls (they return when the bytes are in the caller's buffer), barrier on both sides, max over ranks.
    pcm32 = np.ascontiguousarray(pcm_np.astype(np.int32))
    rows = (C.POINTER(C.c_int32) * CHANNELS)()
    for ch in range(CHANNELS):
        rows[ch] = C.cast(pcm32[ch].ctypes.data, C.POINTER(C.c_int32))
    out_np = np.empty(cap, dtype=np.uint8)
    size = C.c_uint32(0)
    lib.SRLAEncoder_EncodeWhole.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_int32)), C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.c_void_p]

    def step_reference_api():
        rc = lib.SRLAEncoder_EncodeWhole(enc.handle, rows, nsamp, out_np.ctypes.data, min(cap, 0xffffffff), C.byref(size), None)
        assert rc == E.OK, rc

    for _ in range(max(2, args.warmup // 2)):
        step_reference_api()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_reference_api()
    torch.cuda.synchronize()
    ms_e2e = (time.perf_counter() - t0) * 1e3
    barrier()
    if world > 1:
        t = torch.tensor([ms_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    e2e_value = world * samples_per_step * e2e_steps / (ms_e2e * 1e-3) / 1e6
    # what the HOST side of this entry point allows: the reference signature hands over 4-byte samples in pageable memory,
    # so every sample costs 4 bytes read + 2 bytes written of host DRAM traffic before the copy engine can take it.  All
    # ranks run ONLY that narrowing pass (the library's own routine, same thread count) at the same time.
    lib.SRLAB200_TestNarrow.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    lib.SRLAB200_TestNarrow.restype = C.c_uint32
    feed_threads = max(2, min(16, len(os.sched_getaffinity(0))))
    h_narrow = torch.empty((CHANNELS, nsamp), dtype=torch.int16).pin_memory()
    chunk = max(1 << 16, -(-nsamp // (2 * feed_threads)))       # a few large pieces per thread: the Python dispatch must not count
    pieces = [(ch, at) for ch in range(CHANNELS) for at in range(0, nsamp, chunk)]

    def narrow_worker(t):
        for ch, at in pieces[t::feed_threads]:
            cnt = min(chunk, nsamp - at)
            lib.SRLAB200_TestNarrow(pcm32[ch].ctypes.data + 4 * at, h_narrow.data_ptr() + 2 * (ch * nsamp + at), cnt)

    def narrow_pass():
        ths = [threading.Thread(target=narrow_worker, args=(t,)) for t in range(feed_threads)]
        for th in ths:
            th.start()
        for th in ths:
            th.join()

    narrow_pass()
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        narrow_pass()
    ms_feed = (time.perf_counter() - t0) * 1e3
    barrier()
    if world > 1:
        t = torch.tensor([ms_feed], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_feed = float(t.item())
    feed_ceiling = world * samples_per_step * 3 / (ms_feed * 1e-3) / 1e6
    del h_narrow
    h2d = CHANNELS * nsamp * 2          # the host feeder narrows the int32 input to int16 on its way into pinned staging
    d2h = int(size.value)
    same_bytes = bool(bytes(out_np[:size.value]) == bytes(h_out[:int(offs[1])].numpy().tobytes()))

    # ---- decoder on the stream the call above wrote: SRLADecoder_DecodeWhole, checked against the source PCM.
    # Device time from a handle that decodes with ONE pair of launches (SRLA_B200_DECODE_PIPELINE=0: CUDA events around
    # decode_parse_kernel + decode_blocks_kernel); end to end from a default handle, whose third and later long streams
    # take the pipelined path (pageable stream in, pageable planar int32 out: 4 bytes per sample back over PCIe). ----
    decode = None
    if world == 1 and not args.no_decode:
        from srla_b200 import decoder as D
        stream = out_np[:size.value].tobytes()
        got = np.zeros_like(pcm32)
        saved_env = os.environ.get("SRLA_B200_DECODE_PIPELINE")
        os.environ["SRLA_B200_DECODE_PIPELINE"] = "0"
        try:
            with D.Decoder() as dec:
                dev_ms = None
                for _ in range(3):
                    dec.decode_whole(stream, got)
                    dev_ms = dec.kernel_ms() if dev_ms is None else min(dev_ms, dec.kernel_ms())
        finally:
            if saved_env is None:
                del os.environ["SRLA_B200_DECODE_PIPELINE"]
            else:
                os.environ["SRLA_B200_DECODE_PIPELINE"] = saved_env
        identical = bool(np.array_equal(got, pcm32))
        with D.Decoder() as dec:
            best = None
            for it in range(6):
                got[:, :4096] = -1
                t0 = time.perf_counter()
                dec.decode_whole(stream, got)
                dt = time.perf_counter() - t0
                if it >= 2:
                    best = dt if best is None else min(best, dt)
        identical = identical and bool(np.array_equal(got, pcm32))
        decode = {"device": {"value": samples_per_step / (dev_ms * 1e-3) / 1e6, "unit": "Msamples/s", "ms": dev_ms,
                             "what": "decode_parse_kernel + decode_blocks_kernel over all blocks of the stream, CUDA events, stream resident in HBM"},
                  "e2e": {"value": samples_per_step / best / 1e6, "unit": "Msamples/s", "ms": best * 1e3,
                          "h2d_bytes": len(stream), "d2h_bytes": int(pcm32.nbytes),
                          "api": "SRLADecoder_DecodeWhole (include/srla_decoder.h:46-49): pageable stream in, pageable planar int32 PCM out; wall clock, best of 4"},
                  "identical_to_source": identical, "stream_bytes": len(stream)}
        del got
    del pcm32

    # ---- config 5 (BASELINE configs[4]): a batch of 30 s stereo files SHARDED over the ranks (strong scaling) ----
    config5 = None
    if args.files > 0:
        del d_pcm, h_pcm, d_out, h_out
        torch.cuda.empty_cache()
        config5 = run_config5(args, rank, world, local_rank, enc, lib, barrier, torch, dist)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (analyse_kernel): algorithmic bytes / measured duration ----
    with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
        peaks = json.load(f)
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650"
    alg_bytes = int(st.bytes_in) + bytes_out              # PCM in (2 B/sample) + encoded bytes out
    kernel_ms = {k: float(np.mean(v)) for k, v in k_ms.items()}
    dominant = max(("front_kernel", "lpc_kernels(levinson+select)", "residual_kernel"), key=lambda k: kernel_ms[k])
    an = kernel_ms[dominant]
    achieved = alg_bytes / (an * 1e-3) / 1e9
    traffic = load_traffic()
    floors = load_floors() or {}
    front_name = "front16_kernel" if "front16_kernel" in floors else "front_kernel"     # 16-bit PCM without LTP runs front16_kernel
    floor_of = {"front_kernel": [front_name], "lpc_kernels(levinson+select)": ["lpc_levinson_kernel", "lpc_select_kernel"],
                "residual_kernel": ["residual16_kernel"], "emit_kernel(+decide+scan)": ["decide_kernel", "scan_kernel", "emit_kernel"]}
    ncu_dom = (traffic or {}).get(floor_of[dominant][0]) or {}
    issue_floor = {k: (round(sum(floors[n]["issue_floor_ms"] for n in names), 4) if all(n in floors for n in names) else None)
                   for k, names in floor_of.items()}
    step_floor = sum(v for v in issue_floor.values() if v) if all(issue_floor.values()) else None
    roofline = {"bound": "hbm", "kernel": floor_of[dominant][0], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "issue_floor_ms": issue_floor.get(dominant),
                "compute_frac": (issue_floor[dominant] / an) if issue_floor.get(dominant) else None,
                "all_kernels_issue_floor_ms": issue_floor,
                "step_issue_floor_ms": step_floor,
                "step_compute_frac": (step_floor / (ms_total / args.steps)) if step_floor else None,
                "issue_floor_source": "profiles/r2_kernel_floors.json: executed warp instructions per SASS opcode (ncu) x issue cost per "
                                      "instruction class measured with tools/pipe_rates.cu; the kernel cannot run faster than its busiest pipe",
                "frac": achieved / peak, "peak_source": peak_src,
                "traffic": ncu_dom.get("dram_bytes_per_launch"),
                "ncu": {k: v for k, v in ncu_dom.items() if k != "dram_bytes_per_launch"} or None,
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": an,
                "share_of_step": an / float(np.mean(serial_ms)), "all_kernels_ms": kernel_ms,
                "kernel_ms_source": "CUDA events of a handle that runs the call as ONE batch on one stream (SRLA_B200_SPLIT_DEVICE=0): "
                                    "the default call overlaps the kernels of its groups on several lanes, so only the unsplit call "
                                    "has per-kernel times; `value` / `ms_per_step` are the default (overlapped) call",
                "serial_ms_per_step": float(np.mean(serial_ms)), "overlap_gain": float(np.mean(serial_ms)) / (ms_total / args.steps),
                "whole_step_achieved_gbs": alg_bytes / (ms_total / args.steps * 1e-3) / 1e9,
                "note": "compute-bound path (bit-exact non-FMA FP64 FFT + int32 FIR + Rice search, ~100 ops per algorithmic byte): "
                        "the HBM fraction is low by construction; DESIGN.md section 3 gives the issue-rate ceilings that bind"}

    # ---- CPU baseline on a bounded sample (rank 0, N=1 only) ----
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        rate1, kind, _, _ = cpu_throughput(pcm_np, 32, 1)
        bpt = int(max(16, min(args.blocks, rate1 * 12.0 / (BLOCK * CHANNELS))))
        rate, kind, cpu_bytes, cpu_slice = cpu_throughput(pcm_np, bpt, threads)
        # same bytes as the GPU for that slice?
        gpu_bytes = E.encode(cpu_slice, bps=BITS, rate=RATE, max_block=BLOCK, preset=PRESET)
        cpu = {"value": rate / 1e6, "unit": "Msamples/s", "cores": threads, "kind": kind,
               "sample": f"{threads} host threads x {bpt} blocks of {BLOCK} stereo frames (slices of this workload), "
                         f"single-thread rate {rate1 / 1e6:.2f} Msamples/s",
               "gpu_output_identical_on_sample": bool(gpu_bytes == cpu_bytes)}

    hist = np.array(st.order_histogram[:], dtype=np.int64)
    line = {"metric": "encode Msamples/s (mode 4, block 4096, stereo 16-bit; 1 sample = 1 channel-sample)",
            "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32+f64", "data": "synthetic", "config": workload_config(args.blocks),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / e2e_steps, "host_bytes_read_per_step": CHANNELS * nsamp * 4,
                    "api": "SRLAEncoder_EncodeWhole, the reference's own signature (include/srla_encoder.h:77-81): planar int32 PCM and the "
                           "output buffer in pageable host memory; wall clock, max over ranks",
                    "identical_to_batch_api_output": same_bytes,
                    "host_feed_ceiling": {"value": feed_ceiling, "unit": "Msamples/s", "threads_per_rank": feed_threads,
                                          "host_dram_gbs": feed_ceiling * 6e6 / 1e9,
                                          "what": "only the int32 -> int16 narrowing of the input into page-locked staging (4 B read + 2 B written "
                                                  "of host DRAM per sample), all ranks at once, no GPU work: the ceiling the reference's "
                                                  "pageable-int32 signature sets on this host"},
                    "fraction_of_host_feed_ceiling": e2e_value / feed_ceiling,
                    # everything the call moves through host DRAM per sample: 4 B read + 2 B written by the narrowing, 2 B read by the
                    # H2D copy engine, and the encoded bytes three times (written by the D2H copy engine into page-locked staging,
                    # read and written again on their way into the caller's pageable buffer)
                    "host_dram_traffic_gbs": e2e_value * 1e6 * (8.0 + 3.0 * d2h / float(samples_per_step)) / 1e9,
                    "fraction_of_host_dram_ceiling": (e2e_value * (8.0 + 3.0 * d2h / float(samples_per_step))) / (feed_ceiling * 6.0)},
            "e2e_batch_api": batch_api,
            "config5": config5,
            "decode": decode,
            "host_placement": placement,
            "gpu_launches": launches,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "stats": {"compression_ratio": bytes_out / float(st.bytes_in), "bytes_out": bytes_out,
                      "mean_lpc_order": float((hist * np.arange(256)).sum() / max(1, hist.sum())),
                      "order_histogram_by_16": [int(hist[i:i + 16].sum()) for i in range(0, 80, 16)],
                      "stereo_methods_LR_MS_LS_SR": list(st.method_histogram[:]),
                      "block_types_compress_silent_raw": list(st.type_histogram[:])}}
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)
    if world > 1:
        os.dup2(2, 1)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
