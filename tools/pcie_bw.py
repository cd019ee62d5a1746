#!/usr/bin/env python3
"""Pinned host<->device copy bandwidth of this box (the ceiling of bench.py's e2e number)."""
import torch, time
n = 164 * 1000 * 1000
h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device="cuda")
h2 = torch.empty(n // 2, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n // 2, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best
a = t(lambda: d.copy_(h, non_blocking=True)); print(f"H2D {n/a/1e9:.1f} GB/s ({a*1e3:.2f} ms for {n/1e6:.0f} MB)")
b = t(lambda: h2.copy_(d2, non_blocking=True)); print(f"D2H {n/2/b/1e9:.1f} GB/s ({b*1e3:.2f} ms for {n/2e6:.0f} MB)")
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
c = t(both); print(f"H2D {n/1e6:.0f} MB + D2H {n/2e6:.0f} MB concurrently: {c*1e3:.2f} ms")
