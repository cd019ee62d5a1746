"""Synthetic benchmark workloads (BASELINE.json `configs`).

config 2: 10 000 stereo 16-bit blocks of 4096 samples (40.96 M frames) at mode 4.  Generating 40 M
samples of AR-coloured harmonic signal with scipy takes minutes, so the workload is assembled from
`num_templates` independently seeded segments (srla_b200.synth.synth_stereo, SURVEY.md 8d recipe),
each reused with a different integer gain and a circular shift so that no two blocks hold the same
samples.  Deterministic for a given (seed, sizes).
"""
from __future__ import annotations

import numpy as np

from .synth import synth_stereo


def make_blocks_workload(num_blocks: int = 10000, block: int = 4096, channels: int = 2, bits: int = 16,
                         seed: int = 1234, num_templates: int = 8, template_blocks: int = 125) -> np.ndarray:
    """int16 (bits <= 16) or int32 planar PCM [channels, num_blocks * block]."""
    total = num_blocks * block
    dtype = np.int16 if bits <= 16 else np.int32
    out = np.empty((channels, total), dtype=dtype)
    tlen = template_blocks * block
    templates = [synth_stereo(tlen, seed=seed + t, bits=bits, channels=channels) for t in range(num_templates)]
    full = 1 << (bits - 1)
    at, k = 0, 0
    while at < total:
        t = templates[k % num_templates]
        rep = k // num_templates
        gain_num = 16 - (rep % 9)                      # 16/16, 15/16, ... 8/16
        seg = np.roll(t, 977 * rep, axis=1).astype(np.int64) * gain_num // 16
        seg = np.clip(seg, -full, full - 1)
        n = min(tlen, total - at)
        out[:, at:at + n] = seg[:, :n].astype(dtype)
        at += n
        k += 1
    return out
