"""Deterministic synthetic PCM used by bench.py and the parity tests.

Recipe from SURVEY.md section 8(d): harmonic "mid" signal with slow tremolo plus AR(16)-coloured
noise and sparse decaying clicks; an independent AR(8) "side"; L = mid + side, R = mid - side;
peak-normalised to 0.7 full scale and rounded to the requested bit depth.  The mix is chosen so
that the encoder's selected LPC orders spread over the whole 1..64 range the way music does.
"""
from __future__ import annotations

import numpy as np


def _ar_noise(rng: np.random.Generator, n: int, order: int, r_lo: float, r_hi: float) -> np.ndarray:
    """White noise through a random stable all-pole filter (conjugate pole pairs)."""
    poles = []
    for _ in range(order // 2):
        r = rng.uniform(r_lo, r_hi)
        w = rng.uniform(0.02, 0.98) * np.pi
        poles += [r * np.exp(1j * w), r * np.exp(-1j * w)]
    a = np.real(np.poly(poles))
    e = rng.standard_normal(n + 256)
    y = np.zeros(n + 256)
    na = len(a)
    # direct-form recursion, vectorised in chunks is not possible; n is modest for tests and the
    # bench generates one template and tiles it with per-file gain/offset variation.
    from scipy.signal import lfilter  # scipy is available offline in this image
    y = lfilter([1.0], a, e)
    y = y[256:]
    return y / (np.std(y) + 1e-12)


def synth_stereo(num_samples: int, seed: int = 1234, rate: int = 48000, bits: int = 16,
                 channels: int = 2) -> np.ndarray:
    """Returns int32 array [channels, num_samples] holding `bits`-bit PCM."""
    rng = np.random.default_rng(seed)
    n = num_samples
    t = np.arange(n) / rate
    f0 = rng.uniform(80.0, 400.0)
    mid = np.zeros(n)
    for h in range(1, 13):
        mid += (1.0 / h) * np.sin(2 * np.pi * h * f0 * t + rng.uniform(0, 2 * np.pi))
    mid *= 1.0 + 0.3 * np.sin(2 * np.pi * rng.uniform(0.3, 3.0) * t)
    mid /= np.max(np.abs(mid)) + 1e-12
    mid += 10 ** (rng.uniform(-35, -20) / 20) * _ar_noise(rng, n, 16, 0.90, 0.995)
    nclick = max(1, n // 20000)
    for pos in rng.integers(0, n, nclick):
        ln = min(400, n - pos)
        mid[pos:pos + ln] += rng.uniform(0.05, 0.3) * np.exp(-np.arange(ln) / 40.0) * rng.choice([-1.0, 1.0])
    side = 10 ** (-25 / 20) * _ar_noise(rng, n, 8, 0.90, 0.995)
    for h in (1, 2, 3):
        side += 0.02 / h * np.sin(2 * np.pi * h * f0 * 1.003 * t + rng.uniform(0, 2 * np.pi))
    chans = [mid + side, mid - side]
    while len(chans) < channels:
        chans.append(0.5 * _ar_noise(rng, n, 8, 0.9, 0.99) * 0.2 + 0.3 * mid)
    x = np.stack(chans[:channels])
    x *= 0.7 / (np.max(np.abs(x)) + 1e-12)
    full = float(1 << (bits - 1))
    return np.clip(np.round(x * full), -full, full - 1).astype(np.int32)
