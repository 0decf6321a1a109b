"""Synthetic workloads of SURVEY.md §8(d) (configs 1-5 of BASELINE.json).

There is no network for datasets, so bench.py and the tests generate their
inputs here.  All signals are int16-valued float32, exactly what
`read_wav_mono` returns for a 16-bit file (the reference never normalises,
fractal.py:96,113).
"""
from __future__ import annotations

import numpy as np


def sine_noise(seconds=10.0, rate=16000, seed=1234):
    """Config 1: 440 Hz sine at half scale + 5 % white noise."""
    rng = np.random.default_rng(seed)
    t = np.arange(int(seconds * rate)) / rate
    x = np.round(0.5 * 32767 * np.sin(2 * np.pi * 440 * t)
                 + 0.05 * 32767 * rng.standard_normal(len(t)))
    return np.clip(x, -32768, 32767).astype(np.int16).astype(np.float32)


def _one_pole(white, a=0.98):
    from scipy.signal import lfilter
    return lfilter([1.0 - a], [1.0, -a], white)


def music_like(seconds, rate, seed):
    """Configs 2-5: a fundamental that steps every 0.25-0.5 s carrying 6-8
    exponentially decaying harmonic partials, plus -40 dB pink-ish noise; peak
    about 0.7 full scale."""
    rng = np.random.default_rng(seed)
    n = int(seconds * rate)
    tone = np.zeros(n, dtype=np.float32)
    pos = 0
    while pos < n:
        m = min(int(rng.uniform(0.25, 0.5) * rate), n - pos)
        f0 = 110.0 * 2.0 ** (rng.integers(0, 36) / 12.0)
        t = np.arange(m) / rate
        note = np.zeros(m)
        for h in range(1, int(rng.integers(6, 9)) + 1):
            note += (h ** -1.2) * np.exp(-t * (2.0 + 1.5 * h)) * \
                np.sin(2 * np.pi * f0 * h * t + rng.uniform(0, 2 * np.pi))
        tone[pos:pos + m] = note
        pos += m
    tone /= max(1e-9, float(np.max(np.abs(tone))))
    white = rng.standard_normal(n)
    pink = _one_pole(white)
    noise = 0.5 * white + 0.5 * pink / max(1e-9, float(np.std(pink)))
    mix = 0.7 * tone + 0.01 * (noise / max(1e-9, float(np.std(noise)))).astype(np.float32)
    return np.round(32767 * np.clip(mix, -1.0, 1.0)).astype(np.int16).astype(np.float32)


def test_tone(sr=8000, dur=0.12, freq=440.0):
    """The reference's own fixture (test_e2e.py:6-10)."""
    t = np.linspace(0, dur, int(sr * dur), endpoint=False)
    sig = (0.5 * (2 ** 15 - 1) * np.sin(2 * np.pi * freq * t)).astype(np.int16)
    return sig.astype(np.float32), sr, 2


CONFIGS = {
    # name: (generator, kwargs, tile_size, top_k)
    "c1": (sine_noise, dict(seconds=10.0, rate=16000, seed=1234), 1024, 32),
    "c2": (music_like, dict(seconds=180.0, rate=44100, seed=2), 4096, 32),
    "c3": (music_like, dict(seconds=3600.0, rate=48000, seed=3), 4096, 32),
    "c4": (music_like, dict(seconds=1800.0, rate=48000, seed=4), 1024, 64),
}


def make(name, scale=1.0):
    """Signal + (tile_size, top_k) of a named config; `scale` shortens it."""
    gen, kw, tile, k = CONFIGS[name]
    kw = dict(kw)
    kw["seconds"] = kw["seconds"] * scale
    return gen(**kw), kw["rate"], tile, k
