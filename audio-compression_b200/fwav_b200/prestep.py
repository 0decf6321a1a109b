"""Host pre-step of compress_audio: voiced gate, masking, reflect padding and
range framing (reference fractal.py:880-909 and :1070-1112).

SURVEY.md §8(f) row N2 marks this "next": it is a handful of numpy passes over
the signal, kept on the host for now.  The hysteresis scan is vectorised (the
reference walks the frames in a Python loop).
"""
from __future__ import annotations

import numpy as np


def voiced_detection(signal, frame_size=64, energy_threshold=1e-4, smooth_window=5, low_threshold=None):
    """0/1 mask per sample: a frame switches on above `energy_threshold`, off
    below `low_threshold` (default half of it), and otherwise keeps the
    previous frame's state (fractal.py:880-909)."""
    x = np.asarray(signal, dtype=np.float32)
    n = len(x)
    frames = -(-n // frame_size)
    padded = np.pad(x, (0, frames * frame_size - n), mode="reflect")
    energy = np.mean(np.square(padded.reshape(frames, frame_size)), axis=1)
    if smooth_window > 1:
        energy = np.convolve(energy, np.ones(smooth_window, dtype=np.float32) / smooth_window, mode="same")
    low = energy_threshold * 0.5 if low_threshold is None else low_threshold
    rises = energy > energy_threshold
    falls = ~rises & (energy < low)
    # state[i] = rises[last frame <= i that rose or fell]; silent before the first one
    marker = np.where(rises | falls, np.arange(frames), -1)
    np.maximum.accumulate(marker, out=marker)
    state = rises[np.maximum(marker, 0)] & (marker >= 0)
    return np.repeat(state.astype(np.uint8), frame_size)[:n]


def frame_ranges(signal, range_size, energy_thresh):
    """Returns (ranges or None, original_len).  `ranges` is the masked signal,
    reflect-padded to a multiple of range_size and viewed as (n_ranges,
    range_size); None when the reference takes its empty early-out
    (fractal.py:1083 silent input, :1100 no ranges)."""
    mask = voiced_detection(signal, frame_size=2 * range_size, energy_threshold=energy_thresh)
    gated = signal * mask
    original_len = len(gated)
    if np.sum(gated ** 2) < 1e-8:
        return None, original_len
    tail = -original_len % range_size
    if tail:
        gated = np.pad(gated, (0, tail), mode="reflect")
    if len(gated) < range_size:
        return None, original_len
    return np.ascontiguousarray(gated.reshape(-1, range_size), dtype=np.float32), original_len
