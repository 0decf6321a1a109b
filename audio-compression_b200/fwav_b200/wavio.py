"""PCM WAV in/out with the reference's semantics (fractal.py:81-137): 8-bit
unsigned, 16/24-bit signed and 32-bit float, multi-channel averaged to mono,
samples returned at their integer scale (never normalised).

COMPATIBILITY TRANSLITERATION of the reference's two functions (SURVEY 2 marks WAV I/O "reuse semantics verbatim",
BASELINE.json: "read_wav_mono/write_wav ... stay unchanged"); host glue, not part of the hot path."""
from __future__ import annotations

import wave

import numpy as np


def read_wav_mono(path, mmap=False):
    with wave.open(path, "rb") as w:
        channels, width, rate = w.getnchannels(), w.getsampwidth(), w.getframerate()
        if w.getcomptype() != "NONE":
            raise ValueError(f"Unsupported WAV compression type: {w.getcomptype()}")
        raw = w.readframes(w.getnframes())
    if width == 1:
        pcm = np.frombuffer(raw, dtype=np.uint8).astype(np.int16) - 128
    elif width == 2:
        pcm = np.frombuffer(raw, dtype="<i2")
    elif width == 3:
        b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        pcm = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        pcm = pcm - ((pcm & 0x800000) << 1)
    elif width == 4:
        pcm = np.frombuffer(raw, dtype="<f4")
    else:
        raise ValueError(f"Unsupported sample width: {width}")
    if channels > 1:
        pcm = pcm.reshape(-1, channels).mean(axis=1)
    return pcm.astype(np.float32), rate, width


def write_wav(path, data, framerate, sampwidth):
    data = np.asarray(data)
    if sampwidth == 1:
        payload = (data + 128).clip(0, 255).astype(np.uint8)
    elif sampwidth == 2:
        payload = data.clip(-32768, 32767).astype("<i2")
    elif sampwidth == 3:
        v = data.clip(-2 ** 23, 2 ** 23 - 1).astype(np.int32)
        payload = np.stack([v & 0xFF, (v >> 8) & 0xFF, (v >> 16) & 0xFF], axis=1).astype(np.uint8).ravel()
    elif sampwidth == 4:
        payload = data.astype("<f4")
    else:
        raise ValueError(f"Unsupported sample width: {sampwidth}")
    with wave.open(path, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(sampwidth)
        w.setframerate(framerate)
        w.writeframes(payload.tobytes())
