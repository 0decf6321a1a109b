"""PCM WAV in/out with the reference's semantics (fractal.py:81-137): 8-bit
unsigned, 16/24-bit signed and 32-bit float, multi-channel averaged to mono,
samples returned at their integer scale (never normalised).

Restatement of the reference's two functions (SURVEY 2 marks WAV I/O "reuse semantics verbatim", BASELINE.json:
"read_wav_mono/write_wav ... stay unchanged"); host glue, not part of the hot path.  The 24-bit and stereo paths
are single-pass forms with the same results bit for bit (scripts/bench_wav_io.py times both
against the reference's own functions at config-3/4 sizes, profiles/r02_wav_io.json)."""
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
        # three little-endian bytes into the upper three of an int32, arithmetic shift down: the reference's
        # b0 | b1 << 8 | b2 << 16 followed by its sign extension, in one pass instead of five int32 temporaries
        # (row N4: a 30-minute 24-bit stereo file is 518 MB of PCM)
        n = len(raw) // 3
        q = np.empty((n, 4), np.uint8)
        q[:, 0] = 0
        q[:, 1:] = np.frombuffer(raw, dtype=np.uint8, count=3 * n).reshape(n, 3)
        pcm = q.view("<i4").ravel() >> 8
    elif width == 4:
        pcm = np.frombuffer(raw, dtype="<f4")
    else:
        raise ValueError(f"Unsupported sample width: {width}")
    if channels == 2 and pcm.dtype.kind == "i":
        # ndarray.mean over two integers accumulates in float64 and divides by 2: (a + b) / 2, exactly
        pair = pcm.reshape(-1, 2)
        return ((pair[:, 0].astype(np.float64) + pair[:, 1]) * 0.5).astype(np.float32), rate, width
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
        # the low three bytes of the little-endian int32 (the reference masks and shifts them out one by one)
        v = data.clip(-2 ** 23, 2 ** 23 - 1).astype("<i4")
        payload = np.ascontiguousarray(v.view(np.uint8).reshape(-1, 4)[:, :3])
    elif sampwidth == 4:
        payload = data.astype("<f4")
    else:
        raise ValueError(f"Unsupported sample width: {sampwidth}")
    with wave.open(path, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(sampwidth)
        w.setframerate(framerate)
        w.writeframes(payload.tobytes())
