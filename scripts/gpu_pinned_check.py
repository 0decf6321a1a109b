"""One-shot check of the page-locked result buffers of the Python API (no torch import):
pinned and pageable runs must return identical results; times fractal.compress_audio on config 2."""
import os, sys, time, json, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "audio-compression_b200"))
import numpy as np
import fractal
from fwav_b200 import synth
res = {}
t = np.arange(16000 * 2) / 16000.0
rng = np.random.default_rng(1234)
small = (0.5 * np.sin(2 * np.pi * 440 * t) + 0.05 * rng.standard_normal(len(t))).astype(np.float32)
outs = {}
for pinned in ("1", "0"):
    os.environ["FWAV_PINNED"] = pinned
    outs[pinned] = fractal.compress_audio_arrays(small, tile_size=1024)
a, b = outs["1"], outs["0"]
assert np.array_equal(a[0].idx, b[0].idx) and np.array_equal(a[0].s.view(np.uint32), b[0].s.view(np.uint32))
assert np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32)) and a[2:] == b[2:]
with tempfile.TemporaryDirectory() as tmp:
    p = os.path.join(tmp, "x.fwav")
    fractal.save_compressed(p, a[0], a[1], a[3], 16000, 2, a[4], a[5], a[6], a[7])
    m, d, n_r, rs, *_rest = fractal.load_compressed(p)
    rec = fractal.decompress_audio(m, d, n_r, rs, iterations=8, original_len=len(small))
res["small_snr_db"] = float(fractal.compute_snr(small, rec))
res["pinned_equals_pageable"] = True
del outs, a, b
sig = synth.music_like(seconds=180.0, rate=44100, seed=2)
for pinned in ("1", "0"):
    os.environ["FWAV_PINNED"] = pinned
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        out = fractal.compress_audio_arrays(sig, tile_size=4096)
        ts.append((time.perf_counter() - t0) * 1e3)
    t0 = time.perf_counter()
    out = fractal.compress_audio(sig, 44100, 2, tile_size=4096)
    res["pinned" if pinned == "1" else "pageable"] = {"compress_audio_arrays_ms": [round(x, 1) for x in ts],
                                                       "compress_audio_ms": round((time.perf_counter() - t0) * 1e3, 1)}
    del out
print(json.dumps(res))
