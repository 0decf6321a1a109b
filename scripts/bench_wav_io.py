#!/usr/bin/env python
"""WAV input/output at hour scale, timed next to the reference's own read_wav_mono /
write_wav on the same files (CPU only; needs /root/reference, so it runs in the build container, never on the GPU
box).  Outputs are compared bit for bit before a time is reported.

    python scripts/bench_wav_io.py [--scale 1.0] > profiles/r02_wav_io.json

  c4 input : 30 min, 48 kHz, 24-bit STEREO  (518 MB of PCM)  -> read_wav_mono
  c3 input : 1 h,   48 kHz, 16-bit mono     (346 MB)         -> read_wav_mono
  c5 output: 1 h,   48 kHz, 16-bit and 24-bit mono           -> write_wav
"""
import argparse, json, os, sys, tempfile, time, wave

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "audio-compression_b200"))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import numpy as np  # noqa: E402
from fwav_b200 import wavio  # noqa: E402
from bench_host_rows import load_reference, best_of  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()
    ref = load_reference()
    rng = np.random.default_rng(4)
    out = {"scale": args.scale, "host": "build container, one thread", "rows": {}}
    d = tempfile.mkdtemp(dir=os.environ.get("TMPDIR", "/tmp"))

    def bits(a):
        return np.ascontiguousarray(a).view(np.uint32)

    # c4: 24-bit stereo, 30 min at 48 kHz
    n = int(1800 * 48000 * args.scale)
    v = rng.integers(-2 ** 23, 2 ** 23, size=2 * n, dtype=np.int32)
    p = os.path.join(d, "c4.wav")
    with wave.open(p, "wb") as w:
        w.setnchannels(2); w.setsampwidth(3); w.setframerate(48000)
        w.writeframes(np.ascontiguousarray(v.astype("<i4").view(np.uint8).reshape(-1, 4)[:, :3]).tobytes())
    del v
    t_ref, a = best_of(lambda: ref.read_wav_mono(p), args.reps)
    t_new, b = best_of(lambda: wavio.read_wav_mono(p), args.reps)
    assert a[1:] == b[1:] and np.array_equal(bits(a[0]), bits(b[0]))
    out["rows"]["read 24-bit stereo (c4 input)"] = {"frames": n, "pcm_mb": 6 * n / 1e6, "reference_s": t_ref, "here_s": t_new,
                                                    "identical": True}
    os.remove(p)
    # c3: 16-bit mono, 1 h at 48 kHz
    n = int(3600 * 48000 * args.scale)
    sig = (rng.standard_normal(n) * 8000).astype(np.float32)
    p = os.path.join(d, "c3.wav")
    ref.write_wav(p, sig, 48000, 2)
    t_ref, a = best_of(lambda: ref.read_wav_mono(p), args.reps)
    t_new, b = best_of(lambda: wavio.read_wav_mono(p), args.reps)
    assert np.array_equal(bits(a[0]), bits(b[0]))
    out["rows"]["read 16-bit mono (c3 input)"] = {"frames": n, "pcm_mb": 2 * n / 1e6, "reference_s": t_ref, "here_s": t_new,
                                                  "identical": True}
    # c5: decoded hour back to a file
    for width in (2, 3):
        pr, pn = os.path.join(d, "r.wav"), os.path.join(d, "n.wav")
        scale = 1.0 if width == 2 else 300.0
        t_ref, _ = best_of(lambda: ref.write_wav(pr, sig * scale, 48000, width), args.reps)
        t_new, _ = best_of(lambda: wavio.write_wav(pn, sig * scale, 48000, width), args.reps)
        same = open(pr, "rb").read() == open(pn, "rb").read()
        assert same
        out["rows"][f"write {8 * width}-bit mono (c5 output)"] = {"frames": n, "pcm_mb": width * n / 1e6, "reference_s": t_ref,
                                                                 "here_s": t_new, "identical": same}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
