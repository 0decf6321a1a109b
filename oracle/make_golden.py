"""Generate tests/golden/*.npz by EXECUTING the reference (TEST INFRASTRUCTURE).

Run in the build container only (needs /root/reference):

    OMP_NUM_THREADS=1 python oracle/make_golden.py

The reference is imported unmodified; the one blocker, a module-level
`import librosa` whose result is never consumed on the live path (fractal.py:488,
:1210-1213, :648), is satisfied with a stub module.  Two ways of driving it are
recorded:

  * "pipeline": the real multi-process compress_audio / save_compressed /
    load_compressed / decompress_audio calls;
  * "replay":   the same reference functions called in one process
    (build_domains_memmap -> build_domain_embeddings ->
    range_candidates_from_embedding_emb -> pad_candidates -> _process_gpu_batch),
    which also yields the intermediates (domains, embeddings, candidate table).

The fixtures hold inputs and reference outputs only; nothing from the
reference's source is copied.
"""
from __future__ import annotations

import os
import sys
import tempfile
import types

os.environ.setdefault("OMP_NUM_THREADS", "1")
import numpy as np  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, os.path.join(ROOT, "audio-compression_b200"))


def load_reference():
    stub = types.ModuleType("librosa")
    stub.filters = types.ModuleType("librosa.filters")
    stub.filters.mel = lambda sr, n_fft, n_mels=40, fmin=0, fmax=None: \
        np.zeros((n_mels, n_fft // 2 + 1), np.float32)
    sys.modules["librosa"] = stub
    sys.modules["librosa.filters"] = stub.filters
    sys.path.insert(0, "/root/reference")
    import fractal as ref
    import logging
    logging.getLogger().setLevel(logging.WARNING)
    ref.logger.setLevel(logging.WARNING)
    return ref


class _Sink:
    def __init__(self):
        self.items = []

    def put(self, item):
        self.items.append(item)


def replay(ref, signal, tile_size, top_k=32, emb_dim=16, energy_thresh=1e-4, fast_mode=True):
    """Single-process walk through the reference's own functions."""
    signal = np.asarray(signal, dtype=np.float32)
    range_size = max(4, tile_size // 256)
    step = max(1, range_size // 4)
    mask = ref.voiced_detection(signal, frame_size=range_size * 2, energy_threshold=energy_thresh)
    weighted = signal * mask
    original_len = len(weighted)
    pad = (range_size - (original_len % range_size)) % range_size
    if pad:
        weighted = np.pad(weighted, (0, pad), mode="reflect")
    n_ranges = len(weighted) // range_size
    ranges = weighted.reshape(n_ranges, range_size)
    with tempfile.TemporaryDirectory() as tmp:
        dpath, n_domains = ref.build_domains_memmap(signal, tile_size, range_size, step,
                                                    block_size=500, tmpdir=tmp)
        domains = np.array(np.memmap(dpath, dtype="float32", mode="r",
                                     shape=(n_domains, range_size)))
        epath = ref.build_domain_embeddings(dpath, n_domains, range_size, emb_dim=emb_dim,
                                            block_size=4096, tmpdir=tmp)
        embs = np.array(np.memmap(epath, dtype="float32", mode="r", shape=(n_domains, emb_dim)))
    cand = np.empty((n_ranges, top_k), dtype=np.int32)
    for i in range(n_ranges):
        if fast_mode and np.mean(ranges[i] ** 2) < energy_thresh * 0.75:
            c = np.empty(0, dtype=np.int32)
        else:
            c = ref.range_candidates_from_embedding_emb(embs[i], embs, top_k=top_k)
        cand[i] = ref.pad_candidates(c, top_k)
    sink = _Sink()
    for lo in range(0, n_ranges, 512):
        ids = np.arange(lo, min(lo + 512, n_ranges), dtype=np.int32)
        ref._process_gpu_batch(ids, ranges[ids], cand[ids], domains, sink, use_gpu=False)
    res = dict(sorted(sink.items))
    m = [res[i] for i in range(n_ranges)]
    return dict(
        signal=signal, tile_size=tile_size, top_k=top_k, emb_dim=emb_dim,
        energy_thresh=np.float64(energy_thresh), range_size=range_size, domain_step=step,
        original_len=original_len, ranges=ranges, domains=domains, embeddings=embs,
        candidates=cand,
        idx=np.array([t[0] for t in m], np.int32), s=np.array([t[1] for t in m], np.float32),
        o=np.array([t[2] for t in m], np.float32), sym=np.array([t[3] for t in m], np.uint8),
        err=np.array([t[4] for t in m], np.float32),
    )


def add_decodes(ref, g, variants):
    m = list(zip(g["idx"].tolist(), g["s"].tolist(), g["o"].tolist(), g["sym"].tolist(),
                 g["err"].tolist()))
    n_ranges = len(m)
    for tag, kw in variants.items():
        out = ref.decompress_audio(m, g["domains"], n_ranges, int(g["range_size"]),
                                   original_len=int(g["original_len"]), **kw)
        g["dec_" + tag] = np.asarray(out, dtype=np.float32)


DECODES = {
    "default": dict(iterations=8, convergence_eps=1e-3),
    "damp50": dict(iterations=8, convergence_eps=0.0, s_damping=0.5),
    "damp25_clip2": dict(iterations=5, convergence_eps=1e-3, s_damping=0.25, s_clip=2.0),
}


def main():
    from fwav_b200 import synth
    ref = load_reference()
    os.makedirs(GOLD, exist_ok=True)

    # ---- 1. the reference's own test fixture through the real pipeline ----
    sig, sr, sw = synth.test_tone()
    out = ref.compress_audio(sig, sr, sw, tile_size=128, energy_thresh=1e-4, use_gpu=False,
                             fast_mode=True, cpu_workers=2)
    matches, domains, n_ranges, range_size, tile_size, step, thr, orig = out
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "tone.fwav")
        ref.save_compressed(path, matches, domains, range_size, sr, sw, tile_size, step, thr, len(sig))
        blob = np.frombuffer(open(path, "rb").read(), dtype=np.uint8)
        back = ref.load_compressed(path)
    assert back[0] == matches
    rec = np.asarray(ref.decompress_audio(back[0], back[1], back[2], back[3], iterations=8,
                                          convergence_eps=1e-3, original_len=back[9]))
    g = replay(ref, sig, 128)
    assert [tuple(t) for t in matches] == list(zip(g["idx"].tolist(), g["s"].tolist(), g["o"].tolist(),
                                                    g["sym"].tolist(), g["err"].tolist())), \
        "single-process replay must reproduce the multi-process pipeline"
    assert np.array_equal(np.asarray(domains), g["domains"])
    g.update(framerate=sr, sampwidth=sw, fwav_bytes=blob, pipeline_decode=rec,
             snr=np.float64(ref.compute_snr(sig, rec)))
    add_decodes(ref, g, DECODES)
    np.savez_compressed(os.path.join(GOLD, "tone128.npz"), **g)
    print("tone128: ranges", n_ranges, "domains", len(domains), "snr", float(g["snr"]))

    # ---- 2. config-1 style, N=4 / ds=1, through the real pipeline + replay ----
    sig = synth.sine_noise(seconds=0.4, rate=16000, seed=1234)
    out = ref.compress_audio(sig, 16000, 2, tile_size=1024, use_gpu=False, cpu_workers=2)
    g = replay(ref, sig, 1024)
    assert [tuple(t) for t in out[0]] == list(zip(g["idx"].tolist(), g["s"].tolist(), g["o"].tolist(),
                                                   g["sym"].tolist(), g["err"].tolist()))
    add_decodes(ref, g, DECODES)
    np.savez_compressed(os.path.join(GOLD, "sine_t1024.npz"), **g)
    print("sine_t1024: ranges", len(g["idx"]), "domains", len(g["domains"]))

    # ---- 3. config-2 style, N=16 / ds=4 ----
    sig = synth.music_like(seconds=1.0, rate=44100, seed=2)
    g = replay(ref, sig, 4096)
    add_decodes(ref, g, DECODES)
    np.savez_compressed(os.path.join(GOLD, "music_t4096.npz"), **g)
    print("music_t4096: ranges", len(g["idx"]), "domains", len(g["domains"]))

    # ---- 4. gaps of digital silence: voiced gate + energy prune (F6) ----
    sig = synth.sine_noise(seconds=0.5, rate=16000, seed=7)
    sig[1500:3500] = 0.0
    sig[6000:6400] = 0.0
    g = replay(ref, sig, 1024)
    assert np.isinf(g["err"]).any()
    add_decodes(ref, g, DECODES)
    np.savez_compressed(os.path.join(GOLD, "gaps_t1024.npz"), **g)
    print("gaps_t1024: pruned", int(np.isinf(g["err"]).sum()), "of", len(g["idx"]))

    # ---- 5. float-scale input (+-1): most ranges fall under the prune threshold ----
    rng = np.random.default_rng(5)
    t = np.arange(6000) / 16000.0
    sig = (0.02 * np.sin(2 * np.pi * 300 * t) * (1 + np.sin(2 * np.pi * 3 * t))
           + 0.001 * rng.standard_normal(len(t))).astype(np.float32)
    g = replay(ref, sig, 1024)
    add_decodes(ref, g, DECODES)
    np.savez_compressed(os.path.join(GOLD, "float_t1024.npz"), **g)
    print("float_t1024: pruned", int(np.isinf(g["err"]).sum()), "of", len(g["idx"]))

    # ---- 6. top_k = 64 (config 4 style: module global, F7) and K >= n_domains ----
    sig = synth.music_like(seconds=0.25, rate=16000, seed=4)
    g = replay(ref, sig, 1024, top_k=64)
    np.savez_compressed(os.path.join(GOLD, "music_k64.npz"), **g)
    sig, _, _ = synth.test_tone(dur=0.0215)          # 172 samples -> 43 ranges, 45 domains
    g = replay(ref, sig, 128, top_k=64)
    assert (g["candidates"] == -1).any()
    np.savez_compressed(os.path.join(GOLD, "tiny_kfull.npz"), **g)
    print("tiny_kfull: ranges", len(g["idx"]), "domains", len(g["domains"]))

    # ---- 7. odd geometry: tile not a multiple of 256 (run length 275), N=4 ----
    sig = synth.sine_noise(seconds=0.3, rate=16000, seed=11)
    g = replay(ref, sig, 1100)
    np.savez_compressed(os.path.join(GOLD, "sine_t1100.npz"), **g)
    # and a larger odd one: tile 3000 -> N=11, ds=2, run 272
    sig = synth.music_like(seconds=0.5, rate=22050, seed=12)
    g = replay(ref, sig, 3000)
    add_decodes(ref, g, {"default": DECODES["default"]})
    np.savez_compressed(os.path.join(GOLD, "music_t3000.npz"), **g)
    print("music_t3000: N", int(g["range_size"]), "ds", int(g["domain_step"]))

    # ---- 8. decoder with legacy -1 sentinels (fractal.py:1399-1426) ----
    g = dict(np.load(os.path.join(GOLD, "sine_t1024.npz")))
    idx = g["idx"].copy()
    idx[::7] = -1
    m = list(zip(idx.tolist(), g["s"].tolist(), g["o"].tolist(), g["sym"].tolist(), g["err"].tolist()))
    outs = {}
    for tag, kw in DECODES.items():
        outs["dec_" + tag] = np.asarray(ref.decompress_audio(
            m, g["domains"], len(m), int(g["range_size"]), original_len=int(g["original_len"]), **kw),
            dtype=np.float32)
    np.savez_compressed(os.path.join(GOLD, "sentinel_decode.npz"), idx=idx, s=g["s"], o=g["o"],
                        sym=g["sym"], domains=g["domains"], range_size=g["range_size"],
                        original_len=g["original_len"], **outs)

    # ---- 9. voiced gate on its own ----
    sig = synth.sine_noise(seconds=0.5, rate=16000, seed=9)
    sig[2000:2600] *= 1e-4
    sig[5000:5050] = 0
    vm = ref.voiced_detection(sig, frame_size=8, energy_threshold=1e-4)
    vm2 = ref.voiced_detection(sig * 1e-4, frame_size=32, energy_threshold=1e-4)
    np.savez_compressed(os.path.join(GOLD, "voiced.npz"), signal=sig, mask_f8=vm, mask_scaled_f32=vm2)

    total = sum(os.path.getsize(os.path.join(GOLD, f)) for f in os.listdir(GOLD))
    print("golden bytes:", total)


def main_t2048():
    """Added after the first batch (the other fixtures are left untouched): tile 2048 -> N=8, ds=2, the one
    constant-bank geometry (4 / 8 / 16) the first batch did not cover."""
    from fwav_b200 import synth
    ref = load_reference()
    sig = synth.music_like(seconds=0.6, rate=22050, seed=21)
    g = replay(ref, sig, 2048)
    assert int(g["range_size"]) == 8 and int(g["domain_step"]) == 2
    add_decodes(ref, g, DECODES)
    np.savez_compressed(os.path.join(GOLD, "music_t2048.npz"), **g)
    print("music_t2048: ranges", len(g["idx"]), "domains", len(g["domains"]))


def main_tile_embedding():
    """Round 2: the reference's tile_embedding(x, k=EMBED_K=32) (fractal.py:178-208) on its own -- the embedding the
    README / north star describe, not on the live path -- for rows of 4, 8, 16 and 40 samples."""
    ref = load_reference()
    rng = np.random.default_rng(32)
    out = {}
    for n in (4, 8, 16, 40):
        rows = (rng.standard_normal((64, n)) * 10.0 ** rng.uniform(-2, 4, (64, 1))).astype(np.float32)
        rows[0] = 0.0                      # norm below 1e-8: left unnormalised
        rows[1] = 123.0                    # constant tile: DC only, every kept coefficient ~0
        out[f"rows_{n}"] = rows
        out[f"emb_{n}"] = np.stack([ref.tile_embedding(r, k=ref.EMBED_K) for r in rows]).astype(np.float32)
    np.savez_compressed(os.path.join(GOLD, "tile_embedding_k32.npz"), **out)
    print("tile_embedding_k32:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "t2048":
        main_t2048()
    elif len(sys.argv) > 1 and sys.argv[1] == "tile_embedding":
        main_tile_embedding()
    else:
        main()
