#!/usr/bin/env python
"""Host-side rows of SURVEY.md §8(f) (N1 container, N2 pre-step, N3 match hand-off) at config-2 size,
timed next to the reference's own functions.  CPU only; needs /root/reference (build container), so it is a
measurement script, not part of bench.py, and it never runs on the GPU box.  The reference is imported unmodified
under a results-neutral `librosa` stub (its module-level import is never used on these paths).

    OMP_NUM_THREADS=1 python scripts/bench_host_rows.py [--scale 1.0] > profiles/r01_host_rows.json

Every pair of outputs is compared (bytes of the .fwav file, mask, framed ranges, loaded arrays) before a time
is reported.
"""
import argparse, hashlib, json, os, sys, tempfile, time

os.environ.setdefault("OMP_NUM_THREADS", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "audio-compression_b200"))
import numpy as np  # noqa: E402
from fwav_b200 import container, prestep, synth  # noqa: E402


def load_reference():
    import logging, types
    stub = types.ModuleType("librosa")
    stub.filters = types.ModuleType("librosa.filters")
    stub.filters.mel = lambda sr, n_fft, n_mels=40, fmin=0, fmax=None: np.zeros((n_mels, n_fft // 2 + 1), np.float32)
    sys.modules["librosa"], sys.modules["librosa.filters"] = stub, stub.filters
    sys.path.insert(0, "/root/reference")
    ours = sys.modules.pop("fractal", None)          # the drop-in has the same module name
    import fractal as ref
    sys.modules.pop("fractal", None)
    if ours is not None:
        sys.modules["fractal"] = ours
    logging.getLogger().setLevel(logging.WARNING)
    ref.logger.setLevel(logging.WARNING)
    return ref


def best_of(fn, reps):
    ts = []
    out = None
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        ts.append(time.perf_counter() - t0)
    return min(ts), out


def sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 24), b""):
            h.update(blk)
    return h.hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    ref = load_reference()

    rate, tile = 44100, 4096
    N, ds = max(4, tile // 256), max(1, max(4, tile // 256) // 4)
    sig = synth.music_like(seconds=180.0 * args.scale, rate=rate, seed=2).astype(np.float32)
    n = len(sig)
    n_d = (n - tile) // ds + 1
    rng = np.random.default_rng(7)
    res = {"workload": f"config-2 shape x{args.scale}: {n} samples, range_size {N}, {n_d} domains", "rows": {}}

    # ---- N2: voiced gate + masking + reflect pad + framing (fractal.py:880-909, :1074-1112) ----
    def ref_prestep():
        mask = ref.voiced_detection(sig, frame_size=2 * N, energy_threshold=1e-4)
        gated = sig * mask
        pad = (N - len(gated) % N) % N
        if pad:
            gated = np.pad(gated, (0, pad), mode="reflect")
        return mask, gated.reshape(-1, N)

    def our_prestep():
        return prestep.voiced_detection(sig, frame_size=2 * N, energy_threshold=1e-4), prestep.frame_ranges(sig, N, 1e-4)[0]

    t_ref, (m_ref, r_ref) = best_of(ref_prestep, 1)
    t_our, (m_our, r_our) = best_of(our_prestep, args.reps)
    assert np.array_equal(np.asarray(m_ref, dtype=np.uint8), m_our), "voiced mask differs"
    assert np.array_equal(r_ref.view(np.uint32), r_our.view(np.uint32)), "framed ranges differ"
    res["rows"]["N2_prestep"] = {"reference_s": t_ref, "ours_s": t_our, "speedup": t_ref / t_our,
                                 "check": "mask and framed ranges bit-identical",
                                 "note": "ours runs the gate twice here (mask alone + frame_ranges)"}
    n_r = len(r_our)

    # ---- synthetic matches + domains of the right shape ----
    domains = rng.standard_normal((n_d, N), dtype=np.float32) * 0.1
    idx = rng.integers(0, n_d, n_r).astype(np.int32)
    idx[rng.random(n_r) < 0.02] = -1
    s = rng.uniform(-1, 1, n_r).astype(np.float32)
    o = rng.uniform(-0.1, 0.1, n_r).astype(np.float32)
    sym = rng.integers(0, 2, n_r).astype(np.uint8)
    err = rng.uniform(0, 1, n_r).astype(np.float32)
    err[idx < 0] = np.inf
    arrays = container.MatchArrays(idx, s, o, sym, err)

    # ---- N3: match hand-off (list of tuples <-> arrays) ----
    t_list, tuples = best_of(arrays.tolist, args.reps)
    t_back, back = best_of(lambda: container.MatchArrays.from_any(tuples), args.reps)
    assert np.array_equal(back.idx, idx) and np.array_equal(back.s.view(np.uint32), s.view(np.uint32))
    res["rows"]["N3_handoff"] = {"arrays_to_tuple_list_s": t_list, "tuple_list_to_arrays_s": t_back,
                                 "n_matches": n_r,
                                 "note": "compress -> save and load -> decompress keep the arrays (as_arrays=True) and skip both"}

    # ---- N1: .fwav writer / reader (fractal.py:1278-1375) ----
    with tempfile.TemporaryDirectory() as tmp:
        p_ref, p_our, p_arr = (os.path.join(tmp, x) for x in ("ref.fwav", "ours.fwav", "ours_arrays.fwav"))
        meta = (N, rate, 2, tile, ds, 1e-4, n)
        t_wr_ref, _ = best_of(lambda: ref.save_compressed(p_ref, tuples, domains, *meta), 1)
        t_wr_our, _ = best_of(lambda: container.save_compressed(p_our, tuples, domains, *meta), args.reps)
        t_wr_arr, _ = best_of(lambda: container.save_compressed(p_arr, arrays, domains, *meta), args.reps)
        h = sha(p_ref)
        assert h == sha(p_our) == sha(p_arr), "container bytes differ"
        size = os.path.getsize(p_ref)
        t_rd_ref, out_ref = best_of(lambda: ref.load_compressed(p_ref), 1)
        t_rd_our, out_our = best_of(lambda: container.load_compressed(p_ref), args.reps)
        t_rd_arr, out_arr = best_of(lambda: container.load_compressed(p_ref, as_arrays=True), args.reps)
        assert list(out_ref[0]) == list(out_our[0]) and np.array_equal(np.asarray(out_ref[1]).view(np.uint32), out_our[1].view(np.uint32))
        assert tuple(out_ref[2:]) == tuple(out_our[2:])
        assert np.array_equal(out_arr[0].idx, idx)
    res["rows"]["N1_container"] = {
        "file_bytes": size, "sha256_equal": True,
        "write": {"reference_s": t_wr_ref, "ours_from_tuple_list_s": t_wr_our, "ours_from_arrays_s": t_wr_arr,
                  "speedup_same_input": t_wr_ref / t_wr_our, "speedup_arrays": t_wr_ref / t_wr_arr,
                  "ours_arrays_MBps": size / t_wr_arr / 1e6},
        "read": {"reference_s": t_rd_ref, "ours_to_tuple_list_s": t_rd_our, "ours_to_arrays_s": t_rd_arr,
                 "speedup_same_output": t_rd_ref / t_rd_our, "speedup_arrays": t_rd_ref / t_rd_arr,
                 "ours_arrays_MBps": size / t_rd_arr / 1e6},
        "check": "file bytes identical (SHA-256 of the whole file), loaded tuples / domains / header fields equal"}
    res["host"] = {"cpu_count": os.cpu_count(), "threads_used": 1}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
