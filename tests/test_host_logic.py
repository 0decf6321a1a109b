"""Host side of the drop-in (no GPU): pre-step, container, WAV I/O, match
arrays, and that the C-ABI library loads and exports every declared symbol."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, golden


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def test_library_exports_every_declared_symbol():
    from fwav_b200 import _lib
    header = open(os.path.join(ROOT, "include", "fwav_b200.h")).read()
    declared = set(re.findall(r"\b(fwav_[a-z0-9_]+)\s*\(", header))
    declared -= {"fwav_ctx"}
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == {n for n, _, _ in _lib.SIGNATURES}
    typed = _lib.load_library()
    assert b"sm_100a" in typed.fwav_version()
    # geometry helpers are host-only and must agree with the reference's derivation
    for tile, want in [(128, (4, 1)), (1024, (4, 1)), (2048, (8, 2)), (4096, (16, 4)), (8192, (32, 8)), (3000, (11, 2))]:
        assert _lib.geometry(tile) == want
    assert _lib.count_domains(160000, 1024, 1) == 158977
    assert _lib.count_domains(100, 1024, 1) == 0


def test_no_cpu_fallback_without_a_device():
    from fwav_b200 import _lib
    lib = _lib.load_library()
    if lib.fwav_device_count() > 0:
        pytest.skip("a GPU is present")
    import fractal
    sig = golden("tone128")["signal"]
    with pytest.raises(_lib.FwavError):
        fractal.compress_audio(sig, 8000, 2, tile_size=128)
    with pytest.raises(_lib.FwavError):
        fractal.decompress_audio([(0, 1.0, 0.0, 0, 0.0)], np.zeros((1, 4), np.float32), 1, 4)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "audio-compression_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "fwav_oracle" not in text and "import oracle" not in text and "from oracle" not in text, f


@pytest.mark.parametrize("name", ["tone128", "sine_t1024", "music_t4096", "gaps_t1024", "float_t1024", "music_t3000", "music_t2048"])
def test_prestep_matches_reference(name):
    from fwav_b200.prestep import frame_ranges
    g = golden(name)
    ranges, original_len = frame_ranges(g["signal"], int(g["range_size"]), float(g["energy_thresh"]))
    assert original_len == int(g["original_len"])
    assert np.array_equal(bits(ranges), bits(g["ranges"]))


def test_voiced_gate_and_early_outs():
    from fwav_b200.prestep import frame_ranges, voiced_detection
    g = golden("voiced")
    assert np.array_equal(voiced_detection(g["signal"], 8, 1e-4), g["mask_f8"])
    assert np.array_equal(voiced_detection(g["signal"] * 1e-4, 32, 1e-4), g["mask_scaled_f32"])
    r, n = frame_ranges(np.zeros(1000, np.float32), 4, 1e-4)
    assert r is None and n == 1000
    import fractal
    out = fractal.compress_audio(np.zeros(5000, np.float32), 16000, 2, tile_size=1024)
    assert out[0] == [] and out[1].shape == (0, 4) and out[2] == 0 and out[3:] == (4, 1024, 1, 1e-4, 5000)
    out = fractal.compress_audio(np.full(500, 1000, np.float32), 16000, 2, tile_size=1024)   # shorter than a tile
    assert out[0] == [] and out[2] == 0 and out[7] == 500
    assert fractal.decompress_audio([], np.zeros((0, 4), np.float32), 0, 4).shape == (0,)


def test_container_bytes_identical_to_reference(tmp_path):
    import fractal
    g = golden("tone128")
    matches = list(zip(g["idx"].tolist(), g["s"].tolist(), g["o"].tolist(), g["sym"].tolist(), g["err"].tolist()))
    path = str(tmp_path / "a.fwav")
    fractal.save_compressed(path, matches, g["domains"], int(g["range_size"]), int(g["framerate"]),
                            int(g["sampwidth"]), int(g["tile_size"]), int(g["domain_step"]),
                            float(g["energy_thresh"]), int(g["original_len"]))
    assert open(path, "rb").read() == g["fwav_bytes"].tobytes()
    # same bytes from the array fast path
    path2 = str(tmp_path / "b.fwav")
    fractal.save_compressed(path2, fractal.MatchArrays(g["idx"], g["s"], g["o"], g["sym"], g["err"]), g["domains"],
                            int(g["range_size"]), int(g["framerate"]), int(g["sampwidth"]), int(g["tile_size"]),
                            int(g["domain_step"]), float(g["energy_thresh"]), int(g["original_len"]))
    assert open(path2, "rb").read() == g["fwav_bytes"].tobytes()
    back = fractal.load_compressed(path)
    assert len(back) == 10 and back[0] == matches and isinstance(back[0][0][0], int)
    assert np.array_equal(bits(back[1]), bits(g["domains"]))
    assert back[2:] == (len(matches), 4, 8000, 2, 128, 1, pytest.approx(1e-4), int(g["original_len"]))
    raw = bytearray(open(path, "rb").read())
    raw[100] ^= 0xFF
    open(path, "wb").write(bytes(raw))
    with pytest.raises(ValueError, match="Checksum"):
        fractal.load_compressed(path)
    assert fractal.load_compressed(path, verify_checksum=False)[2] == len(matches)
    open(path, "wb").write(b"RIFF" + bytes(raw[4:]))
    with pytest.raises(ValueError, match="Not a FWAV"):
        fractal.load_compressed(path)
    raw[4] = 9
    open(path, "wb").write(b"FWAV" + bytes(raw[4:]))
    with pytest.raises(ValueError, match="version"):
        fractal.load_compressed(path)


def test_container_keeps_inf_and_sentinels(tmp_path):
    import fractal
    g = golden("gaps_t1024")
    m = fractal.MatchArrays(g["idx"], g["s"], g["o"], g["sym"], g["err"])
    assert np.isinf(m.err).any()
    path = str(tmp_path / "g.fwav")
    fractal.save_compressed(path, m, g["domains"], 4, 16000, 2, 1024, 1, 1e-4, int(g["original_len"]))
    back = fractal.load_compressed(path, as_arrays=True)[0]
    assert np.array_equal(back.idx, m.idx) and np.array_equal(bits(back.err), bits(m.err))
    assert m[3] == m.tolist()[3] and len(m) == len(g["idx"]) and list(m)[:2] == m.tolist()[:2]


@pytest.mark.parametrize("width", [1, 2, 3, 4])
def test_wav_roundtrip(tmp_path, width):
    import wave
    import fractal
    rng = np.random.default_rng(width)
    if width == 4:
        x = rng.uniform(-1, 1, 500).astype(np.float32)
    else:
        top = {1: 127, 2: 32767, 3: 2 ** 23 - 1}[width]
        x = rng.integers(-top - 1, top + 1, 500).astype(np.float32)
    p = str(tmp_path / "x.wav")
    fractal.write_wav(p, x, 22050, width)
    y, rate, w = fractal.read_wav_mono(p)
    assert rate == 22050 and w == width and np.array_equal(x, y)
    # stereo is averaged to mono
    if width == 2:
        st = np.stack([x, -x + 2], axis=1).astype("<i2")
        with wave.open(p, "wb") as f:
            f.setnchannels(2); f.setsampwidth(2); f.setframerate(8000); f.writeframes(st.tobytes())
        y, _, _ = fractal.read_wav_mono(p)
        assert np.allclose(y, 1.0)


def _plain_read(path):
    """the reference's decode, step by step (fractal.py:81-112)"""
    import wave
    with wave.open(path, "rb") as w:
        ch, width, raw = w.getnchannels(), w.getsampwidth(), w.readframes(w.getnframes())
    if width == 1:
        d = np.frombuffer(raw, np.uint8).astype(np.int16) - 128
    elif width == 2:
        d = np.frombuffer(raw, np.int16)
    elif width == 3:
        b = np.frombuffer(raw, np.uint8).reshape(-1, 3)
        d = b[:, 0].astype(np.int32) | (b[:, 1].astype(np.int32) << 8) | (b[:, 2].astype(np.int32) << 16)
        d = d - ((d & 0x800000) << 1)
    else:
        d = np.frombuffer(raw, np.float32)
    if ch > 1:
        d = d.reshape(-1, ch).mean(axis=1)
    return d.astype(np.float32)


@pytest.mark.parametrize("width, channels", [(1, 1), (1, 2), (2, 1), (2, 2), (2, 3), (3, 1), (3, 2), (3, 4), (4, 1), (4, 2)])
def test_wav_single_pass_forms_equal_the_plain_decode(tmp_path, width, channels):
    """24-bit and stereo files (config 4's input) through the single-pass forms of wavio: the same float32 bits as
    the reference's byte-by-byte decode, extremes and odd sums (x.5 means) included"""
    import wave
    import fractal
    rng = np.random.default_rng(10 * width + channels)
    n = 4001
    if width == 4:
        raw = rng.standard_normal((n, channels)).astype("<f4").tobytes()
    elif width == 1:
        raw = rng.integers(0, 256, (n, channels), dtype=np.uint8).tobytes()
    else:
        top = 2 ** (8 * width - 1)
        v = rng.integers(-top, top, (n, channels), dtype=np.int64)
        v[:4] = [[-top] * channels, [top - 1] * channels, [0] * channels, [-1] * channels]
        raw = b"".join(int(x).to_bytes(width, "little", signed=True) for x in v.ravel())
    p = str(tmp_path / "x.wav")
    with wave.open(p, "wb") as f:
        f.setnchannels(channels); f.setsampwidth(width); f.setframerate(48000); f.writeframes(raw)
    got, rate, w = fractal.read_wav_mono(p)
    want = _plain_read(p)
    assert rate == 48000 and w == width and got.dtype == np.float32
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    if channels == 1 and width == 3:
        # write_wav's 24-bit packing against the reference's mask-and-shift form
        fractal.write_wav(p, got * 1.5, 48000, 3)           # clips at both ends
        d32 = (got * 1.5).clip(-2 ** 23, 2 ** 23 - 1).astype(np.int32)
        ref = np.column_stack([(d32 & 0xFF).astype(np.uint8), ((d32 >> 8) & 0xFF).astype(np.uint8),
                               ((d32 >> 16) & 0xFF).astype(np.uint8)]).flatten().tobytes()
        with wave.open(p, "rb") as f:
            assert f.readframes(f.getnframes()) == ref


def test_api_surface_matches_reference():
    import inspect
    import fractal
    sig = inspect.signature(fractal.compress_audio)
    assert list(sig.parameters) == ["signal", "framerate", "sampwidth", "tile_size", "emb_dim", "top_k", "ef_search",
                                    "use_gpu", "energy_thresh", "domains_tmpdir", "batch_size_gpu", "batch_size_cpu",
                                    "fast_mode", "transient_weight", "n_mels", "cpu_workers"]
    d = {k: v.default for k, v in sig.parameters.items()}
    assert (d["tile_size"], d["emb_dim"], d["top_k"], d["energy_thresh"], d["batch_size_gpu"]) == (1024, 16, 32, 1e-4, 512)
    sig = inspect.signature(fractal.decompress_audio)
    assert list(sig.parameters) == ["matches", "domains_array", "n_ranges", "range_size", "iterations",
                                    "convergence_eps", "use_gpu", "original_len", "s_clip", "s_damping"]
    for name in ("save_compressed", "load_compressed", "read_wav_mono", "write_wav", "compute_snr",
                 "process_file_compress", "process_file_decompress", "main", "voiced_detection"):
        assert callable(getattr(fractal, name))
    assert (fractal.top_k, fractal.EMBED_K, fractal.FWAV_VERSION) == (32, 32, 1)


def test_experimental_collect_protocol_simulation():
    """The mbarrier protocol of the (not yet device-tested) four-buffer collect kernel, as modelled in
    scripts/sim/quad_protocol_sim.py: random schedules must finish without deadlock or early overwrite."""
    import importlib.util
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts", "sim", "quad_protocol_sim.py")
    spec = importlib.util.spec_from_file_location("quad_protocol_sim", path)
    sim = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sim)
    for n_visit in (2, 6, 34, 70):
        for seed in range(4):
            sim.run(n_visit, seed)


def _select_rounds(lists, want):
    """finalize_kernel's selection, lane by lane (topk_umma.cu): every round emits the heads that beat the largest
    SECOND key of any lane, ranked among themselves; returns (selected keys in output order, first, last, rounds)"""
    lists = [sorted(l, reverse=True) + [0, 0] for l in lists]          # 0 = empty slot
    out, n_sel, first, last, rounds = [None] * want, 0, None, None, 0
    while n_sel < want:
        heads = [l[0] for l in lists]
        s2 = max(l[1] for l in lists)
        emit = [h > s2 for h in heads]
        if not any(emit):
            break
        rounds += 1
        e = sum(emit)
        rank = [sum(1 for m in range(32) if emit[m] and heads[m] > heads[l]) for l in range(32)]
        for l in range(32):
            if emit[l] and n_sel + rank[l] < want:
                out[n_sel + rank[l]] = heads[l]
        if n_sel == 0:
            first = next(heads[l] for l in range(32) if emit[l] and rank[l] == 0)
        if n_sel + e >= want:
            last = next(heads[l] for l in range(32) if emit[l] and n_sel + rank[l] == want - 1)
        for l in range(32):
            if emit[l]:
                lists[l].pop(0)
        n_sel = min(n_sel + e, want)
    return out[:n_sel] if n_sel < want else out, first, last, rounds, n_sel


@pytest.mark.parametrize("want", [32, 64])
def test_selection_by_heads_above_every_second_key_is_a_sort(want):
    """The verification pass selects several keys per round: any distribution of unique keys over the
    32 lanes must come out as the `want` largest in descending order, with the first and the want-th reported --
    random spreads, everything in one lane, fewer keys than wanted, exactly `want` keys."""
    rng = np.random.default_rng(want)
    cases = []
    for n in (5, want - 1, want, want + 1, 200, 256, 320):
        keys = (rng.permutation(10 ** 6)[:n] + 1).tolist()
        cases.append([keys[l::32] for l in range(32)])                               # the kernel's interleaving
        skew = [[] for _ in range(32)]
        for i, k in enumerate(sorted(keys, reverse=True)):                           # the largest keys crowd a few lanes
            skew[(i // 10) % 32].append(k)
        cases.append(skew)
    one = [[] for _ in range(32)]
    one[7] = list(range(1, 11))
    cases.append(one)
    total_rounds = []
    for lists in cases:
        assert all(len(l) <= 10 for l in lists)                                      # kFinRegs
        keys = sorted((k for l in lists for k in l), reverse=True)
        out, first, last, rounds, n_sel = _select_rounds(lists, want)
        assert n_sel == min(want, len(keys)) and out == keys[:n_sel]
        assert first == keys[0]
        if len(keys) >= want:
            assert last == keys[want - 1]
        total_rounds.append(rounds)
    assert max(total_rounds) <= want                                                 # never worse than one key per round
