"""Pins oracle/fwav_oracle.py to outputs of the reference itself (tests/golden,
made by oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import golden
from oracle import fwav_oracle as O

REPLAYS = ["tone128", "sine_t1024", "music_t4096", "gaps_t1024", "float_t1024",
           "music_k64", "tiny_kfull", "sine_t1100", "music_t3000", "music_t2048"]


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.mark.parametrize("name", REPLAYS)
def test_compress_replay_bit_exact(name):
    g = golden(name)
    res = O.compress(g["signal"], tile_size=int(g["tile_size"]), emb_dim=int(g["emb_dim"]),
                     top_k=int(g["top_k"]), energy_thresh=float(g["energy_thresh"]),
                     want_intermediates=True)
    assert res["range_size"] == int(g["range_size"]) and res["domain_step"] == int(g["domain_step"])
    assert res["original_len"] == int(g["original_len"])
    assert np.array_equal(bits(res["ranges"]), bits(g["ranges"]))
    assert np.array_equal(bits(res["domains"]), bits(g["domains"]))
    assert np.array_equal(bits(res["embeddings"]), bits(g["embeddings"]))
    assert np.array_equal(res["candidates"], g["candidates"])
    assert np.array_equal(res["idx"], g["idx"])
    assert np.array_equal(res["sym"], g["sym"])
    for k in ("s", "o", "err"):
        assert np.array_equal(bits(res[k]), bits(g[k])), k


@pytest.mark.parametrize("name", ["tone128", "sine_t1024", "music_t4096", "gaps_t1024",
                                  "float_t1024", "sentinel_decode", "music_t2048"])
def test_decode_bit_exact(name):
    g = golden(name)
    n_ranges = len(g["idx"])
    variants = {
        "default": dict(iterations=8, convergence_eps=1e-3),
        "damp50": dict(iterations=8, convergence_eps=0.0, s_damping=0.5),
        "damp25_clip2": dict(iterations=5, convergence_eps=1e-3, s_damping=0.25, s_clip=2.0),
    }
    for tag, kw in variants.items():
        out = O.decode(g["idx"], g["s"], g["o"], g["sym"], g["domains"], n_ranges,
                       int(g["range_size"]), original_len=int(g["original_len"]), **kw)
        assert np.array_equal(bits(out), bits(g["dec_" + tag])), tag


def test_container_bytes_and_roundtrip():
    g = golden("tone128")
    blob = O.pack_fwav(g["idx"], g["s"], g["o"], g["sym"], g["err"], g["domains"],
                       int(g["range_size"]), int(g["framerate"]), int(g["sampwidth"]),
                       int(g["tile_size"]), int(g["domain_step"]), float(g["energy_thresh"]),
                       int(g["original_len"]))
    assert blob == g["fwav_bytes"].tobytes()
    back = O.unpack_fwav(blob)
    assert np.array_equal(back["idx"], g["idx"]) and np.array_equal(bits(back["domains"]), bits(g["domains"]))
    assert back["original_len"] == int(g["original_len"]) and back["tile_size"] == 128
    bad = bytearray(blob)
    bad[-1] ^= 1
    with pytest.raises(ValueError):
        O.unpack_fwav(bytes(bad))
    with pytest.raises(ValueError):
        O.unpack_fwav(b"XWAV" + blob[4:])
    rec = O.decode(back["idx"], back["s"], back["o"], back["sym"], back["domains"],
                   back["n_ranges"], back["range_size"], original_len=back["original_len"])
    assert np.array_equal(bits(rec), bits(g["pipeline_decode"]))
    assert abs(O.compute_snr(g["signal"], rec) - float(g["snr"])) < 1e-12
    assert float(g["snr"]) > 4.0          # the reference's own assertion (test_e2e.py:38)


def test_voiced_gate():
    g = golden("voiced")
    assert np.array_equal(O.voiced_mask(g["signal"], 8, 1e-4), g["mask_f8"])
    assert np.array_equal(O.voiced_mask(g["signal"] * 1e-4, 32, 1e-4), g["mask_scaled_f32"])
    assert 0 < g["mask_scaled_f32"].mean() < 1


def test_reference_quirks():
    # F4: more ranges than domains is an error on the live path
    x = np.arange(150, dtype=np.float32) * 100
    with pytest.raises(ValueError):
        O.compress(x, tile_size=128)
    # silent input: empty result (fractal.py:1083-1093)
    res = O.compress(np.zeros(4000, np.float32), tile_size=1024)
    assert res["n_ranges"] == 0 and res["domains"].shape == (0, 4) and res["original_len"] == 4000
    # signal shorter than a tile: no domains (fractal.py:1130)
    res = O.compress(np.ones(500, np.float32) * 1000, tile_size=1024)
    assert res["n_ranges"] == 0 and res["original_len"] == 500
    assert O.derive_geometry(4096) == (16, 4) and O.derive_geometry(128) == (4, 1)


def test_tile_embedding_k32_pinned():
    """The tonal-only embedding (the README's EMBED_K = 32 form) against the reference's own tile_embedding."""
    g = golden("tile_embedding_k32")
    for n in (4, 8, 16, 40):
        got = O.embed_rows(g[f"rows_{n}"], 32, head="tonal")
        assert np.array_equal(bits(got), bits(g[f"emb_{n}"])), n
