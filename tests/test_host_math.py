"""The kernels' per-element math (np_math.cuh / fwav_math.cuh), compiled for the
CPU, against numpy and the reference-generated golden vectors.  This is what
makes the GPU results predictable before a GPU is touched."""
import numpy as np
import pytest

from conftest import golden
from host_harness import Harness

H = Harness()
ALL = ["tone128", "sine_t1024", "music_t4096", "gaps_t1024", "float_t1024",
       "music_k64", "tiny_kfull", "sine_t1100", "music_t3000", "music_t2048"]


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def test_pairwise_mean_matches_numpy():
    rng = np.random.default_rng(0)
    for n in list(range(1, 40)) + [63, 64, 100, 127, 128, 129, 136, 255, 256, 257, 272, 275, 300, 511]:
        for _ in range(5):
            a = (rng.standard_normal(n) * 10 ** rng.uniform(-3, 4)).astype(np.float32)
            want = a.reshape(1, n).mean(axis=1, dtype=np.float32)[0]
            assert H.np_mean(a).view(np.uint32) == want.view(np.uint32), n


@pytest.mark.parametrize("name", ALL)
def test_domains_bit_exact(name):
    g = golden(name)
    tile, N, ds = int(g["tile_size"]), int(g["range_size"]), int(g["domain_step"])
    got = H.build_domains(g["signal"], tile, N, ds)
    assert np.array_equal(bits(got), bits(g["domains"]))
    if tile // N == 256:
        assert np.array_equal(bits(H.build_domains(g["signal"], tile, N, ds, force_generic=True)), bits(g["domains"]))


@pytest.mark.parametrize("name", ALL)
def test_embedding_close(name):
    g = golden(name)
    got = H.embed(g["domains"], int(g["emb_dim"]))
    assert np.abs(got - g["embeddings"]).max() <= 2e-6
    # padding layout: [tonal | transient | zeros]
    assert np.array_equal(got == 0, g["embeddings"] == 0) or int(g["range_size"]) >= 9


@pytest.mark.parametrize("name", ALL)
def test_affine_bit_exact_given_candidates(name):
    g = golden(name)
    got = H.affine(g["ranges"], g["domains"], g["candidates"])
    assert np.array_equal(got["idx"], g["idx"])
    assert np.array_equal(got["sym"], g["sym"])
    for k in ("s", "o", "err"):
        assert np.array_equal(bits(got[k]), bits(g[k])), k
    act = H.activity(g["ranges"], float(g["energy_thresh"]))
    assert np.array_equal(act == 0, (g["candidates"] < 0).all(axis=1))


@pytest.mark.parametrize("name", ["tone128", "sine_t1024", "music_t4096", "gaps_t1024",
                                  "float_t1024", "sentinel_decode", "music_t3000", "music_t2048"])
def test_decode_bit_exact(name):
    g = golden(name)
    N = int(g["range_size"])
    variants = {
        "default": dict(iterations=8, eps=1e-3),
        "damp50": dict(iterations=8, eps=0.0, s_damping=0.5),
        "damp25_clip2": dict(iterations=5, eps=1e-3, s_damping=0.25, s_clip=2.0),
    }
    for tag, kw in variants.items():
        if "dec_" + tag not in g:
            continue
        out, it, delta = H.decode(g["domains"], g["idx"], g["s"], g["o"], g["sym"], N, **kw)
        want = g["dec_" + tag]
        assert np.array_equal(bits(out[:len(want)]), bits(want)), tag


def test_prestep_math_matches_the_reference_gate():
    """Row N2: the device pre-step's per-frame math (frame energies in numpy's order, the 5-tap smoothing with its
    two edge forms, the threshold keys whose running maximum is the hysteresis) against masks the reference's
    voiced_detection produced, and against the host pre-step on every fixture's signal."""
    from fwav_b200.prestep import frame_ranges, voiced_detection
    g = golden("voiced")
    mask, _, _ = H.prestep(g["signal"], 4)
    assert np.array_equal(mask, g["mask_f8"])
    mask, _, _ = H.prestep(g["signal"] * 1e-4, 16)
    assert np.array_equal(mask, g["mask_scaled_f32"])
    for name in ALL:
        f = golden(name)
        N, thr = int(f["range_size"]), float(f["energy_thresh"])
        mask, ranges, ssq = H.prestep(f["signal"], N, thr)
        assert np.array_equal(mask, voiced_detection(f["signal"], 2 * N, thr)), name
        assert np.array_equal(bits(ranges), bits(f["ranges"])), name            # what the reference framed
        assert ssq >= 1e-8
    # float-scale signals whose energies straddle the thresholds, odd frame sizes, lengths that need the reflected tail
    rng = np.random.default_rng(5)
    for N, n in [(4, 1003), (5, 777), (11, 4099), (16, 10007), (32, 6400), (7, 71)]:
        for scale in (1e-2, 3e-2, 1.0):
            x = (rng.standard_normal(n) * scale * np.repeat(rng.random(-(-n // 50)) < 0.5, 50)[:n]).astype(np.float32)
            mask, ranges, ssq = H.prestep(x, N, 1e-4)
            assert np.array_equal(mask, voiced_detection(x, 2 * N, 1e-4)), (N, n, scale)
            want, _ = frame_ranges(x, N, 1e-4)
            if want is None:
                assert ssq < 1e-8 or len(x) < N
            else:
                assert np.array_equal(bits(ranges), bits(want)), (N, n, scale)


def test_tonal_embedding_close_to_the_reference():
    """embed_tonal_row (float64 matrix product, one rounding) against the reference's tile_embedding(k=32)."""
    g = golden("tile_embedding_k32")
    for n in (4, 8, 16, 40):
        got = H.embed_tonal(g[f"rows_{n}"], 32)
        want = g[f"emb_{n}"]
        # rows 0 (all zero) and 1 (constant) have no AC content: the reference normalises its float32 DCT's rounding
        # noise to a unit vector there, the float64 product stays below the 1e-8 norm gate -- as for the live path's
        # tonal head, such tiles carry no direction worth matching
        assert np.abs(got[2:] - want[2:]).max() <= 2e-6, n
        assert np.array_equal((got == 0).all(axis=0), (want == 0).all(axis=0)), n      # zero padding past N - 1
