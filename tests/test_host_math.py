"""The kernels' per-element math (np_math.cuh / fwav_math.cuh), compiled for the
CPU, against numpy and the reference-generated golden vectors.  This is what
makes the GPU results predictable before a GPU is touched."""
import ctypes as C

import numpy as np
import pytest

from conftest import golden
from host_harness import Harness, _p

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
def test_domains_through_shared_chains_bit_exact(name):
    """tables.cu's route (every chain of numpy's 128-sample leaf computed once, leaves from eight chains, rows from a
    staged window of leaves), emulated block by block with the kernels' index arithmetic: the reference's bits."""
    g = golden(name)
    tile, N, ds = int(g["tile_size"]), int(g["range_size"]), int(g["domain_step"])
    if tile // N != 256 or 128 % ds:
        pytest.skip("run length is not 256: the generic kernel serves this geometry")
    got, half = H.build_domains_chain(g["signal"], tile, N, ds)
    assert not np.isnan(half).any(), "a leaf no block wrote"
    assert np.array_equal(bits(got), bits(g["domains"]))


@pytest.mark.parametrize("n, stride", [(128, 1), (129, 4), (4343 + 128, 1), (4344 + 128, 4), (4345 + 128, 1),
                                       (3 * 4344 + 127, 1), (3 * 4344 + 131, 2), (20011, 3), (50000, 8), (70001, 4)])
def test_chain_half_sums_equal_the_direct_leaf(n, stride):
    """block edges of half_sums_chain_kernel (4344 leaf starts per block) and strides that are not powers of two"""
    rng = np.random.default_rng(n + stride)
    sig = (rng.standard_normal(n) * 10 ** rng.uniform(-3, 3, n)).astype(np.float32)
    n_half = (n - 128) // stride + 1
    half = np.full(n_half, np.nan, np.float32)
    H.lib.hh_half_sums_chain(_p(sig), C.c_longlong(n), C.c_longlong(n_half), C.c_int(stride), _p(half))
    assert not np.isnan(half).any(), "a leaf no block wrote"
    # np_mean of 128 values = (0 + leaf) / 128 and the division is exact: equal means <=> equal leaves
    edges = {min(n_half - 1, max(0, (b * 4344 + d) // stride)) for b in range(1, 4) for d in (-8, -1, 0, 1, 8)}
    for u in sorted(set(range(0, n_half, max(1, n_half // 300))) | {0, n_half - 1} | edges):
        got = np.float32((np.float32(0.0) + half[u]) / np.float32(128))
        assert bits(got) == bits(H.np_mean(sig[u * stride: u * stride + 128])), u


@pytest.mark.parametrize("name", ALL)
def test_embedding_close(name):
    g = golden(name)
    got = H.embed(g["domains"], int(g["emb_dim"]))
    assert np.abs(got - g["embeddings"]).max() <= 2e-6
    # padding layout: [tonal | transient | zeros]
    assert np.array_equal(got == 0, g["embeddings"] == 0) or int(g["range_size"]) >= 9
    # the kernels' compile-time form (half-length chains over mirrored sums / differences): the same gate, and
    # float64 rounding noise apart from the full-length chains
    if int(g["emb_dim"]) == 16:
        st = H.embed_static(g["domains"])
        if st is not None:
            assert np.abs(st - g["embeddings"]).max() <= 2e-6
            assert np.abs(st - got).max() <= 1.2e-7 and (st != got).mean() < 1e-2      # one float32 ulp, rarely
            assert np.array_equal(st == 0, got == 0)


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
    # the kernel's form: both orientations in one go, tile statistics shared where mirroring cannot change them
    pair = H.affine(g["ranges"], g["domains"], g["candidates"], pair=True)
    assert np.array_equal(pair["idx"], g["idx"]) and np.array_equal(pair["sym"], g["sym"])
    for k in ("s", "o", "err"):
        assert np.array_equal(bits(pair[k]), bits(g[k])), k


@pytest.mark.parametrize("N", [4, 8, 16, 32, 11])
def test_affine_pair_equals_two_single_fits(N):
    """fwm::affine_fit_pair against two fwm::affine_fit calls on adversarial rows: wide dynamic range, constant
    tiles (den = 1e-12), signed zeros, rows equal to their own mirror"""
    rng = np.random.default_rng(N)
    n_d, n_r, K = 400, 300, 24
    dom = (rng.standard_normal((n_d, N)) * 10.0 ** rng.integers(-6, 4, (n_d, 1))).astype(np.float32)
    dom[::7] = dom[::7, :1]                       # constant tiles
    dom[1::7] = (dom[1::7] + dom[1::7, ::-1]) / 2   # palindromes
    dom[2::7] *= 0.0
    dom[3::7] = -dom[2::7]
    rngs = (rng.standard_normal((n_r, N)) * 10.0 ** rng.integers(-5, 3, (n_r, 1))).astype(np.float32)
    rngs[::5] = dom[rng.integers(0, n_d, len(rngs[::5]))][:, ::-1]       # exact mirrored copies: err ties at 0
    cand = rng.integers(-1, n_d, (n_r, K)).astype(np.int32)
    a, b = H.affine(rngs, dom, cand), H.affine(rngs, dom, cand, pair=True)
    assert np.array_equal(a["idx"], b["idx"]) and np.array_equal(a["sym"], b["sym"])
    for k in ("s", "o", "err"):
        assert np.array_equal(bits(a[k]), bits(b[k])), k


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
