"""CPU oracle for the FWAV hot path (TEST INFRASTRUCTURE ONLY).

This file is a numpy restatement of the reference algorithm in
/root/reference/fractal.py for the path BASELINE.json names: derived geometry,
domain construction, the two-head DCT embedding, the linear candidate search,
the batched affine + mirror match, the iterative decoder and the .fwav
container.  Every function cites the reference lines it follows.

It is the CHECKER, never the product:
  * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
    --impl reference legs may import it;
  * nothing under audio-compression_b200/ imports it; the product path fails
    loudly when the CUDA library is missing.

Parity pinning: the reference ships no golden vectors for this path
(test_e2e.py only asserts SNR > 4 dB), so the oracle is pinned against outputs
of the reference itself, executed in the build container under a results-neutral
`librosa` stub by oracle/make_golden.py; the resulting fixtures live in
tests/golden/ and tests/test_oracle_golden.py checks this file against them.

Third-party arithmetic the reference leans on (no version pins upstream; the
fixtures were generated with numpy 2.3.5 / scipy 1.18.1 / OpenBLAS 0.3.30):
scipy.fftpack.dct (pocketfft), numpy pairwise-sum reductions, OpenBLAS sgemv,
np.argpartition (introselect).  The oracle calls the same library routines in
the same order, so on the same library versions it is bit-identical to the
reference's single-process replay.
"""
from __future__ import annotations

import hashlib
import struct

import numpy as np
from scipy.fftpack import dct as _dct

FWAV_MAGIC = b"FWAV"
FWAV_VERSION = 1          # fractal.py:59
DEFAULT_TOP_K = 32        # fractal.py:77 (module global read by cpu_worker :611-622)


# ----------------------------------------------------------------------------
# A0  derived geometry, masking, padding           fractal.py:1070-1112, 880-909
# ----------------------------------------------------------------------------
def derive_geometry(tile_size: int):
    """range_size and domain_step are derived, never passed (fractal.py:1070-1071)."""
    range_size = max(4, tile_size // 256)
    domain_step = max(1, range_size // 4)
    return range_size, domain_step


def voiced_mask(signal, frame_size, energy_threshold, smooth_window=5, low_threshold=None):
    """Frame-energy hysteresis gate, one 0/1 value per sample (fractal.py:880-909).

    The reference walks the frames in a Python loop; the state after frame i is
    decided by the most recent frame whose energy crossed either threshold, so a
    forward fill of "last decisive frame" gives the same mask.
    """
    x = np.asarray(signal, dtype=np.float32)
    n = x.shape[0]
    n_frames = (n + frame_size - 1) // frame_size
    tail = n_frames * frame_size - n
    framed = np.pad(x, (0, tail), mode="reflect").reshape(n_frames, frame_size)
    energy = np.mean(framed * framed, axis=1)                      # :891
    if smooth_window > 1:                                          # :893-895
        box = np.ones(smooth_window, dtype=np.float32) / smooth_window
        energy = np.convolve(energy, box, mode="same")
    if low_threshold is None:                                      # :897-898
        low_threshold = energy_threshold * 0.5
    on = energy > energy_threshold                                 # :903-904
    off = (~on) & (energy < low_threshold)                         # :905-906
    decisive = on | off
    last = np.where(decisive, np.arange(n_frames), -1)
    last = np.maximum.accumulate(last)
    state = np.where(last >= 0, on[np.maximum(last, 0)], False)
    return np.repeat(state.astype(np.uint8), frame_size)[:n]       # :909


def prepare_ranges(signal, tile_size, energy_thresh):
    """Mask, early-out test, reflect-pad and frame the ranges (fractal.py:1070-1112).

    Returns a dict; `ranges` is None when the reference would return its empty
    result (silent input :1083, or no ranges :1100).
    """
    range_size, domain_step = derive_geometry(tile_size)
    mask = voiced_mask(signal, frame_size=range_size * 2, energy_threshold=energy_thresh)
    weighted = signal * mask                                        # :1079
    original_len = len(weighted)
    out = dict(range_size=range_size, domain_step=domain_step,
               original_len=original_len, ranges=None, n_ranges=0)
    if np.sum(weighted ** 2) < 1e-8:                                # :1083
        return out
    pad = (range_size - (original_len % range_size)) % range_size   # :1095
    if pad:
        weighted = np.pad(weighted, (0, pad), mode="reflect")       # :1097
    n_ranges = len(weighted) // range_size
    if n_ranges == 0:                                               # :1100
        return out
    out["ranges"] = weighted.reshape(n_ranges, range_size)          # :1112
    out["n_ranges"] = n_ranges
    return out


# ----------------------------------------------------------------------------
# A1  domain construction                                     fractal.py:285-334
# ----------------------------------------------------------------------------
def count_domains(n, tile_size, domain_step):
    if n < tile_size:
        return 0
    return (n - tile_size) // domain_step + 1


def build_domains(signal, tile_size, range_size, domain_step, block=500):
    """domain[j][k] = f32 mean of the k-th run of `tile_size // range_size`
    samples of the window that starts at j*domain_step (fractal.py:301-331).

    The mean is numpy's float32 pairwise reduction over a C-contiguous last
    axis, which is what the reference gets after `reshape` copies the strided
    window view (:326-327).  Windows are taken from the RAW signal (:1120-1121).
    """
    x = np.asarray(signal, dtype=np.float32)
    n_domains = count_domains(len(x), tile_size, domain_step)
    if n_domains == 0:
        return np.zeros((0, range_size), dtype=np.float32)
    run = tile_size // range_size                                   # :314
    usable = run * range_size                                       # :315
    windows = np.lib.stride_tricks.sliding_window_view(x, tile_size)[::domain_step]
    out = np.empty((n_domains, range_size), dtype=np.float32)
    for lo in range(0, n_domains, block):                           # :318
        chunk = windows[lo:lo + block][:, :usable]
        chunk = chunk.reshape(chunk.shape[0], range_size, run)      # forces a contiguous copy
        out[lo:lo + chunk.shape[0]] = chunk.mean(axis=2, dtype=np.float32)
    return out


# ----------------------------------------------------------------------------
# A2/A3  embeddings                                  fractal.py:154-208, 238-280
# ----------------------------------------------------------------------------
def tonal_head(rows, k, fast_norm=False):
    """`tile_embedding`: f32 orthonormal DCT-II, x linspace(1,2,N) in f64, drop
    DC, keep k, zero-pad to k, f32 L2-normalise when norm > 1e-8 (fractal.py:178-208)."""
    rows = np.asarray(rows, dtype=np.float32)
    n_rows, n = rows.shape
    spec = _dct(rows, axis=1, norm="ortho")                         # :186 (f32 pocketfft)
    spec = spec * np.linspace(1.0, 2.0, n)                          # :188-189 (-> f64)
    take = min(k, max(0, n - 1))                                    # :192-193
    head = np.zeros((n_rows, k), dtype=np.float32)
    head[:, :take] = spec[:, 1:1 + take].astype(np.float32)         # :195
    if fast_norm:       # bulk tables for the CPU baseline leg: same math, may differ in the last bit
        nrm = np.sqrt(np.sum(head * head, axis=1, dtype=np.float32))
        ok = nrm > 1e-8
        head[ok] = head[ok] / nrm[ok, None]
        return head
    for i in range(n_rows):                                         # :205-207, per-row f32 norm
        nrm = np.linalg.norm(head[i])
        if nrm > 1e-8:
            head[i] = head[i] / nrm
    return head


def transient_head(rows, k, fast_norm=False):
    """`transient_embedding`: first difference (prepend x0), x linspace(1,2,N),
    f64 orthonormal DCT-II, keep the first min(k,N) INCLUDING DC, normalise in
    f64 when norm > 1e-8, cast to f32 (fractal.py:154-164)."""
    rows = np.asarray(rows, dtype=np.float32)
    n_rows, n = rows.shape
    diff = np.diff(rows, axis=1, prepend=rows[:, :1])               # :156 (f32)
    diff = diff * np.linspace(1.0, 2.0, n)                          # :158 (-> f64)
    spec = _dct(diff, axis=1, norm="ortho")[:, :k]                  # :159-160
    if fast_norm:
        nrm = np.sqrt(np.sum(spec * spec, axis=1))
        ok = nrm > 1e-8
        spec = spec.copy()
        spec[ok] = spec[ok] / nrm[ok, None]
        return spec.astype(np.float32)
    out = np.empty(spec.shape, dtype=np.float32)
    for i in range(n_rows):                                         # :161-164
        v = spec[i]
        nrm = np.linalg.norm(v)
        if nrm > 1e-8:
            v = v / nrm
        out[i] = v.astype(np.float32)
    return out


def embed_rows(rows, emb_dim=16, fast_norm=False, head="two_head"):
    """`multi_head_embedding(tile, emb_dim//2, emb_dim//2)` for every row
    (fractal.py:166-175 called from :271-277): [tonal | transient | zero pad].
    head="tonal": `tile_embedding(tile, k=emb_dim)` (fractal.py:178-208), the form the README describes
    (EMBED_K = 32); not on the reference's live path."""
    if head == "tonal":
        return tonal_head(rows, emb_dim, fast_norm)
    half = emb_dim // 2
    rows = np.asarray(rows, dtype=np.float32)
    ton = tonal_head(rows, half, fast_norm)
    tra = transient_head(rows, half, fast_norm)
    out = np.zeros((rows.shape[0], emb_dim), dtype=np.float32)
    out[:, :half] = ton
    out[:, half:half + tra.shape[1]] = tra                          # :170-174 pad at the END
    return out


def embed_one(tile, emb_dim=16):
    """Single-tile form, used for timing the reference's per-domain loop."""
    return embed_rows(np.asarray(tile, dtype=np.float32)[None, :], emb_dim)[0]


# ----------------------------------------------------------------------------
# A4/A5  linear candidate search + energy prune     fractal.py:535-552, 598-623
# ----------------------------------------------------------------------------
def search_candidates(q, domain_embs, top_k):
    """Indices of the top_k largest `domain_embs @ q`, best first (fractal.py:535-541)."""
    scores = domain_embs @ q                                        # f32 sgemv
    if top_k >= len(scores):
        return np.argsort(scores)[::-1].astype(np.int32)            # :538-539
    part = np.argpartition(scores, -top_k)[-top_k:]                 # :540
    return part[np.argsort(scores[part])[::-1]].astype(np.int32)    # :541


def pad_candidates(idxs, top_k):
    """Fixed-width candidate row, -1 padded (fractal.py:544-552)."""
    row = np.full(top_k, -1, dtype=np.int32)
    if idxs is None or len(idxs) == 0:
        return row
    idxs = np.asarray(idxs, dtype=np.int32)[:top_k]
    row[:len(idxs)] = idxs
    return row


def is_pruned(range_row, energy_thresh, fast_mode=True):
    """Energy prune of cpu_worker (fractal.py:602): no candidates at all."""
    return bool(fast_mode and np.mean(range_row ** 2) < energy_thresh * 0.75)


def candidates_for_ranges(ranges, query_embs, domain_embs, top_k, energy_thresh,
                          fast_mode=True, which=None):
    """Candidate table (n, top_k) int32 for the given range ids (fractal.py:598-623).

    `query_embs[i]` is what the reference calls the "range embedding" of range
    i; on the live path that is row i of the DOMAIN embedding file (:1190-1195).
    """
    ids = np.arange(len(ranges)) if which is None else np.asarray(which)
    table = np.full((len(ids), top_k), -1, dtype=np.int32)
    for row, i in enumerate(ids):
        if is_pruned(ranges[i], energy_thresh, fast_mode):
            continue
        table[row] = pad_candidates(search_candidates(query_embs[i], domain_embs, top_k), top_k)
    return table


# ----------------------------------------------------------------------------
# A6  batched affine solve + mirror check                     fractal.py:757-850
# ----------------------------------------------------------------------------
def affine_match(ranges, cand, domains, s_clip=16.0, want_all=False):
    """Least-squares R ~ s*D + o over K candidates and their mirrors; first
    argmin of the L2 residual over [plain 0..K-1, mirrored 0..K-1]; s clipped
    after the residual is taken, o not recomputed (fractal.py:772-825)."""
    ranges = np.asarray(ranges, dtype=np.float32)
    cand = np.asarray(cand, dtype=np.int32)
    n_r, k = cand.shape
    safe = np.where(cand < 0, 0, cand)                              # :772-773
    tiles = domains[safe]                                           # :776
    both = np.concatenate([tiles, tiles[:, :, ::-1]], axis=1)       # :779-780
    idx2 = np.concatenate([safe, safe], axis=1)                     # :787
    r_mean = np.mean(ranges, axis=1, keepdims=True)                 # :790
    r_c = ranges - r_mean                                           # :791
    d_mean = np.mean(both, axis=2, keepdims=True)                   # :796
    d_c = both - d_mean                                             # :797
    num = np.sum(d_c * r_c[:, None, :], axis=2)                     # :802
    den = np.sum(d_c * d_c, axis=2) + 1e-12                         # :803
    s = num / den                                                   # :804
    o = r_mean - s * d_mean[:, :, 0]                                # :805
    fit = s[:, :, None] * both + o[:, :, None]                      # :811
    err = np.linalg.norm(fit - ranges[:, None, :], axis=2)          # :812-813
    bad = np.concatenate([cand < 0, cand < 0], axis=1)              # :816
    err = np.where(bad, np.inf, err)                                # :817
    pick = np.argmin(err, axis=1)                                   # :820
    rows = np.arange(n_r)
    res = dict(
        idx=idx2[rows, pick].astype(np.int32),
        s=np.clip(s[rows, pick], -abs(s_clip), abs(s_clip)).astype(np.float32),  # :823
        o=o[rows, pick].astype(np.float32),                         # :824
        sym=(pick >= k).astype(np.uint8),                           # :782-785, :825
        err=err[rows, pick].astype(np.float32),
    )
    if want_all:
        res["all_err"] = err
        res["all_s"] = s
        res["all_o"] = o
    return res


# ----------------------------------------------------------------------------
# A8  compress: single-process replay of the live path      fractal.py:1045-1256
# ----------------------------------------------------------------------------
def compress(signal, tile_size=1024, emb_dim=16, top_k=DEFAULT_TOP_K, energy_thresh=1e-4,
             fast_mode=True, batch=512, query_mode="reference", want_intermediates=False, head="two_head"):
    """Replay of compress_audio without processes or queues.

    query_mode="reference": q_i = E[i] (the live aliasing, fractal.py:1190-1195;
    raises ValueError like np.memmap does when n_ranges > n_domains).
    query_mode="range": q_i = embedding of range i (what the docs describe).
    Returns a dict of arrays (+ intermediates on request).
    """
    signal = np.asarray(signal, dtype=np.float32)
    prep = prepare_ranges(signal, tile_size, energy_thresh)
    n_rng, rsz, step = prep["n_ranges"], prep["range_size"], prep["domain_step"]
    empty = dict(n_ranges=0, range_size=rsz, tile_size=tile_size, domain_step=step,
                 energy_thresh=energy_thresh, original_len=prep["original_len"],
                 domains=np.zeros((0, rsz), np.float32),
                 idx=np.zeros(0, np.int32), s=np.zeros(0, np.float32), o=np.zeros(0, np.float32),
                 sym=np.zeros(0, np.uint8), err=np.zeros(0, np.float32))
    if prep["ranges"] is None:
        return empty
    domains = build_domains(signal, tile_size, rsz, step)
    if len(domains) == 0:                                           # :1130
        return empty
    embs = embed_rows(domains, emb_dim, head=head)
    if query_mode == "reference":
        if n_rng > len(domains):
            raise ValueError("mmap length is greater than file size")   # what :1190 raises
        queries = embs[:n_rng]
    elif query_mode == "range":
        queries = embed_rows(prep["ranges"], emb_dim, head=head)
    else:
        raise ValueError(query_mode)
    cand = candidates_for_ranges(prep["ranges"], queries, embs, top_k, energy_thresh, fast_mode)
    parts = [affine_match(prep["ranges"][a:a + batch], cand[a:a + batch], domains)
             for a in range(0, n_rng, batch)]
    res = {key: np.concatenate([p[key] for p in parts]) for key in ("idx", "s", "o", "sym", "err")}
    res.update(n_ranges=n_rng, range_size=rsz, tile_size=tile_size, domain_step=step,
               energy_thresh=energy_thresh, original_len=prep["original_len"], domains=domains)
    if want_intermediates:
        res.update(ranges=prep["ranges"], embeddings=embs, candidates=cand)
    return res


def matches_as_tuples(res):
    """The reference's list of (int, float, float, int, float) (fractal.py:836-850)."""
    return list(zip(res["idx"].tolist(), res["s"].tolist(), res["o"].tolist(),
                    res["sym"].tolist(), res["err"].tolist()))


# ----------------------------------------------------------------------------
# A9  iterative decoder                                     fractal.py:1378-1473
# ----------------------------------------------------------------------------
def decode(idx, s, o, sym, domains, n_ranges, range_size, iterations=8, convergence_eps=1e-3,
           original_len=None, s_clip=16.0, s_damping=0.0, want_trace=False):
    """Range-local fixed-point iteration with static domain tiles."""
    idx = np.asarray(idx, dtype=np.int32).copy()
    s_st = np.asarray(s, dtype=np.float32).copy()
    o_st = np.asarray(o, dtype=np.float32).copy()
    flip = np.asarray(sym).astype(bool).copy()
    recon = np.zeros(n_ranges * range_size, dtype=np.float32)       # :1388-1389
    dead = idx < 0                                                  # :1399
    idx[dead] = 0
    trace = []
    for _ in range(iterations):                                     # :1411
        cur = recon.reshape(n_ranges, range_size)
        tiles = domains[idx]                                        # :1414
        if dead.any():                                              # :1417-1426
            tiles = np.array(tiles, copy=True)
            tiles[dead] = 0
            s_st[dead] = 0.0
            o_st[dead] = 0.0
            flip[dead] = False
        if flip.any():                                              # :1428-1429
            tiles = np.where(flip[:, None], tiles[:, ::-1], tiles)
        t_c = tiles - tiles.mean(axis=1)[:, None]                   # :1431-1432
        r_c = cur - cur.mean(axis=1)[:, None]                       # :1434-1435
        num = np.sum(r_c * t_c, axis=1)                             # :1437
        den = np.sum(t_c * t_c, axis=1)                             # :1438
        ok = den > 1e-12                                            # :1440
        s_opt = np.zeros_like(den)
        if ok.any():
            s_opt[ok] = num[ok] / den[ok]                           # :1443
        if s_damping > 0:                                           # :1445
            s_use = (1.0 - s_damping) * s_st + s_damping * s_opt
        else:
            s_use = np.where(ok, s_opt, s_st)
        s_use = np.clip(s_use, -abs(s_clip), abs(s_clip))           # :1446
        nxt = (s_use[:, None] * tiles + o_st[:, None]).ravel()      # :1449
        # :1451-1458: bincount over the identity scatter with unit counts is a
        # value-preserving f32 -> f64 -> f32 round trip.
        nxt = nxt.astype(np.float64).astype(np.float32)
        base = np.linalg.norm(recon)                                # :1460
        base = base if base > 0 else 1.0
        delta = float(np.linalg.norm(nxt - recon) / base)           # :1461
        recon = nxt
        trace.append(delta)
        if delta < convergence_eps:                                 # :1465
            break
    if original_len is not None:
        recon = recon[:original_len]                                # :1470-1471
    return (recon, trace) if want_trace else recon


# ----------------------------------------------------------------------------
# A10  .fwav container                                      fractal.py:1278-1375
# ----------------------------------------------------------------------------
_HEADER = "<4sBIIBHHfIII"          # 34 bytes, fractal.py:1291-1301
_MATCH = np.dtype([("idx", "<i4"), ("s", "<f4"), ("o", "<f4"), ("sym", "u1"), ("err", "<f4")])
assert struct.calcsize(_HEADER) == 34 and _MATCH.itemsize == 17


def pack_fwav(idx, s, o, sym, err, domains, range_size, framerate, sampwidth, tile_size,
              domain_step, energy_threshold, original_len) -> bytes:
    """Header, 32-byte SHA-256 of (domains || matches), payload (fractal.py:1289-1322)."""
    recs = np.empty(len(idx), dtype=_MATCH)
    recs["idx"], recs["s"], recs["o"], recs["sym"], recs["err"] = idx, s, o, sym, err
    body = np.ascontiguousarray(domains, dtype="<f4").tobytes() + recs.tobytes()
    head = struct.pack(_HEADER, FWAV_MAGIC, FWAV_VERSION, range_size, framerate, sampwidth,
                       tile_size, domain_step, energy_threshold, len(idx), len(domains),
                       original_len)
    return head + hashlib.sha256(body).digest() + body


def unpack_fwav(blob: bytes, verify_checksum=True):
    """Inverse of pack_fwav; ValueError on bad magic / version / digest (:1331-1370)."""
    (magic, ver, rsz, rate, width, tile, step, thr, n_rng, n_dom, orig) = \
        struct.unpack_from(_HEADER, blob, 0)
    if magic != FWAV_MAGIC:
        raise ValueError("Not a FWAV file")
    if ver != FWAV_VERSION:
        raise ValueError(f"Unsupported FWAV version: {ver}")
    digest = blob[34:66]
    body = blob[66:66 + 4 * rsz * n_dom + 17 * n_rng]
    if verify_checksum and hashlib.sha256(body).digest() != digest:
        raise ValueError("Checksum mismatch — file may be corrupted")
    domains = np.frombuffer(body, dtype="<f4", count=rsz * n_dom).reshape(n_dom, rsz)
    recs = np.frombuffer(body, dtype=_MATCH, count=n_rng, offset=4 * rsz * n_dom)
    return dict(idx=recs["idx"].copy(), s=recs["s"].copy(), o=recs["o"].copy(),
                sym=recs["sym"].copy(), err=recs["err"].copy(), domains=domains,
                n_ranges=n_rng, range_size=rsz, framerate=rate, sampwidth=width,
                tile_size=tile, domain_step=step, energy_threshold=thr, original_len=orig)


def compute_snr(original, reconstructed):
    """float64 SNR in dB (fractal.py:1478-1487)."""
    n = min(len(original), len(reconstructed))
    a = np.asarray(original[:n], dtype=np.float64)
    e = a - np.asarray(reconstructed[:n], dtype=np.float64)
    noise = np.sum(e * e)
    if noise <= 0:
        return float("inf")
    return 10.0 * np.log10(np.sum(a * a) / noise)
