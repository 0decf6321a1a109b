"""Host logic of the range-sharded multi-GPU path, world_size 2 over gloo on the
CPU.  The device work is replaced by a stand-in engine built from the oracle, so
what is tested is the sharding, the gathers, the rank-ordered delta reduction and
the convergence decision — the code the NCCL run on the B200 box goes through."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import golden
from fwav_b200 import distributed as D
from oracle import fwav_oracle as O


class OracleEngine:
    """CPU stand-in for CudaEngine (TEST INFRASTRUCTURE)."""

    def empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype)

    def from_numpy(self, a):
        return torch.from_numpy(np.ascontiguousarray(a))

    def alloc_tables(self, signal, tile_size, emb_dim):
        N, ds = O.derive_geometry(tile_size)
        n_d = O.count_domains(signal.shape[0], tile_size, ds)
        return torch.empty((n_d, N)), torch.empty((n_d, emb_dim))

    def build_tables(self, signal, tile_size, emb_dim, out=None):
        N, ds = O.derive_geometry(tile_size)
        dom = O.build_domains(signal.numpy(), tile_size, N, ds)
        emb = O.embed_rows(dom, emb_dim)
        if out is not None:
            out[0].copy_(torch.from_numpy(dom)); out[1].copy_(torch.from_numpy(emb))
            return out
        return torch.from_numpy(dom), torch.from_numpy(emb)

    def match_slice(self, ranges, lo, hi, domains, embs, tile_size, emb_dim, top_k, energy_thresh, fast_mode, query_mode,
                    out):
        r, e, d = ranges.numpy(), embs.numpy(), domains.numpy()
        q = e if query_mode == 0 else O.embed_rows(r, emb_dim)
        ids = np.arange(lo, hi)
        cand = O.candidates_for_ranges(r, q, e, top_k, energy_thresh, fast_mode, which=ids)
        m = O.affine_match(r[ids], cand, d)
        o = out.numpy()                                   # packed rows, as the CUDA kernel writes them
        for row, k in enumerate(("idx", "s", "o", "err")):
            o[row, :hi - lo] = m[k].view(np.int32)
        o[4].view(np.uint8)[:hi - lo] = m["sym"]
        return out

    def new_state(self):
        return dict(iters=0, done=False, delta=0.0)

    def converge(self, all_sums, parts, eps, state):
        if state["done"]:
            return
        h = all_sums.numpy().reshape(parts, 2)
        dsq = csq = 0.0
        for r in range(parts):                            # rank order
            dsq += float(h[r, 0]); csq += float(h[r, 1])
        state["delta"] = D.delta_from_sums(dsq, csq)
        state["iters"] += 1
        state["done"] = state["delta"] < eps

    def read_state(self, state):
        return state["iters"], state["delta"]

    def decode_iter(self, domains, idx, s, o, sym, N, s_clip, s_damping, first, cur, nxt, state):
        if state["done"]:
            return torch.zeros(2, dtype=torch.float64)
        n = idx.shape[0]
        c = np.zeros(n * N, np.float32) if first else cur.numpy()[:n * N].copy()
        # one oracle iteration from the given state: replay its loop body
        tiles = domains.numpy()[np.maximum(idx.numpy(), 0)].copy()
        dead = idx.numpy() < 0
        tiles[dead] = 0
        flip = sym.numpy().astype(bool) & ~dead
        tiles = np.where(flip[:, None], tiles[:, ::-1], tiles)
        ss = np.where(dead, 0, s.numpy()).astype(np.float32)
        oo = np.where(dead, 0, o.numpy()).astype(np.float32)
        cr = c.reshape(n, N)
        t_c = tiles - tiles.mean(axis=1)[:, None]
        r_c = cr - cr.mean(axis=1)[:, None]
        num = np.sum(r_c * t_c, axis=1); den = np.sum(t_c * t_c, axis=1)
        ok = den > 1e-12
        s_opt = np.zeros_like(den); s_opt[ok] = num[ok] / den[ok]
        s_use = ((1.0 - s_damping) * ss + s_damping * s_opt) if s_damping > 0 else np.where(ok, s_opt, ss)
        s_use = np.clip(s_use, -abs(s_clip), abs(s_clip))
        out = (s_use[:, None] * tiles + oo[:, None]).ravel().astype(np.float32)
        nxt[:n * N] = torch.from_numpy(out)
        d = (out - c).astype(np.float64)
        return torch.tensor([float(np.sum(d * d)), float(np.sum(c.astype(np.float64) ** 2))], dtype=torch.float64)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, name, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = golden(name)
        eng = OracleEngine()
        sig, rng = torch.from_numpy(g["signal"]), torch.from_numpy(g["ranges"])
        idx, s, o, sym, err, dom = D.compress_sharded(eng, sig, rng, int(g["tile_size"]), 16, int(g["top_k"]), 1e-4)
        rec, iters, delta = D.decode_sharded(eng, dom, idx, s, o, sym, int(g["range_size"]), iterations=8,
                                             convergence_eps=0.0, s_damping=0.5)
        rec2, iters2, _ = D.decode_sharded(eng, dom, idx, s, o, sym, int(g["range_size"]), iterations=8,
                                           convergence_eps=1e-3, gather_every_iteration=False)
        q.put((rank, idx.numpy(), s.numpy(), o.numpy(), sym.numpy(), err.numpy(), rec.numpy(), iters, delta,
               rec2.numpy(), iters2))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["music_t3000", "gaps_t1024"])
def test_sharded_compress_and_decode_two_ranks(name):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    g = golden(name)
    N = int(g["range_size"])
    for rank, idx, s, o, sym, err, rec, iters, delta, rec2, iters2 in results:
        assert np.array_equal(idx, g["idx"]) and np.array_equal(sym, g["sym"])
        assert np.array_equal(s.view(np.uint32), g["s"].view(np.uint32))
        assert np.array_equal(o.view(np.uint32), g["o"].view(np.uint32))
        assert np.array_equal(err.view(np.uint32), g["err"].view(np.uint32))
        want = O.decode(g["idx"], g["s"], g["o"], g["sym"], g["domains"], len(g["idx"]), N, iterations=8,
                        convergence_eps=0.0, s_damping=0.5)
        assert iters == 8 and np.array_equal(rec.view(np.uint32), want.view(np.uint32))
        want2, trace = O.decode(g["idx"], g["s"], g["o"], g["sym"], g["domains"], len(g["idx"]), N, iterations=8,
                                convergence_eps=1e-3, want_trace=True)
        assert iters2 == len(trace) and np.array_equal(rec2.view(np.uint32), want2.view(np.uint32))
    # both ranks took identical decisions
    assert results[0][7:9] == results[1][7:9]


def test_shard_bounds_match_array_split():
    for n in (0, 1, 7, 240, 1003, 496125):
        for w in (1, 2, 3, 4, 8):
            want = [(int(a[0]), int(a[-1]) + 1) if len(a) else None for a in np.array_split(np.arange(n), w)]
            got = D.shard_bounds(n, w)
            assert len(got) == w and got[0][0] == 0 and got[-1][1] == n
            for g_, w_ in zip(got, want):
                if w_ is None:
                    assert g_[0] == g_[1]
                else:
                    assert g_ == w_
    assert D.delta_from_sums(0.0, 0.0) == 0.0 and D.delta_from_sums(4.0, 0.0) == 2.0
    assert abs(D.delta_from_sums(1.0, 4.0) - 0.5) < 1e-7
