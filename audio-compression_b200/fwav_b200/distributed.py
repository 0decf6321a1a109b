"""Range-sharded multi-GPU compress and decode.

One process per GPU, `torch.distributed` for the plumbing (NCCL over NVLink on
the B200 box; the host logic is backend-agnostic and is exercised with gloo,
world_size 2, in tests/test_distributed_gloo.py).

Compress (reference: np.array_split of the ranges over cpu_workers,
fractal.py:1180-1207):
    rank 0 builds the domain table and its embeddings, both are broadcast, every
    rank matches its contiguous slice of ranges against the full tables, the five
    match arrays are all-gathered.  No collective sits inside the search.

Decode (reference loop fractal.py:1411-1467, range-local by construction):
    every rank iterates its slice; per iteration the two float64 sums behind
    delta are all-gathered and added in RANK ORDER (bit-stable for a given world
    size), so all ranks take the same convergence decision; the reconstruction is
    all-gathered every iteration (what BASELINE.json's north star specifies) or
    once at the end (`gather_every_iteration=False`).

The device work goes through an "engine" with four methods so the same driver
runs on the CUDA library (CudaEngine, below) and, in the CPU tests, on a stand-in
built from the oracle:
    build_tables(signal, tile) -> (domains, embs)
    match_slice(ranges, lo, hi, domains, embs, ...) -> (idx, s, o, err, sym) tensors
    match_slice(ranges, lo, hi, domains, embs, ..., out=packed (5, cap) int32)   rows idx | s | o | err | sym bytes
    new_state() / decode_iter(..., first, cur, nxt, state) -> sums(2,) f64 / converge(all_sums, parts, eps, state) /
    read_state(state) -> (iterations run, last delta): the decode loop with its convergence decision on the device
    empty(shape, dtype) / from_numpy(...) for allocation on the engine's device
"""
from __future__ import annotations

import os

import numpy as np


def shard_bounds(n, world):
    """Contiguous slices with np.array_split's sizes (fractal.py:1182)."""
    base, extra = divmod(int(n), int(world))
    sizes = [base + (1 if r < extra else 0) for r in range(world)]
    edges = np.concatenate([[0], np.cumsum(sizes)])
    return [(int(edges[r]), int(edges[r + 1])) for r in range(world)]


def delta_from_sums(dsq, csq):
    """fractal.py:1460-1461 from float64 sums of squares, in float32 like numpy's norms."""
    nd = np.float32(np.sqrt(np.float64(dsq)))
    nc = np.float32(np.sqrt(np.float64(csq)))
    return float(nd / (nc if nc > 0 else np.float32(1.0)))


def _dist():
    import torch.distributed as dist
    return dist


def compress_sharded(engine, signal, ranges, tile_size, emb_dim=16, top_k=32, energy_thresh=1e-4,
                     fast_mode=True, query_mode=0, broadcast_tables=False):
    """Returns (idx, s, o, sym, err, domains) as engine tensors on every rank.

    Tables: every rank builds the domain table and its embeddings from the (replicated) signal itself -- the
    kernels are deterministic, so the tables are bit-identical, and 0.2 ms of HBM-bound work beats broadcasting
    254 MB (0.85 ms measured at config 2).  `broadcast_tables=True` keeps the north star's literal form (rank 0
    builds, NCCL broadcasts).  Matches: ONE all-gather of a packed (5, cap) int32 block per rank
    (idx | s | o | err | sym bytes), written in place by the match kernel."""
    import torch
    dist = _dist()
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    n_r = ranges.shape[0]
    bounds = shard_bounds(n_r, world)
    lo, hi = bounds[rank]
    if world > 1 and broadcast_tables:
        domains, embs = engine.alloc_tables(signal, tile_size, emb_dim)
        if rank == 0:
            engine.build_tables(signal, tile_size, emb_dim, out=(domains, embs))
        dist.broadcast(domains, 0)
        dist.broadcast(embs, 0)
    else:
        domains, embs = engine.build_tables(signal, tile_size, emb_dim)
    if query_mode == 0 and n_r > domains.shape[0]:
        raise ValueError("mmap length is greater than file size")       # fractal.py:1190
    cap = max(max(b - a for a, b in bounds), 1)
    cap = -(-cap // 4) * 4                       # the sym row is read back as bytes of whole int32 words
    packed = engine.empty((5, cap), torch.int32)
    packed.zero_()
    engine.match_slice(ranges, lo, hi, domains, embs, tile_size, emb_dim, top_k, energy_thresh, fast_mode,
                       query_mode, out=packed)
    if world == 1:
        g = packed.view(1, 5, cap)
    else:
        g = engine.empty((world * 5 * cap,), torch.int32)     # flat: accepted by both the NCCL and the gloo all-gather
        dist.all_gather_into_tensor(g, packed.view(-1))
        g = g.view(world, 5, cap)
    cols = [torch.cat([g[r, k, :b - a] for r, (a, b) in enumerate(bounds)]) for k in range(4)]
    sym_all = torch.cat([g[r, 4].view(torch.uint8)[:b - a] for r, (a, b) in enumerate(bounds)])
    return (cols[0], cols[1].view(torch.float32), cols[2].view(torch.float32), sym_all,
            cols[3].view(torch.float32), domains)


def decode_sharded(engine, domains, idx, s, o, sym, range_size, iterations=8, convergence_eps=1e-3,
                   s_clip=16.0, s_damping=0.0, gather_every_iteration=True, fused=None):
    """Every rank passes the FULL match arrays; returns (recon tensor of n_ranges*range_size on every rank,
    iterations run, last delta).

    No host read-back inside the loop: per iteration the ranks all-gather their two float64 sums, a one-thread
    kernel adds them in rank order and takes the convergence decision into a device-side state, and the iteration
    kernels that follow a `done` return at once.  The state is read once, after the last launch.

    `fused` (default: when the engine offers it and the reconstruction is exchanged every iteration): the iteration
    kernel itself stores every output sample into the full buffer of every GPU -- one multimem.st to the NVSwitch
    multicast address of a symmetric allocation, or plain stores through peer pointers -- so the exchange the north
    star asks for overlaps the compute and no all-gather of the reconstruction is issued at all."""
    import torch
    dist = _dist()
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    n_r, N = idx.shape[0], range_size
    bounds = shard_bounds(n_r, world)
    lo, hi = bounds[rank]
    cap = max(b - a for a, b in bounds)
    bufs = [engine.empty((cap * N,), torch.float32), engine.empty((cap * N,), torch.float32)]
    for b in bufs:
        b.zero_()
    sym_full = None
    if world > 1 and gather_every_iteration and fused is not False and hasattr(engine, "symmetric_full") and \
            N in (4, 8, 16, 32) and (cap * N) % 4 == 0:
        sym_full = engine.symmetric_full(world * cap * N)           # None when symmetric memory is not available
    if fused and sym_full is None:
        raise RuntimeError("fused decode asked for, but the engine has no symmetric memory for the full buffer")
    if sym_full:
        engine.symmetric_barrier(sym_full)           # nobody still reads the buffer of a previous decode
    full = sym_full[0] if sym_full else (engine.empty((world * cap * N,), torch.float32) if world > 1 else None)
    all_sums = engine.empty((world * 2,), torch.float64)
    state = engine.new_state()
    for it in range(iterations):
        # iteration `it` reads bufs[it & 1] and writes bufs[(it + 1) & 1]
        if sym_full:
            sums = engine.decode_iter_bcast(domains, idx[lo:hi], s[lo:hi], o[lo:hi], sym[lo:hi], N, s_clip, s_damping,
                                            it == 0, bufs[it & 1], bufs[(it + 1) & 1], state, sym_full, rank * cap * N)
        else:
            sums = engine.decode_iter(domains, idx[lo:hi], s[lo:hi], o[lo:hi], sym[lo:hi], N, s_clip, s_damping,
                                      it == 0, bufs[it & 1], bufs[(it + 1) & 1], state)
        if world > 1:
            dist.all_gather_into_tensor(all_sums, sums)
        else:
            all_sums.copy_(sums)
        engine.converge(all_sums, world, convergence_eps, state)        # rank order: bit-stable for a given world size
        if world > 1 and gather_every_iteration and not sym_full:
            dist.all_gather_into_tensor(full, bufs[(it + 1) & 1])
    it_run, delta = engine.read_state(state)                            # the one synchronisation of the decode
    cur = bufs[it_run & 1]
    if world == 1:
        return cur[:n_r * N], it_run, delta
    if sym_full and it_run > 0:
        engine.symmetric_barrier(sym_full)           # every rank's last stores have landed in every full buffer
    elif not gather_every_iteration or it_run < iterations or it_run == 0:
        dist.all_gather_into_tensor(full, cur)       # converged early: the gathers after it carried the stale buffer
    full = full.view(world, cap * N)
    out = torch.cat([full[r, :(b - a) * N] for r, (a, b) in enumerate(bounds)])
    return out, it_run, delta


class CudaEngine:
    """The engine on libfwav_b200.so: torch tensors for memory, the C ABI for work."""

    def __init__(self, device, ctx=None):
        import torch
        from . import _lib
        self.torch = torch
        self.lib = _lib
        self.device = torch.device("cuda", device)
        self.ctx = ctx if ctx is not None else _lib.Context(device)
        self._sums = None

    def _stream(self):
        # torch's NULL stream is the legacy default stream; the C ABI reads NULL as "the
        # context's own stream", so name the legacy stream explicitly (cudaStreamLegacy = 0x1)
        st = self.torch.cuda.current_stream(self.device).cuda_stream
        return st if st else 1

    def empty(self, shape, dtype):
        return self.torch.empty(shape, dtype=dtype, device=self.device)

    def from_numpy(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a)).to(self.device)

    def alloc_tables(self, signal, tile_size, emb_dim):
        N, ds = self.lib.geometry(tile_size)
        n_d = self.lib.count_domains(signal.shape[0], tile_size, ds)
        return (self.empty((n_d, N), self.torch.float32), self.empty((n_d, emb_dim), self.torch.float32))

    def build_tables(self, signal, tile_size, emb_dim, out=None):
        N, ds = self.lib.geometry(tile_size)
        domains, embs = out if out is not None else self.alloc_tables(signal, tile_size, emb_dim)
        self.ctx.build_tables(signal.data_ptr(), signal.shape[0], tile_size, N, ds, emb_dim, domains.data_ptr(),
                              embs.data_ptr(), self._stream())
        return domains, embs

    def match_slice(self, ranges, lo, hi, domains, embs, tile_size, emb_dim, top_k, energy_thresh, fast_mode,
                    query_mode, out):
        """Matches of ranges [lo, hi) into the packed block `out` (5, cap) int32: the kernel writes the rows."""
        t = self.torch
        N = ranges.shape[1]
        cnt = hi - lo
        if cnt:
            rp = ranges.data_ptr() + lo * N * 4
            act = self.empty((cnt,), t.uint8)
            cand = self.empty((cnt, top_k), t.int32)
            st = self._stream()
            self.ctx.range_activity(rp, cnt, N, energy_thresh, fast_mode, act.data_ptr(), st)
            if query_mode == 0:
                qp = embs.data_ptr() + lo * emb_dim * 4
            else:
                q = self.empty((cnt, emb_dim), t.float32)
                self.ctx.embed(rp, cnt, N, emb_dim, q.data_ptr(), st)
                qp = q.data_ptr()
            self.ctx.set_search_range_size(N)        # both tables come from fwav_embed for this geometry
            try:
                self.ctx.topk(qp, cnt, embs.data_ptr(), embs.shape[0], emb_dim, top_k, act.data_ptr(), cand.data_ptr(),
                              None, st)
            finally:
                self.ctx.set_search_range_size(0)
            self.ctx.affine_match(rp, cnt, N, domains.data_ptr(), domains.shape[0], cand.data_ptr(), top_k, 16.0,
                                  out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), out[4].data_ptr(),
                                  out[3].data_ptr(), st)
            self._keep = (act, cand)             # alive until the stream has consumed them
        return out

    def new_state(self):
        return self.torch.zeros(4, dtype=self.torch.int32, device=self.device)      # fwav_decode_state

    def decode_iter(self, domains, idx, s, o, sym, N, s_clip, s_damping, first, cur, nxt, state):
        if self._sums is None:
            self._sums = self.empty((2,), self.torch.float64)
        self.ctx.decode_iter_gated(domains.data_ptr(), domains.shape[0], idx.data_ptr(), s.data_ptr(), o.data_ptr(),
                                   sym.data_ptr(), idx.shape[0], N, s_clip, s_damping, first, cur.data_ptr(),
                                   nxt.data_ptr(), self._sums.data_ptr(), state.data_ptr(), self._stream())
        return self._sums

    def symmetric_full(self, numel):
        """Full reconstruction buffer in symmetric memory (same allocation on every rank, peer-mapped; NVSwitch
        multicast address when the fabric offers one).  Returns (tensor, targets, multimem, handle) or None."""
        want_mc = os.environ.get("FWAV_DECODE_MULTIMEM", "1") != "0"
        cache = self.__dict__.setdefault("_symm", {})
        if (numel, want_mc) in cache:                # allocation + rendezvous are collective and slow: once per size
            return cache[(numel, want_mc)]
        try:
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm
            key = ("buf", numel)
            if key not in cache:
                t = symm.empty(numel, dtype=self.torch.float32, device=self.device)
                cache[key] = (t, symm.rendezvous(t, dist.group.WORLD))
            t, hdl = cache[key]
            mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
            if want_mc and mc:
                res = (t, [mc], True, hdl)
            else:
                res = (t, [int(p) for p in hdl.buffer_ptrs], False, hdl)
        except Exception as e:                       # no NVLink peer access / old torch: the NCCL all-gather path
            import logging
            logging.getLogger("fwavc").warning("symmetric memory unavailable (%s): decode falls back to NCCL all-gather", e)
            res = None
        cache[(numel, want_mc)] = res
        return res

    def symmetric_barrier(self, sym_full):
        sym_full[3].barrier(channel=0)
        self.torch.cuda.current_stream(self.device).synchronize()

    def decode_iter_bcast(self, domains, idx, s, o, sym, N, s_clip, s_damping, first, cur, nxt, state, sym_full, offset):
        if self._sums is None:
            self._sums = self.empty((2,), self.torch.float64)
        _, targets, multimem, _ = sym_full
        self.ctx.decode_iter_bcast(domains.data_ptr(), domains.shape[0], idx.data_ptr(), s.data_ptr(), o.data_ptr(),
                                   sym.data_ptr(), idx.shape[0], N, s_clip, s_damping, first, cur.data_ptr(),
                                   nxt.data_ptr(), self._sums.data_ptr(), state.data_ptr(), targets, multimem, offset,
                                   self._stream())
        return self._sums

    def converge(self, all_sums, parts, eps, state):
        self.ctx.decode_converge(all_sums.data_ptr(), parts, eps, state.data_ptr(), self._stream())

    def read_state(self, state):
        h = state.cpu()
        if int(h[3]):
            raise IndexError("index out of bounds: a match points past the domain table")
        return int(h[0]), float(h[2:3].view(self.torch.float32)[0])
