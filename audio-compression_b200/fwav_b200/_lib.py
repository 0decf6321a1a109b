"""ctypes binding of libfwav_b200.so (the C ABI in include/fwav_b200.h).

This is the only door between the Python host code and the CUDA kernels.  There
is NO fallback: if the shared library is missing or no CUDA device is present
the first call raises FwavError.  The library is loaded lazily so importing the
package (and forking workers, as the reference's batch mode does,
fractal.py:1605) never touches CUDA.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FWAV_LIB") or os.path.join(_HERE, "libfwav_b200.so")    # FWAV_LIB: experiment builds

SEARCH_AUTO, SEARCH_FFMA, SEARCH_UMMA = 0, 1, 2
EMBED_TWO_HEAD, EMBED_TONAL = 0, 1

c_ctx = C.c_void_p
c_ptr = C.c_void_p
i64 = C.c_int64


class FwavError(RuntimeError):
    pass


# (name, restype, argtypes) — must list every symbol include/fwav_b200.h declares
SIGNATURES = [
    ("fwav_version", C.c_char_p, []),
    ("fwav_device_count", C.c_int, []),
    ("fwav_ctx_create", C.c_int, [C.c_int, C.POINTER(c_ctx)]),
    ("fwav_ctx_destroy", C.c_int, [c_ctx]),
    ("fwav_last_error", C.c_char_p, [c_ctx]),
    ("fwav_ctx_sync", C.c_int, [c_ctx]),
    ("fwav_ctx_set_search_impl", C.c_int, [c_ctx, C.c_int]),
    ("fwav_ctx_set_embedding", C.c_int, [c_ctx, C.c_int]),
    ("fwav_ctx_set_search_range_size", C.c_int, [c_ctx, C.c_int]),
    ("fwav_ctx_launch_count", i64, [c_ctx]),
    ("fwav_ctx_search_fallbacks", i64, [c_ctx]),
    ("fwav_ctx_search_timings", C.c_int, [c_ctx, C.POINTER(C.c_float)]),
    ("fwav_ctx_search_route", C.c_int, [c_ctx]),
    ("fwav_geometry", C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    ("fwav_count_domains", i64, [i64, C.c_int, C.c_int]),
    ("fwav_build_domains", C.c_int, [c_ctx, c_ptr, i64, C.c_int, C.c_int, C.c_int, c_ptr, c_ptr]),
    ("fwav_embed", C.c_int, [c_ctx, c_ptr, i64, C.c_int, C.c_int, c_ptr, c_ptr]),
    ("fwav_build_tables", C.c_int, [c_ctx, c_ptr, i64, C.c_int, C.c_int, C.c_int, C.c_int, c_ptr, c_ptr, c_ptr]),
    ("fwav_topk", C.c_int, [c_ctx, c_ptr, i64, c_ptr, i64, C.c_int, C.c_int, c_ptr, c_ptr, c_ptr, c_ptr]),
    ("fwav_range_activity", C.c_int, [c_ctx, c_ptr, i64, C.c_int, C.c_double, C.c_int, c_ptr, c_ptr]),
    ("fwav_affine_match", C.c_int, [c_ctx, c_ptr, i64, C.c_int, c_ptr, i64, c_ptr, C.c_int, C.c_double,
                                    c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    ("fwav_decode", C.c_int, [c_ctx, c_ptr, i64, c_ptr, c_ptr, c_ptr, c_ptr, i64, C.c_int, C.c_int,
                              C.c_double, C.c_double, C.c_double, c_ptr,
                              C.POINTER(C.c_int), C.POINTER(C.c_float), c_ptr]),
    ("fwav_decode_iter", C.c_int, [c_ctx, c_ptr, i64, c_ptr, c_ptr, c_ptr, c_ptr, i64, C.c_int,
                                   C.c_double, C.c_double, C.c_int, c_ptr, c_ptr, c_ptr, c_ptr]),
    ("fwav_decode_iter_gated", C.c_int, [c_ctx, c_ptr, i64, c_ptr, c_ptr, c_ptr, c_ptr, i64, C.c_int,
                                         C.c_double, C.c_double, C.c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    ("fwav_decode_converge", C.c_int, [c_ctx, c_ptr, C.c_int, C.c_double, c_ptr, c_ptr]),
    ("fwav_decode_iter_bcast", C.c_int, [c_ctx, c_ptr, i64, c_ptr, c_ptr, c_ptr, c_ptr, i64, C.c_int,
                                         C.c_double, C.c_double, C.c_int, c_ptr, c_ptr, c_ptr, c_ptr,
                                         C.POINTER(c_ptr), C.c_int, C.c_int, i64, c_ptr]),
    ("fwav_compress_device", C.c_int, [c_ctx, c_ptr, i64, c_ptr, i64, i64, C.c_int, C.c_int, C.c_int,
                                       C.c_double, C.c_int, C.c_int, C.c_int, c_ptr, c_ptr,
                                       c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    ("fwav_compress_host", C.c_int, [c_ctx, c_ptr, i64, c_ptr, i64, C.c_int, C.c_int, C.c_int, C.c_double,
                                     C.c_int, C.c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    ("fwav_prepare_ranges", C.c_int, [c_ctx, c_ptr, i64, C.c_int, C.c_double, c_ptr, c_ptr, c_ptr]),
    ("fwav_compress_signal_host", C.c_int, [c_ctx, c_ptr, i64, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int,
                                            c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, C.POINTER(C.c_int)]),
    ("fwav_decode_host", C.c_int, [c_ctx, c_ptr, i64, c_ptr, c_ptr, c_ptr, c_ptr, i64, C.c_int, C.c_int,
                                   C.c_double, C.c_double, C.c_double, c_ptr,
                                   C.POINTER(C.c_int), C.POINTER(C.c_float)]),
    ("fwav_malloc", C.c_int, [c_ctx, i64, C.POINTER(c_ptr)]),
    ("fwav_free", C.c_int, [c_ctx, c_ptr]),
    ("fwav_host_alloc", C.c_int, [c_ctx, i64, C.POINTER(c_ptr)]),
    ("fwav_host_free", C.c_int, [c_ctx, c_ptr]),
    ("fwav_memcpy_h2d", C.c_int, [c_ctx, c_ptr, c_ptr, i64, c_ptr]),
    ("fwav_memcpy_d2h", C.c_int, [c_ctx, c_ptr, c_ptr, i64, c_ptr]),
]

_lib = None
_lock = threading.Lock()

# page-locked blocks waiting for reuse: {(pid, nbytes): [address, ...]} (a forked child must not touch its parent's)
_POOL = {}
_POOL_PER_SIZE = 2
_POOL_MAX_IDLE = 8 << 30


def _pool_take(nbytes):
    with _lock:
        lst = _POOL.get((os.getpid(), nbytes))
        return lst.pop() if lst else None


def _pool_give(lib, pid, nbytes, addr):
    if pid != os.getpid():
        return
    with _lock:
        lst = _POOL.setdefault((pid, nbytes), [])
        idle = sum(k[1] * len(v) for k, v in _POOL.items() if k[0] == pid)
        if len(lst) < _POOL_PER_SIZE and idle + nbytes <= _POOL_MAX_IDLE:
            lst.append(addr)
            return
    lib.fwav_host_free(None, addr)


def load_library():
    """dlopen libfwav_b200.so and type every entry point (no CUDA call yet)."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise FwavError(
                    f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` "
                    "(or make -C audio-compression_b200/csrc). There is no CPU fallback.")
            lib = C.CDLL(LIB_PATH)
            for name, res, args in SIGNATURES:
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def geometry(tile_size):
    lib = load_library()
    n, d = C.c_int(), C.c_int()
    lib.fwav_geometry(int(tile_size), C.byref(n), C.byref(d))
    return n.value, d.value


def count_domains(n_samples, tile_size, domain_step):
    return int(load_library().fwav_count_domains(int(n_samples), int(tile_size), int(domain_step)))


def _hp(a):
    """Host pointer of a C-contiguous numpy array (or None)."""
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _as(a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


class DeviceBuffer:
    """cudaMalloc'ed block owned through the C ABI; used by the per-kernel
    entry points when the caller has no tensor library of its own."""

    def __init__(self, ctx, nbytes):
        self.ctx = ctx
        self.nbytes = int(nbytes)
        p = c_ptr()
        ctx._check(ctx.lib.fwav_malloc(ctx.h, self.nbytes, C.byref(p)))
        self.ptr = p.value

    @classmethod
    def from_host(cls, ctx, arr):
        arr = np.ascontiguousarray(arr)
        buf = cls(ctx, max(arr.nbytes, 16))
        if arr.nbytes:
            ctx._check(ctx.lib.fwav_memcpy_h2d(ctx.h, buf.ptr, _hp(arr), arr.nbytes, None))
        buf.shape, buf.dtype = arr.shape, arr.dtype
        return buf

    def to_host(self, shape, dtype):
        out = np.empty(shape, dtype=dtype)
        if out.nbytes:
            self.ctx._check(self.ctx.lib.fwav_memcpy_d2h(self.ctx.h, _hp(out), self.ptr, out.nbytes, None))
        return out

    def free(self):
        if self.ptr:
            self.ctx.lib.fwav_free(self.ctx.h, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """One fwav_ctx: a device, its stream and scratch workspace."""

    def __init__(self, device=0):
        self.lib = load_library()
        if self.lib.fwav_device_count() <= 0:
            raise FwavError("no CUDA device visible: the FWAV hot path runs on B200 only "
                            "(there is no CPU fallback)")
        h = c_ctx()
        rc = self.lib.fwav_ctx_create(int(device), C.byref(h))
        if rc != 0:
            raise FwavError(f"fwav_ctx_create(device={device}) failed with {rc}")
        self.h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "h", None):
            self.lib.fwav_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            msg = self.lib.fwav_last_error(self.h)
            text = msg.decode("utf-8", "replace") if msg else ""
            if rc == -1 and text.startswith("mmap length is greater than file size"):
                raise ValueError(text)      # what np.memmap raises in the reference (fractal.py:1190)
            if rc == -1 and text.startswith("index out of bounds"):
                raise IndexError(text)      # what the reference's fancy index raises (fractal.py:1414)
            raise FwavError(f"fwav error {rc}: {text}")

    # ---- control ----
    def sync(self):
        self._check(self.lib.fwav_ctx_sync(self.h))

    def set_search_impl(self, impl):
        self._check(self.lib.fwav_ctx_set_search_impl(self.h, int(impl)))

    def set_embedding(self, kind):
        """EMBED_TWO_HEAD (the reference's live path) or EMBED_TONAL (tile_embedding with k = emb_dim)."""
        self._check(self.lib.fwav_ctx_set_embedding(self.h, int(kind)))

    def set_search_range_size(self, range_size):
        self._check(self.lib.fwav_ctx_set_search_range_size(self.h, int(range_size)))

    def launch_count(self):
        return int(self.lib.fwav_ctx_launch_count(self.h))

    def search_fallbacks(self):
        return int(self.lib.fwav_ctx_search_fallbacks(self.h))

    def search_route(self):
        """0 list kernel, 1 full split, 2 hi*hi / float32 accumulators, 3 hi*hi / fp16 accumulators (last batch)."""
        return int(self.lib.fwav_ctx_search_route(self.h))

    def search_timings(self):
        """Device milliseconds of the last tensor-core search: pack, threshold pass, collect pass,
        finalize, exact list kernel (call after synchronising the stream)."""
        ms = (C.c_float * 5)()
        self._check(self.lib.fwav_ctx_search_timings(self.h, ms))
        return dict(zip(("pack", "threshold", "collect", "finalize", "lists"), [float(x) for x in ms]))

    def upload(self, arr):
        return DeviceBuffer.from_host(self, arr)

    def alloc(self, nbytes):
        return DeviceBuffer(self, nbytes)

    # ---- host-buffer entry points (what fractal.compress_audio / decompress_audio call) ----
    def pinned_empty(self, shape, dtype):
        """Output array of the host-buffer entry points.  Default: a plain pageable array -- the C side stages it
        through the context's page-locked ring, the domain table downloading on a helper thread beside the search
        (config 2 through compress_audio_arrays on B200: 83 ms per call against 78 ms of kernels).  FWAV_PINNED=1:
        page-locked blocks that are the end points of the asynchronous copies themselves; page-locking fresh memory
        costs more than it saves (290-510 ms per call), so the blocks are POOLED per process and reused by the next
        call of the same size (91-99 ms per call: the pool still misses while the previous result is alive)."""
        dtype = np.dtype(dtype)
        shape = (shape,) if np.isscalar(shape) else tuple(shape)
        nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
        if nbytes == 0 or os.environ.get("FWAV_PINNED", "0") != "1":
            return np.empty(shape, dtype)
        addr = _pool_take(nbytes)
        if addr is None:
            p = c_ptr()
            if self.lib.fwav_host_alloc(self.h, nbytes, C.byref(p)) != 0 or not p.value:
                return np.empty(shape, dtype)
            addr = p.value
        buf = (C.c_ubyte * nbytes).from_address(addr)
        weakref.finalize(buf, _pool_give, self.lib, os.getpid(), nbytes, addr)   # numpy keeps `buf` alive through .base
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    def compress_host(self, signal, ranges, tile_size, emb_dim, top_k, energy_thresh, fast_mode=True,
                      query_mode=0, want_domains=True, out=None):
        signal = _as(signal, np.float32)
        ranges = _as(ranges, np.float32)
        n_r = ranges.shape[0]
        rs, ds = geometry(tile_size)
        n_d = count_domains(len(signal), tile_size, ds)
        if out is None:
            pe = self.pinned_empty
            out = dict(
                domains=pe((n_d, rs), np.float32) if want_domains else None,
                idx=pe(n_r, np.int32), s=pe(n_r, np.float32), o=pe(n_r, np.float32),
                sym=pe(n_r, np.uint8), err=pe(n_r, np.float32))
        self._check(self.lib.fwav_compress_host(
            self.h, _hp(signal), len(signal), _hp(ranges), n_r, int(tile_size), int(emb_dim), int(top_k),
            float(energy_thresh), int(bool(fast_mode)), int(query_mode), _hp(out["domains"]),
            _hp(out["idx"]), _hp(out["s"]), _hp(out["o"]), _hp(out["sym"]), _hp(out["err"])))
        return out

    def compress_signal_host(self, signal, tile_size, emb_dim, top_k, energy_thresh, fast_mode=True,
                             query_mode=0, want_domains=True, want_ranges=False, out=None):
        """compress from the raw signal (device pre-step).  Returns the dict of compress_host (+ "ranges" on
        request), or None when the reference would return its empty result for a silent input."""
        signal = _as(signal, np.float32)
        rs, ds = geometry(tile_size)
        n_r = -(-len(signal) // rs)
        n_d = count_domains(len(signal), tile_size, ds)
        if out is None:
            pe = self.pinned_empty
            out = dict(
                domains=pe((n_d, rs), np.float32) if want_domains else None,
                ranges=pe((n_r, rs), np.float32) if want_ranges else None,
                idx=pe(n_r, np.int32), s=pe(n_r, np.float32), o=pe(n_r, np.float32),
                sym=pe(n_r, np.uint8), err=pe(n_r, np.float32))
        silent = C.c_int(0)
        self._check(self.lib.fwav_compress_signal_host(
            self.h, _hp(signal), len(signal), int(tile_size), int(emb_dim), int(top_k), float(energy_thresh),
            int(bool(fast_mode)), int(query_mode), _hp(out.get("ranges")), _hp(out.get("domains")),
            _hp(out["idx"]), _hp(out["s"]), _hp(out["o"]), _hp(out["sym"]), _hp(out["err"]), C.byref(silent)))
        return None if silent.value else out

    def prepare_ranges(self, d_signal, n, range_size, energy_thresh, d_ranges, d_sumsq, stream=None):
        self._check(self.lib.fwav_prepare_ranges(self.h, d_signal, n, int(range_size), float(energy_thresh),
                                                 d_ranges, d_sumsq, stream))

    def decode_host(self, domains, idx, s, o, sym, range_size, iterations=8, convergence_eps=1e-3,
                    s_clip=16.0, s_damping=0.0, out=None):
        domains = _as(domains, np.float32)
        idx, s, o, sym = _as(idx, np.int32), _as(s, np.float32), _as(o, np.float32), _as(sym, np.uint8)
        n_r = len(idx)
        # what the reference's numpy indexing would reject (fractal.py:1391-1414) must not reach the device
        if domains.ndim != 2 or domains.shape[1] != int(range_size):
            raise ValueError(f"domains must be (n_domains, {int(range_size)}), got {domains.shape}")
        if not (len(s) == len(o) == len(sym) == n_r):
            raise ValueError(f"match arrays differ in length: idx {n_r}, s {len(s)}, o {len(o)}, sym {len(sym)}")
        if n_r and int(idx.max()) >= domains.shape[0]:
            raise IndexError(f"index out of bounds: match {int(idx.argmax())} points at domain {int(idx.max())} "
                             f"of {domains.shape[0]}")
        if out is None:
            out = np.empty(n_r * range_size, np.float32)
        it, delta = C.c_int(0), C.c_float(0)
        self._check(self.lib.fwav_decode_host(
            self.h, _hp(domains), domains.shape[0], _hp(idx), _hp(s), _hp(o), _hp(sym), n_r,
            int(range_size), int(iterations), float(convergence_eps), float(s_clip), float(s_damping),
            _hp(out), C.byref(it), C.byref(delta)))
        return out, it.value, delta.value

    # ---- device-pointer entry points (pointers are ints; stream is an int or None) ----
    def build_domains(self, d_signal, n, tile_size, range_size, domain_step, d_domains, stream=None):
        self._check(self.lib.fwav_build_domains(self.h, d_signal, n, tile_size, range_size, domain_step,
                                                d_domains, stream))

    def embed(self, d_rows, rows, range_size, emb_dim, d_emb, stream=None):
        self._check(self.lib.fwav_embed(self.h, d_rows, rows, range_size, emb_dim, d_emb, stream))

    def build_tables(self, d_signal, n, tile_size, range_size, domain_step, emb_dim, d_domains, d_emb, stream=None):
        """domains + embeddings in one pass (fwav_build_tables)"""
        self._check(self.lib.fwav_build_tables(self.h, d_signal, n, tile_size, range_size, domain_step, emb_dim,
                                               d_domains, d_emb, stream))

    def range_activity(self, d_ranges, n_r, range_size, energy_thresh, fast_mode, d_active, stream=None):
        self._check(self.lib.fwav_range_activity(self.h, d_ranges, n_r, range_size, float(energy_thresh),
                                                 int(bool(fast_mode)), d_active, stream))

    def topk(self, d_q, n_q, d_emb, n_d, emb_dim, top_k, d_active, d_cand, d_scores=None, stream=None):
        self._check(self.lib.fwav_topk(self.h, d_q, n_q, d_emb, n_d, emb_dim, top_k, d_active, d_cand,
                                       d_scores, stream))

    def affine_match(self, d_ranges, n_r, range_size, d_domains, n_d, d_cand, top_k, s_clip,
                     d_idx, d_s, d_o, d_sym, d_err, stream=None):
        self._check(self.lib.fwav_affine_match(self.h, d_ranges, n_r, range_size, d_domains, n_d, d_cand,
                                               top_k, float(s_clip), d_idx, d_s, d_o, d_sym, d_err, stream))

    def decode(self, d_domains, n_d, d_idx, d_s, d_o, d_sym, n_r, range_size, iterations, convergence_eps,
               s_clip, s_damping, d_out, stream=None):
        it, delta = C.c_int(0), C.c_float(0)
        self._check(self.lib.fwav_decode(self.h, d_domains, n_d, d_idx, d_s, d_o, d_sym, n_r, range_size,
                                         int(iterations), float(convergence_eps), float(s_clip),
                                         float(s_damping), d_out, C.byref(it), C.byref(delta), stream))
        return it.value, delta.value

    def decode_iter(self, d_domains, n_d, d_idx, d_s, d_o, d_sym, n_r, range_size, s_clip, s_damping,
                    first, d_cur, d_next, d_sums, stream=None):
        self._check(self.lib.fwav_decode_iter(self.h, d_domains, n_d, d_idx, d_s, d_o, d_sym, n_r, range_size,
                                              float(s_clip), float(s_damping), int(bool(first)), d_cur, d_next,
                                              d_sums, stream))

    def decode_iter_gated(self, d_domains, n_d, d_idx, d_s, d_o, d_sym, n_r, range_size, s_clip, s_damping,
                          first, d_cur, d_next, d_sums, d_state, stream=None):
        self._check(self.lib.fwav_decode_iter_gated(self.h, d_domains, n_d, d_idx, d_s, d_o, d_sym, n_r, range_size,
                                                    float(s_clip), float(s_damping), int(bool(first)), d_cur, d_next,
                                                    d_sums, d_state, stream))

    def decode_iter_bcast(self, d_domains, n_d, d_idx, d_s, d_o, d_sym, n_r, range_size, s_clip, s_damping,
                          first, d_cur, d_next, d_sums, d_state, targets, multimem, target_offset, stream=None):
        arr = (c_ptr * len(targets))(*[int(t) for t in targets])
        self._check(self.lib.fwav_decode_iter_bcast(self.h, d_domains, n_d, d_idx, d_s, d_o, d_sym, n_r, range_size,
                                                    float(s_clip), float(s_damping), int(bool(first)), d_cur, d_next,
                                                    d_sums, d_state, arr, len(targets), int(bool(multimem)),
                                                    int(target_offset), stream))

    def decode_converge(self, d_sums_all, n_parts, eps, d_state, stream=None):
        self._check(self.lib.fwav_decode_converge(self.h, d_sums_all, int(n_parts), float(eps), d_state, stream))

    def compress_device(self, d_signal, n, d_ranges, n_r, query_offset, tile_size, emb_dim, top_k,
                        energy_thresh, fast_mode, query_mode, build, d_domains, d_emb,
                        d_idx, d_s, d_o, d_sym, d_err, stream=None):
        self._check(self.lib.fwav_compress_device(
            self.h, d_signal, n, d_ranges, n_r, query_offset, int(tile_size), int(emb_dim), int(top_k),
            float(energy_thresh), int(bool(fast_mode)), int(query_mode), int(bool(build)),
            d_domains, d_emb, d_idx, d_s, d_o, d_sym, d_err, stream))


_default = {}


def default_context(device=0):
    """Process-wide context per device, created on first use (fork-safe: the
    cache is keyed by pid so a forked worker builds its own)."""
    key = (os.getpid(), int(device))
    ctx = _default.get(key)
    if ctx is None:
        ctx = Context(device)
        _default[key] = ctx
    return ctx
