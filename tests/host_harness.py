"""ctypes loader for tests/csrc/libhost_harness.so (TEST INFRASTRUCTURE): the
host build of the kernels' per-element math.  Built on demand with g++."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc", "host_harness.cpp")
LIB = os.path.join(HERE, "csrc", "libhost_harness.so")
INC = os.path.join(ROOT, "audio-compression_b200", "csrc")


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [SRC] + [os.path.join(INC, f) for f in ("np_math.cuh", "fwav_math.cuh", "embed_tables.h", "tables_geom.h", "embed_static.cuh")]
    return any(os.path.getmtime(d) > t for d in deps)


def load():
    if _stale():
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared",
                               "-I" + INC, SRC, "-o", LIB])
    lib = C.CDLL(LIB)
    lib.hh_np_mean.restype = C.c_float
    lib.hh_score.restype = C.c_float
    lib.hh_decode.restype = C.c_int
    lib.hh_prestep.restype = C.c_double
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class Harness:
    def __init__(self):
        self.lib = load()

    def embed_tonal(self, rows, k=32):
        rows = f32(rows)
        out = np.empty((rows.shape[0], k), np.float32)
        self.lib.hh_embed_tonal(_p(rows), C.c_longlong(rows.shape[0]), C.c_int(rows.shape[1]), C.c_int(k), _p(out))
        return out

    def prestep(self, sig, N, thr=1e-4):
        sig = f32(sig)
        n = len(sig)
        mask = np.empty(n, np.uint8)
        ranges = np.empty((-(-n // N), N), np.float32)
        ssq = self.lib.hh_prestep(_p(sig), C.c_longlong(n), C.c_int(N), C.c_double(thr), _p(mask), _p(ranges))
        return mask, ranges, float(ssq)

    def np_mean(self, a):
        a = f32(a)
        return np.float32(self.lib.hh_np_mean(_p(a), C.c_int(len(a))))

    def build_domains(self, sig, tile, N, ds, force_generic=False):
        sig = f32(sig)
        nd = 0 if len(sig) < tile else (len(sig) - tile) // ds + 1
        out = np.empty((nd, N), np.float32)
        self.lib.hh_build_domains(_p(sig), C.c_longlong(len(sig)), C.c_int(tile), C.c_int(N), C.c_int(ds),
                                  _p(out), C.c_int(int(force_generic)))
        return out

    def build_domains_chain(self, sig, tile, N, ds):
        """the route of tables.cu: chain half sums at every ds-th sample, domain rows from a staged window"""
        sig = f32(sig)
        n = len(sig)
        nd = 0 if n < tile else (n - tile) // ds + 1
        n_half = (n - 128) // ds + 1
        half = np.full(max(n_half, 1), np.nan, np.float32)
        self.lib.hh_half_sums_chain(_p(sig), C.c_longlong(n), C.c_longlong(n_half), C.c_int(ds), _p(half))
        out = np.empty((nd, N), np.float32)
        self.lib.hh_domains_from_halves(_p(half), C.c_longlong(n_half), C.c_longlong(nd), C.c_int(N), C.c_int(ds), _p(out))
        return out, half

    def embed(self, rows, emb_dim=16):
        rows = f32(rows)
        out = np.empty((rows.shape[0], emb_dim), np.float32)
        self.lib.hh_embed(_p(rows), C.c_longlong(rows.shape[0]), C.c_int(rows.shape[1]), C.c_int(emb_dim), _p(out))
        return out

    def embed_static(self, rows):
        """embed_static.cuh (the kernels' form for range_size 4 / 8 / 16 / 32 at emb_dim 16); None for other sizes"""
        rows = f32(rows)
        out = np.empty((rows.shape[0], 16), np.float32)
        if self.lib.hh_embed_static(_p(rows), C.c_longlong(rows.shape[0]), C.c_int(rows.shape[1]), _p(out)):
            return None
        return out

    def activity(self, ranges, thr, fast=True):
        ranges = f32(ranges)
        out = np.empty(ranges.shape[0], np.uint8)
        self.lib.hh_activity(_p(ranges), C.c_longlong(ranges.shape[0]), C.c_int(ranges.shape[1]),
                             C.c_double(thr), C.c_int(int(fast)), _p(out))
        return out

    def affine(self, ranges, domains, cand, s_clip=16.0, pair=False):
        """pair=True: the kernel's form (fwm::affine_fit_pair: tile statistics shared between the orientations)"""
        ranges, domains = f32(ranges), f32(domains)
        cand = np.ascontiguousarray(cand, np.int32)
        n = ranges.shape[0]
        idx = np.empty(n, np.int32); s = np.empty(n, np.float32); o = np.empty(n, np.float32)
        sym = np.empty(n, np.uint8); err = np.empty(n, np.float32)
        (self.lib.hh_affine_pair if pair else self.lib.hh_affine)(_p(ranges), C.c_longlong(n), C.c_int(ranges.shape[1]), _p(domains), _p(cand),
                           C.c_int(cand.shape[1]), C.c_double(s_clip), _p(idx), _p(s), _p(o), _p(sym), _p(err))
        return dict(idx=idx, s=s, o=o, sym=sym, err=err)

    def decode(self, domains, idx, s, o, sym, N, iterations=8, eps=1e-3, s_clip=16.0, s_damping=0.0):
        domains = f32(domains)
        idx = np.ascontiguousarray(idx, np.int32); s = f32(s); o = f32(o)
        sym = np.ascontiguousarray(sym, np.uint8)
        n = len(idx)
        out = np.empty(n * N, np.float32)
        delta = C.c_float(0)
        it = self.lib.hh_decode(_p(domains), _p(idx), _p(s), _p(o), _p(sym), C.c_longlong(n), C.c_int(N),
                                C.c_int(iterations), C.c_double(eps), C.c_double(s_clip), C.c_double(s_damping),
                                _p(out), C.byref(delta))
        return out, it, delta.value
