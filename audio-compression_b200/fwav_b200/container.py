"""The .fwav container (reference fractal.py:1278-1375), byte-identical, without
the per-row Python loops.

Layout: 'FWAV', u8 version, u32 range_size, u32 framerate, u8 sampwidth,
u16 tile_size, u16 domain_step, f32 energy_threshold, u32 n_ranges,
u32 n_domains, u32 original_len (34 bytes); 32-byte SHA-256 of everything that
follows; n_domains*range_size f32; n_ranges 17-byte records
(i32 domain, f32 s, f32 o, u8 sym, f32 err).  All little-endian.
"""
from __future__ import annotations

import gc
import hashlib
import struct

import numpy as np

FWAV_VERSION = 1
HEADER = struct.Struct("<4sBIIBHHfIII")
RECORD = np.dtype([("idx", "<i4"), ("s", "<f4"), ("o", "<f4"), ("sym", "u1"), ("err", "<f4")])
assert HEADER.size == 34 and RECORD.itemsize == 17
_CHUNK = 1 << 24


class MatchArrays:
    """Struct-of-arrays view of the match list.  Behaves like the reference's
    list of (idx, s, o, sym, err) tuples (len, indexing, iteration) while
    letting compress -> save and load -> decompress skip the Python tuples
    (SURVEY.md §8(f) row N3)."""

    __slots__ = ("idx", "s", "o", "sym", "err")

    def __init__(self, idx, s, o, sym, err):
        self.idx = np.ascontiguousarray(idx, dtype=np.int32)
        self.s = np.ascontiguousarray(s, dtype=np.float32)
        self.o = np.ascontiguousarray(o, dtype=np.float32)
        self.sym = np.ascontiguousarray(sym, dtype=np.uint8)
        self.err = np.ascontiguousarray(err, dtype=np.float32)

    @classmethod
    def from_any(cls, matches):
        if isinstance(matches, cls):
            return matches
        n = len(matches)
        if n == 0:
            z = np.zeros(0)
            return cls(z, z, z, z, z)
        cols = list(zip(*matches))
        return cls(np.asarray(cols[0], dtype=np.int64).astype(np.int32),
                   np.asarray(cols[1], dtype=np.float64).astype(np.float32),
                   np.asarray(cols[2], dtype=np.float64).astype(np.float32),
                   np.asarray(cols[3], dtype=np.int64).astype(np.uint8),
                   np.asarray(cols[4], dtype=np.float64).astype(np.float32))

    def __len__(self):
        return len(self.idx)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return self.tolist()[i]
        return (int(self.idx[i]), float(self.s[i]), float(self.o[i]), int(self.sym[i]), float(self.err[i]))

    def __iter__(self):
        return iter(self.tolist())

    def __eq__(self, other):
        try:
            return self.tolist() == list(other)
        except TypeError:
            return NotImplemented

    def tolist(self):
        """The reference's representation: Python ints/floats widened from f32
        (fractal.py:836-845)."""
        # half a million tuples are 2.5 M new objects: with the cyclic collector running the build costs 175-440 ms
        # (it rescans the growing list again and again), paused 140 ms; nothing built here can form a cycle
        was_enabled = gc.isenabled()
        gc.disable()
        try:
            return list(zip(self.idx.tolist(), self.s.tolist(), self.o.tolist(),
                            self.sym.tolist(), self.err.tolist()))
        finally:
            if was_enabled:
                gc.enable()

    def records(self):
        rec = np.empty(len(self), dtype=RECORD)
        rec["idx"], rec["s"], rec["o"], rec["sym"], rec["err"] = self.idx, self.s, self.o, self.sym, self.err
        return rec


def save_compressed(filepath, matches, domains_array, range_size, framerate, sampwidth,
                    tile_size, domain_step, energy_threshold, original_len):
    m = MatchArrays.from_any(matches)
    domains = np.ascontiguousarray(domains_array, dtype="<f4")
    dom_bytes = memoryview(domains).cast("B") if domains.size else b""
    rec_bytes = memoryview(m.records()).cast("B") if len(m) else b""
    sha = hashlib.sha256()
    head = HEADER.pack(b"FWAV", FWAV_VERSION, range_size, framerate, sampwidth, tile_size,
                       domain_step, energy_threshold, len(m), len(domains), original_len)
    with open(filepath, "wb") as f:
        f.write(head)
        f.write(bytes(32))
        for blob in (dom_bytes, rec_bytes):
            for lo in range(0, len(blob), _CHUNK):
                piece = blob[lo:lo + _CHUNK]
                f.write(piece)
                sha.update(piece)
        f.seek(HEADER.size)
        f.write(sha.digest())


def load_compressed(filepath, verify_checksum=True, as_arrays=False):
    with open(filepath, "rb") as f:
        head = f.read(HEADER.size)
        if head[:4] != b"FWAV":
            raise ValueError("Not a FWAV file")
        if len(head) < 5 or head[4] != FWAV_VERSION:
            raise ValueError(f"Unsupported FWAV version: {head[4] if len(head) > 4 else None}")
        (_, _, range_size, framerate, sampwidth, tile_size, domain_step, energy_threshold,
         n_ranges, n_domains, original_len) = HEADER.unpack(head)
        digest = f.read(32)
        domains = np.fromfile(f, dtype="<f4", count=n_domains * range_size)
        recs = np.fromfile(f, dtype=RECORD, count=n_ranges)
    if len(domains) != n_domains * range_size or len(recs) != n_ranges:
        raise ValueError("Truncated FWAV file")
    if verify_checksum:
        sha = hashlib.sha256()
        for arr in (domains, recs):
            blob = memoryview(arr).cast("B") if arr.size else b""
            for lo in range(0, len(blob), _CHUNK):
                sha.update(blob[lo:lo + _CHUNK])
        if sha.digest() != digest:
            raise ValueError("Checksum mismatch — file may be corrupted")
    m = MatchArrays(recs["idx"], recs["s"], recs["o"], recs["sym"], recs["err"])
    domains = domains.reshape(n_domains, range_size)
    return (m if as_arrays else m.tolist(), domains, n_ranges, range_size, framerate, sampwidth,
            tile_size, domain_step, energy_threshold, original_len)
