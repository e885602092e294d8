"""Per-chunk preset-dictionary DEFLATE.  TEST INFRASTRUCTURE - see oracle/__init__.py.

The spec's L1 layer (README.md:288, 1159-1198; skeleton call
mz_deflateInit2(..., window_bits=15, ...) at README.md:2374) resolved per
SURVEY.md §0.2 C5: one RFC 1950 (zlib) stream per selected chunk, window 32 KiB,
preset dictionary (FDICT), level 6.  Streams are compared by inflate-equality and
total size only - never byte for byte.
"""
from __future__ import annotations

import zlib

import numpy as np


def compress(data, cuts, select, zdict: bytes = b"", level: int = 6, start0: int = 0):
    """Returns (blob uint8[...], offsets uint64[m+1]); slice j is the zlib stream of chunk select[j]."""
    mv = memoryview(np.ascontiguousarray(data)) if isinstance(data, np.ndarray) else memoryview(data).cast("B")
    cuts = np.asarray(cuts, dtype=np.uint64)
    starts = np.concatenate([[np.uint64(start0)], cuts[:-1]]) if cuts.size else cuts
    parts, offs = [], [0]
    for j in np.asarray(select, dtype=np.int64).tolist():
        if zdict:
            co = zlib.compressobj(level, zlib.DEFLATED, 15, 8, zlib.Z_DEFAULT_STRATEGY, zdict)
        else:
            co = zlib.compressobj(level, zlib.DEFLATED, 15, 8, zlib.Z_DEFAULT_STRATEGY)
        s = co.compress(mv[int(starts[j]):int(cuts[j])]) + co.flush()
        parts.append(s)
        offs.append(offs[-1] + len(s))
    blob = np.frombuffer(b"".join(parts), dtype=np.uint8) if parts else np.zeros(0, dtype=np.uint8)
    return blob, np.array(offs, dtype=np.uint64)


def inflate_all(blob, offsets, zdict: bytes = b""):
    """Inflates every stream with host zlib (the spec's read path, README.md:1638-1640)."""
    raw = bytes(memoryview(np.ascontiguousarray(blob)))
    offs = np.asarray(offsets, dtype=np.uint64).tolist()
    out = []
    for a, b in zip(offs[:-1], offs[1:]):
        do = zlib.decompressobj(15, zdict) if zdict else zlib.decompressobj(15)
        chunk = do.decompress(raw[a:b]) + do.flush()
        if not do.eof or do.unused_data:
            raise ValueError("stream %d..%d is not exactly one complete zlib stream" % (a, b))
        out.append(chunk)
    return out


def make_zdict(samples, size: int = 32768) -> bytes:
    """Deterministic preset dictionary: fragments ordered rarest-first so the most
    frequent text sits at the END (shortest distances), truncated to `size` bytes.
    `samples` is an iterable of (bytes, weight)."""
    seen, frags = set(), []
    for frag, w in samples:
        if frag and frag not in seen:
            seen.add(frag)
            frags.append((w, len(frags), frag))
    frags.sort(key=lambda t: (t[0], -t[1]))
    blob = b"".join(f for _, _, f in frags)
    return blob[-size:]
