"""SHA-256 digests and exact dedup.  TEST INFRASTRUCTURE - see oracle/__init__.py.

digest(): FIPS 180-4 SHA-256 of each RAW chunk (README.md:290, 2543; the spec's
"hash of the compressed chunk", README.md:1289, is resolved to raw bytes in
SURVEY.md §0.2 C2).
dedup(): the index-lookup rule of README.md:1288-1292 / 1542-1551 - the first
chunk with a digest is stored, every later one becomes a pointer to it.
"""
from __future__ import annotations

import hashlib

import numpy as np


def digest(data, cuts, start0: int = 0) -> np.ndarray:
    mv = memoryview(data).cast("B") if not isinstance(data, np.ndarray) else memoryview(np.ascontiguousarray(data))
    cuts = np.asarray(cuts, dtype=np.uint64)
    out = np.empty((cuts.size, 32), dtype=np.uint8)
    s = int(start0)
    for j, e in enumerate(cuts.tolist()):
        out[j] = np.frombuffer(hashlib.sha256(mv[s:e]).digest(), dtype=np.uint8)
        s = e
    return out


def dedup(digests: np.ndarray):
    """canon[i] = smallest j with digests[j] == digests[i]; is_first[i] = (canon[i] == i)."""
    digests = np.ascontiguousarray(digests, dtype=np.uint8).reshape(-1, 32)
    n = digests.shape[0]
    canon = np.empty(n, dtype=np.int64)
    seen = {}
    raw = digests.tobytes()
    for i in range(n):
        canon[i] = seen.setdefault(raw[32 * i:32 * i + 32], i)
    return canon, canon == np.arange(n, dtype=np.int64)


def digest_mt(data, cuts, start0: int = 0, threads: int = 0) -> np.ndarray:
    """digest() over a thread pool (hashlib releases the GIL for inputs >= 2 KiB, i.e. for every chunk of at least
    min_size): the full-size parity checks of tests/ and of bench.py's `verify` leg.  Same result as digest()."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    mv = memoryview(data).cast("B") if not isinstance(data, np.ndarray) else memoryview(np.ascontiguousarray(data))
    ends = np.asarray(cuts, dtype=np.uint64).astype(np.int64)
    n = ends.size
    out = np.empty((n, 32), dtype=np.uint8)
    if n == 0:
        return out
    starts = np.concatenate([[int(start0)], ends[:-1]]).tolist()
    ends = ends.tolist()
    k = threads or min(32, os.cpu_count() or 1)
    step = max(256, -(-n // (k * 8)))

    def work(lo: int) -> None:
        hi = min(n, lo + step)
        buf = bytearray()
        for j in range(lo, hi):
            buf += hashlib.sha256(mv[starts[j]:ends[j]]).digest()
        out[lo:hi] = np.frombuffer(bytes(buf), dtype=np.uint8).reshape(-1, 32)

    with ThreadPoolExecutor(k) as ex:
        list(ex.map(work, range(0, n, step)))
    return out


def dedup_fast(digests: np.ndarray):
    """dedup() for tens of millions of digests (a sort instead of a Python dict): the same canon / is_first -
    np.unique returns the index of the FIRST occurrence of every distinct row."""
    digests = np.ascontiguousarray(digests, dtype=np.uint8).reshape(-1, 32)
    n = digests.shape[0]
    if n == 0:
        return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=bool)
    keys = digests.view(np.dtype((np.void, 32))).reshape(-1)
    _, first_idx, inv = np.unique(keys, return_index=True, return_inverse=True)
    canon = first_idx[inv.reshape(-1)].astype(np.int64)
    return canon, canon == np.arange(n, dtype=np.int64)
