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
