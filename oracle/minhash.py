"""MinHash signatures and LSH banding.  TEST INFRASTRUCTURE - see oracle/__init__.py.

minhash(): the spec's minhash_compute (README.md:2578-2597): for every 4-byte
shingle (little-endian u32 at each byte offset, stride 1) and every seed,
MurmurHash3_x86_32(shingle, seed); keep the running minimum starting from
0xFFFFFFFF.  Chunks shorter than 4 bytes have no shingles (the skeleton's
`len - 3` underflow, SURVEY.md §0.2 C10, resolved to "all 0xFFFFFFFF").
Seeds default to 1..128 (VALIDATION_METHODS.md:122).

band_keys(): banding per README.md:2231-2235 - signature split into `bands`
groups of `rows` values; two chunks are candidates when any band matches.  The
band key is FNV-1a-64 over the band's rows*4 little-endian bytes (the spec's
15/16-bit keys, README.md:1909-1913, 1978, are an SD-card budget artefact;
SURVEY.md §0.2 C7).

buckets(): every (band, key, id) triple sorted lexicographically; equal
(band, key) runs are the LSH buckets.
"""
from __future__ import annotations

import numpy as np

from .config import SimConfig

_C1, _C2 = np.uint32(0xCC9E2D51), np.uint32(0x1B873593)
FNV_OFFSET, FNV_PRIME = 0xCBF29CE484222325, 0x100000001B3


def murmur3_32(key: bytes, seed: int) -> int:
    """Public MurmurHash3_x86_32 (Appleby), any key length - pure Python, for KATs."""
    M = 0xFFFFFFFF
    rotl = lambda x, r: ((x << r) | (x >> (32 - r))) & M  # noqa: E731
    h = seed & M
    nb = len(key) // 4
    for i in range(nb):
        k = int.from_bytes(key[4 * i:4 * i + 4], "little")
        k = (k * 0xCC9E2D51) & M
        k = rotl(k, 15)
        k = (k * 0x1B873593) & M
        h ^= k
        h = rotl(h, 13)
        h = (h * 5 + 0xE6546B64) & M
    tail = key[4 * nb:]
    k = 0
    if len(tail) >= 3:
        k ^= tail[2] << 16
    if len(tail) >= 2:
        k ^= tail[1] << 8
    if len(tail) >= 1:
        k ^= tail[0]
        k = (k * 0xCC9E2D51) & M
        k = rotl(k, 15)
        k = (k * 0x1B873593) & M
        h ^= k
    h ^= len(key)
    h ^= h >> 16
    h = (h * 0x85EBCA6B) & M
    h ^= h >> 13
    h = (h * 0xC2B2AE35) & M
    h ^= h >> 16
    return h


def _shingles(d: np.ndarray) -> np.ndarray:
    n = d.size
    if n < 4:
        return np.zeros(0, dtype=np.uint32)
    b = d.astype(np.uint32)
    return b[:n - 3] | (b[1:n - 2] << np.uint32(8)) | (b[2:n - 1] << np.uint32(16)) | (b[3:] << np.uint32(24))


def _rotl(x, r):
    return (x << np.uint32(r)) | (x >> np.uint32(32 - r))


def minhash_chunk(d: np.ndarray, seeds: np.ndarray) -> np.ndarray:
    k = _shingles(d)
    out = np.full(seeds.size, 0xFFFFFFFF, dtype=np.uint32)
    if k.size == 0:
        return out
    k = _rotl(k * _C1, 15) * _C2  # seed-independent
    for p, sd in enumerate(seeds):
        h = np.uint32(sd) ^ k
        h = _rotl(h, 13) * np.uint32(5) + np.uint32(0xE6546B64)
        h ^= np.uint32(4)
        h ^= h >> np.uint32(16)
        h *= np.uint32(0x85EBCA6B)
        h ^= h >> np.uint32(13)
        h *= np.uint32(0xC2B2AE35)
        h ^= h >> np.uint32(16)
        out[p] = h.min()
    return out


def minhash(data, cuts, cfg: SimConfig = SimConfig(), start0: int = 0) -> np.ndarray:
    d = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    cuts = np.asarray(cuts, dtype=np.uint64)
    seeds = cfg.seed_array
    sig = np.empty((cuts.size, cfg.n_perm), dtype=np.uint32)
    s = int(start0)
    for j, e in enumerate(cuts.tolist()):
        sig[j] = minhash_chunk(d[s:e], seeds)
        s = e
    return sig


def minhash_c(data, cuts, cfg: SimConfig = SimConfig(), start0: int = 0) -> np.ndarray:
    """Same result through oracle/hmse_ref.c (fast enough for MiB-scale parity inputs)."""
    from .cdc import ref_lib
    d = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data)
    cuts = np.ascontiguousarray(cuts, dtype=np.uint64)
    seeds = cfg.seed_array
    sig = np.empty((cuts.size, cfg.n_perm), dtype=np.uint32)
    ref_lib().hmse_ref_minhash(d.ctypes.data, int(start0), cuts.ctypes.data, cuts.size,
                               seeds.ctypes.data, cfg.n_perm, sig.ctypes.data)
    return sig


def band_keys(sig: np.ndarray, cfg: SimConfig = SimConfig()) -> np.ndarray:
    sig = np.ascontiguousarray(sig, dtype=np.uint32).reshape(-1, cfg.n_perm)
    n = sig.shape[0]
    by = sig.view(np.uint8).reshape(n, cfg.bands, cfg.rows * 4).astype(np.uint64)
    h = np.full((n, cfg.bands), FNV_OFFSET, dtype=np.uint64)
    prime = np.uint64(FNV_PRIME)
    for i in range(cfg.rows * 4):
        h = (h ^ by[:, :, i]) * prime
    return h


def buckets(keys: np.ndarray, id_base: int = 0):
    """(band uint32[n*b], key uint64[n*b], id uint64[n*b]) sorted by (band, key, id)."""
    n, b = keys.shape
    band = np.tile(np.arange(b, dtype=np.uint32), n)
    ids = np.repeat(np.arange(n, dtype=np.uint64) + np.uint64(id_base), b)
    k = keys.reshape(-1)
    order = np.lexsort((ids, k, band))
    return band[order], k[order], ids[order]


def similarity(data, cuts, cfg: SimConfig = SimConfig(), start0: int = 0, use_c: bool = False):
    sig = (minhash_c if use_c else minhash)(data, cuts, cfg, start0)
    keys = band_keys(sig, cfg)
    return sig, keys, buckets(keys)
