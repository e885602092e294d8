"""FastCDC oracle.  TEST INFRASTRUCTURE - see oracle/__init__.py.

Restates the paper's Algorithm 1 (FastCDC, Xia et al. ATC'16, the citation the
spec gives at README.md:2753-2755) with the spec's size parameters
(README.md:289, 2444-2446) and min/max clamps (README.md:2482-2483):

    fp = 0; i = MinSize
    if n <= MinSize: return n
    if n >= MaxSize: n = MaxSize  elif n <= NormalSize: NormalSize = n
    for (; i < NormalSize; i++) { fp = (fp<<1) + Gear[src[i]]; if !(fp & MaskS) return i }
    for (; i < n;          i++) { fp = (fp<<1) + Gear[src[i]]; if !(fp & MaskL) return i }
    return i

Cut convention (pinned by tests/test_oracle_cdc.py::test_cut_convention): the
returned i is the chunk LENGTH, i.e. the byte that completed the match starts
the next chunk.  `cuts` are exclusive end offsets; cuts[-1] == len(data).

Three implementations, asserted equal in the tests:
  chunk_naive  literal byte-at-a-time transcription (ground truth, slow)
  chunk        NumPy: full-window candidate scan + per-chunk walk (1 GiB-capable)
  chunk_c      the same loop in C (oracle/hmse_ref.c), for large parity inputs
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from .config import CDCConfig, M64


def _as_u8(data) -> np.ndarray:
    if isinstance(data, np.ndarray):
        if data.dtype != np.uint8:
            raise TypeError("data must be uint8")
        return np.ascontiguousarray(data)
    return np.frombuffer(bytes(data) if not isinstance(data, (bytes, bytearray, memoryview)) else data,
                         dtype=np.uint8)


def next_cut(data: np.ndarray, start: int, n: int, cfg: CDCConfig, gear=None) -> int:
    """One application of Algorithm 1 to data[start:n]; returns the absolute cut."""
    gear = cfg.gear.tolist() if gear is None else gear
    rem = n - start
    if rem <= cfg.min_size:
        return n
    end = start + min(rem, cfg.max_size)
    normal = start + min(cfg.avg_size, end - start)
    fp = 0
    i = start + cfg.min_size
    ms, ml = cfg.mask_s, cfg.mask_l
    while i < normal:
        fp = ((fp << 1) + gear[data[i]]) & M64
        if not (fp & ms):
            return i
        i += 1
    while i < end:
        fp = ((fp << 1) + gear[data[i]]) & M64
        if not (fp & ml):
            return i
        i += 1
    return end


def chunk_naive(data, cfg: CDCConfig = CDCConfig()) -> np.ndarray:
    d = _as_u8(data)
    n = d.size
    lst = d.tolist()
    gear = cfg.gear.tolist()
    cuts = []
    s = 0
    while s < n:
        s = next_cut(lst, s, n, cfg, gear)
        cuts.append(s)
    return np.array(cuts, dtype=np.uint64)


# --------------------------------------------------------------------------- #
# NumPy path: position-only full-window hash, then a per-chunk walk.
# --------------------------------------------------------------------------- #

def full_window_hash(d: np.ndarray, gear: np.ndarray) -> np.ndarray:
    """H(i) = sum_{k=0..min(63,i)} Gear[d[i-k]] << k  (mod 2^64), by doubling:
    H_2m(i) = H_m(i) + (H_m(i-m) << m).  Six passes instead of 64."""
    h = gear[d]
    m = 1
    while m < 64:
        nh = h.copy()
        nh[m:] += h[:-m] << np.uint64(m)
        h = nh
        m *= 2
    return h


def candidates(d: np.ndarray, cfg: CDCConfig, block: int = 1 << 22):
    """Sorted positions where the FULL 64-byte-window hash clears MaskS / MaskL."""
    gear = cfg.gear
    ms, ml = np.uint64(cfg.mask_s), np.uint64(cfg.mask_l)
    n = d.size
    out_s, out_l = [], []
    for b0 in range(0, n, block):
        lo = max(0, b0 - 63)
        h = full_window_hash(d[lo:min(n, b0 + block)], gear)[b0 - lo:]
        out_s.append(np.flatnonzero((h & ms) == 0) + b0)
        out_l.append(np.flatnonzero((h & ml) == 0) + b0)
    cat = lambda xs: np.concatenate(xs) if xs else np.zeros(0, dtype=np.int64)  # noqa: E731
    return cat(out_s).astype(np.int64), cat(out_l).astype(np.int64)


def chunk_shard(data, cfg: CDCConfig, entry: int, n_own: int, eof: bool):
    """Chain of chunk starts s with entry <= s < n_own over data[0:n_avail].

    eof=True : data ends the stream (cuts never exceed len(data); the last is len(data)).
    eof=False: the stream continues past len(data); requires len(data) >= n_own + max_size
               so every next_cut(s) for s < n_own sees a full max_size look-ahead.
    Returns (cuts, exit): cuts = next_cut of every owned start, exit = cuts[-1]
    (the first chunk start >= n_own, or len(data) at eof).
    """
    d = _as_u8(data)
    n = d.size
    if not eof and n < n_own + cfg.max_size:
        raise ValueError("non-final shard needs max_size bytes of look-ahead")
    if eof:
        n_own = n
    cand_s, cand_l = candidates(d, cfg)
    gear = cfg.gear.tolist()
    ms, ml = cfg.mask_s, cfg.mask_l
    mn, av, mx = cfg.min_size, cfg.avg_size, cfg.max_size
    t = cfg.top_bit  # positions s+min .. s+min+t-1 see fewer than t+1 bytes
    cuts = []
    s = entry
    dl = d  # index lazily
    while s < n_own:
        rem = n - s
        if rem <= mn:
            c = n
        else:
            end = s + min(rem, mx)
            normal = s + min(av, end - s)
            c = -1
            # (1) partial-window positions (fp was reset to 0 at s+min)
            fp = 0
            p_end = min(s + mn + t, end)
            seg = dl[s + mn:p_end].tolist()
            i = s + mn
            for b in seg:
                fp = ((fp << 1) + gear[b]) & M64
                if not (fp & (ms if i < normal else ml)):
                    c = i
                    break
                i += 1
            if c < 0:
                # (2) full-window positions: the reset no longer shows in the mask bits
                lo = s + mn + t
                if lo < normal:
                    k = np.searchsorted(cand_s, lo)
                    if k < cand_s.size and cand_s[k] < normal:
                        c = int(cand_s[k])
                if c < 0:
                    lo2 = max(lo, normal)
                    if lo2 < end:
                        k = np.searchsorted(cand_l, lo2)
                        if k < cand_l.size and cand_l[k] < end:
                            c = int(cand_l[k])
                if c < 0:
                    c = end
        cuts.append(c)
        s = c
    cuts = np.array(cuts, dtype=np.uint64)
    return cuts, (int(cuts[-1]) if cuts.size else entry)


def chunk(data, cfg: CDCConfig = CDCConfig()) -> np.ndarray:
    """cuts: uint64[n], strictly increasing, cuts[-1] == len(data) (empty for empty input)."""
    d = _as_u8(data)
    if d.size == 0:
        return np.zeros(0, dtype=np.uint64)
    return chunk_shard(d, cfg, 0, d.size, True)[0]


# --------------------------------------------------------------------------- #
# C restatement (oracle/hmse_ref.c), built by __graft_entry__.build().
# --------------------------------------------------------------------------- #

_REF = None


def ref_lib():
    global _REF
    if _REF is None:
        here = os.path.dirname(os.path.abspath(__file__))
        path = os.path.join(here, "_build", "libhmse_ref.so")
        if not os.path.exists(path):
            from . import build_ref
            build_ref.build()
        lib = ctypes.CDLL(path)
        lib.hmse_ref_chunk.restype = ctypes.c_int64
        lib.hmse_ref_chunk.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64,
                                       ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                       ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_uint64]
        lib.hmse_ref_minhash.restype = None
        lib.hmse_ref_minhash.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64,
                                         ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p]
        _REF = lib
    return _REF


def chunk_c(data, cfg: CDCConfig = CDCConfig(), entry: int = 0, n_own=None, eof: bool = True) -> np.ndarray:
    d = _as_u8(data)
    n = d.size
    if n == 0:
        return np.zeros(0, dtype=np.uint64)
    n_own = n if (eof or n_own is None) else n_own
    cap = n // cfg.min_size + 2
    out = np.zeros(cap, dtype=np.uint64)
    gear = cfg.gear
    k = ref_lib().hmse_ref_chunk(d.ctypes.data, n, entry, n_own, int(eof), cfg.min_size, cfg.avg_size,
                                 cfg.max_size, cfg.mask_s, cfg.mask_l, gear.ctypes.data,
                                 out.ctypes.data, cap)
    if k < 0:
        raise RuntimeError("hmse_ref_chunk capacity")
    return out[:k].copy()
