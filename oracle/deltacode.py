"""L4 delta coding against LSH-selected bases.  TEST INFRASTRUCTURE - see oracle/__init__.py.

Spec: "If LSH match -> compute binary difference.  Store delta only if size <= 20 % of original
chunk" (README.md:1328, 2175), flow README.md:1555-1570, `struct DeltaChunk` (README.md:2182-2189),
reconstruction "read base -> decompress -> apply delta" (README.md:2191-2198), example op list
COPY / INSERT / COPY (README.md:1402-1412).  The spec names xdelta3 (README.md:2162) and, in the
example, bsdiff (README.md:1402): neither is vendored or pinned and neither's byte format is given,
so - **parity unpinned** - this module defines the coding the spec's example describes (a COPY/ADD
op list) completely, and the GPU build must reproduce it byte for byte:

Base selection (the "probe LSH index -> return base chunk" step, README.md:1556-1559)
    head(i, b)  = smallest first-occurrence chunk id j with keys[j][b] == keys[i][b]   (the entry
                  an LSH index that keeps the first chunk of every bucket returns for the probe of
                  band b; only chunks that passed exact dedup are ever inserted, README.md:1553-1556.
                  In one stream a duplicate has an earlier first occurrence with the same keys, so
                  the restriction changes nothing there; it matters for a shard whose duplicates
                  have their first occurrence in another shard)
    votes(i, j) = number of bands b with head(i, b) == j, for j < i
    root(i)     = is_first[i] and max_j votes(i, j) < min_votes       (nothing earlier is similar:
                  the chunk is stored whole and may serve as a base)
    base(i)     = the root j < i with the most votes(i, j) >= min_votes, ties to the smaller j;
                  -1 when there is none or when chunk i is a duplicate.
    Bases are always roots, so reconstruction never chains (README.md:2191-2198 reads one base).

Delta format (target T of n bytes, base B of nb bytes, both <= 32768)
    delta  := op*                       ops until n target bytes are produced
    op     := varint(len << 1 | kind)   kind 0 = ADD: `len` literal bytes follow
                                        kind 1 = COPY: varint(zigzag(q - expect)) follows and
                                        T[..] += B[q : q + len]; expect = q + len (0 at the start)
    varint  = unsigned LEB128.
Encoder (greedy, deterministic)
    H[h]    = smallest base position q <= nb - 8 whose 8 bytes hash to h
              (h = (le64(B[q:q+8]) * 0x9E3779B97F4A7C15) >> 50, 16384 buckets)
    seed(s) = H[h(T[s:s+8])] when that entry exists and its 8 bytes equal T[s:s+8]
    walk: from the end p of the previous COPY, take the first s >= p with a seed; extend it
    backwards while bytes agree (not past p, not past the start of B), then forwards; emit the
    pending literals as one ADD, then the COPY.  Trailing bytes become a final ADD.
    The delta is kept only if 5 * len(delta) <= n (the 20 % rule); otherwise there is none.
"""
from __future__ import annotations

import numpy as np

from .config import SimConfig

SEED_LEN = 8
HASH_BITS = 14
HASH_MUL = 0x9E3779B97F4A7C15
MAX_LEN = 32768
MIN_VOTES = 4


# ---- base selection ------------------------------------------------------------------------

def lsh_heads(keys: np.ndarray) -> np.ndarray:
    """head[i][b] = smallest j with keys[j][b] == keys[i][b] (int64[n][bands])."""
    keys = np.asarray(keys, dtype=np.uint64)
    n, nb = keys.shape
    heads = np.empty((n, nb), dtype=np.int64)
    for b in range(nb):
        _, first, inv = np.unique(keys[:, b], return_index=True, return_inverse=True)
        heads[:, b] = first[inv]
    return heads


def delta_bases(keys: np.ndarray, is_first: np.ndarray, min_votes: int = MIN_VOTES) -> np.ndarray:
    """base int64[n] as defined in the module docstring."""
    keys = np.asarray(keys, dtype=np.uint64)
    n = keys.shape[0]
    is_first = np.asarray(is_first, dtype=bool)
    idx = np.flatnonzero(is_first)
    heads = np.tile(np.arange(n, dtype=np.int64)[:, None], (1, keys.shape[1]))   # duplicates: no head but themselves
    if idx.size:
        heads[idx] = idx[lsh_heads(keys[idx])]
    root = np.zeros(n, dtype=bool)
    votes = []
    for i in range(n):
        h = heads[i]
        js, cnt = np.unique(h[h < i], return_counts=True)
        votes.append((js, cnt))
        root[i] = is_first[i] and (cnt.size == 0 or int(cnt.max()) < min_votes)
    base = np.full(n, -1, dtype=np.int64)
    for i in range(n):
        if not is_first[i]:
            continue
        js, cnt = votes[i]
        ok = root[js] & (cnt >= min_votes)
        if ok.any():
            js, cnt = js[ok], cnt[ok]
            base[i] = int(js[np.argmax(cnt)])  # np.unique sorts js ascending: argmax takes the smallest j of a tie
    return base


# ---- delta coding ----------------------------------------------------------------------------

def _varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _zigzag(v: int) -> int:
    return (v << 1) if v >= 0 else ((-v) << 1) - 1


def _windows64(a: np.ndarray) -> np.ndarray:
    """le64 of every 8-byte window of a (uint64[len(a) - 7])."""
    m = a.size - 7
    w = np.zeros(m, dtype=np.uint64)
    for k in range(8):
        w |= a[k:k + m].astype(np.uint64) << np.uint64(8 * k)
    return w


def _hash(w: np.ndarray) -> np.ndarray:
    return ((w * np.uint64(HASH_MUL)) >> np.uint64(64 - HASH_BITS)).astype(np.int64)


def delta_encode(target, base):
    """bytes of the delta, or None when the 20 % rule (or a size limit) rejects it."""
    T = np.frombuffer(bytes(target), dtype=np.uint8)
    B = np.frombuffer(bytes(base), dtype=np.uint8)
    n, nb = T.size, B.size
    if n == 0 or n > MAX_LEN or nb > MAX_LEN:
        return None
    cap = n // 5
    seed_q = np.full(max(n, 1), -1, dtype=np.int64)
    if n >= SEED_LEN and nb >= SEED_LEN:
        wb = _windows64(B)
        H = np.full(1 << HASH_BITS, 0xFFFF, dtype=np.int64)
        np.minimum.at(H, _hash(wb), np.arange(wb.size, dtype=np.int64))
        wt = _windows64(T)
        q = H[_hash(wt)]
        ok = q != 0xFFFF
        ok[ok] = wb[q[ok]] == wt[ok]
        seed_q[:wt.size][ok] = q[ok]
    seeds = np.flatnonzero(seed_q >= 0)
    out = bytearray()
    p, expect = 0, 0
    while True:
        k = np.searchsorted(seeds, p)
        if k >= seeds.size:
            break
        s = int(seeds[k])
        q = int(seed_q[s])
        while s > p and q > 0 and T[s - 1] == B[q - 1]:
            s -= 1
            q -= 1
        L = SEED_LEN
        lim = min(n - s, nb - q)
        neq = np.flatnonzero(T[s + L:s + lim] != B[q + L:q + lim])
        L = lim if neq.size == 0 else L + int(neq[0])
        if s > p:
            out += _varint((s - p) << 1) + T[p:s].tobytes()
        out += _varint((L << 1) | 1) + _varint(_zigzag(q - expect))
        expect = q + L
        p = s + L
        if len(out) > cap:
            return None
    if p < n:
        out += _varint((n - p) << 1) + T[p:].tobytes()
    return bytes(out) if len(out) <= cap else None


def delta_encode_naive(target, base):
    """The same encoder as a literal byte-at-a-time transcription of the module docstring (pure Python, no NumPy):
    ground truth for delta_encode, like chunk_naive is for chunk."""
    T, B = bytes(target), bytes(base)
    n, nb = len(T), len(B)
    if n == 0 or n > MAX_LEN or nb > MAX_LEN:
        return None
    M64 = (1 << 64) - 1

    def h(buf, i):
        return ((int.from_bytes(buf[i:i + 8], "little") * HASH_MUL) & M64) >> (64 - HASH_BITS)

    H = {}
    for q in range(nb - 7):
        H.setdefault(h(B, q), q)          # positions ascend: the first one seen is the smallest
    out = bytearray()
    p = expect = 0
    s = 0
    while s + 8 <= n:
        q = H.get(h(T, s))
        if q is None or B[q:q + 8] != T[s:s + 8]:
            s += 1
            continue
        while s > p and q > 0 and T[s - 1] == B[q - 1]:
            s -= 1
            q -= 1
        L = 0
        while s + L < n and q + L < nb and T[s + L] == B[q + L]:
            L += 1
        if s > p:
            out += _varint((s - p) << 1) + T[p:s]
        out += _varint((L << 1) | 1) + _varint(_zigzag(q - expect))
        expect = q + L
        p = s = s + L
    if p < n:
        out += _varint((n - p) << 1) + T[p:]
    return bytes(out) if len(out) * 5 <= n else None


def delta_apply(delta: bytes, base: bytes, n: int) -> bytes:
    """Reconstructs the n-byte target; raises ValueError on a malformed delta."""
    out = bytearray()
    i, expect = 0, 0

    def rd():
        nonlocal i
        v = sh = 0
        while True:
            if i >= len(delta) or sh > 28:
                raise ValueError("truncated or oversized varint")
            b = delta[i]
            i += 1
            v |= (b & 0x7F) << sh
            sh += 7
            if not b & 0x80:
                if v > 0xFFFFFFFF:
                    raise ValueError("varint exceeds 32 bits")
                return v

    while len(out) < n:
        tag = rd()
        ln = tag >> 1
        if ln == 0 or len(out) + ln > n:
            raise ValueError("bad op length")
        if tag & 1:
            z = rd()
            q = expect + ((z >> 1) if not z & 1 else -((z + 1) >> 1))
            if q < 0 or q + ln > len(base):
                raise ValueError("copy outside the base")
            out += base[q:q + ln]
            expect = q + ln
        else:
            if i + ln > len(delta):
                raise ValueError("literals past the end")
            out += delta[i:i + ln]
            i += ln
    if i != len(delta):
        raise ValueError("trailing bytes")
    return bytes(out)


def delta(data, cuts, keys, is_first, min_votes: int = MIN_VOTES, start0: int = 0):
    """(base int64[n], blob uint8[...], offsets uint64[n+1]): chunk i is stored as a delta against
    chunk base[i] iff offsets[i+1] > offsets[i]; base[i] is -1 where no delta is kept."""
    d = bytes(data) if not isinstance(data, np.ndarray) else data.tobytes()
    cuts = np.asarray(cuts, dtype=np.uint64)
    n = cuts.size
    starts = np.concatenate([[start0], cuts[:-1]]).astype(np.int64)
    cand = delta_bases(keys, is_first, min_votes)
    base = np.full(n, -1, dtype=np.int64)
    parts, offs = [], np.zeros(n + 1, dtype=np.uint64)
    for i in range(n):
        offs[i + 1] = offs[i]
        j = int(cand[i])
        if j < 0:
            continue
        dl = delta_encode(d[int(starts[i]):int(cuts[i])], d[int(starts[j]):int(cuts[j])])
        if dl is None:
            continue
        base[i] = j
        parts.append(dl)
        offs[i + 1] += np.uint64(len(dl))
    blob = np.frombuffer(b"".join(parts), dtype=np.uint8).copy()
    return base, blob, offs


def delta_records(base, cuts, offsets, lba_of, start0: int = 0) -> np.ndarray:
    """`struct DeltaChunk` headers (README.md:2182-2189), 8 bytes each, one per kept delta in
    chunk order: u32 base_lba (= lba_of[base], the base's ChunkIndex lba), u16 base_length - 1,
    u16 delta_length; the delta bytes live in the delta blob."""
    cuts = np.asarray(cuts, dtype=np.uint64)
    starts = np.concatenate([[start0], cuts[:-1]]).astype(np.uint64)
    ln = (cuts - starts).astype(np.int64)
    offsets = np.asarray(offsets, dtype=np.uint64)
    rec = np.zeros(0, dtype=[("base_lba", "<u4"), ("base_length", "<u2"), ("delta_length", "<u2")])
    idx = np.flatnonzero(np.asarray(base) >= 0)
    rec = np.zeros(idx.size, dtype=rec.dtype)
    b = np.asarray(base)[idx]
    rec["base_lba"] = np.asarray(lba_of)[b]
    rec["base_length"] = ln[b] - 1
    rec["delta_length"] = (offsets[idx + 1] - offsets[idx]).astype(np.uint16)
    return rec
