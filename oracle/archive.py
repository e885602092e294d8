"""CPU restatement of the archive container (hmse_b200/archive.py): the spec's 40-byte ChunkIndex entries
(README.md:1264-1269), 8-byte pointer records (README.md:1312) and packed chunk store (README.md:1879-1887),
and a restore that uses nothing but zlib.  TEST INFRASTRUCTURE - see oracle/__init__.py."""
from __future__ import annotations

import struct
import zlib

import numpy as np

MAGIC = b"HMSEARC1"
HEADER = struct.Struct("<8sIIQQQQIIQ")


def records(digests, canon, cuts, select, offsets, start0: int = 0):
    """(index uint8[m,40], pointers uint8[n,8]) from the outputs of digest/dedup/compress."""
    cuts = np.asarray(cuts, dtype=np.uint64).astype(np.int64)
    canon = np.asarray(canon, dtype=np.int64)
    select = np.asarray(select, dtype=np.int64)
    offsets = np.asarray(offsets, dtype=np.uint64).astype(np.int64)
    n, m = cuts.size, select.size
    slot_of = np.full(n, -1, dtype=np.int64)
    slot_of[select] = np.arange(m)
    slot = slot_of[canon]
    assert (slot >= 0).all()
    refcount = np.minimum(np.bincount(slot, minlength=m), 0xFFFF)
    starts = np.concatenate([[start0], cuts[:-1]])
    raw = cuts - starts
    pos = offsets[:-1]
    clen = np.diff(offsets)
    index = np.zeros((m, 40), dtype=np.uint8)
    index[:, :32] = np.asarray(digests, dtype=np.uint8).reshape(-1, 32)[select]
    index[:, 32:36] = (pos >> 9).astype("<u4").view(np.uint8).reshape(m, 4)
    index[:, 36:38] = clen.astype("<u2").view(np.uint8).reshape(m, 2)
    index[:, 38:40] = refcount.astype("<u2").view(np.uint8).reshape(m, 2)
    ptr = np.zeros((n, 8), dtype=np.uint8)
    ptr[:, 0:4] = (pos[slot] >> 9).astype("<u4").view(np.uint8).reshape(n, 4)
    ptr[:, 4:6] = (pos[slot] & 511).astype("<u2").view(np.uint8).reshape(n, 2)
    ptr[:, 6:8] = (raw - 1).astype("<u2").view(np.uint8).reshape(n, 2)
    return index, ptr


def records_l4(digests, canon, cuts, select, offsets, base, delta_blob, delta_offsets, start0: int = 0):
    """(index, pointers, delta_store, n_delta) when first occurrences with base >= 0 are stored as deltas
    (`struct DeltaChunk`, README.md:2182-2189; container v2 of hmse_b200/archive.py).  `select` lists the chunks of
    the chunk store (first occurrences without a delta)."""
    cuts = np.asarray(cuts, dtype=np.uint64).astype(np.int64)
    canon = np.asarray(canon, dtype=np.int64)
    select = np.asarray(select, dtype=np.int64)
    offsets = np.asarray(offsets, dtype=np.uint64).astype(np.int64)
    base = np.asarray(base, dtype=np.int64)
    doff = np.asarray(delta_offsets, dtype=np.uint64).astype(np.int64)
    dblob = np.asarray(delta_blob, dtype=np.uint8)
    n, m = cuts.size, select.size
    starts = np.concatenate([[start0], cuts[:-1]])
    raw = cuts - starts
    slot_of = np.full(n, -1, dtype=np.int64)
    slot_of[select] = np.arange(m)
    kept = np.diff(doff) > 0
    rank = np.cumsum(kept) - kept
    store_bytes = int(offsets[-1])
    # delta store
    parts = []
    for c in np.flatnonzero(kept):
        b = int(base[c])
        assert slot_of[b] >= 0
        parts.append(struct.pack("<IHH", int(slot_of[b]), int(raw[b]) - 1, int(doff[c + 1] - doff[c])))
        parts.append(dblob[doff[c]:doff[c + 1]].tobytes())
    dstore = np.frombuffer(b"".join(parts), dtype=np.uint8)
    # positions and reference counts
    refcount = np.zeros(m, dtype=np.int64)
    pos = np.zeros(n, dtype=np.int64)
    for i in range(n):
        c = int(canon[i])
        if kept[c]:
            pos[i] = store_bytes + 8 * rank[c] + doff[c]
            if i == c:
                refcount[slot_of[base[c]]] += 1
        else:
            assert slot_of[c] >= 0
            pos[i] = offsets[slot_of[c]]
            refcount[slot_of[c]] += 1
    refcount = np.minimum(refcount, 0xFFFF)
    index = np.zeros((m, 40), dtype=np.uint8)
    index[:, :32] = np.asarray(digests, dtype=np.uint8).reshape(-1, 32)[select]
    index[:, 32:36] = (offsets[:-1] >> 9).astype("<u4").view(np.uint8).reshape(m, 4)
    index[:, 36:38] = np.diff(offsets).astype("<u2").view(np.uint8).reshape(m, 2)
    index[:, 38:40] = refcount.astype("<u2").view(np.uint8).reshape(m, 2)
    ptr = np.zeros((n, 8), dtype=np.uint8)
    ptr[:, 0:4] = (pos >> 9).astype("<u4").view(np.uint8).reshape(n, 4)
    ptr[:, 4:6] = (pos & 511).astype("<u2").view(np.uint8).reshape(n, 2)
    ptr[:, 6:8] = (raw - 1).astype("<u2").view(np.uint8).reshape(n, 2)
    return index, ptr, dstore, int(kept.sum())


def pack(zdict: bytes, index, pointers, store, raw_bytes: int, delta_store=None, n_delta: int = 0) -> bytes:
    pad = (-len(zdict)) % 8
    v2 = n_delta > 0
    hdr = HEADER.pack(MAGIC, 2 if v2 else 1, len(zdict), pointers.shape[0], index.shape[0], raw_bytes, int(np.asarray(store).size),
                      zlib.adler32(zdict) if zdict else 0, n_delta if v2 else 0, int(np.asarray(delta_store).size) if v2 else 0)
    return b"".join([hdr, zdict, b"\0" * pad, index.tobytes(), pointers.tobytes(), np.asarray(store).tobytes(),
                     np.asarray(delta_store).tobytes() if v2 else b""])


def restore(buf: bytes) -> bytes:
    """Pure-Python read path: walks the pointer records, inflates with zlib."""
    from .deltacode import delta_apply
    magic, ver, dlen, n, m, raw_bytes, sb, dad, n_delta, dsb = HEADER.unpack_from(buf, 0)
    assert magic == MAGIC and ver in (1, 2)
    o = HEADER.size
    zd = buf[o:o + dlen]
    o += dlen + ((-dlen) % 8)
    index = np.frombuffer(buf, dtype=np.uint8, count=m * 40, offset=o).reshape(m, 40)
    o += m * 40
    ptr = np.frombuffer(buf, dtype=np.uint8, count=n * 8, offset=o).reshape(n, 8)
    o += n * 8
    store = buf[o:o + sb]
    dstore = buf[o + sb:o + sb + dsb]
    clen = index[:, 36:38].copy().view("<u2").reshape(-1).astype(np.int64)
    upos = np.concatenate([[0], np.cumsum(clen)[:-1]])
    lba = ptr[:, 0:4].copy().view("<u4").reshape(-1).astype(np.int64)
    off = ptr[:, 4:6].copy().view("<u2").reshape(-1).astype(np.int64)
    raw = ptr[:, 6:8].copy().view("<u2").reshape(-1).astype(np.int64) + 1
    pos = lba * 512 + off
    slot = np.searchsorted(upos, np.minimum(pos, sb))
    cache = {}

    def stored(s):
        if s not in cache:
            do = zlib.decompressobj(15, zd) if dlen else zlib.decompressobj(15)
            cache[s] = do.decompress(store[int(upos[s]):int(upos[s] + clen[s])]) + do.flush()
            assert do.eof
        return cache[s]

    out = []
    for i in range(n):
        if pos[i] >= sb:   # DeltaChunk: read base -> inflate -> apply the delta (README.md:2191-2198)
            dp = int(pos[i]) - sb
            bslot, blen, dl = struct.unpack_from("<IHH", dstore, dp)
            b = stored(bslot)
            assert len(b) == blen + 1
            out.append(delta_apply(bytes(dstore[dp + 8:dp + 8 + dl]), b, int(raw[i])))
            continue
        s = int(slot[i])
        assert upos[s] == pos[i]
        assert len(stored(s)) == raw[i]
        out.append(stored(s))
    res = b"".join(out)
    assert len(res) == raw_bytes
    return res
