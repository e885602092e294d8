"""CPU restatement of the archive container (hmse_b200/archive.py): the spec's 40-byte ChunkIndex entries
(README.md:1264-1269), 8-byte pointer records (README.md:1312) and packed chunk store (README.md:1879-1887),
and a restore that uses nothing but zlib.  TEST INFRASTRUCTURE - see oracle/__init__.py."""
from __future__ import annotations

import struct
import zlib

import numpy as np

MAGIC = b"HMSEARC1"
HEADER = struct.Struct("<8sIIQQQQI12x")


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


def pack(zdict: bytes, index, pointers, store, raw_bytes: int) -> bytes:
    pad = (-len(zdict)) % 8
    hdr = HEADER.pack(MAGIC, 1, len(zdict), pointers.shape[0], index.shape[0], raw_bytes, int(np.asarray(store).size),
                      zlib.adler32(zdict) if zdict else 0)
    return b"".join([hdr, zdict, b"\0" * pad, index.tobytes(), pointers.tobytes(), np.asarray(store).tobytes()])


def restore(buf: bytes) -> bytes:
    """Pure-Python read path: walks the pointer records, inflates with zlib."""
    magic, ver, dlen, n, m, raw_bytes, sb, dad = HEADER.unpack_from(buf, 0)
    assert magic == MAGIC and ver == 1
    o = HEADER.size
    zd = buf[o:o + dlen]
    o += dlen + ((-dlen) % 8)
    index = np.frombuffer(buf, dtype=np.uint8, count=m * 40, offset=o).reshape(m, 40)
    o += m * 40
    ptr = np.frombuffer(buf, dtype=np.uint8, count=n * 8, offset=o).reshape(n, 8)
    o += n * 8
    store = buf[o:o + sb]
    clen = index[:, 36:38].copy().view("<u2").reshape(-1).astype(np.int64)
    upos = np.concatenate([[0], np.cumsum(clen)[:-1]])
    lba = ptr[:, 0:4].copy().view("<u4").reshape(-1).astype(np.int64)
    off = ptr[:, 4:6].copy().view("<u2").reshape(-1).astype(np.int64)
    raw = ptr[:, 6:8].copy().view("<u2").reshape(-1).astype(np.int64) + 1
    pos = lba * 512 + off
    slot = np.searchsorted(upos, pos)
    cache = {}
    out = []
    for i in range(n):
        s = int(slot[i])
        if s not in cache:
            do = zlib.decompressobj(15, zd) if dlen else zlib.decompressobj(15)
            cache[s] = do.decompress(store[int(upos[s]):int(upos[s] + clen[s])]) + do.flush()
            assert do.eof
        assert len(cache[s]) == raw[i]
        out.append(cache[s])
    res = b"".join(out)
    assert len(res) == raw_bytes
    return res
