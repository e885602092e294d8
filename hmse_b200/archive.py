"""The ingest result as a usable artefact: the spec's on-disk records (40-byte `ChunkIndex` entries,
README.md:1264-1269; 8-byte pointer records, README.md:1312; the packed chunk store, README.md:1879-1887) in one
container, written from the device results and restored through the device read path (hmse_inflate +
hmse_segment_copy, README.md:1617-1675).

Container v1, little endian:
    0   magic b"HMSEARC1"
    8   u32 version (1), u32 dict_len
    16  u64 n_chunks, u64 n_unique, u64 raw_bytes, u64 store_bytes
    48  u32 dict_adler, 12 reserved bytes                                   (64-byte header)
    64  preset dictionary, zero padded to a multiple of 8
    ..  index:    n_unique x { sha256[32], u32 lba, u16 length, u16 refcount }   lba = store position >> 9
    ..  pointers: n_chunks x { u32 lba, u16 position & 511, u16 raw length - 1 }  (stream order)
    ..  store:    the zlib streams of the unique chunks, back to back
Limits of v1: chunks of 1..65536 bytes, compressed chunks below 64 KiB, stores below 2 TiB.
Container v2 (written only when the ingest ran the L4 layer and kept deltas) appends the delta store after the chunk
store: header bytes 52..55 = u32 n_delta, 56..63 = u64 delta_store_bytes; the delta store holds one `struct
DeltaChunk` (README.md:2182-2189) per kept delta: { u32 base = ChunkIndex entry number of the base chunk, u16 base raw
length - 1, u16 delta length, delta bytes (the COPY/ADD op list of hmse_delta_encode) }.  A pointer record whose
position lba * 512 + offset is >= store_bytes addresses the DeltaChunk at position - store_bytes; the read path inflates
the base, applies the delta (README.md:2191-2198) and checks nothing else - the digests of delta chunks are not stored.
The same layout is restated for the tests in oracle/archive.py."""
from __future__ import annotations

import ctypes as C
import struct
import zlib
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from .api import Context, default_context

MAGIC = b"HMSEARC1"
HEADER = struct.Struct("<8sIIQQQQIIQ")


@dataclass
class Archive:
    zdict: bytes
    index: np.ndarray       # uint8 [n_unique, 40]
    pointers: np.ndarray    # uint8 [n_chunks, 8]
    store: np.ndarray       # uint8 [store_bytes]
    raw_bytes: int
    delta_store: Optional[np.ndarray] = None   # uint8: DeltaChunk records (container v2)
    n_delta: int = 0

    @property
    def n_chunks(self) -> int:
        return int(self.pointers.shape[0])

    @property
    def n_unique(self) -> int:
        return int(self.index.shape[0])

    def tobytes(self) -> bytes:
        pad = (-len(self.zdict)) % 8
        ds = self.delta_store if self.n_delta else None
        hdr = HEADER.pack(MAGIC, 2 if ds is not None else 1, len(self.zdict), self.n_chunks, self.n_unique, self.raw_bytes,
                          int(self.store.size), zlib.adler32(self.zdict) if self.zdict else 0, self.n_delta if ds is not None else 0,
                          int(ds.size) if ds is not None else 0)
        return b"".join([hdr, self.zdict, b"\0" * pad, self.index.tobytes(), self.pointers.tobytes(), self.store.tobytes(),
                         ds.tobytes() if ds is not None else b""])

    def save(self, path: str) -> None:
        with open(path, "wb") as f:
            f.write(self.tobytes())

    @staticmethod
    def frombytes(buf) -> "Archive":
        mv = memoryview(buf)
        magic, ver, dlen, n, m, raw, sb, dad, n_delta, dsb = HEADER.unpack_from(mv, 0)
        if magic != MAGIC or ver not in (1, 2) or (ver == 1 and (n_delta or dsb)):
            raise ValueError("not an HMSE archive v1/v2")
        o = HEADER.size
        zd = bytes(mv[o:o + dlen])
        if dlen and zlib.adler32(zd) != dad:
            raise ValueError("dictionary checksum mismatch")
        o += dlen + ((-dlen) % 8)
        index = np.frombuffer(mv, dtype=np.uint8, count=m * 40, offset=o).reshape(m, 40)
        o += m * 40
        pointers = np.frombuffer(mv, dtype=np.uint8, count=n * 8, offset=o).reshape(n, 8)
        o += n * 8
        store = np.frombuffer(mv, dtype=np.uint8, count=sb, offset=o)
        if o + sb + dsb != len(mv):
            raise ValueError("archive size does not match its header")
        dstore = np.frombuffer(mv, dtype=np.uint8, count=dsb, offset=o + sb) if ver == 2 else None
        return Archive(zd, index, pointers, store, raw, dstore, n_delta)

    @staticmethod
    def load(path: str) -> "Archive":
        with open(path, "rb") as f:
            return Archive.frombytes(f.read())


def _dev(ctx: Context, t, dtype):
    if isinstance(t, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(t))
    t = t.contiguous()
    if t.dtype != dtype:
        t = t.view(dtype) if t.element_size() == torch.empty(0, dtype=dtype).element_size() else t.to(dtype)
    return t if t.is_cuda else t.to(ctx.tdev)


def build(res, zdict: bytes, raw_bytes: Optional[int] = None, ctx: Optional[Context] = None) -> Archive:
    """Archive of one ingest result (`IngestResult` on the device or `HostIngestResult`): index and pointer records
    are produced by hmse_index_build on the device."""
    ctx = ctx or default_context()
    cuts = _dev(ctx, res.cuts, torch.int64)
    n = cuts.numel()
    digests = _dev(ctx, res.digests, torch.uint8).view(-1)
    canon = _dev(ctx, res.canon, torch.int64)
    sel = _dev(ctx, res.select, torch.int64)
    offs = _dev(ctx, res.offsets, torch.int64)
    m = sel.numel()
    entry = int(getattr(res, "entry", 0))
    id_base = int(getattr(res, "id_base", 0))
    index = ctx.empty(max(m, 1) * 40, torch.uint8)
    ptrs = ctx.empty(max(n, 1) * 8, torch.uint8)
    dstore, n_delta = None, 0
    base = getattr(res, "base", None)
    if base is not None and n:
        if id_base or entry:
            raise ValueError("archives with deltas are single-GPU (local chunk indices)")
        base = _dev(ctx, base, torch.int64)
        doffs = _dev(ctx, res.delta_offsets, torch.int64)
        dblob = _dev(ctx, res.delta_blob, torch.uint8)
        n_delta = int((base >= 0).sum())
        cap = int(doffs[-1]) + 8 * n_delta
        dbuf = ctx.empty(cap + 64, torch.uint8)
        got = C.c_uint64(0)
        ctx.check(ctx.lib.hmse_index_build_l4(ctx.h, digests.data_ptr(), canon.data_ptr(), cuts.data_ptr(), 0, n,
                                              sel.data_ptr(), m, offs.data_ptr(), base.data_ptr(), doffs.data_ptr(),
                                              dblob.data_ptr(), index.data_ptr(), ptrs.data_ptr(), dbuf.data_ptr(), cap,
                                              C.byref(got), ctx.stream))
        dstore = dbuf[:got.value].cpu().numpy()
    else:
        ctx.check(ctx.lib.hmse_index_build(ctx.h, digests.data_ptr(), canon.data_ptr(), id_base, cuts.data_ptr(), entry, n,
                                           sel.data_ptr(), m, offs.data_ptr(), index.data_ptr(), ptrs.data_ptr(), ctx.stream))
    blob = res.blob
    store = blob.cpu().numpy() if isinstance(blob, torch.Tensor) else np.asarray(blob)
    if raw_bytes is None:
        raw_bytes = (int(cuts[-1]) - entry) if n else 0
    return Archive(bytes(zdict), index[:m * 40].cpu().numpy().reshape(m, 40), ptrs[:n * 8].cpu().numpy().reshape(n, 8),
                   np.ascontiguousarray(store), int(raw_bytes), dstore, n_delta)


def restore(ar: Archive, ctx: Optional[Context] = None, device_out: bool = False):
    """The read path: every unique chunk is inflated once on the device, the stream is rebuilt from the pointer
    records.  Returns the raw stream (numpy, or a CUDA tensor with device_out) - raises on a damaged archive."""
    ctx = ctx or default_context()
    n, m = ar.n_chunks, ar.n_unique
    if n == 0:
        return ctx.empty(0, torch.uint8) if device_out else np.zeros(0, dtype=np.uint8)
    idx = ar.index
    lba_u = idx[:, 32:36].copy().view(np.uint32).reshape(-1).astype(np.int64)
    clen = idx[:, 36:38].copy().view(np.uint16).reshape(-1).astype(np.int64)
    p = ar.pointers
    lba = p[:, 0:4].copy().view(np.uint32).reshape(-1).astype(np.int64)
    off = p[:, 4:6].copy().view(np.uint16).reshape(-1).astype(np.int64)
    raw = p[:, 6:8].copy().view(np.uint16).reshape(-1).astype(np.int64) + 1
    pos = lba * 512 + off
    is_delta = pos >= ar.store.size
    if is_delta.any() and (ar.delta_store is None or not ar.n_delta):
        raise ValueError("a pointer record addresses a delta store this archive does not have")
    dpos_all = pos - ar.store.size      # DeltaChunk position of the chunks stored as deltas
    pos = np.where(is_delta, 0, pos)    # (delta chunks are resolved below; slot 0 is a placeholder)
    # store positions of the index entries: sector from the entry, exact byte from the running sum of lengths
    upos = np.concatenate([[0], np.cumsum(clen)[:-1]])
    if not np.array_equal(upos >> 9, lba_u) or int(upos[-1] + clen[-1]) != ar.store.size:
        raise ValueError("index entries do not tile the chunk store")
    slot = np.searchsorted(upos, pos)
    if (slot >= m).any() or not np.array_equal(upos[slot], pos):
        raise ValueError("a pointer record does not address a stored chunk")
    # raw length of every unique chunk = raw length of any chunk pointing at it
    ulen = np.zeros(m, dtype=np.int64)
    ulen[slot[~is_delta]] = raw[~is_delta]
    dinfo = None
    if is_delta.any():
        # DeltaChunk headers: { u32 base slot, u16 base raw length - 1, u16 delta length }
        ds = ar.delta_store
        dpos, dinv = np.unique(dpos_all[is_delta], return_inverse=True)
        if (dpos < 0).any() or (dpos + 8 > ds.size).any():
            raise ValueError("a pointer record addresses outside the delta store")
        hdr = ds[(dpos[:, None] + np.arange(8)[None, :])]
        bslot = hdr[:, 0:4].copy().view("<u4").reshape(-1).astype(np.int64)
        blen = hdr[:, 4:6].copy().view("<u2").reshape(-1).astype(np.int64) + 1
        dlen = hdr[:, 6:8].copy().view("<u2").reshape(-1).astype(np.int64)
        if (bslot >= m).any() or (dpos + 8 + dlen > ds.size).any():
            raise ValueError("a DeltaChunk record is damaged")
        traw = np.zeros(dpos.size, dtype=np.int64)
        traw[dinv] = raw[is_delta]
        if not np.array_equal(traw[dinv], raw[is_delta]):
            raise ValueError("pointer records disagree about a delta chunk's length")
        known = ulen[bslot] != 0
        if not np.array_equal(ulen[bslot][known], blen[known]):
            raise ValueError("a DeltaChunk record disagrees with the pointer records about its base's length")
        ulen[bslot] = blen              # a base may be referenced by deltas only
        dinfo = (dpos, dinv, bslot, blen, dlen, traw)
    if (ulen == 0).any() or not np.array_equal(ulen[slot[~is_delta]], raw[~is_delta]):
        raise ValueError("pointer records disagree about a chunk's length")
    if int(raw.sum()) != ar.raw_bytes:
        raise ValueError("pointer records do not add up to the stream length")
    offs = ctx.stage_u64(np.concatenate([upos, [ar.store.size]]).astype(np.uint64))
    uo = np.concatenate([[0], np.cumsum(ulen)]).astype(np.uint64)
    zd = ctx.stage(ar.zdict)
    uniq, status, bad = ctx.inflate(ctx.stage(ar.store), offs, ctx.stage_u64(uo), zd)
    if bad:
        raise ValueError("%d stored chunks failed to inflate (first: %d)" % (bad, int(torch.nonzero(status)[0])))
    src = uo[:-1][slot].astype(np.uint64)
    if dinfo is not None:
        dpos, dinv, bslot, blen, dlen, traw = dinfo
        doff = np.concatenate([[0], np.cumsum(dlen)]).astype(np.uint64)
        dsd = ctx.stage(ar.delta_store)
        # gather the delta bytes of the records (they sit behind their 8-byte headers) into one packed blob
        dblob = ctx.empty(int(doff[-1]) + 64, torch.uint8)[:int(doff[-1])]
        t_src, t_dst = ctx.stage_u64((dpos + 8).astype(np.uint64)), ctx.stage_u64(doff)   # named: they must outlive the call
        ctx.check(ctx.lib.hmse_segment_copy(ctx.h, dsd.data_ptr(), t_src.data_ptr(), dblob.data_ptr(), t_dst.data_ptr(),
                                            dpos.size, ctx.stream))
        to = np.concatenate([[0], np.cumsum(traw)]).astype(np.uint64)
        bl = torch.from_numpy(blen.astype(np.uint32).view(np.int32).copy()).to(ctx.tdev)
        tout, dstatus, dbad = ctx.delta_apply(dblob, t_dst, uniq, ctx.stage_u64(uo[:-1][bslot].astype(np.uint64)), bl,
                                              ctx.stage_u64(to))
        if dbad:
            raise ValueError("%d deltas failed to apply (first: %d)" % (dbad, int(torch.nonzero(dstatus)[0])))
        total_u = int(uo[-1])
        both = ctx.empty(total_u + int(to[-1]) + 64, torch.uint8)
        both[:total_u].copy_(uniq)
        both[total_u:total_u + int(to[-1])].copy_(tout)
        uniq = both
        src[is_delta] = np.uint64(total_u) + to[:-1][dinv]
    out = ctx.empty(ar.raw_bytes + 64, torch.uint8)[:ar.raw_bytes]
    src_off = ctx.stage_u64(src)
    dst_off = ctx.stage_u64(np.concatenate([[0], np.cumsum(raw)]).astype(np.uint64))
    ctx.check(ctx.lib.hmse_segment_copy(ctx.h, uniq.data_ptr(), src_off.data_ptr(), out.data_ptr(), dst_off.data_ptr(), n,
                                        ctx.stream))
    return out if device_out else out.cpu().numpy()
