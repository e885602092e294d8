"""The ingest pipeline the spec's write path describes (README.md:1507-1583, resolved per
SURVEY.md §0.2 C1): CDC -> SHA-256 -> exact dedup -> DEFLATE of first occurrences, on one GPU
(`Ingest`) or on byte-range shards of one stream across GPUs (`ShardedIngest`)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import torch

from . import sharding
from .api import Context
from .config import CDCConfig


@dataclass
class IngestResult:
    cuts: torch.Tensor        # int64 (uint64 bit patterns) exclusive end offsets, local to the shard buffer
    digests: torch.Tensor     # uint8 [n, 32]
    canon: torch.Tensor       # int64 [n]: index (global id when sharded) of the first equal chunk
    is_first: torch.Tensor    # bool [n]
    select: torch.Tensor      # int64 [m] chunks that were compressed
    blob: torch.Tensor        # uint8 packed zlib streams
    offsets: torch.Tensor     # int64 [m + 1]
    entry: int = 0            # offset of the first owned chunk (sharded)
    id_base: int = 0          # global id of local chunk 0 (sharded)

    @property
    def n_chunks(self) -> int:
        return int(self.cuts.numel())


class Ingest:
    def __init__(self, ctx: Context, cdc: CDCConfig = CDCConfig(), zdict=b"", level: int = 6):
        self.ctx, self.cdc, self.level = ctx, cdc, level
        self.zdict = ctx.stage(zdict) if not isinstance(zdict, torch.Tensor) else zdict

    def select_first(self, first_u8: torch.Tensor) -> torch.Tensor:
        n = first_u8.numel()
        sel = self.ctx.empty(max(n, 1), torch.int64)
        m = C.c_uint64(0)
        self.ctx.check(self.ctx.lib.hmse_dedup_select(self.ctx.h, first_u8.data_ptr(), n, sel.data_ptr(), n, C.byref(m),
                                                      self.ctx.stream))
        return sel[:m.value]

    def run(self, d: torch.Tensor, compress: bool = True) -> IngestResult:
        ctx = self.ctx
        cuts = ctx.chunk(d, self.cdc)
        digests = ctx.digest(d, cuts)
        canon, first = ctx.dedup(digests)
        sel = self.select_first(first.view(torch.uint8))
        if compress:
            blob, offs = ctx.compress(d, cuts, sel, self.zdict, self.level)
        else:
            blob, offs = ctx.empty(0, torch.uint8), ctx.empty(1, torch.int64).zero_()
        return IngestResult(cuts, digests, canon, first, sel, blob, offs)


class ShardedIngest(Ingest):
    """Rank r holds stream bytes [lo_r, hi_r + max_size) (the last rank up to the stream end).
    Chunks are owned by the shard they start in; global ids follow stream order."""

    def __init__(self, ctx: Context, cdc: CDCConfig = CDCConfig(), zdict=b"", level: int = 6, group=None):
        super().__init__(ctx, cdc, zdict, level)
        self.group = group

    def run(self, d: torch.Tensor, n_own: int, eof: bool, compress: bool = True) -> IngestResult:
        import torch.distributed as dist
        ctx = self.ctx
        dev = ctx.tdev
        world = dist.get_world_size(self.group)
        ctx.chunk_scan(d, self.cdc)
        own = d.numel() if eof else n_own
        cuts, entry, _ = sharding.stitch_cuts(lambda e: ctx.chunk_resolve(d, self.cdc, n_own, eof, e), own, dev, self.group)
        n = cuts.numel()
        digests = ctx.digest(d, cuts, start0=entry)
        counts_all = sharding._all_gather_i64(n, dev, self.group)
        id_base = sum(counts_all[:dist.get_rank(self.group)])
        # partition -> all-to-all -> owner table -> all-to-all back -> scatter
        rec = ctx.empty(max(n, 1) * 40, torch.uint8)
        perm = ctx.empty(max(n, 1), torch.int32)
        cnt = (C.c_uint64 * world)()
        ctx.check(ctx.lib.hmse_dedup_partition(ctx.h, digests.data_ptr(), n, id_base, world, rec.data_ptr(), perm.data_ptr(),
                                               cnt, ctx.stream))

        def owner(recv: torch.Tensor, m: int) -> torch.Tensor:
            out = ctx.empty(max(m, 1), torch.int64)
            ctx.check(ctx.lib.hmse_dedup_records(ctx.h, recv.data_ptr(), m, out.data_ptr(), ctx.stream))
            return out[:m]

        reply = sharding.exchange_dedup(rec[:n * 40], [int(c) for c in cnt], owner, self.group)
        canon = ctx.empty(max(n, 1), torch.int64)
        first = ctx.empty(max(n, 1), torch.uint8)
        ctx.check(ctx.lib.hmse_dedup_scatter(ctx.h, reply.data_ptr(), perm.data_ptr(), n, id_base, canon.data_ptr(),
                                             first.data_ptr(), ctx.stream))
        canon, first = canon[:n], first[:n]
        sel = self.select_first(first)
        if compress:
            blob, offs = ctx.compress(d, cuts, sel, self.zdict, self.level, start0=entry)
        else:
            blob, offs = ctx.empty(0, torch.uint8), ctx.empty(1, torch.int64).zero_()
        return IngestResult(cuts, digests, canon, first.view(torch.bool), sel, blob, offs, entry, id_base)
