"""GPU parity: FastCDC cut points bit-exact against the oracle (oracle/cdc.py), through the C ABI."""
import numpy as np
import pytest

import oracle
from oracle import corpus
from oracle.config import CDCConfig as OCfg

pytestmark = pytest.mark.gpu


def _pcfg(o: OCfg):
    import hmse_b200
    return hmse_b200.CDCConfig(o.min_size, o.avg_size, o.max_size, o.mask_s, o.mask_l, o.gear_seed)


def _check(data, ocfg=OCfg(), ctx=None):
    import hmse_b200
    data = np.ascontiguousarray(data, dtype=np.uint8)
    want = oracle.chunk_c(data, ocfg)
    got = hmse_b200.chunk(data, _pcfg(ocfg), ctx=ctx)
    assert got.dtype == np.uint64
    if not np.array_equal(got, want):
        k = min(got.size, want.size)
        bad = np.flatnonzero(got[:k] != want[:k])
        first = int(bad[0]) if bad.size else k
        raise AssertionError("cuts differ: n=%d got %d want %d first mismatch @%d got %s want %s" % (
            data.size, got.size, want.size, first, got[max(0, first - 1):first + 3], want[max(0, first - 1):first + 3]))
    return got


def test_corpus_default(ctx, corpus8):
    cuts = _check(corpus8, ctx=ctx)
    assert int(cuts[-1]) == corpus8.size
    lens = np.diff(np.concatenate([[0], cuts]).astype(np.int64))
    assert lens[:-1].min() >= 2048 and lens.max() <= 32768


@pytest.mark.parametrize("avg,nc", [(4096, 2), (8192, 1), (8192, 3), (16384, 2), (65536, 2)])
def test_corpus_configs(ctx, corpus8, avg, nc):
    _check(corpus8[:4 << 20], OCfg.for_avg(avg, nc), ctx=ctx)


def test_spec_sizes_1_4_16(ctx, corpus8):
    # the spec's own parameters: min 1 KiB / avg 4 KiB / max 16 KiB (README.md:289, 2444-2446)
    _check(corpus8[:4 << 20], OCfg.for_avg(4096), ctx=ctx)


def test_tiny_config(ctx, corpus8):
    _check(corpus8[:1 << 20], OCfg(64, 256, 1024, oracle.config.spread_mask(10), oracle.config.spread_mask(6)), ctx=ctx)


@pytest.mark.parametrize("n", [0, 1, 15, 16, 63, 64, 65, 2047, 2048, 2049, 2111, 2112, 2113, 8192, 32767, 32768, 32769,
                               65535, 65536, 65537, 100000, 262143, 262144, 262145, 1 << 20])
def test_lengths(ctx, corpus8, n):
    _check(corpus8[1000:1000 + n], ctx=ctx)


def test_random_bytes(ctx):
    _check(corpus.random_bytes(4 << 20), ctx=ctx)


@pytest.mark.parametrize("fill", [0, 0xFF, 0x20])
def test_constant(ctx, fill):
    _check(np.full(3 << 20, fill, dtype=np.uint8), ctx=ctx)
    _check(np.full((3 << 20) + 12345, fill, dtype=np.uint8), ctx=ctx)


@pytest.mark.parametrize("period", [2, 3, 64, 65, 1000])
def test_periodic(ctx, period):
    base = corpus.random_bytes(period, seed=period)
    _check(np.tile(base, (2 << 20) // period + 1)[:2 << 20], ctx=ctx)


def test_zero_run_inside_text(ctx, corpus8):
    d = corpus8[:6 << 20].copy()
    d[(1 << 20) + 777:(4 << 20) + 123] = 0      # forced max-size cuts at offsets no guess can know
    _check(d, ctx=ctx)


def test_deterministic(ctx, corpus8):
    import hmse_b200
    a = hmse_b200.chunk(corpus8, ctx=ctx)
    b = hmse_b200.chunk(corpus8, ctx=ctx)
    assert np.array_equal(a, b)


def test_shift_resistance(ctx, corpus8):
    # README.md:1254: inserting 100 bytes at the front leaves >= 99 % of chunks unchanged
    import hmse_b200
    d = corpus8[:4 << 20]
    shifted = np.concatenate([corpus.random_bytes(100, seed=7), d])
    a = hmse_b200.chunk(d, ctx=ctx).astype(np.int64)
    b = hmse_b200.chunk(shifted, ctx=ctx).astype(np.int64) - 100
    common = np.intersect1d(a, b).size
    assert common >= 0.99 * a.size


@pytest.mark.parametrize("world", [2, 3, 8])
def test_virtual_shards_equal_single_stream(ctx, corpus8, world):
    """Byte-range shards with max_size look-ahead, entry chained from the previous shard's exit,
    reproduce the single-stream cut list."""
    import torch
    import hmse_b200
    cfg = hmse_b200.CDCConfig()
    d = corpus8[:6 << 20]
    n = d.size
    want = oracle.chunk_c(d)
    dev = ctx.stage(d)
    per = (n + world - 1) // world
    got, entry = [], 0
    for r in range(world):
        lo, hi = r * per, min(n, (r + 1) * per)
        eof = hi == n
        avail_hi = n if eof else min(n, hi + cfg.max_size)
        if not eof and avail_hi - lo < (hi - lo) + cfg.max_size:
            eof = True  # look-ahead would cross the stream end: treat the tail as part of this shard
            avail_hi = n
        buf = ctx.stage(dev[lo:avail_hi].clone())
        ctx.chunk_scan(buf, cfg)
        # speculative first, then the true entry (exercises the incremental re-resolve)
        ctx.chunk_resolve(buf, cfg, hi - lo, eof, 0)
        cuts, ex = ctx.chunk_resolve(buf, cfg, hi - lo, eof, entry)
        got.append(cuts.cpu().numpy().view(np.uint64) + np.uint64(lo))
        if eof:
            break
        entry = ex - (hi - lo)
    got = np.concatenate(got)
    assert np.array_equal(got, want)
    torch.cuda.synchronize()


@pytest.mark.parametrize("kind", ["text", "random"])
def test_scan_candidate_bitmaps(ctx, corpus8, kind):
    """K1 alone: the MaskS / MaskL bitmaps equal the oracle's full-window candidates."""
    import hmse_b200
    from oracle.cdc import candidates
    d = corpus8[:3 << 20] if kind == "text" else corpus.random_bytes(3 << 20)
    d = np.ascontiguousarray(d[:d.size - 777])       # ragged tail
    ocfg = OCfg()
    ctx.chunk_scan(ctx.stage(d), _pcfg(ocfg))
    bs, bl = ctx.chunk_candidates(d.size)
    cs, cl = candidates(d, ocfg)
    for name, bits, want in (("S", bs, cs), ("L", bl, cl)):
        got = np.flatnonzero(np.unpackbits(bits.cpu().numpy().view(np.uint8), bitorder="little"))
        got, want = got[got >= 64], want[want >= 64]
        if not np.array_equal(got, want):
            miss = np.setdiff1d(want, got)[:8]
            extra = np.setdiff1d(got, want)[:8]
            raise AssertionError("%s bitmap (%s): got %d want %d missing %s extra %s" % (name, kind, got.size, want.size,
                                                                                         miss, extra))

