"""hmse_inflate (the read path, README.md:1617-1675) against the raw bytes: streams from hmse_compress, streams
from stock zlib (every level, with and without the preset dictionary, stored / fixed / dynamic blocks, multi-block
streams), and malformed streams, which must be reported without touching other streams' output."""
import zlib

import numpy as np
import pytest

import oracle
from oracle import corpus

pytestmark = pytest.mark.gpu


def _sizes(cuts, select=None):
    starts = np.concatenate([[0], cuts[:-1]]).astype(np.int64)
    lens = cuts.astype(np.int64) - starts
    return lens if select is None else lens[select]


def _raw(data, cuts, select=None):
    starts = np.concatenate([[0], cuts[:-1]]).astype(np.int64)
    idx = range(cuts.size) if select is None else select
    return b"".join(data[starts[j]:int(cuts[j])].tobytes() for j in idx)


def test_roundtrip_of_gpu_streams(ctx, corpus8):
    import hmse_b200
    d = corpus8
    cuts = oracle.chunk_c(d)
    zd = corpus.zdict()
    _, first = oracle.dedup(oracle.digest(d, cuts))
    sel = np.flatnonzero(first)
    blob, offs = hmse_b200.compress(d, cuts, sel, zd, ctx=ctx)
    out, status = hmse_b200.inflate(blob, offs, _sizes(cuts, sel), zd, ctx=ctx)
    assert (status == 0).all()
    assert out.tobytes() == _raw(d, cuts, sel)
    # no dictionary
    blob, offs = hmse_b200.compress(d[:1 << 20], oracle.chunk_c(d[:1 << 20]), None, b"", ctx=ctx)
    c1 = oracle.chunk_c(d[:1 << 20])
    out, status = hmse_b200.inflate(blob, offs, _sizes(c1), b"", ctx=ctx)
    assert (status == 0).all() and out.tobytes() == d[:1 << 20].tobytes()


@pytest.mark.parametrize("level", [0, 1, 6, 9])
@pytest.mark.parametrize("use_dict", [True, False])
def test_stock_zlib_streams(ctx, corpus8, level, use_dict):
    import hmse_b200
    rng = np.random.default_rng(level)
    d = corpus8[:3 << 20].copy()
    d[1 << 20:(1 << 20) + 200000] = rng.integers(0, 256, 200000, dtype=np.uint8)     # stored blocks inside
    lens = [0, 1, 2, 5, 40, 300, 4096, 70000, 200000, 32768, 65536] + [8192] * 200 + [1100000]
    cuts = np.cumsum(lens).astype(np.uint64)
    assert int(cuts[-1]) <= d.size
    zd = corpus.zdict() if use_dict else b""
    blob, offs = oracle.compress(d, cuts, np.arange(cuts.size), zd, level=level)
    out, status = hmse_b200.inflate(blob, offs, _sizes(cuts), zd, ctx=ctx)
    assert (status == 0).all(), np.flatnonzero(status)[:10]
    assert out.tobytes() == d[:int(cuts[-1])].tobytes()


def test_fixed_huffman_and_long_gpu_streams(ctx, corpus8):
    import hmse_b200
    # Z_FIXED streams from zlib; long (multi-block) and constant chunks from the GPU encoder
    d = corpus8[:400000]
    co = zlib.compressobj(6, zlib.DEFLATED, 15, 8, zlib.Z_FIXED)
    s = co.compress(d.tobytes()) + co.flush()
    out, status = hmse_b200.inflate(np.frombuffer(s, dtype=np.uint8), np.array([0, len(s)], dtype=np.uint64),
                                    np.array([d.size]), b"", ctx=ctx)
    assert status.tolist() == [0] and out.tobytes() == d.tobytes()
    big = np.concatenate([corpus8[:300000], np.zeros(100000, dtype=np.uint8), corpus8[300000:700000]])
    cuts = np.array([70000, 300000, 400000, 800000], dtype=np.uint64)
    zd = corpus.zdict()
    blob, offs = hmse_b200.compress(big, cuts, None, zd, ctx=ctx)
    out, status = hmse_b200.inflate(blob, offs, _sizes(cuts), zd, ctx=ctx)
    assert (status == 0).all() and out.tobytes() == big.tobytes()


def test_malformed_streams_are_reported(ctx, corpus8):
    import hmse_b200
    d = corpus8[:1 << 20]
    cuts = oracle.chunk_c(d)
    zd = corpus.zdict()
    blob, offs = hmse_b200.compress(d, cuts, None, zd, ctx=ctx)
    sizes = _sizes(cuts)
    raw = _raw(d, cuts)
    o = offs.astype(np.int64)
    bad = blob.copy()
    victims = {3: "flip", 10: "trailer", 20: "header", 30: "truncate"}
    bad[(o[3] + o[4]) // 2] ^= 0x5A            # somewhere in the body: bad code, wrong length or wrong checksum
    bad[o[11] - 1] ^= 1                        # Adler-32 trailer
    bad[o[20]] = 0x79                          # CMF
    offs2 = offs.copy()
    out, status = hmse_b200.inflate(bad, offs2, sizes, zd, ctx=ctx)
    for j in (3, 10, 20):
        assert status[j] != 0, j
    ok = np.ones(cuts.size, dtype=bool)
    ok[[3, 10, 20]] = False
    assert (status[ok] == 0).all()
    oo = np.concatenate([[0], np.cumsum(sizes)])
    for j in np.flatnonzero(ok)[:200].tolist() + [2, 4, 9, 11, 19, 21]:
        assert out[oo[j]:oo[j + 1]].tobytes() == raw[oo[j]:oo[j + 1]], j
    # wrong dictionary -> every stream fails the DICTID check; wrong expected sizes -> length / overrun
    out, status = hmse_b200.inflate(blob, offs, sizes, zd[:-1] + b"x", ctx=ctx)
    assert (status == 1).all()
    s2 = sizes.copy()
    s2[5] += 1
    s2[6] -= 1
    out, status = hmse_b200.inflate(blob, offs, s2, zd, ctx=ctx)
    assert status[5] == 5 and status[6] == 4 and (np.delete(status, [5, 6]) == 0).all()
    # a stream cut short
    cut = blob[:o[31] - 7].copy()
    offs3 = offs[:32].copy()
    offs3[31] = cut.size
    out, status = hmse_b200.inflate(cut, offs3, sizes[:31], zd, ctx=ctx)
    assert status[30] != 0 and (status[:30] == 0).all()


def test_digest_of_inflated_equals_digest_of_source(ctx, corpus8):
    # the size-independent property bench.py checks at full scale: SHA-256 of what the read path returns
    import torch
    import hmse_b200
    d = ctx.stage(corpus8)
    pipe = hmse_b200.Ingest(ctx, hmse_b200.CDCConfig(), corpus.zdict())
    r = pipe.run(d)
    bad, same = hmse_b200.verify_roundtrip(ctx, r, pipe.zdict)
    assert bad == 0 and same
