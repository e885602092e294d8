"""GPU parity: SHA-256 digests, exact dedup, MinHash signatures and LSH keys vs the oracle."""
import hashlib

import numpy as np
import pytest

import oracle
from oracle import corpus

pytestmark = pytest.mark.gpu


def test_sha256_known_answers(ctx):
    import hmse_b200
    msgs = [b"", b"abc", b"abcdbcdecdefdefgefghfghighijhijkijkljklmklmnlmnomnopnopq", b"a" * 55, b"a" * 56, b"a" * 63,
            b"a" * 64, b"a" * 65, b"a" * 119, b"a" * 120, b"a" * 127, b"a" * 128, b"a" * 1000]
    data = b"".join(msgs)
    cuts = np.cumsum([len(m) for m in msgs]).astype(np.uint64)
    # the empty message is chunk 0 with cut 0: lengths of zero are legal for digest()
    got = hmse_b200.digest(data + b"\0" * 8, cuts, ctx=ctx)
    for j, m in enumerate(msgs):
        assert bytes(got[j]) == hashlib.sha256(m).digest(), "message %d (len %d)" % (j, len(m))


@pytest.mark.parametrize("offset", [0, 1, 2, 3, 5])
def test_sha256_corpus_unaligned(ctx, corpus8, offset):
    import hmse_b200
    d = corpus8[offset:(2 << 20) + offset]
    cuts = oracle.chunk_c(d)
    assert np.array_equal(hmse_b200.digest(d, cuts, ctx=ctx), oracle.digest(d, cuts))


def test_sha256_all_lengths(ctx, corpus8):
    import hmse_b200
    lens = np.arange(0, 300)
    cuts = np.cumsum(lens).astype(np.uint64)
    d = corpus8[:int(cuts[-1]) + 16]
    assert np.array_equal(hmse_b200.digest(d, cuts, ctx=ctx), oracle.digest(d, cuts))


def test_sha256_start0(ctx, corpus8):
    import hmse_b200
    d = corpus8[:1 << 20]
    cuts = oracle.chunk_c(d)
    assert np.array_equal(hmse_b200.digest(d, cuts[3:], start0=int(cuts[2]), ctx=ctx),
                          oracle.digest(d, cuts[3:], start0=int(cuts[2])))


def test_dedup_corpus(ctx, corpus8):
    import hmse_b200
    cuts = oracle.chunk_c(corpus8)
    dg = oracle.digest(corpus8, cuts)
    canon, first = hmse_b200.dedup(dg, ctx=ctx)
    wc, wf = oracle.dedup(dg)
    assert np.array_equal(canon, wc) and np.array_equal(first, wf)
    assert first.sum() < first.size  # the corpus does contain duplicates


def test_dedup_five_copies(ctx, corpus8):
    # README.md:1210, 1299-1315: 5 copies => physical ~ 1 copy
    import hmse_b200
    one = corpus8[:2 << 20]
    d = np.tile(one, 5)
    cuts = hmse_b200.chunk(d, ctx=ctx)
    dg = hmse_b200.digest(d, cuts, ctx=ctx)
    canon, first = hmse_b200.dedup(dg, ctx=ctx)
    wc, wf = oracle.dedup(dg)
    assert np.array_equal(canon, wc) and np.array_equal(first, wf)
    lens = np.diff(np.concatenate([[0], cuts]).astype(np.int64))
    assert lens[first].sum() <= 1.1 * one.size


def test_dedup_heavy_collisions(ctx):
    import hmse_b200
    rng = np.random.default_rng(1)
    base = rng.integers(0, 256, (50, 32), dtype=np.uint8)
    dg = base[rng.integers(0, 50, 20000)]
    canon, first = hmse_b200.dedup(dg, ctx=ctx)
    wc, wf = oracle.dedup(dg)
    assert np.array_equal(canon, wc) and np.array_equal(first, wf)
    assert hmse_b200.dedup(np.zeros((0, 32), np.uint8), ctx=ctx)[0].size == 0


def test_dedup_sharded_records(ctx, corpus8):
    """partition -> (exchange emulated in-process) -> owner table -> scatter == global dedup."""
    import ctypes as C
    import torch
    cuts = oracle.chunk_c(corpus8)
    dg = oracle.digest(corpus8, cuts)
    n = dg.shape[0]
    world = 4
    bounds = np.linspace(0, n, world + 1).astype(np.int64)
    lib, h = ctx.lib, ctx.h
    send, perms, counts = [], [], []
    for r in range(world):
        part = torch.from_numpy(dg[bounds[r]:bounds[r + 1]].copy()).cuda()
        m = part.shape[0]
        rec = torch.empty(m * 40, dtype=torch.uint8, device="cuda")
        perm = torch.empty(m, dtype=torch.int32, device="cuda")
        cnt = (C.c_uint64 * world)()
        ctx.check(lib.hmse_dedup_partition(h, part.data_ptr(), m, int(bounds[r]), world, rec.data_ptr(), perm.data_ptr(),
                                           cnt, ctx.stream))
        send.append(rec.view(m, 40))
        perms.append(perm)
        counts.append(list(cnt))
        assert sum(cnt) == m
    canon = np.empty(n, dtype=np.int64)
    first = np.empty(n, dtype=bool)
    replies = [[None] * world for _ in range(world)]
    for o in range(world):      # owner o receives its slice from every sender, in rank order
        pieces = []
        for r in range(world):
            off = sum(counts[r][:o])
            pieces.append(send[r][off:off + counts[r][o]])
        rec = torch.cat(pieces).contiguous()
        m = rec.shape[0]
        out = torch.empty(m, dtype=torch.int64, device="cuda")
        ctx.check(lib.hmse_dedup_records(h, rec.data_ptr(), m, out.data_ptr(), ctx.stream))
        pos = 0
        for r in range(world):
            replies[r][o] = out[pos:pos + counts[r][o]]
            pos += counts[r][o]
    for r in range(world):
        rep = torch.cat(replies[r]).contiguous()
        m = int(bounds[r + 1] - bounds[r])
        c = torch.empty(m, dtype=torch.int64, device="cuda")
        f = torch.empty(m, dtype=torch.uint8, device="cuda")
        ctx.check(lib.hmse_dedup_scatter(h, rep.data_ptr(), perms[r].data_ptr(), m, int(bounds[r]), c.data_ptr(),
                                         f.data_ptr(), ctx.stream))
        canon[bounds[r]:bounds[r + 1]] = c.cpu().numpy()
        first[bounds[r]:bounds[r + 1]] = f.cpu().numpy().astype(bool)
    wc, wf = oracle.dedup(dg)
    assert np.array_equal(canon, wc) and np.array_equal(first, wf)


def test_minhash_and_keys(ctx, corpus8):
    import hmse_b200
    d = corpus8[:1 << 20]
    cuts = oracle.chunk_c(d)
    sig = ctx.minhash(ctx.stage(d), ctx.stage_u64(cuts), hmse_b200.SimConfig()).cpu().numpy().view(np.uint32)
    want = oracle.minhash_c(d, cuts)
    assert np.array_equal(sig, want)
    import torch
    keys = ctx.lsh_keys(torch.from_numpy(sig.view(np.int32).copy()).cuda(), hmse_b200.SimConfig())
    keys = keys.cpu().numpy().view(np.uint64)
    assert np.array_equal(keys, oracle.band_keys(want))


def test_minhash_short_chunks_and_seed0(ctx, corpus8):
    import hmse_b200
    lens = np.array([0, 1, 2, 3, 4, 5, 35, 36, 37, 100, 4099])
    cuts = np.cumsum(lens).astype(np.uint64)
    d = corpus8[7:7 + int(cuts[-1]) + 16]
    ocfg = oracle.SimConfig(seeds=tuple(range(0, 128)))       # the skeleton's 0..127 variant
    pcfg = hmse_b200.SimConfig(seeds=tuple(range(0, 128)))
    sig = ctx.minhash(ctx.stage(d), ctx.stage_u64(cuts), pcfg).cpu().numpy().view(np.uint32)
    want = oracle.minhash(d, cuts, ocfg)
    assert np.array_equal(sig, want)
    assert (sig[:4] == 0xFFFFFFFF).all()


@pytest.mark.parametrize("n_perm,bands", [(32, 8), (64, 16), (256, 32)])
def test_minhash_other_shapes(ctx, corpus8, n_perm, bands):
    import torch
    import hmse_b200
    d = corpus8[:256 << 10]
    cuts = oracle.chunk_c(d)
    seeds = tuple(range(1, n_perm + 1))
    ocfg = oracle.SimConfig(n_perm, bands, n_perm // bands, seeds)
    pcfg = hmse_b200.SimConfig(n_perm, bands, n_perm // bands, seeds)
    sig = ctx.minhash(ctx.stage(d), ctx.stage_u64(cuts), pcfg)
    want = oracle.minhash_c(d, cuts, ocfg)
    assert np.array_equal(sig.cpu().numpy().view(np.uint32), want)
    keys = ctx.lsh_keys(sig, pcfg).cpu().numpy().view(np.uint64)
    assert np.array_equal(keys, oracle.band_keys(want, ocfg))
    torch.cuda.synchronize()
