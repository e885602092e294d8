"""GPU parity of L4 delta coding (csrc/delta.cu) against oracle/deltacode.py: base selection (index work) and
delta bytes are compared bit for bit; the read path must rebuild every target and reject malformed deltas."""
import numpy as np
import pytest

import oracle
from oracle import deltacode as D

pytestmark = pytest.mark.gpu


def _t64(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64).copy()).cuda()


def _bases_gpu(ctx, keys, first, min_votes=4):
    import torch
    n, bands = keys.shape
    band, key, ids = ctx.lsh_buckets(_t64(keys))
    f = torch.from_numpy(np.ascontiguousarray(first, dtype=np.uint8)).cuda()
    return ctx.delta_bases(band, key, ids, n, bands, f, min_votes).cpu().numpy()


def test_bases_match_oracle_on_corpus(ctx, corpus8):
    d = corpus8[:6 << 20]
    cuts = oracle.chunk_c(d)
    _, first = oracle.dedup(oracle.digest(d, cuts))
    keys = oracle.band_keys(oracle.minhash_c(d, cuts))
    want = D.delta_bases(keys, first)
    assert (want >= 0).sum() > 20
    assert np.array_equal(_bases_gpu(ctx, keys, first), want)
    for mv in (1, 2, 9, 32):
        assert np.array_equal(_bases_gpu(ctx, keys, first, mv), D.delta_bases(keys, first, mv)), mv


@pytest.mark.parametrize("n,bands,vals", [(1, 32, 3), (40, 32, 2), (700, 32, 5), (3000, 32, 40), (2500, 8, 6), (900, 1, 4)])
def test_bases_heavy_ties(ctx, n, bands, vals):
    rng = np.random.default_rng(n * 31 + bands)
    keys = rng.integers(0, vals, (n, bands)).astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    first = rng.integers(0, 4, n) > 0
    for mv in (1, 4, 7):
        assert np.array_equal(_bases_gpu(ctx, keys, first, mv), D.delta_bases(keys, first, mv)), mv


def test_delta_api_matches_oracle_on_corpus(ctx, corpus8):
    import hmse_b200
    d = corpus8
    cuts = oracle.chunk_c(d)
    _, first = oracle.dedup(oracle.digest(d, cuts))
    keys = oracle.band_keys(oracle.minhash_c(d, cuts))
    wbase, wblob, woffs = oracle.delta(d, cuts, keys, first)
    base, blob, offs = hmse_b200.delta(d, cuts, keys, first, ctx=ctx)
    assert np.array_equal(base, wbase)
    assert np.array_equal(offs, woffs)
    assert np.array_equal(blob, wblob)
    assert (wbase >= 0).sum() > 50
    # read path on the device
    kept = np.flatnonzero(base >= 0)
    starts = np.concatenate([[0], cuts[:-1]]).astype(np.int64)
    ln = cuts.astype(np.int64) - starts
    bases = np.concatenate([d[starts[j]:int(cuts[j])] for j in base[kept]])
    out, status = hmse_b200.delta_apply(blob, _dense_offsets(offs, kept), bases, ln[base[kept]], ln[kept], ctx=ctx)
    assert not status.any()
    assert np.array_equal(out, np.concatenate([d[starts[i]:int(cuts[i])] for i in kept]))


def _dense_offsets(offs, kept):
    """offsets of the kept deltas only (they are packed back to back in chunk order)."""
    return np.concatenate([offs[kept], offs[-1:]]).astype(np.uint64)


def _pairs_to_stream(pairs):
    """data = base0 target0 base1 target1 ...; cuts; base[] with target 2k+1 -> 2k."""
    parts, cuts, base, pos = [], [], [], 0
    for b, t in pairs:
        for x in (b, t):
            parts.append(np.frombuffer(x, dtype=np.uint8))
            pos += len(x)
            cuts.append(pos)
        base += [-1, len(base)]
    return np.concatenate(parts), np.array(cuts, dtype=np.uint64), np.array(base, dtype=np.int64)


def _edit(rng, base, n_edits, alphabet=(97, 123)):
    t = bytearray(base)
    for _ in range(n_edits):
        pos = int(rng.integers(0, max(1, len(t))))
        op = int(rng.integers(0, 3))
        if op == 0:
            t[pos:pos] = bytes(rng.integers(alphabet[0], alphabet[1], int(rng.integers(1, 40)), dtype=np.uint8))
        elif op == 1:
            del t[pos:pos + int(rng.integers(1, 40))]
        elif len(t):
            t[pos] = 33
    return bytes(t[:32768]) or b"x"


def _check_pairs(ctx, pairs):
    import torch
    data, cuts, base = _pairs_to_stream(pairs)
    dd = ctx.stage(data)
    bt = torch.from_numpy(base.copy()).cuda()
    blob, offs = ctx.delta_encode(dd, _t64(cuts), bt)
    blob, offs, got_base = blob.cpu().numpy(), offs.cpu().numpy().view(np.uint64), bt.cpu().numpy()
    for k, (b, t) in enumerate(pairs):
        want = D.delta_encode(t, b)
        got = blob[int(offs[2 * k + 1]):int(offs[2 * k + 2])].tobytes()
        assert offs[2 * k] == offs[2 * k + 1]
        if want is None:
            assert got == b"" and got_base[2 * k + 1] == -1, (k, len(b), len(t))
        else:
            assert got == want, (k, len(b), len(t), got[:16].hex(), want[:16].hex())
            assert got_base[2 * k + 1] == 2 * k
    return blob, offs, got_base


def test_encode_random_edit_pairs(ctx):
    rng = np.random.default_rng(7)
    pairs = []
    for k in range(300):
        n = int(rng.integers(1, 32769)) if k % 3 else int(rng.integers(1, 600))
        lo, hi = ((97, 123), (0, 256), (97, 99))[k % 3]          # text, binary, low-entropy (long false runs)
        b = bytes(rng.integers(lo, hi, n, dtype=np.uint8))
        pairs.append((b, _edit(rng, b, int(rng.integers(0, 25)), (lo, hi))))
    _check_pairs(ctx, pairs)


def test_encode_edge_cases(ctx):
    rng = np.random.default_rng(8)
    r = lambda n: bytes(rng.integers(0, 256, n, dtype=np.uint8))  # noqa: E731
    a = r(32768)
    pairs = [
        (a, a),                                     # maximum size, one COPY
        (a, a[1:] + b"\x00"),                       # shifted by one
        (a[:20000], a[10000:30000]),                # half overlap: rejected or not, must agree
        (bytes(32768), bytes(32768)),               # constant runs: every window hashes alike
        (bytes(5000), bytes(4999) + b"\x01"),
        (b"ab" * 3000, b"ab" * 2999 + b"ba"),       # period 2
        (r(7), r(40)),                              # base shorter than a seed
        (r(40), r(7)),                              # target shorter than a seed
        (r(8), r(8)),
        (a[:8], a[:8] * 5),                         # 40 bytes made of one seed
        (a[:100], a[:100]),                         # 100 bytes: cap 20, needs 3
        (a[:39], a[:39]),                           # cap 7
        (a[:4], a[:4]),                             # cap 0
        (a[:1000], r(1000)),                        # nothing in common
        (a[:1000], a[:790] + r(210)),               # just over 20 %
        (a[:1000], a[:810] + r(190)),               # just under
        (a[:3000], a[2000:3000] + a[1000:2000] + a[:1000]),   # backward offsets
        (a[:3000], a[:1000] + a[:1000] + a[:1000]),           # repeated source
    ]
    _check_pairs(ctx, pairs)


def test_encode_no_candidates_and_empty(ctx):
    import torch
    data = np.frombuffer(b"hello world, hello world, hello world", dtype=np.uint8)
    cuts = np.array([10, 20, len(data)], dtype=np.uint64)
    bt = torch.full((3,), -1, dtype=torch.int64).cuda()
    blob, offs = ctx.delta_encode(ctx.stage(data), _t64(cuts), bt)
    assert blob.numel() == 0 and not offs.cpu().numpy().any()
    blob, offs = ctx.delta_encode(ctx.stage(data), _t64(cuts[:0]), bt[:0])
    assert blob.numel() == 0 and offs.cpu().numpy().tolist() == [0]


def test_apply_rejects_malformed_like_the_oracle(ctx):
    import hmse_b200
    base = bytes(range(200))
    good = D.delta_encode(base, base)
    cases = [good, b"", good[:-1], good + b"\x00", bytes([0x00]), bytes([0x80]),
             bytes([0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0x01]), bytes([0xFF, 0xFF, 0xFF, 0xFF, 0x1F]),
             bytes([0x91, 0x03, 0x00]), bytes([0x93, 0x03, 0x00]), bytes([0x91, 0x03, 0x02]), bytes([0x91, 0x03, 0x01]),
             bytes([0x90, 0x03]), bytes([0x90, 0x03]) + base[:199], bytes([0x90, 0x03]) + base]
    offs = np.concatenate([[0], np.cumsum([len(c) for c in cases])]).astype(np.uint64)
    blob = np.frombuffer(b"".join(cases), dtype=np.uint8)
    m = len(cases)
    out, status = hmse_b200.delta_apply(blob, offs, np.frombuffer(base * m, dtype=np.uint8), np.full(m, 200), np.full(m, 200),
                                        ctx=ctx)
    for k, c in enumerate(cases):
        try:
            want = D.delta_apply(c, base, 200)
        except ValueError:
            want = None
        assert (status[k] == 0) == (want is not None), (k, c.hex(), int(status[k]))
        if want is not None:
            assert out[200 * k:200 * (k + 1)].tobytes() == want
    assert status[0] == 0 and status[-1] == 0


def test_roundtrip_larger_stream(ctx):
    """Property at a larger size (no oracle in the loop): every kept delta rebuilds its chunk on the device."""
    import torch
    import hmse_b200
    from hmse_b200 import corpus as pc
    n = 64 << 20
    d = pc.DeviceCorpus(ctx).generate(n)
    cuts = ctx.chunk(d, hmse_b200.CDCConfig())
    canon, first = ctx.dedup(ctx.digest(d, cuts))
    sig = ctx.minhash(d, cuts, hmse_b200.SimConfig())
    keys = ctx.lsh_keys(sig, hmse_b200.SimConfig())
    base, blob, offs = hmse_b200.delta(d, cuts, keys, first, ctx=ctx)
    kept = torch.nonzero(base >= 0).view(-1)
    assert kept.numel() > 500
    starts = torch.cat([torch.zeros(1, dtype=torch.int64, device=cuts.device), cuts[:-1]])
    ln = cuts - starts
    bj = base[kept]
    assert bool((bj < kept).all()) and bool(first[kept].all()) and bool((base[bj] == -1).all())
    dl = (offs[1:] - offs[:-1])
    assert bool((dl[kept] * 5 <= ln[kept]).all()) and bool((dl[base < 0] == 0).all())
    out_off = torch.cat([torch.zeros(1, dtype=torch.int64, device=cuts.device), torch.cumsum(ln[kept], 0)])
    doff = torch.cat([offs[kept], offs[-1:]])
    out, status, bad = ctx.delta_apply(blob, doff, d, starts[bj].contiguous(), ln[bj].to(torch.int32).contiguous(), out_off)
    assert bad == 0
    want = torch.cat([d[int(s):int(e)] for s, e in zip(starts[kept].tolist(), cuts[kept].tolist())])
    assert torch.equal(out, want)


def test_minhash_select_rows_equal_full_rows(ctx, corpus8):
    """hmse_minhash_select: signatures of the selected chunks only, equal to the oracle's rows; odd and even shingle
    counts and tiny chunks exercise the pairwise inner loop."""
    import torch
    import hmse_b200
    d = corpus8[:2 << 20]
    cuts = oracle.chunk_c(d)
    want = oracle.minhash_c(d, cuts)
    sel = np.flatnonzero(np.arange(cuts.size) % 3 != 1)
    dd = ctx.stage(d)
    got = ctx.minhash(dd, _t64(cuts), hmse_b200.SimConfig(), select=torch.from_numpy(sel.astype(np.int64)).cuda())
    assert np.array_equal(got.cpu().numpy().view(np.uint32), want[sel])
    # chunk lengths 0..70 bytes: 0, 1, 2 ... shingles
    small = np.cumsum(np.arange(1, 71)).astype(np.uint64)
    sd = corpus8[:int(small[-1])]
    got = ctx.minhash(ctx.stage(sd), _t64(small), hmse_b200.SimConfig())
    assert np.array_equal(got.cpu().numpy().view(np.uint32), oracle.minhash_c(sd, small))


def test_gpu_matches_committed_delta_golden(ctx):
    """GPU vs tests/golden/delta_golden.json directly (not through the oracle)."""
    import hashlib
    import json
    import os
    import sys
    import hmse_b200
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden"))
    import make_delta_golden as mk
    g = json.load(open(os.path.join(here, "golden", "delta_golden.json")))
    pairs = mk.pairs()
    blob, offs, got_base = _check_pairs(ctx, pairs)
    for k, want in enumerate(g["pairs"]):
        got = blob[int(offs[2 * k + 1]):int(offs[2 * k + 2])].tobytes()
        assert (got.hex() if got else None) == want["delta_hex"], k
    d = hmse_b200.corpus.DeviceCorpus(ctx).generate(g["n"])
    assert hashlib.sha256(d.cpu().numpy().tobytes()).hexdigest() == g["input_sha256"]
    r = hmse_b200.Ingest(ctx).run(d, compress=False, l4=hmse_b200.SimConfig())
    base = r.base.cpu().numpy()
    assert r.n_chunks == g["chunks"]
    assert hashlib.sha256(base.astype("<i8").tobytes()).hexdigest() == g["base_sha256"]
    assert [[int(i), int(base[i])] for i in np.flatnonzero(base >= 0)] == g["kept"]
    dblob = r.delta_blob.cpu().numpy()
    doffs = r.delta_offsets.cpu().numpy().view(np.uint64)
    assert dblob.size == g["delta_bytes"] and hashlib.sha256(dblob.tobytes()).hexdigest() == g["blob_sha256"]
    assert hashlib.sha256(doffs.astype("<u8").tobytes()).hexdigest() == g["offsets_sha256"]
    for i, hx in zip(np.flatnonzero(base >= 0)[:3], g["first_deltas_hex"]):
        assert dblob[int(doffs[i]):int(doffs[i + 1])].tobytes().hex() == hx


def test_l4_pipeline_edge_cases(ctx, corpus8):
    """Ingest.run(l4=...) on degenerate streams: empty, one chunk, every chunk a duplicate, chunks above the 32 KiB
    delta limit (they are never delta coded), and an archive round trip of each."""
    import torch
    import hmse_b200
    from hmse_b200 import archive
    zd = oracle.corpus.zdict()
    sim = hmse_b200.SimConfig()
    ing = hmse_b200.Ingest(ctx, hmse_b200.CDCConfig(), zd)
    # empty input
    r = ing.run(ctx.stage(np.zeros(0, dtype=np.uint8)), l4=sim)
    assert r.n_chunks == 0 and r.base is None
    # one chunk
    one = corpus8[:1500]
    r = ing.run(ctx.stage(one), l4=sim)
    assert r.n_chunks == 1 and r.base.cpu().tolist() == [-1] and r.delta_blob.numel() == 0
    assert archive.restore(archive.build(r, zd, ctx=ctx), ctx=ctx).tobytes() == one.tobytes()
    # every chunk after the first copy is an exact duplicate: nothing is similar-but-different
    rep = np.concatenate([corpus8[:300000]] * 3)
    r = ing.run(ctx.stage(rep), l4=sim)
    base = r.base.cpu().numpy()
    first = r.is_first.cpu().numpy()
    assert (base[~first] == -1).all()
    assert archive.restore(archive.build(r, zd, ctx=ctx), ctx=ctx).tobytes() == rep.tobytes()
    # 64 KiB maximum chunk size: near-duplicate chunks longer than 32768 bytes keep no delta, shorter ones do
    m32 = hmse_b200.CDCConfig.for_avg(32768)
    big_cfg = hmse_b200.CDCConfig(8192, 32768, 65536, m32.mask_s, m32.mask_l)     # chunks of 8 .. 64 KiB
    a = corpus8[:1 << 20]
    b = a.copy()
    b[5000::40000] ^= 0x20
    both = np.concatenate([a, b])
    r = hmse_b200.Ingest(ctx, big_cfg, zd).run(ctx.stage(both), l4=sim)
    cuts = r.cuts.cpu().numpy().view(np.uint64)
    ln = np.diff(np.concatenate([[0], cuts]).astype(np.int64))
    base = r.base.cpu().numpy()
    assert (ln > 32768).any() and (base >= 0).any()
    assert (base[ln > 32768] == -1).all()
    assert (ln[base[base >= 0]] <= 32768).all()
    keys = oracle.band_keys(oracle.minhash_c(both, cuts))
    wbase, wblob, _ = oracle.delta(both, cuts, keys, r.is_first.cpu().numpy())
    assert np.array_equal(base, wbase) and np.array_equal(r.delta_blob.cpu().numpy(), wblob)
    assert archive.restore(archive.build(r, zd, ctx=ctx), ctx=ctx).tobytes() == both.tobytes()


def test_delta_api_errors(ctx):
    import torch
    import hmse_b200
    n = 8
    keys = torch.arange(n * 64, dtype=torch.int64).view(n, 64).cuda()
    band, key, ids = ctx.lsh_buckets(keys)
    ones = torch.ones(n, dtype=torch.uint8).cuda()
    with pytest.raises(hmse_b200.HmseError):
        ctx.delta_bases(band, key, ids, n, 64, ones, 4)          # more than 32 bands
    k32 = torch.arange(n * 32, dtype=torch.int64).view(n, 32).cuda()
    band, key, ids = ctx.lsh_buckets(k32)
    with pytest.raises(hmse_b200.HmseError):
        ctx.delta_bases(band, key, ids, n, 32, ones, 0)          # min_votes 0
    assert ctx.delta_bases(band, key, ids, n, 32, ones, 4).cpu().tolist() == [-1] * n


def test_delta_encode_capacity_protocol(ctx):
    """Too small an output buffer: HMSE_E_CAPACITY with the required size, the candidate list untouched; a second call
    with that size succeeds and gives the oracle's bytes."""
    import ctypes as C
    import torch
    from hmse_b200 import _lib
    rng = np.random.default_rng(21)
    pairs = []
    for _ in range(8):
        b = bytes(rng.integers(97, 123, 6000, dtype=np.uint8))
        t = bytearray(b)
        t[3000:3000] = b"0123456789" * 5
        pairs.append((b, bytes(t)))
    data, cuts, base = _pairs_to_stream(pairs)
    dd, ct = ctx.stage(data), _t64(cuts)
    bt = torch.from_numpy(base.copy()).cuda()
    n = cuts.size
    offs = ctx.empty(n + 1, torch.int64)
    total = C.c_uint64(0)
    small = ctx.empty(64, torch.uint8)
    rc = ctx.lib.hmse_delta_encode(ctx.h, dd.data_ptr(), 0, ct.data_ptr(), n, bt.data_ptr(), small.data_ptr(), 16,
                                   offs.data_ptr(), C.byref(total), ctx.stream)
    assert rc == _lib.HMSE_E_CAPACITY and total.value > 16
    want = [D.delta_encode(t, b) for b, t in pairs]
    assert total.value == sum(len(w) for w in want)
    assert np.array_equal(bt.cpu().numpy(), base)          # still the candidates: nothing was rejected
    out = ctx.empty(int(total.value) + 64, torch.uint8)
    ctx.check(ctx.lib.hmse_delta_encode(ctx.h, dd.data_ptr(), 0, ct.data_ptr(), n, bt.data_ptr(), out.data_ptr(), int(total.value),
                                        offs.data_ptr(), C.byref(total), ctx.stream))
    assert out[:total.value].cpu().numpy().tobytes() == b"".join(want)
