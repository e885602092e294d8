"""Archive container (hmse_b200/archive.py): the device-built ChunkIndex entries and pointer records equal the
oracle's restatement byte for byte; an archive restores to the input both through the device read path and through
the pure zlib walker of the oracle; damaged archives are rejected."""
import numpy as np
import pytest

import oracle
from oracle import corpus

pytestmark = pytest.mark.gpu


def _ingest(ctx, data, zd):
    import hmse_b200
    return hmse_b200.Ingest(ctx, hmse_b200.CDCConfig(), zd).run(ctx.stage(data))


def test_records_equal_oracle_and_roundtrip(ctx, corpus8, tmp_path):
    import hmse_b200
    from hmse_b200 import archive
    zd = corpus.zdict()
    data = np.concatenate([corpus8[:3 << 20], corpus8[1 << 20:2 << 20], corpus8[:1 << 20]])   # plenty of repeats
    r = _ingest(ctx, data, zd)
    ar = archive.build(r, zd, ctx=ctx)
    cuts = r.cuts.cpu().numpy().view(np.uint64)
    idx, ptr = oracle.archive.records(r.digests.cpu().numpy(), r.canon.cpu().numpy(), cuts, r.select.cpu().numpy(),
                                      r.offsets.cpu().numpy().view(np.uint64))
    assert np.array_equal(ar.index, idx) and np.array_equal(ar.pointers, ptr)
    assert int(ar.index[:, 38:40].copy().view("<u2").max()) >= 2            # duplicates are reference counted
    buf = ar.tobytes()
    assert buf == oracle.archive.pack(zd, idx, ptr, r.blob.cpu().numpy(), data.size)
    assert len(buf) < data.size // 3
    # file round trip, device read path, and the zlib-only walker
    path = tmp_path / "a.hmse"
    ar.save(str(path))
    back = archive.Archive.load(str(path))
    assert archive.restore(back, ctx=ctx).tobytes() == data.tobytes()
    assert oracle.archive.restore(buf) == data.tobytes()


def test_archive_from_stream_and_small_inputs(ctx, corpus8):
    import torch
    import hmse_b200
    from hmse_b200 import archive
    zd = corpus.zdict()
    st = hmse_b200.IngestStream(ctx, hmse_b200.CDCConfig(), zd, piece_bytes=512 << 10)
    data = corpus8[:(2 << 20) + 777]
    h = st.run(torch.from_numpy(data.copy()).pin_memory())
    ar = archive.build(h, zd, ctx=ctx)
    assert archive.restore(ar, ctx=ctx).tobytes() == data.tobytes()
    for n in (1, 2047, 5000):
        d = corpus8[:n]
        ar = archive.build(_ingest(ctx, d, b""), b"", ctx=ctx)
        assert ar.n_chunks == oracle.chunk(d).size
        assert archive.restore(archive.Archive.frombytes(ar.tobytes()), ctx=ctx).tobytes() == d.tobytes()


def test_damaged_archives_are_rejected(ctx, corpus8):
    from hmse_b200 import archive
    zd = corpus.zdict()
    data = corpus8[:1 << 20]
    ar = archive.build(_ingest(ctx, data, zd), zd, ctx=ctx)
    buf = bytearray(ar.tobytes())
    bad = bytearray(buf)
    bad[-100] ^= 0xFF                                   # inside the last stored chunk
    with pytest.raises(ValueError):
        archive.restore(archive.Archive.frombytes(bytes(bad)), ctx=ctx)
    bad = bytearray(buf)
    o = 64 + len(zd) + ((-len(zd)) % 8) + ar.n_unique * 40
    bad[o + 4] ^= 1                                     # a pointer record's in-sector offset
    with pytest.raises(ValueError):
        archive.restore(archive.Archive.frombytes(bytes(bad)), ctx=ctx)
    with pytest.raises(ValueError):
        archive.Archive.frombytes(bytes(buf[:-1]))
    with pytest.raises(ValueError):
        archive.Archive.frombytes(b"XXXXXXXX" + bytes(buf[8:]))


def test_l4_archive_records_equal_oracle_and_roundtrip(ctx, corpus8):
    """Ingest with the similarity layer: near-duplicate first occurrences become DeltaChunk records
    (README.md:2182-2189); records equal the oracle's byte for byte; both read paths rebuild the stream
    (README.md:1329 "Read Base -> Retrieve Delta -> Apply Delta -> Decompress ... 100 % checksum pass")."""
    import hmse_b200
    from hmse_b200 import archive
    zd = corpus.zdict()
    data = np.concatenate([corpus8[:6 << 20], corpus8[2 << 20:3 << 20]])
    r = hmse_b200.Ingest(ctx, hmse_b200.CDCConfig(), zd).run(ctx.stage(data), l4=hmse_b200.SimConfig())
    base = r.base.cpu().numpy()
    kept = base >= 0
    assert kept.sum() > 20
    first = r.is_first.cpu().numpy()
    sel = r.select.cpu().numpy()
    assert np.array_equal(sel, np.flatnonzero(first & ~kept))           # deltas are not compressed
    cuts = r.cuts.cpu().numpy().view(np.uint64)
    # the L4 outputs equal the oracle's on the same input
    keys = oracle.band_keys(oracle.minhash_c(data, cuts))
    wbase, wblob, woffs = oracle.delta(data, cuts, keys, first)
    assert np.array_equal(base, wbase) and np.array_equal(r.delta_blob.cpu().numpy(), wblob)
    assert np.array_equal(r.delta_offsets.cpu().numpy().view(np.uint64), woffs)
    ar = archive.build(r, zd, ctx=ctx)
    idx, ptr, dstore, nd = oracle.archive.records_l4(r.digests.cpu().numpy(), r.canon.cpu().numpy(), cuts, sel,
                                                     r.offsets.cpu().numpy().view(np.uint64), base, wblob, woffs)
    assert nd == kept.sum() == ar.n_delta
    assert np.array_equal(ar.index, idx) and np.array_equal(ar.pointers, ptr) and np.array_equal(ar.delta_store, dstore)
    buf = ar.tobytes()
    assert buf == oracle.archive.pack(zd, idx, ptr, r.blob.cpu().numpy(), data.size, dstore, nd)
    # smaller than the archive without L4
    plain = archive.build(hmse_b200.Ingest(ctx, hmse_b200.CDCConfig(), zd).run(ctx.stage(data)), zd, ctx=ctx)
    assert len(buf) < len(plain.tobytes())
    back = archive.Archive.frombytes(buf)
    assert back.n_delta == nd
    assert archive.restore(back, ctx=ctx).tobytes() == data.tobytes()
    assert oracle.archive.restore(buf) == data.tobytes()
    # damage inside the delta store is caught (a COPY that runs outside its base, or a wrong length)
    bad = bytearray(buf)
    bad[len(buf) - ar.delta_store.size + 8] = 0xFF
    bad[len(buf) - ar.delta_store.size + 9] = 0xFF
    with pytest.raises(ValueError):
        archive.restore(archive.Archive.frombytes(bytes(bad)), ctx=ctx)


def test_cross_shard_canon_is_rejected_not_dereferenced(ctx, corpus8):
    """A shard result of ShardedIngest carries GLOBAL ids in canon: a chunk whose first occurrence lies in an earlier
    shard (canon < id_base) or a later range has no record in this shard's store.  hmse_index_build must report it
    (HMSE_E_INVAL) instead of indexing its slot table with a wrapped value (archive.cu:25)."""
    import copy
    import hmse_b200
    from hmse_b200 import archive
    zd = corpus.zdict()
    data = np.concatenate([corpus8[:2 << 20], corpus8[:1 << 20]])
    r = _ingest(ctx, data, zd)
    # the same result seen as a shard whose global ids start at 1000: canon shifted with it - fine, records unchanged
    ok = copy.copy(r)
    ok.id_base = 1000
    ok.canon = r.canon + 1000
    a0, a1 = archive.build(r, zd, ctx=ctx), archive.build(ok, zd, ctx=ctx)
    assert np.array_equal(a0.index, a1.index) and np.array_equal(a0.pointers, a1.pointers)
    for bad_value in (999, 5, 1000 + r.n_chunks, -1):
        bad = copy.copy(ok)
        bad.canon = ok.canon.clone()
        bad.canon[r.n_chunks // 2] = bad_value
        with pytest.raises(hmse_b200.HmseError) as ei:
            archive.build(bad, zd, ctx=ctx)
        assert ei.value.code == -1 and "cross-shard" in str(ei.value)
    # a chunk that resolves to a local chunk which is not in the store (not selected)
    bad = copy.copy(r)
    dup = int((~r.is_first).nonzero()[0])
    bad.canon = r.canon.clone()
    bad.canon[0] = dup
    with pytest.raises(hmse_b200.HmseError):
        archive.build(bad, zd, ctx=ctx)
    assert archive.restore(archive.build(r, zd, ctx=ctx), ctx=ctx).tobytes() == data.tobytes()   # the ctx still works
