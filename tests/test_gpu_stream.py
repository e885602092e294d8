"""IngestStream (pieces, three CUDA streams, streaming dedup table) must leave in host memory exactly
what the one-shot pipeline computes, and both must equal the oracle.  Also the raw
hmse_dedup_begin/append entry points against oracle.dedup."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _one_shot(ctx, data, zd):
    import hmse_b200
    pipe = hmse_b200.Ingest(ctx, hmse_b200.CDCConfig(), zd)
    r = pipe.run(ctx.stage(data))
    return r


@pytest.mark.parametrize("overlap", [True, False])
@pytest.mark.parametrize("piece_kib", [128, 1024, 3000, 1 << 14])
def test_stream_equals_one_shot_and_oracle(ctx, corpus8, piece_kib, overlap):
    import torch
    import hmse_b200
    import oracle
    from oracle import corpus
    zd = corpus.zdict()
    data = corpus8[:6 * (1 << 20) + 12345]
    host = torch.from_numpy(data.copy()).pin_memory()
    st = hmse_b200.IngestStream(ctx, hmse_b200.CDCConfig(), zd, piece_bytes=(piece_kib << 10) & ~15, overlap=overlap)
    for _ in range(2):  # second run reuses every buffer and a cleared table
        h = st.run(host)
    r = _one_shot(ctx, data, zd)
    assert h.cuts.numpy().tolist() == r.cuts.cpu().numpy().tolist()
    assert np.array_equal(h.digests.numpy(), r.digests.cpu().numpy())
    assert np.array_equal(h.canon.numpy(), r.canon.cpu().numpy())
    assert np.array_equal(h.offsets.numpy(), r.offsets.cpu().numpy())
    assert np.array_equal(h.blob.numpy(), r.blob.cpu().numpy())
    # against the oracle
    cuts = oracle.chunk(data, oracle.CDCConfig())
    assert h.cuts.numpy().view(np.uint64).tolist() == cuts.tolist()
    dg = oracle.digest(data, cuts)
    assert np.array_equal(h.digests.numpy(), dg)
    canon, first = oracle.dedup(dg)
    assert np.array_equal(h.canon.numpy(), canon)
    sel = np.nonzero(first)[0]
    chunks = oracle.deflate.inflate_all(h.blob.numpy(), h.offsets.numpy().view(np.uint64), zd)
    starts = np.concatenate([[0], cuts[:-1]]).astype(np.int64)
    assert len(chunks) == sel.size
    for j, c in zip(sel.tolist(), chunks):
        assert c == data[starts[j]:int(cuts[j])].tobytes()
    assert h.h2d_bytes == data.size and h.d2h_bytes == h.n_chunks * 48 + sel.size * 8 + h.blob.numel()


def test_stream_small_inputs(ctx):
    import torch
    import hmse_b200
    import oracle
    rng = np.random.default_rng(5)
    st = hmse_b200.IngestStream(ctx, hmse_b200.CDCConfig(), b"", piece_bytes=128 << 10)
    for n in (1, 100, 2048, 40000, (128 << 10) + 1, (256 << 10)):
        data = rng.integers(0, 256, n, dtype=np.uint8)
        h = st.run(torch.from_numpy(data).pin_memory())
        cuts = oracle.chunk(data, oracle.CDCConfig())
        assert h.cuts.numpy().view(np.uint64).tolist() == cuts.tolist(), n
        assert np.array_equal(h.digests.numpy(), oracle.digest(data, cuts))


def test_dedup_append_matches_oracle(ctx):
    import torch
    import oracle
    rng = np.random.default_rng(11)
    base = rng.integers(0, 256, (700, 32), dtype=np.uint8)
    dg = base[rng.integers(0, 700, 5000)]
    canon_ref, first_ref = oracle.dedup(dg)
    d = torch.from_numpy(dg.copy()).to(ctx.tdev)
    lib = ctx.lib
    ctx.check(lib.hmse_dedup_begin(ctx.h, 5000, ctx.stream))
    canon = ctx.empty(5000, torch.int64)
    first = ctx.empty(5000, torch.uint8)
    pos = 0
    for step in (1, 999, 0, 2500, 1500):
        ctx.check(lib.hmse_dedup_append(ctx.h, d.data_ptr(), pos, step, canon[pos:].data_ptr(), first[pos:].data_ptr(),
                                        ctx.stream))
        pos += step
    torch.cuda.synchronize()
    assert np.array_equal(canon.cpu().numpy(), canon_ref)
    assert np.array_equal(first.cpu().numpy().astype(bool), first_ref)
    # contract errors: wrong n_prev, over capacity
    assert lib.hmse_dedup_append(ctx.h, d.data_ptr(), 17, 1, canon.data_ptr(), first.data_ptr(), ctx.stream) != 0
    assert lib.hmse_dedup_append(ctx.h, d.data_ptr(), 5000, 1 << 20, canon.data_ptr(), first.data_ptr(), ctx.stream) != 0


def test_run_many_prefetch_equals_single_runs(ctx, corpus8):
    """IngestStream.run_many: the next stream's input is copied in behind the current one's (two device input
    buffers); every result equals a plain run of that stream, in order, including streams of different sizes."""
    import torch
    import hmse_b200
    from oracle import corpus
    zd = corpus.zdict()
    streams = [corpus8[:3 << 20], corpus8[1 << 20:(5 << 20) + 777], corpus8[:4096], corpus8[2 << 20:6 << 20]]
    hosts = [torch.from_numpy(s.copy()).pin_memory() for s in streams]
    st = hmse_b200.IngestStream(ctx, hmse_b200.CDCConfig(), zd, piece_bytes=512 << 10)
    ref = hmse_b200.IngestStream(ctx, hmse_b200.CDCConfig(), zd, piece_bytes=512 << 10)
    k = 0
    for res in st.run_many(hosts):
        want = ref.run(hosts[k])
        assert np.array_equal(res.cuts.numpy(), want.cuts.numpy()) and np.array_equal(res.canon.numpy(), want.canon.numpy())
        assert np.array_equal(res.digests.numpy(), want.digests.numpy())
        assert np.array_equal(res.offsets.numpy(), want.offsets.numpy()) and np.array_equal(res.blob.numpy(), want.blob.numpy())
        k += 1
    assert k == len(hosts)
    # a plain run still works afterwards, and prefetching two streams ahead is refused
    assert np.array_equal(st.run(hosts[0]).cuts.numpy(), ref.run(hosts[0]).cuts.numpy())
    st.prefetch(hosts[1])
    st.prefetch(hosts[2])
    with pytest.raises(RuntimeError):
        st.prefetch(hosts[3])
    assert np.array_equal(st.run(hosts[1]).blob.numpy(), ref.run(hosts[1]).blob.numpy())
    assert np.array_equal(st.run(hosts[2]).blob.numpy(), ref.run(hosts[2]).blob.numpy())
