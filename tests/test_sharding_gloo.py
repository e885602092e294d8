"""CPU (gloo, world_size 2 and 3) tests of the multi-GPU protocol in hmse_b200/sharding.py.  The
local compute is supplied by the oracle, so what is tested here is the exchange logic itself:
shard-edge resync reproduces the single-stream cut list, and the digest-prefix all-to-all
reproduces the global first-occurrence dedup."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_bytes, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from oracle import corpus
        from oracle.cdc import chunk_shard
        from hmse_b200 import sharding
        cfg = oracle.CDCConfig()
        data = corpus.generate(n_bytes)
        per = (n_bytes + world - 1) // world
        lo, hi = rank * per, min(n_bytes, (rank + 1) * per)
        eof = rank == world - 1
        local = data[lo:n_bytes if eof else hi + cfg.max_size]
        n_own = hi - lo
        calls = []

        def resolve(entry):
            calls.append(entry)
            cuts, ex = chunk_shard(local, cfg, entry, n_own, eof)
            return cuts, ex

        dev = torch.device("cpu")
        cuts, entry, rounds = sharding.stitch_cuts(resolve, local.size if eof else n_own, dev)
        # --- global dedup through the all-to-all protocol ---
        dg = oracle.digest(local, cuts, start0=entry)
        n = dg.shape[0]
        counts_all = sharding._all_gather_i64(n, dev)
        id_base = sum(counts_all[:rank])
        owner = dg[:, :4].copy().view(np.uint32).reshape(-1) % world
        order = np.argsort(owner, kind="stable")
        rec = np.zeros((n, 40), dtype=np.uint8)
        rec[:, :32] = dg[order]
        rec[:, 32:] = (np.arange(n, dtype=np.uint64)[order] + np.uint64(id_base)).view(np.uint8).reshape(-1, 8)
        counts = [int((owner == o).sum()) for o in range(world)]

        def owner_resolve(recv, m):
            r = recv.numpy().reshape(m, 40)
            gids = r[:, 32:].copy().view(np.uint64).reshape(-1)
            best = {}
            for i in range(m):
                k = r[i, :32].tobytes()
                best[k] = min(best.get(k, 1 << 63), int(gids[i]))
            return torch.tensor([best[r[i, :32].tobytes()] for i in range(m)], dtype=torch.int64)

        reply = sharding.exchange_dedup(torch.from_numpy(rec.reshape(-1)), counts, owner_resolve)
        canon = np.empty(n, dtype=np.int64)
        canon[order] = reply.numpy()
        q.put((rank, (cuts + np.uint64(lo)).tolist(), canon.tolist(), id_base, rounds, calls))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_stitch_and_dedup_over_gloo(world):
    import oracle
    from oracle import corpus
    n_bytes = 3 << 20
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_bytes, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    data = corpus.generate(n_bytes)
    want = oracle.chunk(data)
    cuts = np.array(sum((g[1] for g in got), []), dtype=np.uint64)
    assert np.array_equal(cuts, want)
    wc, _ = oracle.dedup(oracle.digest(data, want))
    canon = np.array(sum((g[2] for g in got), []), dtype=np.int64)
    assert np.array_equal(canon, wc)
    assert [g[3] for g in got] == np.concatenate([[0], np.cumsum([len(g[1]) for g in got])[:-1]]).tolist()
    # rank 0 never re-resolves; later ranks re-resolve at most a few times
    assert got[0][5] == [0] and all(len(g[5]) <= 3 for g in got)


def _lsh_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import importlib; om = importlib.import_module("oracle.minhash")
        from hmse_b200 import sharding
        rng = np.random.default_rng(77)
        n_total, bands = 1000, 32
        keys = rng.integers(0, 50, (n_total, bands)).astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)  # many collisions
        bounds = np.linspace(0, n_total, world + 1).astype(int)
        if world == 3:
            bounds[1] = bounds[0]     # a rank without chunks
        lo, hi = bounds[rank], bounds[rank + 1]
        local = torch.from_numpy(keys[lo:hi].view(np.int64).copy())
        owned, mine, id_base = sharding.exchange_lsh(local)
        assert id_base == lo and mine == list(range(rank, bands, world))
        assert np.array_equal(owned.numpy().view(np.uint64), keys[:, rank::world])
        b, k, i = om.buckets(owned.numpy().view(np.uint64))
        q.put((rank, (b.astype(np.int64) * world + rank).tolist(), k.tolist(), i.tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_lsh_band_exchange_over_gloo(world):
    import importlib; om = importlib.import_module("oracle.minhash")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_lsh_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(77)
    keys = rng.integers(0, 50, (1000, 32)).astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    wb, wk, wi = om.buckets(keys)
    triples = sorted(zip(sum((g[1] for g in got), []), sum((g[2] for g in got), []), sum((g[3] for g in got), [])))
    assert triples == list(zip(wb.tolist(), wk.tolist(), wi.tolist()))


def _votes_np(heads, ubase, min_votes, root_all):
    """The votes rule of oracle/deltacode.py::delta_bases on given heads (ids in the global first-occurrence space):
    pass 0 (root_all None) -> root flags of these chunks; pass 1 -> base id or -1."""
    h = heads.numpy().astype(np.int64) & 0xFFFFFFFF
    m = h.shape[0]
    if root_all is None:
        out = np.zeros(m, dtype=np.uint8)
        for i in range(m):
            js, cnt = np.unique(h[i][h[i] < ubase + i], return_counts=True)
            out[i] = cnt.size == 0 or int(cnt.max()) < min_votes
        return torch.from_numpy(out)
    root = root_all.numpy().astype(bool)
    base = np.full(m, -1, dtype=np.int64)
    for i in range(m):
        js, cnt = np.unique(h[i][h[i] < ubase + i], return_counts=True)
        ok = root[js] & (cnt >= min_votes)
        if ok.any():
            js, cnt = js[ok], cnt[ok]
            base[i] = int(js[np.argmax(cnt)])
    return torch.from_numpy(base)


def _l4_worker(rank, world, port, n_bytes, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from oracle import corpus
        from hmse_b200 import sharding
        data = corpus.generate(n_bytes)
        cuts = oracle.chunk_c(data)
        n_all = cuts.size
        starts = np.concatenate([[0], cuts[:-1]]).astype(np.int64)
        _, first = oracle.dedup(oracle.digest(data, cuts))
        keys_all = oracle.band_keys(oracle.minhash_c(data, cuts))
        bounds = np.linspace(0, n_all, world + 1).astype(int)       # the shard of a rank = a contiguous range of chunks
        if world == 3:
            bounds[2] = bounds[1]                                    # a rank without chunks
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        sel = np.flatnonzero(first[lo:hi])                           # local indices of the globally-first chunks
        keys = torch.from_numpy(keys_all[lo:hi][sel].view(np.int64).copy()).view(-1, 32)

        class Ops:
            @staticmethod
            def heads(owned):
                return torch.from_numpy(oracle.lsh_heads(owned.numpy().view(np.uint64)).astype(np.int32))

            @staticmethod
            def votes(heads, ubase, mv, root_all):
                return _votes_np(heads, ubase, mv, root_all)

            @staticmethod
            def chunk_bytes(want_j):
                js = (want_j.numpy() + lo).tolist()
                parts = [data[starts[j]:int(cuts[j])] for j in js]
                lens = torch.tensor([p.size for p in parts], dtype=torch.int64)
                blob = np.concatenate(parts) if parts else np.zeros(0, dtype=np.uint8)
                return torch.from_numpy(blob.copy()), lens

        g = sharding.global_delta_bases(sharding.TorchFabric(), Ops, keys, torch.from_numpy(sel.astype(np.int64)), hi - lo, lo, 32, 4)
        bl = g["base_loc"].numpy()
        n = hi - lo
        ext_gid = g["ext_gid"].numpy()
        gid = np.where(bl < 0, -1, np.where(bl < n, bl + lo, ext_gid[np.clip(bl - n, 0, max(0, ext_gid.size - 1))] if ext_gid.size else -1))
        # the fetched bytes are the base chunks' bytes
        eo = g["ext_off"].numpy()
        ed = g["ext_data"].numpy()
        for e, j in enumerate(ext_gid.tolist()):
            assert ed[eo[e]:eo[e + 1]].tobytes() == data[starts[j]:int(cuts[j])].tobytes()
        q.put((rank, gid.tolist(), int(ext_gid.size)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_global_l4_bases_over_gloo(world):
    """sharding.global_delta_bases (the orchestration the GPU path runs over the C fabric) with torch.distributed / gloo
    and the oracle as local compute: bases chosen through the owner-side bucket heads, the gathered root flags and the
    remote-base fetch equal oracle.delta_bases over the whole stream, including bases that live on another rank."""
    import oracle
    from oracle import corpus
    n_bytes = 5 << 19
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_l4_worker, args=(r, world, port, n_bytes, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    data = corpus.generate(n_bytes)
    cuts = oracle.chunk_c(data)
    _, first = oracle.dedup(oracle.digest(data, cuts))
    keys_all = oracle.band_keys(oracle.minhash_c(data, cuts))
    want = oracle.delta_bases(keys_all, first)
    base = np.array(sum((g[1] for g in got), []), dtype=np.int64)
    assert np.array_equal(base, want)
    assert (want >= 0).sum() >= 5 and sum(g[2] for g in got) >= 1      # some bases crossed a rank boundary
