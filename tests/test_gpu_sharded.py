"""ShardedIngest / ShardedSimilarity over NCCL (the exchange steps inside the library, csrc/comm.cu) reproduce the
single-stream cut list, digests, global dedup, LSH buckets and - with the global LSH index - the delta coding of the
oracle run over the whole stream.  world 2 needs two GPUs; world 1 runs the same code on any GPU box.  The protocol
spelled with torch.distributed is covered on CPU by tests/test_sharding_gloo.py."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %r)
import hmse_b200
from hmse_b200 import corpus as pc
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
ctx = hmse_b200.Context(rank)
cfg = hmse_b200.CDCConfig()
total = 24 << 20
per = total // world
gen = pc.DeviceCorpus(ctx)
eof = rank == world - 1
n_avail = (total - rank * per) if eof else per + cfg.max_size
d = gen.generate(n_avail, byte_off=rank * per)
zd = ctx.stage(pc.zdict())
res = hmse_b200.ShardedIngest(ctx, cfg, zd).run(d, per, eof)
assert hmse_b200.ShardedIngest(ctx, cfg, zd).transport == "c"      # the exchange runs inside the library (csrc/comm.cu)
xs = ctx.exchange_stats()
assert world == 1 or xs["bytes_sent"] > 0
# the same protocol spelled with torch.distributed collectives (the form the gloo tests cover) gives the same result
rt = hmse_b200.ShardedIngest(ctx, cfg, zd, transport="torch").run(d, per, eof)
assert rt.entry == res.entry and rt.id_base == res.id_base
assert torch.equal(rt.cuts, res.cuts) and torch.equal(rt.canon, res.canon) and torch.equal(rt.is_first, res.is_first)
assert torch.equal(rt.digests, res.digests) and torch.equal(rt.blob, res.blob)
st_, kt_, (bt_, kkt_, it_) = hmse_b200.ShardedSimilarity(ctx, transport="torch").run(d, res.cuts, start0=res.entry)
sc_, kc_, (bc_, kkc_, ic_) = hmse_b200.ShardedSimilarity(ctx, transport="c").run(d, res.cuts, start0=res.entry)
assert torch.equal(bt_, bc_) and torch.equal(kkt_, kkc_) and torch.equal(it_, ic_)
pipe = hmse_b200.ShardedIngest(ctx, cfg, zd)
hres = pipe.run(d, per, eof, host=pipe.host_buffers(n_avail), groups=3)
assert np.array_equal(hres.cuts.numpy(), res.cuts.cpu().numpy()) and np.array_equal(hres.canon.numpy(), res.canon.cpu().numpy())
assert np.array_equal(hres.digests.numpy(), res.digests.cpu().numpy())
assert np.array_equal(hres.offsets.numpy(), res.offsets.cpu().numpy()) and np.array_equal(hres.blob.numpy(), res.blob.cpu().numpy())
# double-buffered batches from pinned host memory: every batch equals the plain run
hin = torch.empty(n_avail, dtype=torch.uint8, pin_memory=True); hin.copy_(d)
nb = 0
for hb in pipe.run_batches([hin, hin, hin], per, eof, host=pipe.host_buffers(n_avail), groups=2):
    assert np.array_equal(hb.cuts.numpy(), res.cuts.cpu().numpy()) and np.array_equal(hb.canon.numpy(), res.canon.cpu().numpy())
    assert np.array_equal(hb.offsets.numpy(), res.offsets.cpu().numpy()) and np.array_equal(hb.blob.numpy(), res.blob.cpu().numpy())
    assert hb.h2d_bytes == n_avail
    nb += 1
assert nb == 3
# shard-local L4 (torch transport): deltas against bases of the same shard, first-occurrence flags from the global dedup
l4 = hmse_b200.ShardedIngest(ctx, cfg, zd, transport="torch").run(d, per, eof, l4=hmse_b200.SimConfig())
assert np.array_equal(l4.cuts.cpu().numpy(), res.cuts.cpu().numpy()) and np.array_equal(l4.canon.cpu().numpy(), res.canon.cpu().numpy())
# global L4 (C transport): one LSH index over the whole stream, bases may live on the other GPU; `base` = global chunk ids
gp = hmse_b200.ShardedIngest(ctx, cfg, zd)
g4 = gp.run(d, per, eof, l4=hmse_b200.SimConfig())
assert np.array_equal(g4.canon.cpu().numpy(), res.canon.cpu().numpy())
gl = gp.last_l4
# read path of the cross-shard deltas: apply every kept delta to its base (local chunk or fetched bytes), digests must match
kept = torch.nonzero(g4.base >= 0).view(-1)
if kept.numel():
    starts_ = torch.cat([torch.full((1,), int(res.entry), dtype=torch.int64, device=d.device), res.cuts[:-1]])
    lens_ = res.cuts - starts_
    bl_ = gl["base_loc"][kept]
    n_ = res.cuts.numel()
    is_ext = bl_ >= n_
    both = torch.cat([d, gl["ext_data"]])
    e_ = (bl_ - n_).clamp(min=0)
    eo_ = torch.cat([gl["ext_off"], gl["ext_off"][-1:]])       # (torch.where evaluates both sides: keep e_ + 1 in range)
    boff = torch.where(is_ext, d.numel() + eo_[e_], starts_[bl_.clamp(max=n_ - 1)])
    blen = torch.where(is_ext, eo_[e_ + 1] - eo_[e_], lens_[bl_.clamp(max=n_ - 1)]).to(torch.int32)
    out_off = torch.cat([torch.zeros(1, dtype=torch.int64, device=d.device), torch.cumsum(lens_[kept], 0)])
    doff = torch.cat([g4.delta_offsets[kept], g4.delta_offsets[-1:]])
    rebuilt, st_, bad_ = ctx.delta_apply(g4.delta_blob, doff, both, boff.contiguous(), blen.contiguous(), out_off)
    assert bad_ == 0
    assert torch.equal(ctx.digest(rebuilt, out_off[1:].contiguous()), res.digests[kept])
sig, keys, (lb, lk, li) = hmse_b200.ShardedSimilarity(ctx).run(d, res.cuts, start0=res.entry)
torch.cuda.synchronize()
out = dict(rank=rank, cuts=(res.cuts.cpu().numpy().view(np.uint64) + np.uint64(rank * per)).tolist(),
           canon=res.canon.cpu().numpy().tolist(), first=res.is_first.cpu().numpy().astype(int).tolist(),
           digests=res.digests.cpu().numpy().tobytes().hex(), id_base=res.id_base, entry=res.entry,
           blob=res.blob.cpu().numpy().tobytes().hex(), offs=res.offsets.cpu().numpy().tolist(),
           sel=res.select.cpu().numpy().tolist(), l4_base=l4.base.cpu().numpy().tolist(), l4_sel=l4.select.cpu().numpy().tolist(),
           l4_dblob=l4.delta_blob.cpu().numpy().tobytes().hex(), l4_doffs=l4.delta_offsets.cpu().numpy().tolist(),
           l4_nstreams=int(l4.offsets.numel() - 1), g4_base=g4.base.cpu().numpy().tolist(), g4_sel=g4.select.cpu().numpy().tolist(),
           g4_dblob=g4.delta_blob.cpu().numpy().tobytes().hex(), g4_doffs=g4.delta_offsets.cpu().numpy().tolist(),
           g4_remote=int(gl["n_remote"]), keys=keys.cpu().numpy().view(np.uint64).tolist(),
           lsh=[lb.cpu().numpy().tolist(), lk.cpu().numpy().view(np.uint64).tolist(), li.cpu().numpy().tolist()])
json.dump(out, open(os.path.join(%r, "shard_%%d.json" %% rank), "w"))
dist.destroy_process_group()
'''


@pytest.mark.parametrize("world", [1, 2])
def test_sharded_ingest_over_nccl(tmp_path, world):
    """world 2 (two GPUs): the real thing.  world 1 runs the SAME code - communicator, all-to-alls to self, global L4 - on
    a one-GPU box, so the sharded path is exercised wherever the GPU suite runs."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import json
    import numpy as np
    import oracle
    from oracle import corpus
    script = tmp_path / "worker.py"
    script.write_text(WORKER % (ROOT, str(tmp_path)))
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    outs = [json.load(open(tmp_path / ("shard_%d.json" % k))) for k in range(world)]
    total = 24 << 20
    data = corpus.generate(total)
    want_cuts = oracle.chunk_c(data)
    cuts = np.array(sum((o["cuts"] for o in outs), []), dtype=np.uint64)
    assert np.array_equal(cuts, want_cuts)
    dg = np.frombuffer(bytes.fromhex("".join(o["digests"] for o in outs)), dtype=np.uint8).reshape(-1, 32)
    want_dg = oracle.digest(data, want_cuts)
    assert np.array_equal(dg, want_dg)
    wc, wf = oracle.dedup(want_dg)
    assert np.array_equal(np.array(sum((o["canon"] for o in outs), []), dtype=np.int64), wc)
    assert np.array_equal(np.array(sum((o["first"] for o in outs), []), dtype=bool), wf)
    # LSH: band-partitioned buckets over the whole stream equal the oracle's, band by band
    import importlib; om = importlib.import_module("oracle.minhash")
    _, want_keys, (wb, wk, wi) = om.similarity(data, want_cuts, use_c=True)
    keys = np.array(sum((o["keys"] for o in outs), []), dtype=np.uint64)
    assert np.array_equal(keys, want_keys)
    triples = sorted(zip(sum((o["lsh"][0] for o in outs), []), sum((o["lsh"][1] for o in outs), []),
                         sum((o["lsh"][2] for o in outs), [])))
    assert triples == list(zip(wb.tolist(), wk.tolist(), wi.tolist()))
    for r, o in enumerate(outs):
        assert set(o["lsh"][0]) <= set(range(r, 32, world))
    # every rank compressed exactly its globally-first chunks, and they inflate to the raw bytes
    zd = corpus.zdict()
    starts = np.concatenate([[0], want_cuts[:-1]]).astype(np.int64)
    raw = data.tobytes()
    for o in outs:
        blob = np.frombuffer(bytes.fromhex(o["blob"]), dtype=np.uint8)
        offs = np.array(o["offs"], dtype=np.int64).astype(np.uint64)
        streams = oracle.inflate_all(blob, offs, zd)
        assert len(streams) == len(o["sel"])
        for k, j in enumerate(o["sel"]):
            g = o["id_base"] + j
            assert wf[g]
            assert streams[k] == raw[int(starts[g]):int(want_cuts[g])]
    # shard-local L4 equals the oracle's delta() run on each shard's chunks with the global first-occurrence flags
    pos = 0
    n_delta = 0
    for o in outs:
        nloc = len(o["canon"])
        lc = want_cuts[pos:pos + nloc]
        s0 = int(starts[pos])
        shard = data[s0:int(lc[-1])]
        rel = (lc - np.uint64(s0)).astype(np.uint64)
        keys_loc = want_keys[pos:pos + nloc]
        wbase, wblob, woffs = oracle.delta(shard, rel, keys_loc, wf[pos:pos + nloc])
        assert o["l4_base"] == wbase.tolist()
        assert bytes.fromhex(o["l4_dblob"]) == wblob.tobytes() and o["l4_doffs"] == woffs.astype(np.int64).tolist()
        assert o["l4_sel"] == np.flatnonzero(wf[pos:pos + nloc] & (wbase < 0)).tolist() and o["l4_nstreams"] == len(o["l4_sel"])
        n_delta += int((wbase >= 0).sum())
        pos += nloc
    assert n_delta > 20
    # GLOBAL L4 (C transport): equal to oracle.delta over the WHOLE stream - bases cross the shard edge
    gbase, gblob, goffs = oracle.delta(data, want_cuts, want_keys, wf)
    got_base = np.array(sum((o["g4_base"] for o in outs), []), dtype=np.int64)
    assert np.array_equal(got_base, gbase)
    assert b"".join(bytes.fromhex(o["g4_dblob"]) for o in outs) == gblob.tobytes()
    pos, shift, all_offs = 0, 0, [0]
    for o in outs:
        offs_r = np.array(o["g4_doffs"], dtype=np.int64)
        all_offs += (offs_r[1:] + shift).tolist()
        shift += int(offs_r[-1])
        nloc = len(o["canon"])
        assert o["g4_sel"] == np.flatnonzero(wf[pos:pos + nloc] & (gbase[pos:pos + nloc] < 0)).tolist()
        pos += nloc
    assert np.array_equal(np.array(all_offs, dtype=np.uint64), goffs)
    # the input is built so that near duplicates of shard 0's chunks sit in shard 1
    assert int((gbase >= 0).sum()) >= n_delta
    if world > 1:
        n0 = len(outs[0]["canon"])
        cross = int(((gbase[n0:] >= 0) & (gbase[n0:] < n0)).sum())
        assert cross >= 5 and outs[1]["g4_remote"] >= 5
