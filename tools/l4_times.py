"""Per-stage device times of the L4 layer (MinHash of first occurrences, LSH keys + buckets, base selection, delta
encode, delta apply) on the device-generated corpus, best of N.  Usage: python tools/l4_times.py [GB] [reps]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hmse_b200  # noqa: E402
from hmse_b200 import corpus as pc  # noqa: E402


def main():
    gb = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    ctx = hmse_b200.default_context(0)
    dev = ctx.tdev
    n = int(gb * 1e9) & ~15
    d = pc.DeviceCorpus(ctx).generate(n)
    cuts = ctx.chunk(d, hmse_b200.CDCConfig())
    canon, first = ctx.dedup(ctx.digest(d, cuts))
    ing = hmse_b200.Ingest(ctx)
    sel = ing.select_first(first.view(torch.uint8))
    sim = hmse_b200.SimConfig()
    starts = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), cuts[:-1]])
    lens = cuts - starts
    names = ["minhash", "keys_buckets", "bases", "delta_encode", "delta_apply"]
    best = {k: 1e30 for k in names}
    info = {}
    for _ in range(reps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(10)]
        torch.cuda.synchronize()
        ev[0].record()
        sig = ctx.minhash(d, cuts, sim, select=sel)
        ev[1].record()
        keys = ctx.lsh_keys(sig, sim)
        band, key, ids = ctx.lsh_buckets(keys)
        ev[2].record()
        ones = torch.ones(sel.numel(), dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        ev[3].record()
        base_u = ctx.delta_bases(band, key, ids, sel.numel(), sim.bands, ones, 4)
        ev[4].record()
        base = torch.full((cuts.numel(),), -1, dtype=torch.int64, device=dev)
        base[sel] = torch.where(base_u >= 0, sel[base_u.clamp(min=0)], base_u)
        n_cand = int((base >= 0).sum())
        torch.cuda.synchronize()
        ev[5].record()
        dblob, doffs = ctx.delta_encode(d, cuts, base)
        ev[6].record()
        kept = torch.nonzero(base >= 0).view(-1)
        bj = base[kept]
        out_off = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), torch.cumsum(lens[kept], 0)])
        doff_k = torch.cat([doffs[kept], doffs[-1:]])
        bo, bl = starts[bj].contiguous(), lens[bj].to(torch.int32).contiguous()
        torch.cuda.synchronize()
        ev[7].record()
        out, status, bad = ctx.delta_apply(dblob, doff_k, d, bo, bl, out_off)
        ev[8].record()
        torch.cuda.synchronize()
        for k, (a, b) in zip(names, [(0, 1), (1, 2), (3, 4), (5, 6), (7, 8)]):
            best[k] = min(best[k], ev[a].elapsed_time(ev[b]))
        info = {"chunks": int(cuts.numel()), "unique": int(sel.numel()), "unique_bytes": int(lens[sel].sum()), "candidates": n_cand,
                "kept": int(kept.numel()), "kept_raw_bytes": int(lens[kept].sum()), "delta_bytes": int(dblob.numel()), "bad": bad}
    info["ms"] = best
    info["minhash_GB/s"] = info["unique_bytes"] / best["minhash"] / 1e6
    info["delta_encode_GB/s_of_candidates"] = info["kept_raw_bytes"] / best["delta_encode"] / 1e6
    print(json.dumps(info, indent=1))


if __name__ == "__main__":
    main()
