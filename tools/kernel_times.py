"""Per-kernel device timings on a host-generated corpus (CUDA events, best of N).
Usage: python tools/kernel_times.py [MiB] [out.json]"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import hmse_b200  # noqa: E402
from oracle import corpus  # noqa: E402  (input generation only)


def timed(fn, reps=5):
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best, out


def main():
    mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    out_path = sys.argv[2] if len(sys.argv) > 2 else None
    ctx = hmse_b200.default_context(0)
    t = time.time()
    base = corpus.generate(min(mib, 64) << 20)
    host = np.tile(base, (mib + 63) // 64)[:mib << 20]
    gen_s = time.time() - t
    d = ctx.stage(host)
    n = d.numel()
    cfg = hmse_b200.CDCConfig()
    res = {"MiB": mib, "gen_s": gen_s, "gpu": torch.cuda.get_device_name(0), "sm": ctx.lib and torch.cuda.get_device_properties(0).multi_processor_count}
    ms, _ = timed(lambda: ctx.chunk_scan(d, cfg))
    res["scan_ms"], res["scan_GBps"] = ms, n / ms / 1e6
    ms, (cuts, _) = timed(lambda: ctx.chunk_resolve(d, cfg, n, True, 0))
    res["resolve_ms"], res["resolve_GBps"], res["chunks"] = ms, n / ms / 1e6, int(cuts.numel())
    res["resolve_rounds"] = ctx.lib.hmse_chunk_last_rounds(ctx.h)
    ms, cuts = timed(lambda: ctx.chunk(d, cfg))
    res["chunk_ms"], res["chunk_GBps"] = ms, n / ms / 1e6
    ms, dg = timed(lambda: ctx.digest(d, cuts))
    res["sha_ms"], res["sha_GBps"] = ms, n / ms / 1e6
    ms, (canon, first) = timed(lambda: ctx.dedup(dg))
    res["dedup_ms"] = ms
    res["unique_frac"] = float(first.float().mean())
    sub = cuts[:min(cuts.numel(), 20000)]
    nb = int(sub[-1])
    ms, sig = timed(lambda: ctx.minhash(d, sub, hmse_b200.SimConfig()), reps=3)
    res["minhash_ms"], res["minhash_GBps"] = ms, nb / ms / 1e6
    try:
        from oracle import corpus as oc
        zd = ctx.stage(oc.zdict())
        sel = torch.nonzero(first).view(-1)
        lens = torch.diff(cuts, prepend=torch.zeros(1, dtype=torch.int64, device=cuts.device))
        ub = int(lens[sel].sum())
        import ctypes as C
        prof = (C.c_uint64 * 16)()
        ctx.compress(d, cuts, sel, zd)
        ctx.lib.hmse_debug_deflate_prof(prof, 1)
        ctx.compress(d, cuts, sel, zd)
        torch.cuda.synchronize()
        ctx.lib.hmse_debug_deflate_prof(prof, 1)
        tot = sum(prof[i] for i in range(12))
        names = ["P0 stage", "adler", "P1-2 hist+scan", "P3 scatter", "P3b bucket sort", "P4 match", "P5 dp", "P6 hop",
                 "P7 hist+tokens", "P4c runs->match", "-", "-"]
        res["parse_phase_pct"] = {names[i]: round(100.0 * prof[i] / max(1, tot), 1) for i in range(10)}
        res["parse_cycles_per_chunk"] = tot / max(1, prof[15])
        ms, (blob, offs) = timed(lambda: ctx.compress(d, cuts, sel, zd), reps=3)
        res["deflate_ms"], res["deflate_GBps"], res["deflate_ratio"] = ms, ub / ms / 1e6, ub / max(1, blob.numel())
        res["unique_bytes"] = ub
    except Exception as e:  # noqa: BLE001
        res["deflate_error"] = str(e)[:200]
    print(json.dumps(res, indent=1))
    if out_path:
        with open(out_path, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
