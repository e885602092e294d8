"""Gear scan variants (csrc/cdc.cu, HMSE_SCAN_VARIANT - read at every call): device time of hmse_chunk_scan over N GB, and
a checksum of the candidate bitmaps plus the number of chunks (must be the same for every variant).
Usage: python tools/scan_variants.py [GB] [variants, e.g. 0,3,4]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

WHAT = {"0": "warm-up per run, per-thread bulk copies, ONE table, two staging buffers, 3 CTAs per SM (round 1)",
        "1": "as 0 with the 16-fold table (PRMT + IMAD addressing), one buffer, 3 CTAs (the fallback)",
        "2": "as 1 with two buffers, 2 CTAs",
        "3": "runs that continue across tiles, one bulk copy per thread",
        "4": "continuing runs, 16-byte cp.async copies (eight lanes per line)",
        "5": "continuing runs, one tensor-map TMA copy per tile and CTA (the default)"}
# (profiles/r02P_scan_variants.txt and r02Q_* were taken with experimental builds whose variants 3-5 were something else:
# one branch per four bytes, and the hash update on the fma pipe - both slower, removed from the source)


def main():
    import torch
    import hmse_b200
    from hmse_b200 import corpus as pc
    gb = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
    variants = sys.argv[2].split(",") if len(sys.argv) > 2 else list(WHAT)
    ctx = hmse_b200.default_context(0)
    cfg = hmse_b200.CDCConfig()
    n = int(gb * 1e9)
    d = pc.DeviceCorpus(ctx).generate(n)
    for v in variants:
        os.environ["HMSE_SCAN_VARIANT"] = v       # (putenv: the library's getenv sees it)
        best = 1e9
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            ctx.chunk_scan(d, cfg)
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        bs, bl = ctx.chunk_candidates(n)
        cs = int(bs.sum()) ^ (int(bl.sum()) << 1)
        cuts, _ = ctx.chunk_resolve(d, cfg, n, True, 0)
        print(json.dumps({"variant": v, "what": WHAT.get(v, "?"), "bytes": n, "scan_ms": round(best, 3), "GBps": round(n / best / 1e6, 1),
                          "frac_of_hbm_6558": round(n / best / 1e6 / 6558.1, 3), "bitmap_checksum": cs & 0xFFFFFFFFFFFF,
                          "chunks": int(cuts.numel())}), flush=True)
    os.environ.pop("HMSE_SCAN_VARIANT", None)


if __name__ == "__main__":
    main()
