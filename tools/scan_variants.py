"""Gear scan variants (csrc/cdc.cu, HMSE_SCAN_VARIANT): device time of hmse_chunk_scan over 10 GB, and a checksum of the
candidate bitmaps (must be the same for every variant).  One child process per variant (the library reads the variable once).
Usage: python tools/scan_variants.py"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import torch
    import hmse_b200
    from hmse_b200 import corpus as pc
    ctx = hmse_b200.default_context(0)
    cfg = hmse_b200.CDCConfig()
    n = 10_000_000_000
    d = pc.DeviceCorpus(ctx).generate(n)
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
    print(json.dumps({"variant": os.environ.get("HMSE_SCAN_VARIANT", "default"), "scan_ms": round(best, 3),
                      "GBps": round(n / best / 1e6, 1), "frac_of_hbm_6558": round(n / best / 1e6 / 6558.1, 3),
                      "bitmap_checksum": cs & 0xFFFFFFFFFFFF, "chunks": int(cuts.numel())}))


if __name__ == "__main__":
    if os.environ.get("SCAN_VARIANTS_CHILD"):
        child()
    else:
        for v in ("0", "1", "2"):
            subprocess.run([sys.executable, os.path.abspath(__file__)], env=dict(os.environ, SCAN_VARIANTS_CHILD="1", HMSE_SCAN_VARIANT=v),
                           check=False)
