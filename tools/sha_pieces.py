"""SHA-256 on stream pieces: hmse_digest over 1 / 2 / 10 GB of chunks for several resident-block counts per SM
(HMSE_SHA_BLOCKS_PER_SM overrides the library's choice; one process per setting because the library reads it once).
Usage: python tools/sha_pieces.py            -> runs itself once per setting, prints one JSON line each"""
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
    out = {"blocks_per_sm": os.environ.get("HMSE_SHA_BLOCKS_PER_SM", "auto")}
    gen = pc.DeviceCorpus(ctx)
    for gb in (1.0, 2.0, 10.0):
        n = int(gb * (1 << 30)) if gb < 10 else 10_000_000_000
        d = gen.generate(n)
        cuts = ctx.chunk(d, cfg)
        best = 1e9
        for _ in range(4):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            ctx.digest(d, cuts)
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        out["%g" % gb] = {"ms": round(best, 3), "GBps": round(n / best / 1e6, 1), "chunks": int(cuts.numel())}
        del d, cuts
    print(json.dumps(out))


if __name__ == "__main__":
    if os.environ.get("SHA_PIECES_CHILD"):
        child()
    else:
        for v in ("auto", "8", "6", "4", "3", "2"):
            env = dict(os.environ, SHA_PIECES_CHILD="1")
            if v != "auto":
                env["HMSE_SHA_BLOCKS_PER_SM"] = v
            subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, check=False)
