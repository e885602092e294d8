"""DEFLATE throughput on inputs that stress the bucket sort of parse_kernel (every position of a tile in ONE bucket):
constant bytes, a short period, a few distinct words - next to the text corpus.  Streams are inflated with stock zlib.
Usage: python tools/pathological.py [MiB]"""
import json
import os
import sys
import zlib

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hmse_b200  # noqa: E402
from oracle import corpus  # noqa: E402


def main():
    mib = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    n = mib << 20
    ctx = hmse_b200.default_context(0)
    zd = corpus.zdict()
    rng = np.random.default_rng(7)
    words = [bytes(rng.integers(97, 123, size=int(k), dtype=np.uint8)) for k in rng.integers(3, 9, size=8)]
    few = b" ".join(words[i] for i in rng.integers(0, 8, size=n // 4))[:n]
    cases = {
        "zeros": np.zeros(n, dtype=np.uint8),
        "period7": np.frombuffer((b"abcdefg" * (n // 7 + 1))[:n], dtype=np.uint8),
        "period300": np.frombuffer((bytes(rng.integers(97, 123, size=300, dtype=np.uint8)) * (n // 300 + 1))[:n], dtype=np.uint8),
        "eight_words": np.frombuffer(few, dtype=np.uint8),
        "text": corpus.generate(n),
    }
    out = {}
    for name, data in cases.items():
        for size in (8192, 32768):
            cuts = np.arange(size, n + 1, size, dtype=np.uint64)
            d = ctx.stage(data)
            dc = ctx.stage_u64(cuts)
            dz = ctx.stage(zd)
            ctx.compress(d, dc, None, dz)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            blob, offs = ctx.compress(d, dc, None, dz)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b)
            hb, ho = blob.cpu().numpy(), offs.cpu().numpy().view(np.uint64)
            ok = True
            for k in range(0, cuts.size, max(1, cuts.size // 64)):
                do = zlib.decompressobj(zdict=zd)
                raw = do.decompress(hb[int(ho[k]):int(ho[k + 1])].tobytes())
                ok &= raw == data[k * size:(k + 1) * size].tobytes()
            out["%s/%d" % (name, size)] = {"ms": round(ms, 2), "GB/s": round(n / ms / 1e6, 2), "ratio": round(n / max(1, blob.numel()), 1),
                                           "inflates": bool(ok)}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
