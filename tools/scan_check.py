"""One experimental gear-scan variant (csrc/cdc.cu, HMSE_SCAN_VARIANT) against the default, in one process: candidate bitmaps
and cut lists must be EQUAL on ragged sizes (text and random bytes, sizes around every tile / region / run-length
boundary), then the device time of both over the whole buffer.  Prints one JSON line; exit code 1 on any difference.
Usage: python tools/scan_check.py VARIANT [GB]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import hmse_b200
    from hmse_b200 import corpus as pc
    v = sys.argv[1]
    gb = float(sys.argv[2]) if len(sys.argv) > 2 else 4.0
    ctx = hmse_b200.default_context(0)
    cfg = hmse_b200.CDCConfig()
    n_max = int(gb * 1e9)
    text = pc.DeviceCorpus(ctx).generate(n_max + 4096)
    g = torch.Generator(device="cuda")
    g.manual_seed(7)
    rnd = torch.randint(0, 256, ((96 << 20) + 4096,), dtype=torch.uint8, device="cuda", generator=g)
    K, M = 1 << 10, 1 << 20
    small = [1, 15, 16, 17, 63, 64, 65, 127, 128, 129, 191, 4095, 4096, 4097, 32 * K - 1, 32 * K, 32 * K + 1, 64 * K - 1, 96 * K + 5,
             M - 1, M, M + 1, 8 * M + 13, 32 * M, 58 * M - 3, 59 * M + 77, 64 * M + 4097, 96 * M - 17]
    big = [n for n in (130 * M + 11, 250 * M + 9, 500 * M + 7, 10 ** 9 + 3, 2 * 10 ** 9 + 1, n_max - 12345, n_max) if n <= n_max]
    bad = []

    def run(buf, n, variant):
        os.environ["HMSE_SCAN_VARIANT"] = variant       # (putenv: the library's getenv sees it at every call)
        d = buf[:n]
        ctx.chunk_scan(d, cfg)
        bs, bl = ctx.chunk_candidates(n)
        cuts, _ = ctx.chunk_resolve(d, cfg, n, True, 0)
        return bs, bl, cuts

    for name, buf, sizes in (("text", text, small + big), ("random", rnd, small)):
        for n in sizes:
            a = run(buf, n, "1")
            b = run(buf, n, v)
            if not (torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])):
                bad.append([name, n])

    def best(variant):
        os.environ["HMSE_SCAN_VARIANT"] = variant
        d = text[:n_max]
        t = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            ctx.chunk_scan(d, cfg)
            e1.record()
            torch.cuda.synchronize()
            t = min(t, e0.elapsed_time(e1))
        return t

    t1, tv = best("1"), best(v)
    os.environ.pop("HMSE_SCAN_VARIANT", None)
    print(json.dumps({"variant": v, "equal": not bad, "differences": bad, "sizes_checked": 2 * len(small) + len(big), "bytes": n_max,
                      "default_ms": round(t1, 3), "variant_ms": round(tv, 3), "default_GBps": round(n_max / t1 / 1e6, 1),
                      "variant_GBps": round(n_max / tv / 1e6, 1)}), flush=True)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
