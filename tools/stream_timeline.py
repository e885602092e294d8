"""Timeline of IngestStream on a pinned host buffer: when the copies in finish, per-piece compute spans.
Usage: python tools/stream_timeline.py [GB] [piece MiB]"""
import sys
import os
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hmse_b200  # noqa: E402
from hmse_b200 import corpus as pcorpus  # noqa: E402


def main():
    gb = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
    piece = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    ctx = hmse_b200.Context(0)
    n = int(gb * 1e9) & ~15
    d = pcorpus.DeviceCorpus(ctx).generate(n)
    host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    host.copy_(d)
    zd = ctx.stage(pcorpus.zdict())
    # plain copies
    for _ in range(2):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); d.copy_(host, non_blocking=True); b.record(); torch.cuda.synchronize()
        print("H2D %.1f GB/s" % (n / a.elapsed_time(b) / 1e6))
    one = hmse_b200.Ingest(ctx, hmse_b200.CDCConfig(), zd)
    for _ in range(2):
        torch.cuda.synchronize(); t = time.perf_counter(); r = one.run(d); torch.cuda.synchronize()
        print("one-shot device %.1f ms" % ((time.perf_counter() - t) * 1e3))
    hb = torch.empty(r.blob.numel(), dtype=torch.uint8, pin_memory=True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); hb.copy_(r.blob, non_blocking=True); b.record(); torch.cuda.synchronize()
    print("D2H %.1f GB/s (%d MB)" % (hb.numel() / a.elapsed_time(b) / 1e6, hb.numel() >> 20))
    del r, d
    st = hmse_b200.IngestStream(ctx, hmse_b200.CDCConfig(), zd, piece_bytes=piece << 20, overlap=len(sys.argv) > 3)
    st.trace = []
    for _ in range(3):
        st.trace = []
        torch.cuda.synchronize(); t = time.perf_counter(); h = st.run(host); torch.cuda.synchronize()
        print("stream total %.1f ms" % ((time.perf_counter() - t) * 1e3))
    for name, t0 in st.trace:
        print("  %-22s %.1f ms" % (name, t0))
    # kernel spans (library events) summed over pieces
    import ctypes as C
    lib = ctx.lib
    lib.hmse_timing(ctx.h, 1)
    names = ["scan", "resolve", "sha256", "dedup", "deflate", "pack"]
    acc = {k: 0.0 for k in names}
    orig_mark = None
    st.trace = []
    st.timing_hook = lambda: [acc.__setitem__(k, acc[k] + _ms(lib, ctx, i)) for i, k in enumerate(names)]
    torch.cuda.synchronize(); t = time.perf_counter(); h = st.run(host); torch.cuda.synchronize()
    print("stream total (timing on) %.1f ms; kernel spans: %s ; sum %.1f" % ((time.perf_counter() - t) * 1e3,
          {k: round(v, 2) for k, v in acc.items()}, sum(acc.values())))


def _ms(lib, ctx, i):
    import ctypes as C
    f = C.c_float(0)
    return f.value if lib.hmse_timing_ms(ctx.h, i, C.byref(f)) == 0 else 0.0


if __name__ == "__main__":
    main()
