"""Per-phase cycle shares of parse_kernel (thread 0 clocks, hmse_debug_deflate_prof) for a given FastCDC geometry, so that
one size class can be looked at alone.  Usage: python tools/parse_phases.py [min avg max] [MiB]"""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hmse_b200  # noqa: E402
from hmse_b200 import corpus as pc  # noqa: E402

NAMES = ["P0 stage", "adler", "P1-2 hist+scan", "P3 scatter", "P3b bucket sort", "P4 match", "P5 dp", "P6 hop", "P7 hist+tokens",
         "P4c runs->match"]


def main():
    mn, av, mx = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (2048, 8192, 32768)
    mib = int(sys.argv[4]) if len(sys.argv) > 4 else 256
    ctx = hmse_b200.default_context(0)
    bits = av.bit_length() - 1
    m = hmse_b200.CDCConfig.for_avg(1 << bits)
    cfg = hmse_b200.CDCConfig(mn, av, mx, m.mask_s, m.mask_l)
    d = pc.DeviceCorpus(ctx).generate(mib << 20)
    zd = ctx.stage(pc.zdict())
    cuts = ctx.chunk(d, cfg)
    prof = (C.c_uint64 * 16)()
    ctx.compress(d, cuts, None, zd)
    ctx.lib.hmse_debug_deflate_prof(prof, 1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    blob, offs = ctx.compress(d, cuts, None, zd)
    b.record()
    torch.cuda.synchronize()
    ctx.lib.hmse_debug_deflate_prof(prof, 1)
    tot = sum(prof[i] for i in range(10))
    print(json.dumps({"cdc": [mn, av, mx], "chunks": int(cuts.numel()), "mean_len": (mib << 20) / max(1, cuts.numel()),
                      "deflate_ms": a.elapsed_time(b), "GB/s": (mib << 20) / a.elapsed_time(b) / 1e6,
                      "cycles_per_byte": tot / (mib << 20), "ratio": (mib << 20) / max(1, blob.numel()),
                      "phase_pct": {NAMES[i]: round(100.0 * prof[i] / max(1, tot), 1) for i in range(10)},
                      "phase_cycles_per_byte": {NAMES[i]: round(prof[i] / (mib << 20), 2) for i in range(10)}}, indent=1))


if __name__ == "__main__":
    main()
