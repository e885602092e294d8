"""GPU parity, last file of the suite: every gear-scan variant kept in csrc/cdc.cu (HMSE_SCAN_VARIANT) against the default
one, which tests/test_gpu_cdc.py compares with the oracle (oracle/cdc.py)."""
import pytest

pytestmark = pytest.mark.gpu


def test_scan_variants_agree(ctx, monkeypatch):
    """K1: the default scan (tensor-map TMA, runs that continue across tiles) and every measured alternative kept in the
    source (HMSE_SCAN_VARIANT 0-4, read by the library at every call) give the same bitmaps and the same cut lists on
    ragged sizes around every tile / region / run-length boundary - text up to 130 MiB (two tiles per thread and trip
    start at 116 MB), random bytes up to 8 MiB.  The default itself is compared with the oracle above."""
    import torch
    import hmse_b200
    from hmse_b200 import corpus as pc
    cfg = hmse_b200.CDCConfig()
    K, M = 1 << 10, 1 << 20
    text = pc.DeviceCorpus(ctx).generate(130 * M + 11 + 4096)
    g = torch.Generator(device="cuda")
    g.manual_seed(7)
    rnd = torch.randint(0, 256, (8 * M + 13 + 4096,), dtype=torch.uint8, device="cuda", generator=g)
    small = [1, 15, 16, 17, 63, 64, 65, 127, 128, 129, 4095, 4096, 4097, 32 * K - 1, 32 * K, 32 * K + 1, 96 * K + 5, M - 1, M + 1,
             8 * M + 13]
    big = [58 * M - 3, 59 * M + 77, 130 * M + 11]

    def run(buf, n, variant):
        if variant is None:
            monkeypatch.delenv("HMSE_SCAN_VARIANT", raising=False)
        else:
            monkeypatch.setenv("HMSE_SCAN_VARIANT", variant)
        d = buf[:n]
        ctx.chunk_scan(d, cfg)
        bs, bl = ctx.chunk_candidates(n)
        cuts, _ = ctx.chunk_resolve(d, cfg, n, True, 0)
        return bs, bl, cuts

    for name, buf, sizes in (("text", text, small + big), ("random", rnd, small)):
        for n in sizes:
            want = run(buf, n, None)
            for v in ("0", "1", "2", "3", "4"):
                got = run(buf, n, v)
                assert torch.equal(want[0], got[0]) and torch.equal(want[1], got[1]), "bitmaps of variant %s differ (%s, %d bytes)" % (v, name, n)
                assert torch.equal(want[2], got[2]), "cuts of variant %s differ (%s, %d bytes)" % (v, name, n)
    monkeypatch.delenv("HMSE_SCAN_VARIANT", raising=False)
    torch.cuda.synchronize()
