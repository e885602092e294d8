"""Ratio / work study of match-search variants on the CPU model (tests/model).  TEST INFRASTRUCTURE.
Usage: python tools/deflate_study.py [MiB]"""
import ctypes as C
import sys
import os
import zlib

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cdc, corpus  # noqa: E402
from oracle.config import CDCConfig  # noqa: E402
from tests.model.build import Params, build  # noqa: E402


def main():
    mib = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    lib = build()
    data = corpus.generate(mib << 20)
    zd = corpus.zdict()
    cuts = cdc.chunk_c(data, CDCConfig()) if hasattr(cdc, "chunk_c") else cdc.chunk(data, CDCConfig())
    starts = np.concatenate([[0], cuts[:-1]]).astype(np.int64)
    seen = set()
    sel = []
    for s, e in zip(starts.tolist(), cuts.tolist()):
        b = bytes(data[s:e])
        if b not in seen:
            seen.add(b)
            sel.append((s, e))
    zdn = np.frombuffer(zd, dtype=np.uint8)
    ztot = 0
    for s, e in sel:
        co = zlib.compressobj(6, zlib.DEFLATED, 15, 8, zlib.Z_DEFAULT_STRATEGY, zd)
        ztot += len(co.compress(bytes(data[s:e])) + co.flush())
    raw = sum(e - s for s, e in sel)
    print("chunks", len(sel), "raw", raw, "zlib6", ztot, "ratio", raw / ztot)
    out = np.zeros(70000, dtype=np.uint8)
    variants = [
        ("runs exact 4/4", dict(mode=1)),
        ("runs uncond 4/4", dict(mode=2)),
        ("runs sat 4/4", dict(mode=3)),
        ("runs uncond 8/4", dict(mode=2, chain_own=8)),
        ("runs uncond 4/8", dict(mode=2, chain_dict=8)),
        ("runs uncond 6/6", dict(mode=2, chain_own=6, chain_dict=6)),
        ("runs sat 6/6", dict(mode=3, chain_own=6, chain_dict=6)),
    ]
    for name, kw in variants:
        pr = Params(hash_bytes=4, chain_own=4, chain_dict=4, lazy=1, too_far=0, dict_hash_bits=15, mode=0, min_len=0)
        for k, v in kw.items():
            setattr(pr, k, v)
        tot = 0
        st = (C.c_uint32 * 8)()
        acc = np.zeros(8, dtype=np.int64)
        for s, e in sel:
            ch = np.ascontiguousarray(data[s:e])
            for i in range(8):
                st[i] = 0
            r = lib.model_compress(ch.ctypes.data, e - s, zdn.ctypes.data, len(zd), C.byref(pr), out.ctypes.data, out.size, st)
            assert r > 0, r
            if len(sel) < 400 or (s % 7 == 0):
                do = zlib.decompressobj(15, zd)
                assert do.decompress(bytes(out[:r])) == bytes(ch)
            tot += r
            acc += np.array(list(st), dtype=np.int64)
        print("%-24s size %10d  vs zlib6 %.4f  tokens/B %.3f pairs/B %.3f heads/B %.3f steps_all/B %.3f steps_head/B %.3f" % (
            name, tot, tot / ztot, acc[0] / raw, acc[4] / raw, acc[5] / raw, acc[6] / raw, acc[7] / raw))


if __name__ == "__main__":
    main()
