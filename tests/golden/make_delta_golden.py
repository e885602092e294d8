"""Regenerates tests/golden/delta_golden.json from the CPU oracle (run from the repo root:
`python tests/golden/make_delta_golden.py`).  The reference names xdelta3 / bsdiff for L4 (README.md:2162, 1402)
without vendoring or pinning either, so - parity unpinned - these are the outputs of oracle/deltacode.py on the first
3 MiB of the seed-42 corpus plus synthetic pairs, committed so that oracle and CUDA path cannot drift together."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from oracle import corpus  # noqa: E402
from oracle import deltacode as D  # noqa: E402


def pairs():
    """Deterministic (base, target) pairs: edits of corpus text."""
    text = corpus.generate(64 << 10).tobytes()
    out = []
    for k, (off, n) in enumerate([(0, 4096), (5000, 9000), (20000, 32768), (100, 700), (40000, 2048)]):
        b = text[off:off + n]
        t = bytearray(b)
        t[n // 3:n // 3] = b"[[inserted %d]]" % k
        del t[n // 2:n // 2 + 7 * k]
        t[-1] ^= 1
        out.append((b, bytes(t)))
    out.append((text[:3000], text[3000:6000]))      # unrelated: no delta
    return out


def main():
    d = corpus.generate(3 << 20)
    cuts = oracle.chunk_c(d)
    _, first = oracle.dedup(oracle.digest(d, cuts))
    keys = oracle.band_keys(oracle.minhash_c(d, cuts))
    base, blob, offs = oracle.delta(d, cuts, keys, first)
    kept = np.flatnonzero(base >= 0)
    out = {"input_sha256": hashlib.sha256(d.tobytes()).hexdigest(), "n": int(d.size), "chunks": int(cuts.size),
           "min_votes": D.MIN_VOTES, "base_sha256": hashlib.sha256(base.astype("<i8").tobytes()).hexdigest(),
           "kept": [[int(i), int(base[i])] for i in kept],
           "delta_bytes": int(blob.size), "blob_sha256": hashlib.sha256(blob.tobytes()).hexdigest(),
           "offsets_sha256": hashlib.sha256(offs.astype("<u8").tobytes()).hexdigest(),
           "first_deltas_hex": [blob[int(offs[i]):int(offs[i + 1])].tobytes().hex() for i in kept[:3]],
           "pairs": []}
    for b, t in pairs():
        dl = D.delta_encode(t, b)
        out["pairs"].append({"base_sha256": hashlib.sha256(b).hexdigest(), "target_sha256": hashlib.sha256(t).hexdigest(),
                             "delta_hex": None if dl is None else dl.hex()})
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "delta_golden.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", len(kept), "kept deltas,", blob.size, "bytes;", len(out["pairs"]), "pairs")


if __name__ == "__main__":
    main()
