"""Regenerates tests/golden/hotpath_golden.json from the CPU oracle (run from the repo root:
`python tests/golden/make_golden.py`).  The reference repo holds no vectors for this path
(SURVEY.md §8c), so these are the oracle's own outputs on a small fixed input, pinned so that
neither the oracle nor the CUDA path can drift silently.  The input is the first 192 KiB of the
seed-42 procedural corpus plus two adversarial buffers; inputs are identified by SHA-256."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from oracle import corpus  # noqa: E402


def case(name, data, cfg):
    cuts = oracle.chunk_naive(data, cfg)
    assert np.array_equal(cuts, oracle.chunk(data, cfg)) and np.array_equal(cuts, oracle.chunk_c(data, cfg))
    dg = oracle.digest(data, cuts)
    canon, first = oracle.dedup(dg)
    sig = oracle.minhash(data, cuts[:3])
    keys = oracle.band_keys(sig)
    return {"name": name, "input_sha256": hashlib.sha256(data.tobytes()).hexdigest(), "n": int(data.size),
            "cfg": [cfg.min_size, cfg.avg_size, cfg.max_size, cfg.mask_s, cfg.mask_l, cfg.gear_seed],
            "cuts": [int(c) for c in cuts], "digests_sha256": hashlib.sha256(dg.tobytes()).hexdigest(),
            "digest0": dg[0].tobytes().hex() if dg.shape[0] else "", "canon": [int(c) for c in canon],
            "sig0": [int(x) for x in sig[0]] if sig.shape[0] else [], "keys0": [int(x) for x in keys[0]] if keys.shape[0] else []}


def inputs():
    text = corpus.generate(192 << 10)
    dup = np.concatenate([text[:60000], text[:60000], text[1000:50000]])
    rnd = corpus.random_bytes(100000)
    return {"text": text, "dup": dup, "random": rnd}


def main():
    ins = inputs()
    out = {"gear_sha256": hashlib.sha256(oracle.gear_table().tobytes()).hexdigest(),
           "zdict_sha256": hashlib.sha256(corpus.zdict()).hexdigest(), "cases": []}
    out["cases"].append(case("text", ins["text"], oracle.CDCConfig()))
    out["cases"].append(case("text_4k", ins["text"], oracle.CDCConfig.for_avg(4096)))
    out["cases"].append(case("dup", ins["dup"], oracle.CDCConfig()))
    out["cases"].append(case("random", ins["random"], oracle.CDCConfig()))
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "hotpath_golden.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
