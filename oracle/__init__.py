"""CPU oracle for the HMSE data-reduction hot path.  TEST INFRASTRUCTURE ONLY.

This package is the parity reference and the timed CPU baseline for
`hmse_b200`.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import it; the product package
`hmse_b200` never does (it fails loudly when its CUDA library is missing).

The upstream repo (1Jamie/HMSE) ships a design spec and no implementation of
this path (README.md:18-28, 293), so the oracle is a restatement *of the spec*
with the contradictions resolved as SURVEY.md §0.2 records:

* chunk()      FastCDC (Xia et al., ATC'16, cited README.md:2753-2755) with the
               spec's min/avg/max parameters (README.md:289, 2444-2446).
* digest()     SHA-256 of the raw chunk bytes (README.md:290, 2543) - hashlib.
* dedup()      first-instance-wins exact dedup (README.md:1288-1292, 1542-1551).
* compress()   per-chunk zlib stream with a preset dictionary, level 6
               (README.md:288, 2374; RFC 1950/1951) - zlib.
* similarity() 128 x MurmurHash3_x86_32 over 4-byte shingles, running minimum
               from 0xFFFFFFFF (README.md:2578-2597), seeds 1..128
               (VALIDATION_METHODS.md:122), banding (README.md:2231-2235).
* delta()      L4 delta coding against LSH-selected bases with the 20 % rule
               (README.md:1328, 2160-2198); byte format defined in oracle/deltacode.py.

Pinning status: the reference holds no golden vectors for any
of these.  SHA-256, MurmurHash3 and zlib are pinned against their public
known-answer vectors (tests/test_oracle_kat.py).  FastCDC cut points are
"parity unpinned": no reference implementation or vector exists; the oracle's
own byte-at-a-time loop (`chunk_naive`, a literal transcription of the paper's
Algorithm 1) is the ground truth and its outputs are committed under
tests/golden/.
"""
from .config import CDCConfig, SimConfig, gear_table, PAPER_MASK_S, PAPER_MASK_L  # noqa: F401
from .cdc import chunk, chunk_naive, chunk_c, next_cut  # noqa: F401
from .sha import digest, dedup, digest_mt, dedup_fast  # noqa: F401
from .deflate import compress, inflate_all, make_zdict  # noqa: F401
from .minhash import (murmur3_32, minhash, minhash_c, band_keys, buckets,  # noqa: F401
                      similarity)
from .deltacode import delta_bases, delta_encode, delta_encode_naive, delta_apply, delta, lsh_heads  # noqa: F401
from . import archive, corpus  # noqa: F401
