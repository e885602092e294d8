"""Deterministic synthetic templated-wiki corpus (NumPy twin).  TEST INFRASTRUCTURE.

The spec benchmarks on Wikipedia-like text: infobox / cite / category templates
with varying fields (README.md:1176-1178) and a redundancy mix of exact
duplicates, near duplicates and unique articles (README.md:2123-2127); seed 42
(VALIDATION_METHODS.md:119-120).  No corpus ships with the repo and there is no
network, so the corpus is *procedural*: every byte is a pure function of
(seed, article index, token index) built from 32-bit integer hashes only, which
lets `hmse_b200`'s device generator (csrc/corpus.cu) reproduce it bit for bit at
10-100 GB while this twin regenerates any prefix on the CPU.

Stream  = article(0) ++ article(1) ++ ...           truncated to n bytes
article(a) = render(content_id c, edit seed e)      (c, e) = article_meta(a)
render     = concatenation of lexicon entries, one per token
"""
from __future__ import annotations

from dataclasses import dataclass
from functools import lru_cache

import numpy as np

M32 = 0xFFFFFFFF

INFOBOX = [
    b"{{Infobox settlement\n| name = ",
    b"\n| native_name = ",
    b"\n| settlement_type = ",
    b"\n| image_skyline = ",
    b"\n| subdivision_type = [[Country]]\n| subdivision_name = ",
    b"\n| established_title = Founded\n| established_date = ",
    b"\n| population_total = ",
    b"\n| population_as_of = ",
    b"\n| area_total_km2 = ",
    b"\n| timezone = [[UTC+1]]\n| coordinates = {{coord|",
    b"\n| website = {{URL|http://www.",
    b"}}\n}}\n\n'''",
]
BODY = [
    b"<ref>{{cite web |url=http://www.example.org/",
    b" |title=",
    b" |publisher=",
    b" |accessdate=2025-10-20}}</ref> ",
    b"\n\n== History ==\n",
    b"\n\n== Geography ==\n",
    b"\n\n== Demographics ==\n",
    b"\n\n== References ==\n{{reflist}}\n",
    b"\n[[Category:",
    b"]] ",
    b"{{cite journal |last=",
    b"\n* [[",
]
PUNCT = [b". ", b", ", b".\n\n", b"; "]
TERM = b"\n\n\n"
N_WORDS = 4096

NI, NBODY, NPUNCT = len(INFOBOX), len(BODY), len(PUNCT)
ID_BODY0 = NI
ID_PUNCT0 = NI + NBODY
ID_TERM = ID_PUNCT0 + NPUNCT
ID_WORD0 = ID_TERM + 1

K_TOKEN, K_CLASS, K_PICK, K_EDIT, K_NTOK, K_EPOS, K_EWORD, K_WORD = (
    0x5BD1E995, 0x000A11CE, 0x00000D0B, 0x0000ED17, 0x0000070C, 0x00001234, 0x00004321, 0x00000077)


def mix32(x: int) -> int:
    x &= M32
    x ^= x >> 16
    x = (x * 0x85EBCA6B) & M32
    x ^= x >> 13
    x = (x * 0xC2B2AE35) & M32
    x ^= x >> 16
    return x


def H(seed: int, a: int, b: int) -> int:
    return mix32(mix32((seed ^ (a * 0x9E3779B1)) & M32) + ((b * 0x85EBCA77) & M32))


def _mix32_v(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint32, copy=True)
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x85EBCA6B)
    x ^= x >> np.uint32(13)
    x *= np.uint32(0xC2B2AE35)
    x ^= x >> np.uint32(16)
    return x


def H_v(seed: int, a: int, b: np.ndarray) -> np.ndarray:
    base = np.uint32(mix32((seed ^ (a * 0x9E3779B1)) & M32))
    return _mix32_v(base + b.astype(np.uint32) * np.uint32(0x85EBCA77))


def zipf_word(v):
    """Skewed word rank in [0, 4096) from 26 hash bits (integer-only 'Zipf')."""
    a = v & 0xFF
    b = (v >> 8) & 0xFF
    c = (v >> 16) & 0x3FF
    return (((a * b) >> 8) * c) >> 6


@dataclass(frozen=True)
class CorpusConfig:
    """dup/near thresholds are per-1024 article-class probabilities.
    Default mix ~18 % exact-dup, ~35 % near-dup, ~47 % unique articles
    (README.md:2123-2127); `high_redundancy()` is config 4's >= 60 % exact."""
    seed: int = 42
    dup_thr: int = 184
    near_thr: int = 358
    pick_tries: int = 4     # attempts to find an earlier article of the unique class to copy from

    @staticmethod
    def high_redundancy(seed: int = 42) -> "CorpusConfig":
        """83 % of the articles are copies, and with 32 tries nearly all of them copy an article that really occurred
        earlier (its content is on the stream already): >= 60 % of the CHUNKS are exact duplicates (measured; chunks that
        straddle an article boundary are not)."""
        return CorpusConfig(seed, 850, 70, 32)


@lru_cache(maxsize=1)
def lexicon():
    """(blob uint8[], off uint32[L+1]): boilerplate, punctuation, terminator, words."""
    entries = list(INFOBOX) + list(BODY) + list(PUNCT) + [TERM]
    for i in range(N_WORDS):
        ln = 2 + H(K_WORD, i, 0) % (3 + min(i >> 6, 7))
        entries.append(bytes(97 + H(K_WORD, i, 1 + j) % 26 for j in range(ln)) + b" ")
    off = np.zeros(len(entries) + 1, dtype=np.uint32)
    off[1:] = np.cumsum([len(e) for e in entries])
    return np.frombuffer(b"".join(entries), dtype=np.uint8).copy(), off


def article_meta(cfg: CorpusConfig, a: int):
    """(content id, edit seed) of article a."""
    if a == 0:
        return 0, 0
    cls = H(cfg.seed, a, K_CLASS) & 1023
    if cls >= cfg.dup_thr + cfg.near_thr:
        return a, 0
    c = 0
    for k in range(cfg.pick_tries):  # prefer a content id whose own article is of the unique class
        c = (H(cfg.seed, a, K_PICK + k) * a) >> 32
        if c == 0 or (H(cfg.seed, c, K_CLASS) & 1023) >= cfg.dup_thr + cfg.near_thr:
            break
    if cls < cfg.dup_thr:
        return c, 0
    return c, H(cfg.seed, a, K_EDIT) | 1


def n_tokens(cfg: CorpusConfig, c: int) -> int:
    return 800 + (H(cfg.seed, c, K_NTOK) & 16383)


def token_ids(cfg: CorpusConfig, c: int, e: int) -> np.ndarray:
    nt = n_tokens(cfg, c)
    t = np.arange(nt, dtype=np.uint32)
    u = H_v(cfg.seed ^ K_TOKEN, c, t)
    v = (u >> np.uint32(6)).astype(np.int64)
    sel = (u & np.uint32(63)).astype(np.int64)
    word = ID_WORD0 + zipf_word(v)
    ids = np.where(sel == 0, ID_BODY0 + (v % NBODY), np.where(sel <= 6, ID_PUNCT0 + (v & 3), word))
    head = t < 2 * NI
    ids = np.where(head, np.where((t & 1) == 0, (t >> 1).astype(np.int64), word), ids)
    if e:
        for k in range(1 + (e & 3)):
            tp = H(e, k, K_EPOS) % nt
            ids[tp] = ID_WORD0 + zipf_word(H(e, k, K_EWORD) >> 6)
    ids[nt - 1] = ID_TERM
    return ids


def render(cfg: CorpusConfig, c: int, e: int) -> np.ndarray:
    blob, off = lexicon()
    ids = token_ids(cfg, c, e)
    lens = (off[ids + 1] - off[ids]).astype(np.int64)
    total = int(lens.sum())
    starts = np.cumsum(lens) - lens
    # gather: out[starts[t] + j] = blob[off[ids[t]] + j]
    src = np.repeat(off[ids].astype(np.int64) - starts, lens) + np.arange(total, dtype=np.int64)
    return blob[src]


def article(cfg: CorpusConfig, a: int) -> np.ndarray:
    return render(cfg, *article_meta(cfg, a))


def generate(n_bytes: int, cfg: CorpusConfig = CorpusConfig(), first_article: int = 0) -> np.ndarray:
    """First n_bytes of the stream that starts at article `first_article`."""
    parts, total, a = [], 0, first_article
    while total < n_bytes:
        p = article(cfg, a)
        parts.append(p)
        total += p.size
        a += 1
    return np.concatenate(parts)[:n_bytes] if parts else np.zeros(0, dtype=np.uint8)


def zdict(size: int = 32768) -> bytes:
    """Preset dictionary from the generator's own lexicon: rare words first,
    common words and the template boilerplate last (closest to the data)."""
    from .deflate import make_zdict
    blob, off = lexicon()
    b = blob.tobytes()
    ent = [b[off[i]:off[i + 1]] for i in range(off.size - 1)]
    samples = [(ent[ID_WORD0 + i], N_WORDS - i) for i in range(N_WORDS - 1, -1, -1)]
    samples += [(x, 1 << 20) for x in ent[:ID_WORD0]]
    return make_zdict(samples, size)


def random_bytes(n: int, seed: int = 0xDEADBEEF) -> np.ndarray:
    """Incompressible control (VALIDATION_METHODS.md:213; stress seed :121)."""
    return np.random.default_rng(seed).integers(0, 256, n, dtype=np.uint8)
