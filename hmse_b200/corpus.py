"""Procedural templated-wiki corpus, device generator (bench/test INPUT, not part of the
reference path).  The lexicon and the hash-defined article structure are restated here for the
product side; oracle/corpus.py is the independent NumPy twin the tests compare against.
Spec: README.md:1176-1178 (templated infobox/cite text), 2123-2127 (redundancy mix),
VALIDATION_METHODS.md:119-120 (seed 42)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from functools import lru_cache

import numpy as np
import torch

from ._lib import CorpusCfg

M32 = 0xFFFFFFFF
INFOBOX = [
    b"{{Infobox settlement\n| name = ", b"\n| native_name = ", b"\n| settlement_type = ", b"\n| image_skyline = ",
    b"\n| subdivision_type = [[Country]]\n| subdivision_name = ", b"\n| established_title = Founded\n| established_date = ",
    b"\n| population_total = ", b"\n| population_as_of = ", b"\n| area_total_km2 = ",
    b"\n| timezone = [[UTC+1]]\n| coordinates = {{coord|", b"\n| website = {{URL|http://www.", b"}}\n}}\n\n'''",
]
BODY = [
    b"<ref>{{cite web |url=http://www.example.org/", b" |title=", b" |publisher=", b" |accessdate=2025-10-20}}</ref> ",
    b"\n\n== History ==\n", b"\n\n== Geography ==\n", b"\n\n== Demographics ==\n", b"\n\n== References ==\n{{reflist}}\n",
    b"\n[[Category:", b"]] ", b"{{cite journal |last=", b"\n* [[",
]
PUNCT = [b". ", b", ", b".\n\n", b"; "]
TERM = b"\n\n\n"
N_WORDS = 4096
ID_WORD0 = len(INFOBOX) + len(BODY) + len(PUNCT) + 1
K_WORD = 0x00000077


def _mix32(x: int) -> int:
    x &= M32
    x ^= x >> 16
    x = (x * 0x85EBCA6B) & M32
    x ^= x >> 13
    x = (x * 0xC2B2AE35) & M32
    x ^= x >> 16
    return x


def _H(seed: int, a: int, b: int) -> int:
    return _mix32(_mix32((seed ^ (a * 0x9E3779B1)) & M32) + ((b * 0x85EBCA77) & M32))


@lru_cache(maxsize=1)
def lexicon():
    """(blob uint8[], off uint32[L+1]): boilerplate, punctuation, terminator, 4096 words."""
    entries = list(INFOBOX) + list(BODY) + list(PUNCT) + [TERM]
    for i in range(N_WORDS):
        ln = 2 + _H(K_WORD, i, 0) % (3 + min(i >> 6, 7))
        entries.append(bytes(97 + _H(K_WORD, i, 1 + j) % 26 for j in range(ln)) + b" ")
    off = np.zeros(len(entries) + 1, dtype=np.uint32)
    off[1:] = np.cumsum([len(e) for e in entries])
    return np.frombuffer(b"".join(entries), dtype=np.uint8).copy(), off


def zdict(size: int = 32768) -> bytes:
    """Preset dictionary from the lexicon: rare words first, common words and the template
    boilerplate last (closest to the data, shortest distances)."""
    blob, off = lexicon()
    b = blob.tobytes()
    ent = [b[off[i]:off[i + 1]] for i in range(off.size - 1)]
    words, boiler, seen = [], [], set()
    for f in (ent[ID_WORD0 + i] for i in range(N_WORDS - 1, -1, -1)):
        if f and f not in seen:
            seen.add(f)
            words.append(f)
    for f in ent[:ID_WORD0]:
        if f and f not in seen:
            seen.add(f)
            boiler.append(f)
    return b"".join(words + boiler[::-1])[-size:]


@dataclass(frozen=True)
class CorpusConfig:
    """dup/near thresholds are per-1024 article-class probabilities (default ~18 % exact-dup,
    ~35 % near-dup articles; high_redundancy() is BASELINE config 4's >= 60 % exact)."""
    seed: int = 42
    dup_thr: int = 184
    near_thr: int = 358
    pick_tries: int = 4     # attempts to find an earlier article of the unique class to copy from

    @staticmethod
    def high_redundancy(seed: int = 42) -> "CorpusConfig":
        """83 % of the articles are copies of articles that occurred earlier: >= 60 % of the chunks are exact duplicates."""
        return CorpusConfig(seed, 850, 70, 32)


class DeviceCorpus:
    """Renders any byte window of the stream article(0) ++ article(1) ++ ... in HBM."""

    MEAN_ARTICLE = 40000  # bytes, used only to guess how many article lengths to compute

    def __init__(self, ctx, cfg: CorpusConfig = CorpusConfig()):
        self.ctx, self.cfg = ctx, cfg
        blob, off = lexicon()
        self.lex_blob = torch.from_numpy(blob).to(ctx.tdev)
        self.lex_off = torch.from_numpy(off.view(np.int32).copy()).to(ctx.tdev)
        self.c = CorpusCfg(cfg.seed, cfg.dup_thr, cfg.near_thr, off.size - 1, cfg.pick_tries)
        self._offs = None  # int64 stream offsets of articles 0.._n_art (device), grown on demand

    def _ensure(self, end_byte: int):
        n_art = 0 if self._offs is None else self._offs.numel() - 1
        total = 0 if self._offs is None else int(self._offs[-1])
        while total < end_byte:
            want = max(64, int((end_byte - total) / self.MEAN_ARTICLE * 1.2) + 64)
            lens = self.ctx.empty(want, torch.int32)
            self.ctx.check(self.ctx.lib.hmse_corpus_lengths(self.ctx.h, C.byref(self.c), self.lex_off.data_ptr(), n_art,
                                                            want, lens.data_ptr(), self.ctx.stream))
            cs = torch.cumsum(lens.to(torch.int64), 0) + total
            head = torch.zeros(1, dtype=torch.int64, device=self.ctx.tdev) if self._offs is None else self._offs
            self._offs = torch.cat([head, cs])
            n_art += want
            total = int(self._offs[-1])

    def generate(self, n_bytes: int, byte_off: int = 0, pad: int = 64) -> torch.Tensor:
        """uint8 device tensor with stream bytes [byte_off, byte_off + n_bytes) (+ `pad` slack)."""
        self._ensure(byte_off + n_bytes)
        offs = self._offs
        lo = int(torch.searchsorted(offs, torch.tensor([byte_off], device=offs.device), right=True)[0]) - 1
        hi = int(torch.searchsorted(offs, torch.tensor([byte_off + n_bytes], device=offs.device), right=False)[0])
        hi = min(max(hi, lo + 1), offs.numel() - 1)
        win = offs[lo:hi + 1].contiguous()
        out = self.ctx.empty(n_bytes + pad, torch.uint8)
        self.ctx.check(self.ctx.lib.hmse_corpus_render(self.ctx.h, C.byref(self.c), self.lex_blob.data_ptr(),
                                                       self.lex_off.data_ptr(), lo, hi - lo, win.data_ptr(), byte_off,
                                                       n_bytes, out.data_ptr(), self.ctx.stream))
        return out[:n_bytes]
