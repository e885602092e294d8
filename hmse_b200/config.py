"""Frozen configuration of the hot path (product side; the oracle keeps its own copy and
tests assert the two agree).  Spec: README.md:289, 2444-2446 (chunk sizes), FastCDC paper cited
at README.md:2753-2755 (Gear hash, normalised-chunking masks), VALIDATION_METHODS.md:122
(MinHash seeds 1..128), BASELINE.json config 5 (32 bands)."""
from __future__ import annotations

from dataclasses import dataclass, field
from functools import lru_cache

import numpy as np

M64 = (1 << 64) - 1
PAPER_MASK_S = 0x0003590703530000
PAPER_MASK_L = 0x0000D90003530000
GEAR_SEED_DEFAULT = 0x484D5345  # "HMSE"


@lru_cache(maxsize=8)
def _gear_tuple(seed: int) -> tuple:
    out, state = [], seed & M64
    for _ in range(256):
        state = (state + 0x9E3779B97F4A7C15) & M64
        z = state
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
        out.append(z ^ (z >> 31))
    return tuple(out)


def gear_table(seed: int = GEAR_SEED_DEFAULT) -> np.ndarray:
    """256 u64: the splitmix64 stream started at `seed`."""
    return np.array(_gear_tuple(seed), dtype=np.uint64)


def spread_mask(nbits: int) -> int:
    """`nbits` one-bits spread over bit positions 16..47 (bit k of a Gear hash depends on the
    last k+1 bytes only, so mask bits sit high)."""
    if not 1 <= nbits <= 32:
        raise ValueError("nbits must be in 1..32")
    m = 0
    for j in range(nbits):
        m |= 1 << (16 + (j * 32) // nbits)
    return m


@dataclass(frozen=True)
class CDCConfig:
    min_size: int = 2048
    avg_size: int = 8192
    max_size: int = 32768
    mask_s: int = PAPER_MASK_S
    mask_l: int = PAPER_MASK_L
    gear_seed: int = GEAR_SEED_DEFAULT

    def __post_init__(self):
        if not (64 <= self.min_size <= self.avg_size <= self.max_size):
            raise ValueError("need 64 <= min <= avg <= max")
        if self.max_size > (1 << 20):
            raise ValueError("max_size must be <= 1 MiB")
        if self.mask_s == 0 or self.mask_l == 0:
            raise ValueError("masks must be non-zero")

    @staticmethod
    def for_avg(avg: int, nc_level: int = 2, gear_seed: int = GEAR_SEED_DEFAULT) -> "CDCConfig":
        bits = avg.bit_length() - 1
        if avg != 1 << bits:
            raise ValueError("avg must be a power of two")
        if avg == 8192 and nc_level == 2:
            ms, ml = PAPER_MASK_S, PAPER_MASK_L
        else:
            ms, ml = spread_mask(bits + nc_level), spread_mask(bits - nc_level)
        return CDCConfig(avg // 4, avg, avg * 4, ms, ml, gear_seed)

    @property
    def gear(self) -> np.ndarray:
        return gear_table(self.gear_seed)


@dataclass(frozen=True)
class SimConfig:
    n_perm: int = 128
    bands: int = 32
    rows: int = 4
    seeds: tuple = field(default_factory=lambda: tuple(range(1, 129)))

    def __post_init__(self):
        if self.bands * self.rows != self.n_perm:
            raise ValueError("bands * rows must equal n_perm")
        if len(self.seeds) != self.n_perm:
            raise ValueError("need n_perm seeds")
        if self.n_perm % 32 or not 32 <= self.n_perm <= 256:
            raise ValueError("n_perm must be a multiple of 32 in 32..256")

    @property
    def seed_array(self) -> np.ndarray:
        return np.array(self.seeds, dtype=np.uint32)
