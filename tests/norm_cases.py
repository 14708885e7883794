"""BC1 blocks for the normalization tests: the reference's own vectors (experimental/normalize_blocks/normalize.rs tests)
and generators that make normalizable blocks common (solid colours through every index value, equal endpoints,
punch-through transparency, near-miss interpolations)."""
from __future__ import annotations

import numpy as np

# (name, block bytes, {mode: expected bytes}) — normalize.rs:507-700
RED = bytes([0x00, 0xF8])
REFERENCE_VECTORS = [
    ("solid_red", RED + bytes([0x01, 0x01, 0, 0, 0, 0]),
     {1: RED + bytes(6), 2: RED + RED + bytes(4)}),
    ("transparent", bytes([0x00, 0x80, 0x00, 0xF8, 0xFF, 0xFF, 0xFF, 0xFF]), {1: b"\xFF" * 8, 2: b"\xFF" * 8}),
    ("mixed_red_blue", RED + bytes([0x1F, 0x00, 0x11, 0x11, 0x11, 0x11]), {1: None, 2: None}),          # preserved
    ("non_roundtrippable", RED + bytes([0x1F, 0x00, 0xAA, 0xAA, 0xAA, 0xAA]), {1: None, 2: None}),      # (170,0,85) != (173,0,82)
    ("solid_red_c1_zero", RED + bytes(6), {1: RED + bytes(6), 2: RED + RED + bytes(4)}),
]


def crafted_blocks(n: int, seed: int) -> np.ndarray:
    """n BC1 blocks, about half of them normalizable in some way."""
    rng = np.random.default_rng(seed)
    c0 = rng.integers(0, 1 << 16, n, dtype=np.uint32)
    c1 = rng.integers(0, 1 << 16, n, dtype=np.uint32)
    kind = rng.integers(0, 8, n)
    c1 = np.where(kind == 0, c0, c1)                                   # equal endpoints (3-colour mode)
    c1 = np.where(kind == 1, (c0 + rng.integers(-2, 3, n)) & 0xFFFF, c1)  # near-equal endpoints
    c1 = np.where(kind == 2, 0, c1)
    c0 = np.where(kind == 3, np.minimum(c0, c1), c0)                    # force punch-through mode sometimes
    # index patterns: one repeated value, two values, or random
    rep = rng.integers(0, 4, n).astype(np.uint32) * 0x55555555
    two = rep ^ (rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32) & 0x55555555 & rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32))
    rnd = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    pat = rng.integers(0, 4, n)
    idx = np.where(pat <= 1, rep, np.where(pat == 2, two, rnd)).astype(np.uint32)
    out = np.empty((n, 8), np.uint8)
    out[:, 0:2] = c0.astype("<u2").view(np.uint8).reshape(n, 2)
    out[:, 2:4] = c1.astype("<u2").view(np.uint8).reshape(n, 2)
    out[:, 4:8] = idx.astype("<u4").view(np.uint8).reshape(n, 4)
    return out.reshape(-1)
