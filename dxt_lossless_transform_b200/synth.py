"""Synthetic BCn payloads: counter-based, so any block range of any payload can be
generated independently of the rest (shards never need the whole payload).

* ``random_blocks``  — distribution U: uniform random bytes.
* ``texture_blocks`` — distribution T: blocks on a 2-D grid whose endpoints follow smooth
  low-frequency colour fields plus a little noise (c1 a slightly darker c0), random indices.
"""
from __future__ import annotations

import numpy as np

BASE_SEED = 0xD17BC100


def splitmix64(seed: int, start: int, count: int) -> np.ndarray:
    """SplitMix64 outputs number start .. start+count-1 of the stream `seed` (vectorised)."""
    with np.errstate(over="ignore"):
        idx = np.arange(start + 1, start + count + 1, dtype=np.uint64)
        z = np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def block_bytes(fmt: int) -> int:
    return 8 if fmt == 1 else 16


def random_blocks(fmt: int, num_blocks: int, seed: int = BASE_SEED, first_block: int = 0) -> np.ndarray:
    words_per_block = block_bytes(fmt) // 8
    w = splitmix64(seed, first_block * words_per_block, num_blocks * words_per_block)
    return w.view(np.uint8).copy()


def _rgb565(r, g, b):
    return ((r.astype(np.uint32) >> 3) << 11) | ((g.astype(np.uint32) >> 2) << 5) | (b.astype(np.uint32) >> 3)


def texture_blocks(fmt: int, num_blocks: int, seed: int = BASE_SEED, first_block: int = 0, grid_w: int = 1024,
                   smooth: float = 1.0) -> np.ndarray:
    i = np.arange(first_block, first_block + num_blocks, dtype=np.int64)
    x = (i % grid_w).astype(np.float64)
    y = (i // grid_w).astype(np.float64)
    rnd = splitmix64(seed ^ 0x7E57, first_block * 2, num_blocks * 2).reshape(num_blocks, 2)
    ph = (splitmix64(seed ^ 0xC0105, 0, 9).astype(np.float64) / 2.0**64) * 2 * np.pi
    chans = []
    for c in range(3):
        f = (
            np.sin(x * (0.011 + 0.004 * c) * smooth + ph[3 * c])
            + np.sin(y * (0.017 - 0.003 * c) * smooth + ph[3 * c + 1])
            + np.sin((x + y) * 0.007 * smooth + ph[3 * c + 2])
        )
        v = 127.5 + 40.0 * f + ((rnd[:, 0] >> np.uint64(8 * c)) & np.uint64(3)).astype(np.float64)
        chans.append(np.clip(v, 0, 255))
    delta = 8 + ((rnd[:, 0] >> np.uint64(32)) & np.uint64(15)).astype(np.float64)
    c0 = _rgb565(*[c.astype(np.uint32) for c in chans])
    c1 = _rgb565(*[np.clip(c - delta, 0, 255).astype(np.uint32) for c in chans])
    idx = (rnd[:, 1] & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    out = np.zeros((num_blocks, block_bytes(fmt)), np.uint8)
    col = (c0 | (c1 << 16)).astype(np.uint32)
    off = 0 if fmt == 1 else 8
    out[:, off:off + 4] = col.view(np.uint8).reshape(num_blocks, 4)
    out[:, off + 4:off + 8] = idx.view(np.uint8).reshape(num_blocks, 4)
    if fmt == 2:
        out[:, 0:8] = rnd[:, 1].copy().view(np.uint8).reshape(num_blocks, 8)
    elif fmt == 3:
        a = np.clip(200 + 50 * np.sin(x * 0.013 * smooth + y * 0.009 * smooth), 0, 255)
        out[:, 0] = a.astype(np.uint8)
        out[:, 1] = np.clip(a - 16 - (rnd[:, 0] >> np.uint64(40) & np.uint64(7)).astype(np.float64), 0, 255).astype(np.uint8)
        out[:, 2:8] = (rnd[:, 1] >> np.uint64(16)).copy().view(np.uint8).reshape(num_blocks, 8)[:, :6]
    return out.reshape(-1)
