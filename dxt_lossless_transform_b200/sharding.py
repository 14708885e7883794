"""Block-range sharding of one payload across GPUs.

Blocks are independent, so a payload of N blocks is cut into contiguous block ranges, one per rank.
Rank g reads input bytes [B*b_g, B*b_{g+1}) and owns, in every output stream s (element width w_s,
base N*prefix_s in the reference layout), the slice [N*prefix_s + w_s*b_g, N*prefix_s + w_s*b_{g+1}).
The only shared quantities are N and the prefix b_g, both computed on the host — there is no
collective on the data path.
"""
from __future__ import annotations

from . import _native as N


def stream_widths(fmt: int, settings) -> list[int]:
    """Per-block element width of each stream, in output order (csrc/bcn_layout.h)."""
    sc = bool(settings.split_colour_endpoints)
    colour = [2, 2] if sc else [4]
    if fmt == 1:
        return colour + [4]
    if fmt == 2:
        return [8] + colour + [4]
    sa = bool(settings.split_alpha_endpoints)
    return ([1, 1] if sa else [2]) + [6] + colour + [4]


def stream_slices(fmt: int, settings, total_blocks: int, first_block: int, num_blocks: int) -> list[tuple[int, int]]:
    """(byte offset, byte length) of the slice of every stream that blocks [first, first+num) own."""
    out, prefix = [], 0
    for w in stream_widths(fmt, settings):
        out.append((total_blocks * prefix + w * first_block, w * num_blocks))
        prefix += w
    return out


def shard_ranges(fmt: int, total_blocks: int, num_shards: int) -> list[tuple[int, int]]:
    """(first_block, num_blocks) per shard; boundaries are multiples of the kernel tile."""
    f = N.lib().dltcuda_shard_first_block
    bounds = [f(fmt, total_blocks, g, num_shards) for g in range(num_shards + 1)]
    return [(bounds[g], bounds[g + 1] - bounds[g]) for g in range(num_shards)]
