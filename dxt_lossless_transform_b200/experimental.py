"""Mirror of dxt_lossless_transform_bc1::experimental (block normalization) over the additive C entry points
``dltcuda_bc1_*`` (core/dxt-lossless-transform-bc1/src/experimental/{mod.rs,normalize_blocks/normalize.rs,
normalize_blocks/transform.rs}).  Same names and argument meaning as the Rust functions; the work runs on the GPU."""
from __future__ import annotations

import ctypes as C
import enum
from dataclasses import dataclass
from typing import Iterator

import numpy as np

from . import _native as N
from .api import (Bc1TransformSettings, InvalidLength, OutputBufferTooSmall, YCoCgVariant, _check_device, _ro, _rw)


class ColorNormalizationMode(enum.IntEnum):
    """normalize.rs:487-500, `all_values()` order."""

    NONE = 0
    Color0Only = 1
    ReplicateColor = 2


@dataclass(frozen=True)
class Bc1TransformDetailsWithNormalization:
    """experimental/mod.rs:18-35; default (None, Variant1, split) :47-55."""

    color_normalization_mode: ColorNormalizationMode = ColorNormalizationMode.NONE
    decorrelation_mode: YCoCgVariant = YCoCgVariant.Variant1
    split_colour_endpoints: bool = True

    @classmethod
    def all_combinations(cls) -> Iterator["Bc1TransformDetailsWithNormalization"]:
        for n in ColorNormalizationMode:
            for v in YCoCgVariant:
                for s in (True, False):
                    yield cls(n, v, s)

    def untransform_settings(self) -> Bc1TransformSettings:
        """`impl From<Bc1TransformDetailsWithNormalization> for Bc1UntransformSettings`."""
        return Bc1TransformSettings(self.decorrelation_mode, self.split_colour_endpoints)


def _bufs(input, output):
    ip, il, ka = _ro(input)
    op, ol, kb = _rw(output)
    if il % 8:
        raise InvalidLength(il)
    if ol < il:
        raise OutputBufferTooSmall(il, ol)
    return ip, il, op, (ka, kb)


def normalize_blocks(input, output, color_mode: ColorNormalizationMode) -> None:
    """normalize.rs:38 — `output` may be the same buffer as `input`."""
    ip, il, op, _k = _bufs(input, output)
    _check_device(N.lib().dltcuda_bc1_normalize_blocks(ip, op, il, int(color_mode)))


def normalize_blocks_all_modes(input, outputs) -> bool:
    """normalize.rs:417 — outputs[m] for m in ColorNormalizationMode order; returns `any_normalized`."""
    ip, il, _ka = _ro(input)
    if il % 8:
        raise InvalidLength(il)
    ptrs, keep = [], []
    for o in outputs:
        op, ol, k = _rw(o)
        if ol < il:
            raise OutputBufferTooSmall(il, ol)
        ptrs.append(op)
        keep.append(k)
    any_ = C.c_bool(False)
    _check_device(N.lib().dltcuda_bc1_normalize_blocks_all_modes(ip, ptrs[0], ptrs[1], ptrs[2], il, C.byref(any_)))
    return bool(any_.value)


def normalize_split_blocks_in_place(colors, indices, num_blocks: int, color_mode: ColorNormalizationMode) -> None:
    """normalize.rs:286"""
    cp, cl, _a = _rw(colors)
    xp, xl, _b = _rw(indices)
    if cl < 4 * num_blocks or xl < 4 * num_blocks:
        raise OutputBufferTooSmall(4 * num_blocks, min(cl, xl))
    _check_device(N.lib().dltcuda_bc1_normalize_split_blocks_in_place(cp, xp, num_blocks, int(color_mode)))


def transform_bc1_with_normalize_blocks(input, output, details: Bc1TransformDetailsWithNormalization) -> None:
    """transform.rs:65 (the reference's work buffer is not needed: normalization is fused into the transform kernel)."""
    ip, il, op, _k = _bufs(input, output)
    _check_device(N.lib().dltcuda_bc1_transform_with_normalize_blocks(
        ip, op, il, int(details.color_normalization_mode), int(details.decorrelation_mode), bool(details.split_colour_endpoints)))


def transform_bc1_auto_with_normalization(input, output, use_all_decorrelation_modes: bool = False, return_estimates: bool = False):
    """transform.rs:222 with the LTU-semantics estimator, on the GPU."""
    ip, il, op, _k = _bufs(input, output)
    norm, mode, split = C.c_int(), C.c_uint8(), C.c_bool()
    sizes = (C.c_size_t * 24)()
    _check_device(N.lib().dltcuda_bc1_transform_auto_with_normalization(
        ip, op, il, bool(use_all_decorrelation_modes), C.byref(norm), C.byref(mode), C.byref(split), sizes))
    best = Bc1TransformDetailsWithNormalization(ColorNormalizationMode(norm.value), YCoCgVariant(mode.value), bool(split.value))
    return (best, list(sizes)) if return_estimates else best


def normalize_blocks_device(d_in: int, d_out: int, nbytes: int, color_mode: ColorNormalizationMode, stream: int = 0) -> None:
    _check_device(N.lib().dltcuda_bc1_normalize_blocks_device(d_in, d_out, nbytes, int(color_mode), stream))


def transform_bc1_with_normalize_blocks_device(d_in: int, d_out: int, nbytes: int, details: Bc1TransformDetailsWithNormalization,
                                               stream: int = 0) -> None:
    _check_device(N.lib().dltcuda_bc1_transform_with_normalize_blocks_device(
        d_in, d_out, nbytes, int(details.color_normalization_mode), int(details.decorrelation_mode),
        bool(details.split_colour_endpoints), stream))
