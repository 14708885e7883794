"""Python mirror of the reference's Rust interface for the BCn transform path, over the C ABI.

Names, argument meaning and error behaviour follow the reference so the parity tests read like the
reference's own tests:

* ``transform_bc1_with_settings`` / ``untransform_bc1_with_settings`` (+ bc2, bc3) —
  core/dxt-lossless-transform-bc1/src/transform/safe/transform_with_settings.rs:88,192
* ``transform_bc1_auto`` (+ bc2, bc3) — core/.../transform/safe/transform_auto.rs:95
* ``Bc1ManualTransformBuilder`` / ``Bc1AutoTransformBuilder`` (+ Bc2) —
  api/dxt-lossless-transform-bc1-api/src/transform/{manual,auto}_transform_builder.rs
* ``LosslessTransformUtilsSizeEstimation`` — extensions/estimators/dxt-lossless-transform-ltu/src/lib.rs:49

Buffers are anything exposing a writable/readable C-contiguous buffer (numpy uint8 arrays, bytearray,
pinned buffers from :func:`alloc_pinned`).  Everything computes on the GPU through
libdxt_lossless_transform_cuda.so; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import enum
from dataclasses import dataclass
from typing import Callable, Iterator, Optional

import numpy as np

from . import _native as N


# --------------------------------------------------------------------------------------------------
# Enums / settings
# --------------------------------------------------------------------------------------------------
class YCoCgVariant(enum.IntEnum):
    """Internal numbering (common/src/color_565/decorrelate.rs:72-84)."""

    NONE = 0
    Variant1 = 1
    Variant2 = 2
    Variant3 = 3

    def to_stable(self) -> int:
        """Stable API numbering (api-common/src/reexports/color_565.rs:65-85)."""
        return 3 if self == YCoCgVariant.NONE else int(self) - 1

    @staticmethod
    def from_stable(v: int) -> "YCoCgVariant":
        return YCoCgVariant.NONE if v == 3 else YCoCgVariant(v + 1)


@dataclass(frozen=True)
class Bc1TransformSettings:
    """bc1/src/transform/settings.rs:16-27; default (Variant1, split) :35-43."""

    decorrelation_mode: YCoCgVariant = YCoCgVariant.Variant1
    split_colour_endpoints: bool = True
    FORMAT = 1

    @classmethod
    def all_combinations(cls) -> Iterator["Bc1TransformSettings"]:
        for v in YCoCgVariant:
            for s in (True, False):
                yield cls(v, s)


@dataclass(frozen=True)
class Bc2TransformSettings(Bc1TransformSettings):
    """bc2/src/transform/settings.rs:16-27."""

    FORMAT = 2


@dataclass(frozen=True)
class Bc3TransformSettings:
    """bc3/src/transform/settings.rs:16-31; default (Variant1, true, true) :39-48."""

    decorrelation_mode: YCoCgVariant = YCoCgVariant.Variant1
    split_alpha_endpoints: bool = True
    split_colour_endpoints: bool = True
    FORMAT = 3

    @classmethod
    def all_combinations(cls) -> Iterator["Bc3TransformSettings"]:
        for v in YCoCgVariant:
            for a in (True, False):
                for c in (True, False):
                    yield cls(v, a, c)


Bc1UntransformSettings = Bc1TransformSettings
Bc2UntransformSettings = Bc2TransformSettings
Bc3UntransformSettings = Bc3TransformSettings

_SETTINGS = {1: Bc1TransformSettings, 2: Bc2TransformSettings, 3: Bc3TransformSettings}


def block_bytes(fmt: int) -> int:
    return 8 if fmt == 1 else 16


# --------------------------------------------------------------------------------------------------
# Errors
# --------------------------------------------------------------------------------------------------
class BcnValidationError(ValueError):
    """Bc{1,2,3}ValidationError (safe/transform_with_settings.rs:18-31)."""


class InvalidLength(BcnValidationError):
    def __init__(self, length: int):
        super().__init__(f"Invalid input length: {length}")
        self.length = length


class OutputBufferTooSmall(BcnValidationError):
    def __init__(self, needed: int, actual: int):
        super().__init__(f"Output buffer too small: needed {needed}, got {actual}")
        self.needed, self.actual = needed, actual


class SizeEstimationError(RuntimeError):
    """DetermineBestTransformError::SizeEstimationError (transform_auto.rs:15-23)."""


class TransformationError(RuntimeError):
    """A CUDA failure inside the library (core code 8 / stable code 3)."""


# core/dxt-lossless-transform-bc1/src/c_api/transform_auto.rs:37-58
CORE_SUCCESS, CORE_NULL_DATA, CORE_NULL_OUTPUT, CORE_NULL_ESTIMATOR, CORE_NULL_SETTINGS = 0, 1, 2, 3, 4
CORE_INVALID_LENGTH, CORE_OUTPUT_TOO_SMALL, CORE_SIZE_ESTIMATION, CORE_TRANSFORMATION = 5, 6, 7, 8


def _raise_core(code: int, in_len: int, out_len: int) -> None:
    if code == CORE_SUCCESS:
        return
    if code == CORE_INVALID_LENGTH:
        raise InvalidLength(in_len)
    if code == CORE_OUTPUT_TOO_SMALL:
        raise OutputBufferTooSmall(in_len, out_len)
    if code == CORE_SIZE_ESTIMATION:
        raise SizeEstimationError("size estimation failed")
    msg = N.lib().dltcuda_last_error().decode()
    raise TransformationError(f"core error code {code}: {msg}")


# --------------------------------------------------------------------------------------------------
# Buffers
# --------------------------------------------------------------------------------------------------
def _ro(buf) -> tuple[int, int, object]:
    a = np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf
    if a.dtype != np.uint8 or not a.flags.c_contiguous:
        raise TypeError("expected a C-contiguous uint8 buffer")
    # a 0-length numpy array still has a non-null data pointer
    return a.ctypes.data, a.size, a


def _rw(buf) -> tuple[int, int, object]:
    a = np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf
    if a.dtype != np.uint8 or not a.flags.c_contiguous or not a.flags.writeable:
        raise TypeError("expected a writable C-contiguous uint8 buffer")
    return a.ctypes.data, a.size, a


class PinnedBuffer:
    """Page-locked host memory (the analogue of the reference's allocate_align_64 for buffers that
    feed the GPU path at full host-link speed).  ``.array`` is a numpy uint8 view."""

    def __init__(self, nbytes: int):
        self._ptr = N.lib().dltcuda_alloc_pinned(nbytes)
        if not self._ptr:
            raise MemoryError(f"cudaMallocHost({nbytes}) failed: {N.lib().dltcuda_last_error().decode()}")
        self.nbytes = nbytes
        self.array = np.ctypeslib.as_array((C.c_uint8 * max(nbytes, 1)).from_address(self._ptr))[:nbytes]

    def free(self) -> None:
        if self._ptr:
            self.array = None
            N.lib().dltcuda_free_pinned(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def alloc_pinned(nbytes: int) -> PinnedBuffer:
    return PinnedBuffer(nbytes)


# --------------------------------------------------------------------------------------------------
# with-settings entry points (host buffers)
# --------------------------------------------------------------------------------------------------
def _core_settings(fmt: int, s):
    if fmt == 3:
        return N.CoreBc3Settings(bool(s.split_alpha_endpoints), bool(s.split_colour_endpoints), int(s.decorrelation_mode))
    return N.CoreSettings(bool(s.split_colour_endpoints), int(s.decorrelation_mode))


def _run_with_settings(fmt: int, inverse: bool, input, output, settings) -> None:
    ip, il, _ka = _ro(input)
    op, ol, _kb = _rw(output)
    fn = getattr(N.lib(), f"dltbc{fmt}core_{'untransform' if inverse else 'transform'}")
    r = fn(ip, il, op, ol, _core_settings(fmt, settings))
    _raise_core(r.error_code, il, ol)


def transform_bc1_with_settings(input, output, settings: Bc1TransformSettings = Bc1TransformSettings()) -> None:
    _run_with_settings(1, False, input, output, settings)


def untransform_bc1_with_settings(input, output, settings: Bc1TransformSettings = Bc1TransformSettings()) -> None:
    _run_with_settings(1, True, input, output, settings)


def transform_bc2_with_settings(input, output, settings: Bc2TransformSettings = Bc2TransformSettings()) -> None:
    _run_with_settings(2, False, input, output, settings)


def untransform_bc2_with_settings(input, output, settings: Bc2TransformSettings = Bc2TransformSettings()) -> None:
    _run_with_settings(2, True, input, output, settings)


def transform_bc3_with_settings(input, output, settings: Bc3TransformSettings = Bc3TransformSettings()) -> None:
    _run_with_settings(3, False, input, output, settings)


def untransform_bc3_with_settings(input, output, settings: Bc3TransformSettings = Bc3TransformSettings()) -> None:
    _run_with_settings(3, True, input, output, settings)


def transform_with_settings(fmt: int, input, output, settings) -> None:
    _run_with_settings(fmt, False, input, output, settings)


def untransform_with_settings(fmt: int, input, output, settings) -> None:
    _run_with_settings(fmt, True, input, output, settings)


# --------------------------------------------------------------------------------------------------
# Size estimators (SizeEstimationOperations, api-common/src/estimate/mod.rs:24-64)
# --------------------------------------------------------------------------------------------------
class SizeEstimator:
    """Holds a DltSizeEstimator the C ABI can call."""

    def c_estimator(self) -> "C.POINTER(N.DltSizeEstimator)":  # pragma: no cover - interface
        raise NotImplementedError


class LosslessTransformUtilsSizeEstimation(SizeEstimator):
    """The LTU estimator; its callbacks run the GPU match estimator (csrc/estimator.cu)."""

    def __init__(self):
        self._e = N.lib().dltltu_new_size_estimator()
        if not self._e:
            raise MemoryError("dltltu_new_size_estimator failed")

    def c_estimator(self):
        return self._e

    def max_compressed_size(self, len_bytes: int) -> int:
        out = C.c_size_t(0)
        rc = self._e.contents.max_compressed_size(self._e.contents.context, len_bytes, C.byref(out))
        if rc:
            raise SizeEstimationError(f"code {rc}")
        return out.value

    def estimate_compressed_size(self, data) -> int:
        p, n, _k = _ro(data)
        out = C.c_size_t(0)
        rc = self._e.contents.estimate_compressed_size(self._e.contents.context, p, n, None, 0, C.byref(out))
        if rc:
            raise SizeEstimationError(f"code {rc}")
        return out.value

    def __del__(self):
        try:
            if self._e:
                N.lib().dltltu_free_size_estimator(self._e)
                self._e = None
        except Exception:
            pass


class ZStandardError(RuntimeError):
    """extensions/compressors/dxt-lossless-transform-zstd/src/lib.rs:23-44."""


class InvalidLevel(ZStandardError):
    def __init__(self, level: int):
        super().__init__(f"Invalid compression level: {level}")
        self.level = level


class ZStandardSizeEstimation(SizeEstimator):
    """``ZStandardSizeEstimation`` (dxt-lossless-transform-zstd/src/lib.rs:54-140): the estimate is the size real zstd
    produces (magicless frame, no content size / checksum / dict id).  zstd is bound from the system's libzstd at run
    time (csrc/zstd_estimator.cu); inside ``transform_bcN_auto`` the candidates are transformed on the GPU and compressed
    concurrently, one host thread per candidate."""

    def __init__(self, compression_level: int):
        if not 1 <= compression_level <= 22:   # lib.rs:62-66
            raise InvalidLevel(compression_level)
        self.compression_level = compression_level
        self._e = N.lib().dltzstd_new_size_estimator(compression_level)
        if not self._e:
            raise ZStandardError("no libzstd could be loaded (set DLTCUDA_LIBZSTD)")

    @classmethod
    def new_fast(cls):
        return cls(1)

    @classmethod
    def new_default(cls):
        return cls(3)

    @classmethod
    def new_best(cls):
        return cls(22)

    @staticmethod
    def library_version() -> int:
        return int(N.lib().dltzstd_version_number())

    def c_estimator(self):
        return self._e

    def max_compressed_size(self, len_bytes: int) -> int:
        out = C.c_size_t(0)
        rc = self._e.contents.max_compressed_size(self._e.contents.context, len_bytes, C.byref(out))
        if rc:
            raise SizeEstimationError(f"code {rc}")
        return out.value

    def estimate_compressed_size(self, data) -> int:
        p, n, _k = _ro(data)
        cap = self.max_compressed_size(n)
        scratch = (C.c_uint8 * max(cap, 1))()
        out = C.c_size_t(0)
        rc = self._e.contents.estimate_compressed_size(self._e.contents.context, p, n, scratch, cap, C.byref(out))
        if rc:
            raise SizeEstimationError(f"code {rc}")
        return out.value

    def __del__(self):
        try:
            if self._e:
                N.lib().dltzstd_free_size_estimator(self._e)
                self._e = None
        except Exception:
            pass


class CallbackSizeEstimator(SizeEstimator):
    """A caller-supplied estimator: ``estimate(data: np.ndarray) -> int`` sees HOST memory, exactly as a
    Rust ``SizeEstimationOperations`` implementation would.  Raise to signal failure."""

    def __init__(self, estimate: Callable[[np.ndarray], int], max_compressed_size: Callable[[int], int] = lambda n: 0):
        self.calls: list[int] = []

        def _max(_ctx, n, out):
            try:
                out[0] = int(max_compressed_size(n))
                return 0
            except Exception:
                return 2

        def _est(_ctx, ptr, n, _scratch, _scratch_len, out):
            try:
                self.calls.append(n)
                arr = (
                    np.ctypeslib.as_array((C.c_uint8 * n).from_address(ptr)) if n and ptr else np.zeros(0, np.uint8)
                )
                out[0] = int(estimate(arr))
                return 0
            except Exception:
                return 3

        self._max, self._est = N.MaxCompressedSizeFn(_max), N.EstimateCompressedSizeFn(_est)
        self._struct = N.DltSizeEstimator(None, self._max, self._est)

    def c_estimator(self):
        return C.pointer(self._struct)


@dataclass
class Bc1EstimateSettings:
    """bc1/src/transform/transform_auto.rs:27-69."""

    size_estimator: SizeEstimator
    use_all_decorrelation_modes: bool = False


Bc2EstimateSettings = Bc1EstimateSettings
Bc3EstimateSettings = Bc1EstimateSettings


def _auto(fmt: int, input, output, options: Bc1EstimateSettings):
    ip, il, _ka = _ro(input)
    op, ol, _kb = _rw(output)
    fn = getattr(N.lib(), f"dltbc{fmt}core_transform_auto")
    details = N.CoreBc3Settings() if fmt == 3 else N.CoreSettings()
    r = fn(ip, il, op, ol, options.size_estimator.c_estimator(),
           N.CoreAutoSettings(bool(options.use_all_decorrelation_modes)), C.byref(details))
    _raise_core(r.error_code, il, ol)
    v = YCoCgVariant(details.decorrelation_mode)
    if fmt == 3:
        return Bc3TransformSettings(v, bool(details.split_alpha_endpoints), bool(details.split_colour_endpoints))
    return _SETTINGS[fmt](v, bool(details.split_colour_endpoints))


def transform_bc1_auto(input, output, options: Bc1EstimateSettings) -> Bc1TransformSettings:
    return _auto(1, input, output, options)


def transform_bc2_auto(input, output, options: Bc1EstimateSettings) -> Bc2TransformSettings:
    return _auto(2, input, output, options)


def transform_bc3_auto(input, output, options: Bc1EstimateSettings) -> Bc3TransformSettings:
    return _auto(3, input, output, options)


# --------------------------------------------------------------------------------------------------
# Stable builders (bc1-api / bc2-api), through the dltbcN_* C symbols
# --------------------------------------------------------------------------------------------------
class BcnError(RuntimeError):
    """Bc{1,2}Error (api/dxt-lossless-transform-bc1-api/src/error.rs:11-35) with the C error code."""

    def __init__(self, code: int, message: str):
        super().__init__(message)
        self.code = code


class _ManualBuilder:
    FORMAT = 1

    def __init__(self, _handle: Optional[int] = None):
        L = N.lib()
        self._h = _handle or getattr(L, f"dltbc{self.FORMAT}_new_ManualTransformBuilder")()

    def _fn(self, name):
        return getattr(N.lib(), f"dltbc{self.FORMAT}_{name}")

    def decorrelation_mode(self, mode: YCoCgVariant):
        self._fn("ManualTransformBuilder_SetDecorrelationMode")(self._h, YCoCgVariant(mode).to_stable())
        return self

    def split_colour_endpoints(self, split: bool):
        self._fn("ManualTransformBuilder_SetSplitColourEndpoints")(self._h, bool(split))
        return self

    def reset_to_defaults(self):
        self._fn("ManualTransformBuilder_ResetToDefaults")(self._h)
        return self

    def clone(self):
        return type(self)(self._fn("clone_ManualTransformBuilder")(self._h))

    def get_settings(self):
        mode, split = C.c_uint8(), C.c_bool()
        N.lib().dltcuda_ManualTransformBuilder_GetSettings(self._h, C.byref(mode), C.byref(split))
        return _SETTINGS[self.FORMAT](YCoCgVariant.from_stable(mode.value), split.value)

    def _run(self, name, input, output):
        ip, il, _ka = _ro(input)
        op, ol, _kb = _rw(output)
        r = self._fn(name)(ip, il, op, ol, self._h)
        if r.error_code:
            raise BcnError(r.error_code, self._fn("error_message")(r.error_code).decode())

    def transform(self, input, output) -> None:
        self._run("ManualTransformBuilder_Transform", input, output)

    def untransform(self, input, output) -> None:
        self._run("ManualTransformBuilder_Untransform", input, output)

    def __del__(self):
        try:
            if self._h:
                self._fn("free_ManualTransformBuilder")(self._h)
                self._h = None
        except Exception:
            pass


class Bc1ManualTransformBuilder(_ManualBuilder):
    FORMAT = 1


class Bc2ManualTransformBuilder(_ManualBuilder):
    FORMAT = 2


class _AutoBuilder:
    FORMAT = 1
    MANUAL = Bc1ManualTransformBuilder

    def __init__(self, estimator: SizeEstimator):
        self._estimator = estimator  # keep callbacks alive
        self._h = getattr(N.lib(), f"dltbc{self.FORMAT}_new_AutoTransformBuilder")(estimator.c_estimator())
        if not self._h:
            raise BcnError(6, "Null pointer provided for DltSizeEstimator parameter")

    @classmethod
    def new_ultra(cls, estimator: SizeEstimator):
        return cls(estimator).use_all_decorrelation_modes(True)

    def _fn(self, name):
        return getattr(N.lib(), f"dltbc{self.FORMAT}_{name}")

    def use_all_decorrelation_modes(self, use_all: bool):
        self._fn("AutoTransformBuilder_SetUseAllDecorrelationModes")(self._h, bool(use_all))
        return self

    def transform(self, input, output):
        """Returns the manual builder configured with the winning settings (auto_transform_builder.rs:123)."""
        ip, il, _ka = _ro(input)
        op, ol, _kb = _rw(output)
        out = C.c_void_p()
        r = self._fn("AutoTransformBuilder_Transform")(self._h, ip, il, op, ol, C.byref(out))
        if r.error_code:
            raise BcnError(r.error_code, self._fn("error_message")(r.error_code).decode())
        return self.MANUAL(out.value)

    def __del__(self):
        try:
            if self._h:
                self._fn("free_AutoTransformBuilder")(self._h)
                self._h = None
        except Exception:
            pass


class Bc1AutoTransformBuilder(_AutoBuilder):
    FORMAT = 1
    MANUAL = Bc1ManualTransformBuilder


class Bc2AutoTransformBuilder(_AutoBuilder):
    FORMAT = 2
    MANUAL = Bc2ManualTransformBuilder


# --------------------------------------------------------------------------------------------------
# Device-resident entry points (additive dltcuda_* API); pointers are raw CUDA device addresses
# --------------------------------------------------------------------------------------------------
def _dsettings(fmt: int, s) -> N.DltcudaSettings:
    return N.DltcudaSettings(fmt, int(s.decorrelation_mode), bool(getattr(s, "split_alpha_endpoints", False)),
                             bool(s.split_colour_endpoints))


def _check_device(rc: int) -> None:
    if rc:
        raise TransformationError(f"dltcuda status {rc}: {N.lib().dltcuda_last_error().decode()}")


def transform_device(fmt: int, d_in: int, d_out: int, nbytes: int, settings, stream: int = 0) -> None:
    _check_device(N.lib().dltcuda_transform_device(d_in, d_out, nbytes, _dsettings(fmt, settings), stream))


def untransform_device(fmt: int, d_in: int, d_out: int, nbytes: int, settings, stream: int = 0) -> None:
    _check_device(N.lib().dltcuda_untransform_device(d_in, d_out, nbytes, _dsettings(fmt, settings), stream))


def transform_device_range(fmt, d_blocks, d_streams_base, total_blocks, first_block, num_blocks, settings, stream=0):
    _check_device(N.lib().dltcuda_transform_device_range(d_blocks, d_streams_base, total_blocks, first_block,
                                                         num_blocks, _dsettings(fmt, settings), stream))


def untransform_device_range(fmt, d_streams_base, d_blocks, total_blocks, first_block, num_blocks, settings, stream=0):
    _check_device(N.lib().dltcuda_untransform_device_range(d_streams_base, d_blocks, total_blocks, first_block,
                                                           num_blocks, _dsettings(fmt, settings), stream))


def transform_device_streams(fmt, d_blocks, d_streams, num_blocks, settings, stream=0):
    arr = (C.c_void_p * 6)(*d_streams)
    _check_device(N.lib().dltcuda_transform_device_streams(d_blocks, arr, num_blocks, _dsettings(fmt, settings), stream))


def untransform_device_streams(fmt, d_streams, d_blocks, num_blocks, settings, stream=0):
    arr = (C.c_void_p * 6)(*d_streams)
    _check_device(N.lib().dltcuda_untransform_device_streams(arr, d_blocks, num_blocks, _dsettings(fmt, settings), stream))


def split_color_endpoints(colors, colors_out) -> None:
    """split_565_color_endpoints::split_color_endpoints on host buffers: [c0 c1] x n -> c0 x n | c1 x n."""
    ip, il, _ka = _ro(colors)
    op, ol, _kb = _rw(colors_out)
    if ol < il:
        raise OutputBufferTooSmall(il, ol)
    rc = N.lib().dltcuda_split_color_endpoints(ip, op, il)
    if rc == 1:
        raise InvalidLength(il)
    _check_device(rc)


def split_color_endpoints_device(d_colors: int, d_colors_out: int, nbytes: int, stream: int = 0) -> None:
    _check_device(N.lib().dltcuda_split_color_endpoints_device(d_colors, d_colors_out, nbytes, stream))


def transform_batch(items, untransform: bool = False, devices=None) -> None:
    """items: iterable of (fmt, input buffer, output buffer, settings).  One pipelined pass over the whole
    batch on the current device, or dealt out over `devices` (payload-granular multi-GPU)."""
    items = list(items)
    arr = (N.DltcudaPayload * max(len(items), 1))()
    keep = []
    for i, (fmt, src, dst, settings) in enumerate(items):
        ip, il, ka = _ro(src)
        op, ol, kb = _rw(dst)
        if ol < il:
            raise OutputBufferTooSmall(il, ol)
        keep += [ka, kb]
        arr[i] = N.DltcudaPayload(ip, op, il, _dsettings(fmt, settings))
    if devices is None:
        rc = N.lib().dltcuda_transform_batch(arr, len(items), untransform)
    else:
        dev = (C.c_int * len(devices))(*devices)
        rc = N.lib().dltcuda_transform_batch_multi_gpu(arr, len(items), untransform, dev, len(devices))
    if rc == 1:
        raise InvalidLength(-1)
    _check_device(rc)


def shard_first_block(fmt: int, total_blocks: int, shard: int, num_shards: int) -> int:
    return N.lib().dltcuda_shard_first_block(fmt, total_blocks, shard, num_shards)


def ltu_estimate_device(d_data: int, nbytes: int) -> int:
    out = C.c_size_t(0)
    _check_device(N.lib().dltcuda_ltu_estimate_device(d_data, nbytes, C.byref(out)))
    return out.value


def ltu_set_params(hash_bits: int = 16, index_top: bool = True, group: int = 4) -> None:
    """Process-wide parameters of the LTU-semantics estimator (dltcuda_ltu_set_params): the parts of the restated
    third-party algorithm that are unverified.  Defaults = the restatement."""
    if N.lib().dltcuda_ltu_set_params(hash_bits, index_top, group) != 0:
        raise ValueError(f"unsupported LTU parameters: hash_bits={hash_bits} index_top={index_top} group={group}")


def ltu_get_params():
    h, t, g = C.c_int(0), C.c_bool(False), C.c_int(0)
    N.lib().dltcuda_ltu_get_params(C.byref(h), C.byref(t), C.byref(g))
    return h.value, t.value, g.value


def auto_candidates(fmt: int, use_all: bool):
    arr = (N.DltcudaSettings * 16)()
    k = N.lib().dltcuda_auto_candidates(fmt, use_all, arr)
    out = []
    for i in range(k):
        v = YCoCgVariant(arr[i].decorrelation_mode)
        if fmt == 3:
            out.append(Bc3TransformSettings(v, bool(arr[i].split_alpha_endpoints), bool(arr[i].split_colour_endpoints)))
        else:
            out.append(_SETTINGS[fmt](v, bool(arr[i].split_colour_endpoints)))
    return out


def transform_auto_device(fmt: int, d_in: int, d_out: int, nbytes: int, use_all: bool):
    """Returns (winning settings, per-candidate estimates in test order)."""
    s = N.DltcudaSettings()
    est = (C.c_size_t * 16)()
    _check_device(N.lib().dltcuda_transform_auto_device(fmt, d_in, d_out, nbytes, use_all, C.byref(s), est))
    v = YCoCgVariant(s.decorrelation_mode)
    k = (16 if use_all else 8) if fmt == 3 else (8 if use_all else 4)
    if fmt == 3:
        best = Bc3TransformSettings(v, bool(s.split_alpha_endpoints), bool(s.split_colour_endpoints))
    else:
        best = _SETTINGS[fmt](v, bool(s.split_colour_endpoints))
    return best, [est[i] for i in range(k)]


def transform_auto_batch(items, use_all: bool = False, devices=None) -> list:
    """transform_bcN_auto (GPU LTU estimator) for a batch of independent host payloads: ``items`` is a list of
    ``(fmt, input, output)``; returns the winning settings per payload.  One set of estimator launches serves all
    candidates of all payloads (dltcuda_transform_auto_batch); ``devices`` deals whole payloads out over several GPUs."""
    jobs = (N.DltcudaAutoJob * len(items))()
    keep = []
    for i, (fmt, inp, out) in enumerate(items):
        ip, il, ka = _ro(inp)
        op, ol, kb = _rw(out)
        if il % block_bytes(fmt):
            raise InvalidLength(il)
        if ol < il:
            raise OutputBufferTooSmall(il, ol)
        keep += [ka, kb]
        jobs[i] = N.DltcudaAutoJob(fmt, ip, op, il, N.DltcudaSettings(), 0)
    if devices:
        dev = (C.c_int * len(devices))(*devices)
        _check_device(N.lib().dltcuda_transform_auto_batch_multi_gpu(jobs, len(items), bool(use_all), dev, len(devices)))
    else:
        _check_device(N.lib().dltcuda_transform_auto_batch(jobs, len(items), bool(use_all)))
    out = []
    for j in jobs:
        s = j.out_settings
        if j.format == 3:
            out.append(Bc3TransformSettings(YCoCgVariant(s.decorrelation_mode), bool(s.split_alpha_endpoints), bool(s.split_colour_endpoints)))
        else:
            out.append(_SETTINGS[j.format](YCoCgVariant(s.decorrelation_mode), bool(s.split_colour_endpoints)))
    return out


def kernel_launch_count() -> int:
    return N.lib().dltcuda_kernel_launch_count()


def device_count() -> int:
    return N.lib().dltcuda_device_count()


def set_device(device: int) -> None:
    N.lib().dltcuda_set_device(device)
