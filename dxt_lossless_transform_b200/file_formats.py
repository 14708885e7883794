"""Host-side mirror of the reference's file-format layer over the C ABI of include/dxt_lossless_transform_file_formats.h
and include/dxt_lossless_transform_dds.h.

Same names, argument meaning and error behaviour as the Rust items
(api/dxt-lossless-transform-file-formats-api/src: embed/, bundle/, handlers/dispatch.rs, api.rs, error.rs;
extensions/file-formats/dxt-lossless-transform-dds/src: dds/parse_dds.rs, handler/).  The texture data always
goes through the CUDA library; there is no CPU path here.
"""
from __future__ import annotations

import ctypes as C
import enum
from dataclasses import dataclass
from typing import Iterable, Optional, Sequence

from . import _native as N
from .api import (
    Bc1AutoTransformBuilder,
    Bc1ManualTransformBuilder,
    Bc1TransformSettings,
    Bc2AutoTransformBuilder,
    Bc2ManualTransformBuilder,
    Bc2TransformSettings,
    YCoCgVariant,
    _ro,
    _rw,
)

TRANSFORM_HEADER_SIZE = 4  # embed/mod.rs:87


class TransformFormat(enum.IntEnum):
    """embed/transform_format.rs:10-32"""

    Bc1 = 0
    Bc2 = 1
    Bc3 = 2
    Bc7 = 3
    Bc6H = 4
    Rgba8888 = 5
    Bgra8888 = 6
    Bgr888 = 7
    Bc4 = 8
    Bc5 = 9


class DdsFormat(enum.IntEnum):
    """dds/parse_dds.rs:7-33"""

    NotADds = 0
    Unknown = 1
    BC1 = 2
    BC2 = 3
    BC3 = 4
    BC6H = 5
    BC7 = 6
    RGBA8888 = 7
    BGRA8888 = 8
    BGR888 = 9
    BC4 = 10
    BC5 = 11


# --------------------------------------------------------------------------------------------------
# Errors (error.rs:19-87, embed/embed_error.rs:7-15)
# --------------------------------------------------------------------------------------------------
class TransformError(Exception):
    """Base of everything the file-format layer raises; ``code`` is the DltffErrorCode."""

    code = -1


class EmbedError(TransformError):
    pass


class CorruptedEmbeddedData(EmbedError):
    code = 1


class UnknownFormat(EmbedError):
    code = 2


class FormatHandlerError(TransformError):
    pass


class UnknownFileFormat(FormatHandlerError):
    code = 3


class InvalidInputFileHeader(FormatHandlerError):
    code = 4


class InvalidRestoredFileHeader(FormatHandlerError):
    code = 5


class _WithFormat(FormatHandlerError):
    def __init__(self, format: int):
        self.format = TransformFormat(format)
        super().__init__(f"{type(self).__name__}({self.format.name})")


class FormatNotImplemented(_WithFormat):
    code = 6


class NoBuilderForFormat(_WithFormat):
    code = 7


class _RequiredActual(FormatHandlerError):
    def __init__(self, required: int, actual: int):
        self.required, self.actual = required, actual
        super().__init__(f"{type(self).__name__}: required {required} bytes, got {actual} bytes")


class OutputBufferTooSmall(_RequiredActual):
    code = 8


class InputTooShort(_RequiredActual):
    code = 9


class InputTooShortForStatedTextureSize(_RequiredActual):
    code = 10


class _Wrapped(TransformError):
    """TransformError::Bc1 / ::Bc2: ``inner_code`` is the Dltbc{1,2}ErrorCode of the stable API."""

    def __init__(self, inner_code: int, payload: int = 0):
        self.inner_code, self.payload = inner_code, payload
        super().__init__(f"{type(self).__name__}(code {inner_code})")


class Bc1TransformError(_Wrapped):
    code = 11


class Bc2TransformError(_Wrapped):
    code = 12


class UnknownTransformFormat(TransformError):
    code = 13


class InvalidDataAlignment(TransformError):
    code = 14

    def __init__(self, size: int, required_divisor: int):
        self.size, self.required_divisor = size, required_divisor
        super().__init__(f"Invalid data alignment: size {size} is not divisible by {required_divisor}")


class NoSupportedHandler(TransformError):
    code = 15


class NullPointer(TransformError):
    code = 16


_PLAIN = {c.code: c for c in (CorruptedEmbeddedData, UnknownFormat, UnknownFileFormat, InvalidInputFileHeader,
                              InvalidRestoredFileHeader, UnknownTransformFormat, NoSupportedHandler, NullPointer)}


def error_from_result(r: N.DltffResult) -> Optional[TransformError]:
    """The exception a DltffResult stands for (None for success)."""
    code, a, b = r.error_code, r.detail_a, r.detail_b
    if code == 0:
        return None
    if code in _PLAIN:
        return _PLAIN[code](N.lib().dltff_error_message(code).decode())
    if code in (6, 7):
        return (FormatNotImplemented if code == 6 else NoBuilderForFormat)(a)
    if code in (8, 9, 10):
        return {8: OutputBufferTooSmall, 9: InputTooShort, 10: InputTooShortForStatedTextureSize}[code](a, b)
    if code in (11, 12):
        return (Bc1TransformError if code == 11 else Bc2TransformError)(a, b)
    if code == 14:
        return InvalidDataAlignment(a, b)
    return TransformError(f"unknown DltffErrorCode {code}")


def _check(r: N.DltffResult) -> None:
    e = error_from_result(r)
    if e is not None:
        raise e


# --------------------------------------------------------------------------------------------------
# TransformHeader (embed/mod.rs:107-160)
# --------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class TransformHeader:
    value: int = 0

    @staticmethod
    def new(format: TransformFormat, data: int) -> "TransformHeader":
        return TransformHeader(N.lib().dltff_TransformHeader_new(int(format), data & 0xFFFFFFFF))

    def format(self) -> Optional[TransformFormat]:
        raw = C.c_int32()
        known = N.lib().dltff_TransformHeader_format(self.value, C.byref(raw))
        return TransformFormat(raw.value) if known else None

    def format_data(self) -> int:
        return N.lib().dltff_TransformHeader_format_data(self.value)

    @staticmethod
    def read_from(buf) -> "TransformHeader":
        p, n, _k = _ro(buf)
        if n < TRANSFORM_HEADER_SIZE:
            raise InputTooShort(TRANSFORM_HEADER_SIZE, n)
        return TransformHeader(N.lib().dltff_TransformHeader_read(p))

    def write_to(self, buf) -> None:
        p, n, _k = _rw(buf)
        if n < TRANSFORM_HEADER_SIZE:
            raise OutputBufferTooSmall(TRANSFORM_HEADER_SIZE, n)
        N.lib().dltff_TransformHeader_write(self.value, p)

    # EmbeddableBc{1,2}Details
    @staticmethod
    def from_bc1_settings(s: Bc1TransformSettings) -> "TransformHeader":
        return TransformHeader(N.lib().dltff_bc1_header_from_settings(s.decorrelation_mode.to_stable(), s.split_colour_endpoints))

    @staticmethod
    def from_bc2_settings(s: Bc2TransformSettings) -> "TransformHeader":
        return TransformHeader(N.lib().dltff_bc2_header_from_settings(s.decorrelation_mode.to_stable(), s.split_colour_endpoints))

    def _settings(self, n: int):
        mode, split = C.c_uint8(), C.c_bool()
        _check(getattr(N.lib(), f"dltff_bc{n}_settings_from_header")(self.value, C.byref(mode), C.byref(split)))
        cls = Bc1TransformSettings if n == 1 else Bc2TransformSettings
        return cls(YCoCgVariant.from_stable(mode.value), split.value)

    def to_bc1_settings(self) -> Bc1TransformSettings:
        return self._settings(1)

    def to_bc2_settings(self) -> Bc2TransformSettings:
        return self._settings(2)


# --------------------------------------------------------------------------------------------------
# TransformBundle (bundle/mod.rs)
# --------------------------------------------------------------------------------------------------
class TransformBundle:
    def __init__(self, _handle: Optional[int] = None):
        self._h = _handle or N.lib().dltff_new_TransformBundle()
        self._keep = []  # estimators of auto builders must outlive the bundle (their callbacks are borrowed)

    @classmethod
    def new(cls) -> "TransformBundle":
        return cls()

    @classmethod
    def default_all(cls) -> "TransformBundle":
        return cls(N.lib().dltff_TransformBundle_default_all())

    def _with(self, name: str, builder, kind) -> "TransformBundle":
        if not isinstance(builder, kind):
            raise TypeError(f"{name} expects a {kind.__name__}")
        _check(getattr(N.lib(), f"dltff_TransformBundle_{name}")(self._h, builder._h))
        self._keep.append(builder)
        return self

    def with_bc1_manual(self, builder: Bc1ManualTransformBuilder) -> "TransformBundle":
        return self._with("with_bc1_manual", builder, Bc1ManualTransformBuilder)

    def with_bc1_auto(self, builder: Bc1AutoTransformBuilder) -> "TransformBundle":
        return self._with("with_bc1_auto", builder, Bc1AutoTransformBuilder)

    def with_bc2_manual(self, builder: Bc2ManualTransformBuilder) -> "TransformBundle":
        return self._with("with_bc2_manual", builder, Bc2ManualTransformBuilder)

    def with_bc2_auto(self, builder: Bc2AutoTransformBuilder) -> "TransformBundle":
        return self._with("with_bc2_auto", builder, Bc2AutoTransformBuilder)

    def __del__(self):
        try:
            if self._h:
                N.lib().dltff_free_TransformBundle(self._h)
                self._h = None
        except Exception:
            pass


def dispatch_transform(format: TransformFormat, input_texture_data, output_texture_data, bundle: TransformBundle) -> TransformHeader:
    """handlers/dispatch.rs:131-143"""
    ip, il, _a = _ro(input_texture_data)
    op, ol, _b = _rw(output_texture_data)
    hdr = C.c_uint32()
    _check(N.lib().dltff_dispatch_transform(int(format), ip, il, op, ol, bundle._h, C.byref(hdr)))
    return TransformHeader(hdr.value)


def dispatch_untransform(header: TransformHeader, input_texture_data, output_texture_data) -> None:
    """handlers/dispatch.rs:41-101"""
    ip, il, _a = _ro(input_texture_data)
    op, ol, _b = _rw(output_texture_data)
    _check(N.lib().dltff_dispatch_untransform(header.value, ip, il, op, ol))


# --------------------------------------------------------------------------------------------------
# DDS (dxt-lossless-transform-dds)
# --------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class DdsInfo:
    format: DdsFormat
    data_offset: int
    data_length: int


def is_dds(data) -> bool:
    p, n, _k = _ro(data)
    return bool(N.lib().is_dds(p, n))


def _info(raw: N.DdsInfo) -> Optional[DdsInfo]:
    if raw.format == DdsFormat.NotADds:
        return None
    return DdsInfo(DdsFormat(raw.format), raw.data_offset, raw.data_length)


def parse_dds(data) -> Optional[DdsInfo]:
    """dds/parse_dds.rs:58-64 — None when `data` is not a DDS."""
    p, n, _k = _ro(data)
    return _info(N.lib().parse_dds(p, n))


def parse_dds_ignore_magic(data) -> Optional[DdsInfo]:
    """dds/parse_dds.rs:78-172"""
    p, n, _k = _ro(data)
    return _info(N.lib().dltdds_parse_dds_ignore_magic(p, n))


class DdsHandler:
    """FileFormatHandler + FileFormatDetection + FileFormatUntransformDetection for DDS (handler/*.rs)."""

    def transform_bundle(self, input, output, bundle: TransformBundle) -> None:
        ip, il, _a = _ro(input)
        op, ol, _b = _rw(output)
        _check(N.lib().dltdds_transform_bundle(ip, il, op, ol, bundle._h))

    def untransform(self, input, output) -> None:
        ip, il, _a = _ro(input)
        op, ol, _b = _rw(output)
        _check(N.lib().dltdds_untransform(ip, il, op, ol))

    def can_handle(self, input, file_extension: Optional[str] = None) -> bool:
        p, n, _k = _ro(input)
        return bool(N.lib().dltdds_can_handle(p, n, None if file_extension is None else file_extension.encode()))

    def can_handle_untransform(self, input, file_extension: Optional[str] = None) -> bool:
        p, n, _k = _ro(input)
        return bool(N.lib().dltdds_can_handle_untransform(p, n, None if file_extension is None else file_extension.encode()))

    # ---- batched (no reference counterpart: one GPU pipeline for a whole directory) ----
    @staticmethod
    def _files(pairs):
        keep, arr = [], (N.DltddsFile * len(pairs))()
        for i, (inp, out) in enumerate(pairs):
            ip, il, a = _ro(inp)
            op, ol, b = _rw(out)
            keep += [a, b]
            arr[i] = N.DltddsFile(ip, il, op, ol)
        return arr, keep

    @staticmethod
    def _devices(devices):
        if not devices:
            return None, 0
        return (C.c_int * len(devices))(*devices), len(devices)

    def transform_bundle_batch(self, pairs: Sequence[tuple], bundle: TransformBundle,
                               devices: Optional[Sequence[int]] = None) -> list[Optional[TransformError]]:
        """[(input, output), ...] -> one entry per file: None, or the error transform_bundle would have raised."""
        arr, _keep = self._files(pairs)
        res = (N.DltffResult * len(pairs))()
        dev, nd = self._devices(devices)
        if N.lib().dltdds_transform_bundle_batch(arr, len(pairs), bundle._h, res, dev, nd):
            raise NullPointer("null argument")
        return [error_from_result(r) for r in res]

    def untransform_batch(self, pairs: Sequence[tuple], devices: Optional[Sequence[int]] = None) -> list[Optional[TransformError]]:
        arr, _keep = self._files(pairs)
        res = (N.DltffResult * len(pairs))()
        dev, nd = self._devices(devices)
        if N.lib().dltdds_untransform_batch(arr, len(pairs), res, dev, nd):
            raise NullPointer("null argument")
        return [error_from_result(r) for r in res]


# --------------------------------------------------------------------------------------------------
# api.rs: slice-level convenience functions over handler objects
# --------------------------------------------------------------------------------------------------
def _size_check(input, output) -> None:
    _ip, il, _a = _ro(input)
    _op, ol, _b = _rw(output)
    if ol < il:
        raise OutputBufferTooSmall(il, ol)


def transform_slice_with_bundle(handler, input, output, bundle: TransformBundle) -> None:
    """api.rs `transform_slice_with_bundle`"""
    _size_check(input, output)
    handler.transform_bundle(input, output, bundle)


def untransform_slice(handler, input, output) -> None:
    """api.rs `untransform_slice`"""
    _size_check(input, output)
    handler.untransform(input, output)


def transform_slice_with_multiple_handlers(handlers: Iterable, input, output, bundle: TransformBundle):
    """api.rs `transform_slice_with_multiple_handlers`: the first handler whose can_handle(input, None) accepts."""
    _size_check(input, output)
    for h in handlers:
        if h.can_handle(input, None):
            h.transform_bundle(input, output, bundle)
            return h
    raise NoSupportedHandler("No file format handler can process the file")


def untransform_slice_with_multiple_handlers(handlers: Iterable, input, output):
    """api.rs `untransform_slice_with_multiple_handlers`"""
    _size_check(input, output)
    for h in handlers:
        if h.can_handle_untransform(input, None):
            h.untransform(input, output)
            return h
    raise NoSupportedHandler("No file format handler can process the file")
