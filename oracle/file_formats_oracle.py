"""CPU restatement of the reference's file-format layer — TEST INFRASTRUCTURE ONLY (see oracle/bcn_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product path is
dxt_lossless_transform_b200/csrc/file_formats.cu.  Every function cites the reference lines it follows
(paths relative to /root/reference/src).  Payload transforms go through the C oracle (`tests/oracle.py`).

Pinned by the reference's own tests: embed/mod.rs tests (header bit layout, 0x12345673 byte order),
embed/formats/bc1.rs tests, dds/parse_dds.rs tests (format tables, data lengths 32768 / 43704 / 160 / 0 /
1024 / 84), handler/file_format_handler.rs tests (error variants), and the reference's DDS fixtures
(assets/tests/r2-256-bc{1,2,3,7}.dds headers, stored in tests/golden/known_answers.json).
"""
from __future__ import annotations

import struct
from typing import Optional

U32 = 0xFFFFFFFF

# ---- embed/transform_format.rs:10-32 ------------------------------------------------------------
BC1, BC2, BC3, BC7, BC6H, RGBA8888, BGRA8888, BGR888, BC4, BC5 = range(10)

# ---- dds/parse_dds.rs:7-33 ------------------------------------------------------------------------
DDS_NOT_A_DDS, DDS_UNKNOWN, DDS_BC1, DDS_BC2, DDS_BC3, DDS_BC6H, DDS_BC7, DDS_RGBA8888, DDS_BGRA8888, DDS_BGR888, DDS_BC4, DDS_BC5 = range(12)


class OracleError(Exception):
    """(variant name, *fields) of the TransformError the reference would return."""

    def __init__(self, variant: str, *fields):
        super().__init__(variant, *fields)
        self.variant, self.fields = variant, fields


# ---- TransformHeader: embed/mod.rs:107-160 -------------------------------------------------------
def header_new(fmt: int, data: int) -> int:
    return (fmt & 0xF) | ((data & 0x0FFFFFFF) << 4)


def header_format(h: int) -> Optional[int]:
    raw = h & 0xF
    return raw if raw <= 9 else None  # transform_format.rs:39-55


def header_data(h: int) -> int:
    return (h >> 4) & 0x0FFFFFFF


# ---- BC1 / BC2 details: embed/formats/bc1.rs:33-118, bc2.rs:30-118 -------------------------------
# `variant` is the INTERNAL numbering (None=0, Variant1..3=1..3); the header stores Variant1=0..None=3.
def pack_bc12(variant: int, split_colour: bool) -> int:
    stable = 3 if variant == 0 else variant - 1
    return 0 | (int(bool(split_colour)) << 2) | (stable << 3)


def unpack_bc12(data: int) -> tuple[int, bool]:
    if data & 3:
        raise OracleError("Embed::CorruptedEmbeddedData")
    stable = (data >> 3) & 3
    return (0 if stable == 3 else stable + 1), bool((data >> 2) & 1)


# ---- DDS: dds/constants.rs, dds/likely_dds.rs, dds/parse_dds.rs ----------------------------------
DDS_MAGIC = 0x20534444
HEADER, DX10 = 0x80, 20


def _u32(d: bytes, off: int) -> int:
    return struct.unpack_from("<I", d, off)[0]


def likely_dds(d: bytes) -> bool:
    return len(d) >= HEADER and _u32(d, 0) == DDS_MAGIC


_DXGI = {}
for _lo, _hi, _f in ((70, 72, DDS_BC1), (73, 75, DDS_BC2), (76, 78, DDS_BC3), (79, 81, DDS_BC4), (82, 84, DDS_BC5),
                     (94, 96, DDS_BC6H), (97, 99, DDS_BC7), (27, 32, DDS_RGBA8888)):
    for _v in range(_lo, _hi + 1):
        _DXGI[_v] = _f
for _v in (87, 90, 91):
    _DXGI[_v] = DDS_BGRA8888

_FOURCC = {b"DXT1": DDS_BC1, b"DXT2": DDS_BC2, b"DXT3": DDS_BC2, b"DXT4": DDS_BC3, b"DXT5": DDS_BC3,
           b"BC4U": DDS_BC4, b"BC4S": DDS_BC4, b"ATI1": DDS_BC4, b"BC5U": DDS_BC5, b"BC5S": DDS_BC5, b"ATI2": DDS_BC5}


def _mip_chain(w: int, h: int, mips: int, level) -> int:
    total = 0
    for i in range(mips):
        if w == 1 and h == 1:  # constant tail; avoids looping over a forged 2^32-1 level count
            return min(U32, total + level(1, 1) * (mips - i))
        total = min(U32, total + level(w, h))  # saturating_add
        w, h = max(w // 2, 1), max(h // 2, 1)
    return total


def _blocks(w, h, mips, bs):  # parse_dds.rs:283-340 (u32 products wrap in release builds)
    return _mip_chain(w, h, mips, lambda lw, lh: (-(-lw // 4) * -(-lh // 4) * bs) & U32)


def _pixels(w, h, mips, bpp):  # parse_dds.rs:374-400
    return _mip_chain(w, h, mips, lambda lw, lh: (lw * lh * bpp) & U32)


def _data_length(fmt: int, d: bytes) -> int:  # parse_dds.rs:241-281
    flags, height, width, raw_mips = _u32(d, 0x08), _u32(d, 0x0C), _u32(d, 0x10), _u32(d, 0x1C)
    mips = max(raw_mips, 1) if flags & 0x20000 else 1
    if fmt in (DDS_BC1, DDS_BC4):
        return _blocks(width, height, mips, 8)
    if fmt in (DDS_BC2, DDS_BC3, DDS_BC5, DDS_BC6H, DDS_BC7):
        return _blocks(width, height, mips, 16)
    if fmt in (DDS_RGBA8888, DDS_BGRA8888):
        return _pixels(width, height, mips, 4)
    if fmt == DDS_BGR888:
        return _pixels(width, height, mips, 3)
    # Unknown: parse_dds.rs:343-371
    pf, bits = _u32(d, 0x50), _u32(d, 0x58)
    if pf & (0x40 | 0x20000 | 0x200 | 0x2) == 0 or bits % 8 or bits // 8 == 0:
        return 0
    return _pixels(width, height, mips, bits // 8)


def parse_dds_ignore_magic(d: bytes) -> Optional[tuple[int, int, int]]:
    """(format, data_offset, data_length) or None — parse_dds.rs:78-172."""
    if len(d) < HEADER:
        return None
    cc = bytes(d[0x54:0x58])
    if cc == b"DX10":
        if len(d) < HEADER + DX10:
            return None
        fmt, off = _DXGI.get(_u32(d, 0x80), DDS_UNKNOWN), HEADER + DX10
    else:
        pf = _u32(d, 0x50)
        if pf & 0x4:
            fmt = _FOURCC.get(cc, DDS_UNKNOWN)
        elif pf & 0x40:
            bits = _u32(d, 0x58)
            masks = tuple(_u32(d, o) for o in (0x5C, 0x60, 0x64, 0x68))
            fmt = DDS_UNKNOWN
            if bits == 24 and masks == (0x00FF0000, 0x0000FF00, 0x000000FF, 0):
                fmt = DDS_BGR888
            elif bits == 32 and pf & 0x1:
                if masks == (0x000000FF, 0x0000FF00, 0x00FF0000, 0xFF000000):
                    fmt = DDS_RGBA8888
                elif masks == (0x00FF0000, 0x0000FF00, 0x000000FF, 0xFF000000):
                    fmt = DDS_BGRA8888
        else:
            fmt = DDS_UNKNOWN
        off = HEADER
    return fmt, off, _data_length(fmt, d)


def parse_dds(d: bytes) -> Optional[tuple[int, int, int]]:
    return parse_dds_ignore_magic(d) if likely_dds(d) else None  # parse_dds.rs:58-64


def dds_to_transform_format(dds_format: int) -> int:  # handler/format_conversion.rs:45-107, allow_unimplemented = false
    if dds_format == DDS_BC1:
        return BC1
    if dds_format == DDS_BC2:
        return BC2
    not_impl = {DDS_BC3: BC3, DDS_BC4: BC4, DDS_BC5: BC5, DDS_BC6H: BC6H, DDS_BC7: BC7}
    if dds_format in not_impl:
        raise OracleError("FormatHandler::FormatNotImplemented", not_impl[dds_format])
    passthrough = {DDS_RGBA8888: RGBA8888, DDS_BGRA8888: BGRA8888, DDS_BGR888: BGR888}
    if dds_format in passthrough:
        return passthrough[dds_format]
    raise OracleError("FormatHandler::UnknownFileFormat")


# ---- bundle + dispatch: bundle/mod.rs:125-174, handlers/dispatch.rs:41-101 ----------------------
# A bundle is {BC1: settings | callable | None, BC2: ...}: `(variant, split_colour)` stands for a manual
# builder, a callable `(fmt, payload) -> (out_bytes, (variant, split_colour))` for an auto builder.
def dispatch_transform(fmt: int, payload: bytes, out_len: int, bundle: dict, transform) -> tuple[bytes, int]:
    if out_len < len(payload):
        raise OracleError("FormatHandler::OutputBufferTooSmall", len(payload), out_len)
    if fmt not in (BC1, BC2):
        raise OracleError("UnknownTransformFormat")
    builder = bundle.get(fmt)
    if builder is None:
        raise OracleError("FormatHandler::NoBuilderForFormat", fmt)
    n = 1 if fmt == BC1 else 2
    if len(payload) % (8 if n == 1 else 16):  # safe/transform_with_settings.rs:88-105 through the builder
        raise OracleError("Bc1" if n == 1 else "Bc2", "InvalidLength", len(payload))
    if callable(builder):
        out, (variant, split) = builder(n, payload)
    else:
        variant, split = builder
        out = transform(n, payload, variant, split)
    return out, header_new(fmt, pack_bc12(variant, split))


def dispatch_untransform(header: int, payload: bytes, out_len: int, untransform) -> bytes:
    if out_len < len(payload):
        raise OracleError("FormatHandler::OutputBufferTooSmall", len(payload), out_len)
    fmt = header_format(header)
    if fmt not in (BC1, BC2):
        raise OracleError("UnknownTransformFormat")
    variant, split = unpack_bc12(header_data(header))
    n = 1 if fmt == BC1 else 2
    div = 8 if n == 1 else 16
    if len(payload) % div:
        raise OracleError("InvalidDataAlignment", len(payload), div)
    return untransform(n, payload, variant, split)


# ---- DdsHandler: handler/file_format_handler.rs:17-145 -------------------------------------------
def dds_transform_bundle(inp: bytes, out_len: int, bundle: dict, transform) -> bytes:
    if out_len < len(inp):
        raise OracleError("FormatHandler::OutputBufferTooSmall", len(inp), out_len)
    info = parse_dds(inp)
    if info is None:
        raise OracleError("FormatHandler::InvalidInputFileHeader")
    dds_format, off, length = info
    if len(inp) < off + length:
        raise OracleError("FormatHandler::InputTooShortForStatedTextureSize", off + length, len(inp))
    fmt = dds_to_transform_format(dds_format)
    payload, header = dispatch_transform(fmt, inp[off:off + length], length, bundle, transform)
    return struct.pack("<I", header) + inp[4:off] + payload + inp[off + length:]


def dds_untransform(inp: bytes, out_len: int, untransform) -> bytes:
    if len(inp) < 4:
        raise OracleError("FormatHandler::InputTooShort", 4, len(inp))
    if out_len < len(inp):
        raise OracleError("FormatHandler::OutputBufferTooSmall", len(inp), out_len)
    header = _u32(inp, 0)
    info = parse_dds_ignore_magic(inp)
    if info is None:
        raise OracleError("FormatHandler::InvalidRestoredFileHeader")
    _f, off, length = info
    if len(inp) < off + length:
        raise OracleError("FormatHandler::InputTooShortForStatedTextureSize", off + length, len(inp))
    payload = dispatch_untransform(header, inp[off:off + length], length, untransform)
    return struct.pack("<I", DDS_MAGIC) + inp[4:off] + payload + inp[off + length:]


def can_handle(inp: bytes, ext: Optional[str]) -> bool:  # handler/file_format_detection.rs:7-17
    return (ext is None or ext == "dds") and parse_dds(inp) is not None


def can_handle_untransform(inp: bytes, ext: Optional[str]) -> bool:  # handler/file_format_untransform_detection.rs:7-22
    return (ext is None or ext == "dds") and len(inp) >= 4 and parse_dds_ignore_magic(inp) is not None
