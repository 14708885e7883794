"""DDS files for the tests: the reference's real fixtures (header from tests/golden + payload) and synthetic
headers built field by field the way the reference's test prelude does
(extensions/file-formats/dxt-lossless-transform-dds/src/test_prelude.rs:40-330)."""
from __future__ import annotations

import json
import struct
import sys
import zlib
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden"
sys.path.insert(0, str(ROOT / "oracle"))
import file_formats_oracle as FO  # noqa: E402

KNOWN = json.loads((GOLDEN / "known_answers.json").read_text())

DDSD = dict(CAPS=0x1, HEIGHT=0x2, WIDTH=0x4, PIXELFORMAT=0x1000, LINEARSIZE=0x80000, MIPMAPCOUNT=0x20000)
FOURCC = {FO.DDS_BC1: b"DXT1", FO.DDS_BC2: b"DXT3", FO.DDS_BC3: b"DXT5", FO.DDS_BC4: b"BC4U", FO.DDS_BC5: b"BC5U",
          FO.DDS_UNKNOWN: b"UNKN"}
DXGI = {FO.DDS_BC6H: 95, FO.DDS_BC7: 98}


def real_fixture(name: str) -> np.ndarray:
    """assets/tests/r2-256-<name>.dds, name in bc1 / bc2 / bc3 (256x256, no mips, 4096 blocks)."""
    hdr = bytes.fromhex(KNOWN["dds_fixtures"][name]["header"])
    payload = zlib.decompress((GOLDEN / f"r2-256-{name}.payload.zlib").read_bytes())
    data = np.frombuffer(hdr + payload, np.uint8).copy()
    assert data.size == KNOWN["dds_fixtures"][name]["file_len"]
    return data


def header_base(width: int, height: int, mips: int, dx10: bool) -> bytearray:
    d = bytearray(148 if dx10 else 128)
    flags = DDSD["CAPS"] | DDSD["HEIGHT"] | DDSD["WIDTH"] | DDSD["PIXELFORMAT"] | DDSD["LINEARSIZE"]
    if mips > 1:
        flags |= DDSD["MIPMAPCOUNT"]
    struct.pack_into("<IIIII", d, 0, FO.DDS_MAGIC, 124, flags, height, width)
    if mips > 1:
        struct.pack_into("<I", d, 0x1C, mips)
    return d


def make_dds(dds_format: int, width: int, height: int, mips: int = 1, payload: bytes | None = None,
             leftover: bytes = b"") -> np.ndarray:
    """A valid DDS of the given format; payload defaults to the reference's `(i % 256)` test pattern."""
    dx10 = dds_format in DXGI
    d = header_base(width, height, mips, dx10)
    if dx10:
        d[0x54:0x58] = b"DX10"
        struct.pack_into("<I", d, 0x50, 0x4)
        struct.pack_into("<I", d, 0x80, DXGI[dds_format])
    elif dds_format in FOURCC:
        d[0x54:0x58] = FOURCC[dds_format]
        struct.pack_into("<I", d, 0x50, 0x4)
    elif dds_format in (FO.DDS_RGBA8888, FO.DDS_BGRA8888):
        masks = (0xFF, 0xFF00, 0xFF0000, 0xFF000000) if dds_format == FO.DDS_RGBA8888 else (0xFF0000, 0xFF00, 0xFF, 0xFF000000)
        struct.pack_into("<I", d, 0x50, 0x40 | 0x1)
        struct.pack_into("<IIIII", d, 0x58, 32, *masks)
    elif dds_format == FO.DDS_BGR888:
        struct.pack_into("<I", d, 0x50, 0x40)
        struct.pack_into("<IIIII", d, 0x58, 24, 0xFF0000, 0xFF00, 0xFF, 0)
    else:
        raise ValueError(dds_format)
    info = FO.parse_dds_ignore_magic(bytes(d))
    length = info[2]
    if payload is None:
        payload = bytes(i % 256 for i in range(length))
    assert len(payload) == length, (len(payload), length)
    return np.frombuffer(bytes(d) + payload + leftover, np.uint8).copy()
