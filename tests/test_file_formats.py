"""CPU tests of the file-format layer: the TransformHeader bit layout, the DDS parser
and every validation path of the bundle / dispatch / DdsHandler calls that returns before any CUDA work.

Cases follow the reference's own tests (api/dxt-lossless-transform-file-formats-api/src/embed/mod.rs tests,
embed/formats/bc1.rs tests, handlers/dispatch.rs tests; extensions/file-formats/dxt-lossless-transform-dds/src/
dds/parse_dds.rs tests, dds/likely_dds.rs tests, handler/*.rs tests); the product (C ABI) is additionally
compared against the Python oracle on randomized headers."""
import struct

import numpy as np
import pytest

import dxt_lossless_transform_b200 as dlt
from dxt_lossless_transform_b200 import file_formats as ff
from dxt_lossless_transform_b200.file_formats import DdsFormat, TransformFormat, TransformHeader

from dds_fixtures import FO, KNOWN, make_dds, real_fixture


# ---- embed/mod.rs tests ----------------------------------------------------------------------------
def test_transform_format_values():
    want = dict(Bc1=0, Bc2=1, Bc3=2, Bc7=3, Bc6H=4, Rgba8888=5, Bgra8888=6, Bgr888=7, Bc4=8, Bc5=9)
    assert {f.name: int(f) for f in TransformFormat} == want
    for raw in range(16):
        h = TransformHeader(raw)
        assert h.format() == (TransformFormat(raw) if raw <= 9 else None)
        assert FO.header_format(raw) == (raw if raw <= 9 else None)


def test_transform_header_bitfield():
    h = TransformHeader.new(TransformFormat.Bc1, 0x0ABCDEF0)
    assert h.format() == TransformFormat.Bc1 and h.format_data() == 0x0ABCDEF0
    h2 = TransformHeader.new(TransformFormat.Bc3, 0xFFFFFFFF)  # data is masked to 28 bits
    assert h2.format() == TransformFormat.Bc3 and h2.format_data() == 0x0FFFFFFF
    assert h.value == FO.header_new(0, 0x0ABCDEF0) and h2.value == FO.header_new(2, 0xFFFFFFFF)


def test_header_little_endian_byte_order():
    h = TransformHeader.new(TransformFormat.Bc7, 0x1234567)
    assert h.value == 0x12345673
    buf = np.zeros(5, np.uint8)
    h.write_to(buf[1:])  # unaligned on purpose
    assert bytes(buf[1:]) == bytes([0x73, 0x56, 0x34, 0x12])
    assert TransformHeader.read_from(buf[1:]) == h
    with pytest.raises(ff.InputTooShort):
        TransformHeader.read_from(np.zeros(3, np.uint8))


# ---- embed/formats/bc1.rs + bc2.rs tests ----------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 2])
def test_details_roundtrip_all_settings_and_bit_layout(n):
    cls = dlt.Bc1TransformSettings if n == 1 else dlt.Bc2TransformSettings
    make = TransformHeader.from_bc1_settings if n == 1 else TransformHeader.from_bc2_settings
    for s in cls.all_combinations():
        h = make(s)
        assert h.format() == (TransformFormat.Bc1 if n == 1 else TransformFormat.Bc2)
        back = h.to_bc1_settings() if n == 1 else h.to_bc2_settings()
        assert back == s
        data = h.format_data()
        assert data & 3 == 0                                      # header version
        assert bool((data >> 2) & 1) == s.split_colour_endpoints   # bit 2
        assert (data >> 3) & 3 == s.decorrelation_mode.to_stable() # bits 3-4: Variant1=0 .. None=3
        assert data >> 5 == 0                                      # reserved
        assert h.value == FO.header_new(n - 1, FO.pack_bc12(int(s.decorrelation_mode), s.split_colour_endpoints))
    # default settings (Variant1, split): version 0, split 1, variant 0
    assert make(cls()).format_data() == 0b00100


def test_details_reject_bad_version_and_wrong_format():
    for version in (1, 2, 3):
        with pytest.raises(ff.CorruptedEmbeddedData):
            TransformHeader.new(TransformFormat.Bc1, version).to_bc1_settings()
        with pytest.raises(ff.CorruptedEmbeddedData):
            TransformHeader.new(TransformFormat.Bc2, version | 0b11100).to_bc2_settings()
    with pytest.raises(ff.UnknownFormat):
        TransformHeader.new(TransformFormat.Bc2, 0).to_bc1_settings()
    with pytest.raises(ff.UnknownFormat):
        TransformHeader(0xF).to_bc2_settings()
    # reserved bits are ignored on the way in
    assert TransformHeader.new(TransformFormat.Bc1, 0x0FFFFFE4).to_bc1_settings() == dlt.Bc1TransformSettings(
        dlt.YCoCgVariant.Variant1, True)


# ---- dds/likely_dds.rs + parse_dds.rs tests -----------------------------------------------------------
def test_likely_dds():
    magic = bytes([0x44, 0x44, 0x53, 0x20])
    assert ff.is_dds(np.frombuffer(magic + bytes(124), np.uint8))
    assert not ff.is_dds(np.frombuffer(magic + bytes(123), np.uint8))
    assert not ff.is_dds(np.zeros(128, np.uint8))
    assert not ff.is_dds(np.frombuffer(magic, np.uint8))
    assert not ff.is_dds(np.zeros(0, np.uint8))


@pytest.mark.parametrize("fourcc,want", [(b"DXT1", DdsFormat.BC1), (b"DXT2", DdsFormat.BC2), (b"DXT3", DdsFormat.BC2),
                                         (b"DXT4", DdsFormat.BC3), (b"DXT5", DdsFormat.BC3), (b"BC4U", DdsFormat.BC4),
                                         (b"BC4S", DdsFormat.BC4), (b"ATI1", DdsFormat.BC4), (b"BC5U", DdsFormat.BC5),
                                         (b"BC5S", DdsFormat.BC5), (b"ATI2", DdsFormat.BC5), (b"UNKN", DdsFormat.Unknown)])
def test_parse_dds_legacy_formats(fourcc, want):
    data = make_dds(FO.DDS_BC1, 4, 4)
    data[0x54:0x58] = np.frombuffer(fourcc, np.uint8)
    info = ff.parse_dds(data)
    assert info.format == want and info.data_offset == 128
    # transformed file: magic overwritten
    data[0:4] = np.frombuffer(struct.pack("<I", 0xDEADBEEF), np.uint8)
    assert ff.parse_dds(data) is None
    info = ff.parse_dds_ignore_magic(data)
    assert info.format == want and info.data_offset == 128


@pytest.mark.parametrize("dxgi,want", [(70, DdsFormat.BC1), (71, DdsFormat.BC1), (72, DdsFormat.BC1), (73, DdsFormat.BC2),
                                       (74, DdsFormat.BC2), (75, DdsFormat.BC2), (76, DdsFormat.BC3), (77, DdsFormat.BC3),
                                       (78, DdsFormat.BC3), (79, DdsFormat.BC4), (80, DdsFormat.BC4), (81, DdsFormat.BC4),
                                       (82, DdsFormat.BC5), (83, DdsFormat.BC5), (84, DdsFormat.BC5), (94, DdsFormat.BC6H),
                                       (95, DdsFormat.BC6H), (96, DdsFormat.BC6H), (97, DdsFormat.BC7), (98, DdsFormat.BC7),
                                       (99, DdsFormat.BC7), (27, DdsFormat.RGBA8888), (28, DdsFormat.RGBA8888),
                                       (32, DdsFormat.RGBA8888), (87, DdsFormat.BGRA8888), (90, DdsFormat.BGRA8888),
                                       (91, DdsFormat.BGRA8888), (0x12345678, DdsFormat.Unknown), (33, DdsFormat.Unknown),
                                       (88, DdsFormat.Unknown)])
def test_parse_dds_dx10_formats(dxgi, want):
    data = make_dds(FO.DDS_BC7, 4, 4)
    data[0x80:0x84] = np.frombuffer(struct.pack("<I", dxgi), np.uint8)
    info = ff.parse_dds(data[:148])
    assert info.format == want and info.data_offset == 148


def test_parse_dds_too_short():
    assert ff.parse_dds(np.zeros(127, np.uint8)) is None
    d = np.zeros(147, np.uint8)
    d[0:4] = np.frombuffer(struct.pack("<I", FO.DDS_MAGIC), np.uint8)
    d[0x54:0x58] = np.frombuffer(b"DX10", np.uint8)
    assert ff.parse_dds(d) is None
    assert ff.parse_dds(np.zeros(0, np.uint8)) is None


def test_parse_dds_unaligned_buffer():
    valid = make_dds(FO.DDS_BC1, 64, 64)
    buf = np.zeros(valid.size + 1, np.uint8)
    buf[1:] = valid
    info = ff.parse_dds(buf[1:])
    assert info.format == DdsFormat.BC1 and info.data_offset == 128 and info.data_length == 2048


@pytest.mark.parametrize("fmt,w,h,mips,want", [
    (FO.DDS_BC1, 256, 256, 1, 32768), (FO.DDS_BC1, 256, 256, 9, 43704), (FO.DDS_BC1, 17, 13, 1, 160),
    (FO.DDS_RGBA8888, 16, 16, 1, 1024), (FO.DDS_RGBA8888, 4, 4, 3, 84), (FO.DDS_BGR888, 5, 3, 2, 45 + 6),
    (FO.DDS_BC2, 256, 256, 1, 65536), (FO.DDS_BC3, 8, 8, 4, 64 + 16 + 16 + 16)])
def test_data_length_known_answers(fmt, w, h, mips, want):
    data = make_dds(fmt, w, h, mips)
    assert ff.parse_dds(data).data_length == want
    assert data.size == 128 + want


def test_data_length_zero_dimensions_and_forged_mip_counts():
    data = make_dds(FO.DDS_BC1, 256, 256)
    data[0x10:0x14] = 0  # width = 0
    assert ff.parse_dds(data).data_length == 0
    # a header that claims 2^32-1 mip levels: saturates instead of looping for minutes
    data = make_dds(FO.DDS_BC1, 4, 4)
    struct.pack_into("<I", data, 0x08, struct.unpack_from("<I", data, 0x08)[0] | 0x20000)
    struct.pack_into("<I", data, 0x1C, 0xFFFFFFFF)
    assert ff.parse_dds(data).data_length == 0xFFFFFFFF == FO.parse_dds(bytes(data))[2]
    struct.pack_into("<I", data, 0x1C, 1000)
    assert ff.parse_dds(data).data_length == 8 * 1000


def test_reference_fixture_headers():
    for name, (fmt, off, length) in dict(bc1=(DdsFormat.BC1, 128, 32768), bc2=(DdsFormat.BC2, 128, 65536),
                                         bc3=(DdsFormat.BC3, 128, 65536), bc7=(DdsFormat.BC7, 148, 65536)).items():
        hdr = np.frombuffer(bytes.fromhex(KNOWN["dds_fixtures"][name]["header"]), np.uint8)
        info = ff.parse_dds(hdr)
        assert (info.format, info.data_offset, info.data_length) == (fmt, off, length)
        assert off + length == KNOWN["dds_fixtures"][name]["file_len"]


def test_parser_matches_oracle_on_random_headers():
    rng = np.random.default_rng(1234)
    fourccs = [b"DXT1", b"DXT2", b"DXT3", b"DXT4", b"DXT5", b"BC4U", b"BC4S", b"ATI1", b"BC5U", b"BC5S", b"ATI2", b"DX10",
               b"\0\0\0\0", b"ABCD"]
    masks = [0, 0xFF, 0xFF00, 0xFF0000, 0xFF000000, 0x12345678]
    for it in range(4000):
        n = int(rng.choice([100, 127, 128, 140, 147, 148, 200]))
        d = bytearray(rng.integers(0, 256, n, dtype=np.uint8).tobytes()) if it % 4 == 0 else bytearray(n)
        def put(off, v):
            if off + 4 <= n:
                struct.pack_into("<I", d, off, v & 0xFFFFFFFF)
        if it % 5:
            put(0, FO.DDS_MAGIC)
        put(0x08, int(rng.choice([0, 0x20000, 0x1007, 0x2100F, 0xFFFFFFFF])))
        put(0x0C, int(rng.choice([0, 1, 3, 4, 17, 256, 4096, 70000, 0xFFFFFFFF])))
        put(0x10, int(rng.choice([0, 1, 5, 16, 255, 1024, 65536, 0xFFFFFFFF])))
        put(0x1C, int(rng.choice([0, 1, 2, 9, 13, 40, 0xFFFFFFFF])))
        put(0x50, int(rng.choice([0, 0x4, 0x40, 0x41, 0x2, 0x200, 0x20000, 0x45])))
        if 0x58 <= n:
            d[0x54:0x58] = fourccs[int(rng.integers(len(fourccs)))]
        put(0x58, int(rng.choice([0, 8, 12, 16, 24, 32, 64])))
        if it % 3 == 0:
            for off, m in zip((0x5C, 0x60, 0x64, 0x68), [(0xFF0000, 0xFF00, 0xFF, 0), (0xFF, 0xFF00, 0xFF0000, 0xFF000000),
                                                         (0xFF0000, 0xFF00, 0xFF, 0xFF000000)][it % 9 // 3]):
                put(off, m)
        else:
            for off in (0x5C, 0x60, 0x64, 0x68):
                put(off, masks[int(rng.integers(len(masks)))])
        put(0x80, int(rng.choice([0, 27, 32, 33, 70, 75, 78, 81, 84, 87, 90, 91, 93, 95, 99, 100])))
        a = np.frombuffer(bytes(d), np.uint8)
        for mine, theirs in ((ff.parse_dds, FO.parse_dds), (ff.parse_dds_ignore_magic, FO.parse_dds_ignore_magic)):
            got, want = mine(a), theirs(bytes(d))
            got_t = None if got is None else (int(got.format), got.data_offset, got.data_length)
            # the C ABI folds "None" and NotADds together
            assert got_t == want, (it, bytes(d).hex(), got_t, want)
        h = ff.DdsHandler()
        assert h.can_handle(a, None) == FO.can_handle(bytes(d), None)
        assert h.can_handle_untransform(a, "dds") == FO.can_handle_untransform(bytes(d), "dds")


# ---- handler/file_format_detection.rs + file_format_untransform_detection.rs tests -------------------------
def test_can_handle():
    h = ff.DdsHandler()
    valid = make_dds(FO.DDS_BC1, 4, 4)
    assert h.can_handle(valid, "dds") and h.can_handle(valid, None)
    assert not h.can_handle(valid, "txt") and not h.can_handle(valid, "DDS")
    assert not h.can_handle(np.zeros(128, np.uint8), "dds")
    assert not h.can_handle(valid[:127], "dds")
    t = valid.copy()
    t[0:4] = [0xAB, 0xCD, 0xEF, 0x12]
    assert h.can_handle_untransform(t, "dds") and h.can_handle_untransform(t, None)
    assert not h.can_handle_untransform(t, "txt")
    assert not h.can_handle_untransform(np.zeros(127, np.uint8), "dds")
    assert not h.can_handle_untransform(np.zeros(3, np.uint8), None)


# ---- handler/file_format_handler.rs tests: every error that is decided before the GPU is touched --------
def test_transform_bundle_validation_errors():
    h, bundle = ff.DdsHandler(), ff.TransformBundle.default_all()
    incomplete = make_dds(FO.DDS_BC1, 4, 4)[:128].copy()
    with pytest.raises(ff.OutputBufferTooSmall) as e:
        h.transform_bundle(incomplete, np.zeros(127, np.uint8), bundle)
    assert (e.value.required, e.value.actual) == (128, 127)
    with pytest.raises(ff.InvalidInputFileHeader):
        h.transform_bundle(np.zeros(128, np.uint8), np.zeros(128, np.uint8), bundle)
    # no builder for the detected format
    inp = make_dds(FO.DDS_BC1, 64, 64)
    with pytest.raises(ff.NoBuilderForFormat) as e:
        h.transform_bundle(inp, np.zeros_like(inp), ff.TransformBundle.new())
    assert e.value.format == TransformFormat.Bc1
    inp2 = make_dds(FO.DDS_BC2, 4, 4)
    with pytest.raises(ff.NoBuilderForFormat) as e:
        h.transform_bundle(inp2, np.zeros_like(inp2), ff.TransformBundle.new().with_bc1_manual(dlt.Bc1ManualTransformBuilder()))
    assert e.value.format == TransformFormat.Bc2
    # known but unimplemented formats, exactly as in the reference
    for dds_fmt, tf in ((FO.DDS_BC3, TransformFormat.Bc3), (FO.DDS_BC6H, TransformFormat.Bc6H), (FO.DDS_BC7, TransformFormat.Bc7),
                        (FO.DDS_BC4, TransformFormat.Bc4), (FO.DDS_BC5, TransformFormat.Bc5)):
        x = make_dds(dds_fmt, 4, 4)
        with pytest.raises(ff.FormatNotImplemented) as e:
            h.transform_bundle(x, np.zeros_like(x), bundle)
        assert e.value.format == tf
    x = make_dds(FO.DDS_UNKNOWN, 1, 1)
    with pytest.raises(ff.UnknownFileFormat):
        h.transform_bundle(x, np.zeros_like(x), bundle)
    # uncompressed formats pass the handler's conversion and are refused by the bundle
    for dds_fmt in (FO.DDS_RGBA8888, FO.DDS_BGRA8888, FO.DDS_BGR888):
        x = make_dds(dds_fmt, 4, 4)
        with pytest.raises(ff.UnknownTransformFormat):
            h.transform_bundle(x, np.zeros_like(x), bundle)
    # declared texture larger than the file
    x = make_dds(FO.DDS_BC1, 64, 64)[:-100].copy()
    with pytest.raises(ff.InputTooShortForStatedTextureSize) as e:
        h.transform_bundle(x, np.zeros_like(x), bundle)
    assert (e.value.required, e.value.actual) == (128 + 2048, 128 + 2048 - 100)


def test_untransform_validation_errors():
    h = ff.DdsHandler()
    with pytest.raises(ff.InputTooShort) as e:
        h.untransform(np.zeros(3, np.uint8), np.zeros(128, np.uint8))
    assert (e.value.required, e.value.actual) == (4, 3)
    with pytest.raises(ff.OutputBufferTooSmall) as e:
        h.untransform(np.zeros(128, np.uint8), np.zeros(127, np.uint8))
    assert (e.value.required, e.value.actual) == (128, 127)
    with pytest.raises(ff.InvalidRestoredFileHeader):
        h.untransform(np.zeros(64, np.uint8), np.zeros(128, np.uint8))
    x = make_dds(FO.DDS_BC1, 64, 64)
    x[0:4] = 0xFF
    x = x[:-50].copy()
    with pytest.raises(ff.InputTooShortForStatedTextureSize):
        h.untransform(x, np.zeros_like(x))
    # corrupted structure: all-zero header with 0xFFFFFFFF in place of the magic -> format 15 is unknown
    c = np.zeros(128, np.uint8)
    c[0:4] = 0xFF
    with pytest.raises(ff.TransformError):
        h.untransform(c, np.zeros(128, np.uint8))
    # a BC1 file whose embedded header carries a bad version
    x = make_dds(FO.DDS_BC1, 4, 4)
    x[0:4] = np.frombuffer(struct.pack("<I", FO.header_new(FO.BC1, 1)), np.uint8)
    with pytest.raises(ff.CorruptedEmbeddedData):
        h.untransform(x, np.zeros_like(x))
    # a header that names a format the dispatcher does not untransform
    x[0:4] = np.frombuffer(struct.pack("<I", FO.header_new(FO.BC3, 0)), np.uint8)
    with pytest.raises(ff.UnknownTransformFormat):
        h.untransform(x, np.zeros_like(x))


def test_dispatch_validation_errors():
    bundle = ff.TransformBundle.default_all()
    hdr = TransformHeader.from_bc1_settings(dlt.Bc1TransformSettings())
    with pytest.raises(ff.InvalidDataAlignment) as e:  # handlers/dispatch.rs test: 15 bytes of BC1
        ff.dispatch_untransform(hdr, np.zeros(15, np.uint8), np.zeros(15, np.uint8))
    assert (e.value.size, e.value.required_divisor) == (15, 8)
    with pytest.raises(ff.InvalidDataAlignment) as e:
        ff.dispatch_untransform(TransformHeader.from_bc2_settings(dlt.Bc2TransformSettings()), np.zeros(24, np.uint8), np.zeros(24, np.uint8))
    assert (e.value.size, e.value.required_divisor) == (24, 16)
    with pytest.raises(ff.OutputBufferTooSmall):
        ff.dispatch_untransform(hdr, np.zeros(16, np.uint8), np.zeros(8, np.uint8))
    with pytest.raises(ff.CorruptedEmbeddedData):
        ff.dispatch_untransform(TransformHeader.new(TransformFormat.Bc1, 2), np.zeros(16, np.uint8), np.zeros(16, np.uint8))
    for f in (TransformFormat.Bc3, TransformFormat.Bc7, TransformFormat.Rgba8888):
        with pytest.raises(ff.UnknownTransformFormat):
            ff.dispatch_untransform(TransformHeader.new(f, 0), np.zeros(16, np.uint8), np.zeros(16, np.uint8))
        with pytest.raises(ff.UnknownTransformFormat):
            ff.dispatch_transform(f, np.zeros(16, np.uint8), np.zeros(16, np.uint8), bundle)
    with pytest.raises(ff.UnknownTransformFormat):
        ff.dispatch_untransform(TransformHeader(0xF), np.zeros(16, np.uint8), np.zeros(16, np.uint8))
    with pytest.raises(ff.OutputBufferTooSmall):
        ff.dispatch_transform(TransformFormat.Bc1, np.zeros(16, np.uint8), np.zeros(8, np.uint8), bundle)
    with pytest.raises(ff.NoBuilderForFormat):
        ff.dispatch_transform(TransformFormat.Bc2, np.zeros(16, np.uint8), np.zeros(16, np.uint8), ff.TransformBundle.new())
    with pytest.raises(ff.Bc1TransformError) as e:  # Bc1Error::InvalidLength(12)
        ff.dispatch_transform(TransformFormat.Bc1, np.zeros(12, np.uint8), np.zeros(12, np.uint8), bundle)
    assert (e.value.inner_code, e.value.payload) == (1, 12)
    # a bundle refuses a builder of the other format
    with pytest.raises(TypeError):
        ff.TransformBundle.new().with_bc1_manual(dlt.Bc2ManualTransformBuilder())


def test_batch_reports_per_file_validation_errors_without_a_gpu():
    h, bundle = ff.DdsHandler(), ff.TransformBundle.default_all()
    bad = [np.zeros(128, np.uint8), make_dds(FO.DDS_BC3, 4, 4), make_dds(FO.DDS_BC1, 64, 64)[:-100].copy(), make_dds(FO.DDS_RGBA8888, 4, 4)]
    res = h.transform_bundle_batch([(b, np.zeros_like(b)) for b in bad], bundle)
    assert [type(r) for r in res] == [ff.InvalidInputFileHeader, ff.FormatNotImplemented, ff.InputTooShortForStatedTextureSize,
                                      ff.UnknownTransformFormat]
    res = h.untransform_batch([(np.zeros(3, np.uint8), np.zeros(3, np.uint8)), (np.zeros(64, np.uint8), np.zeros(64, np.uint8))])
    assert [type(r) for r in res] == [ff.InputTooShort, ff.InvalidRestoredFileHeader]
    assert h.transform_bundle_batch([], bundle) == []


# ---- api.rs tests (mock handlers, as in the reference) -----------------------------------------------------
class MockHandler:
    def __init__(self, accept_ext=None, accept=True):
        self.accept_ext, self.accept = accept_ext, accept
        self.can_handle_calls, self.transform_called, self.untransform_called = [], False, False

    def can_handle(self, inp, ext):
        self.can_handle_calls.append(ext)
        return self.accept and (self.accept_ext is None or ext == self.accept_ext)

    can_handle_untransform = can_handle

    def transform_bundle(self, i, o, b):
        self.transform_called = True

    def untransform(self, i, o):
        self.untransform_called = True


def test_slice_api_over_handlers():
    bundle = ff.TransformBundle.default_all()
    inp, out = np.arange(64, dtype=np.uint8), np.zeros(64, np.uint8)
    h = MockHandler()
    ff.transform_slice_with_bundle(h, inp, out, bundle)
    ff.untransform_slice(h, inp, out)
    assert h.transform_called and h.untransform_called
    with pytest.raises(ff.OutputBufferTooSmall):
        ff.transform_slice_with_bundle(h, inp, np.zeros(63, np.uint8), bundle)
    # multiple handlers: extension is None for slices; a handler that insists on "dds" rejects
    picky = MockHandler(accept_ext="dds")
    with pytest.raises(ff.NoSupportedHandler):
        ff.transform_slice_with_multiple_handlers([picky], inp, out, bundle)
    assert picky.can_handle_calls == [None]
    h1, h2 = MockHandler(accept=False), MockHandler()
    assert ff.transform_slice_with_multiple_handlers([h1, h2], inp, out, bundle) is h2
    assert len(h1.can_handle_calls) == 1 and not h1.transform_called and h2.transform_called
    h1, h2 = MockHandler(accept=False), MockHandler()
    assert ff.untransform_slice_with_multiple_handlers([h1, h2], inp, out) is h2 and h2.untransform_called
    with pytest.raises(ff.NoSupportedHandler):
        ff.untransform_slice_with_multiple_handlers([MockHandler(accept=False)], inp, out)
    with pytest.raises(ff.OutputBufferTooSmall):
        ff.untransform_slice_with_multiple_handlers([h2], inp, np.zeros(1, np.uint8))


def test_oracle_handler_agrees_with_reference_fixture_layout():
    """The oracle's DDS transform of the real BC1 fixture: header rewritten, payload = C-oracle transform."""
    import oracle

    def tr(n, payload, variant, split):
        return oracle.transform(n, np.frombuffer(payload, np.uint8), variant, False, split).tobytes()

    def un(n, payload, variant, split):
        return oracle.untransform(n, np.frombuffer(payload, np.uint8), variant, False, split).tobytes()

    for name, key in (("bc1", FO.BC1), ("bc2", FO.BC2)):
        dds = real_fixture(name).tobytes()
        out = FO.dds_transform_bundle(dds, len(dds), {FO.BC1: (1, True), FO.BC2: (1, True)}, tr)
        assert out[4:128] == dds[4:128] and len(out) == len(dds)
        assert struct.unpack_from("<I", out, 0)[0] == FO.header_new(key, 0b00100)
        assert FO.dds_untransform(out, len(out), un) == dds
    with pytest.raises(FO.OracleError) as e:
        FO.dds_transform_bundle(real_fixture("bc3").tobytes(), 1 << 20, {}, tr)
    assert e.value.variant == "FormatHandler::FormatNotImplemented"
