"""GPU parity tests of the file-format layer: DDS files transformed through the C ABI
(texture data on the GPU) must equal, byte for byte, what the oracle's restatement of the reference's DdsHandler /
TransformBundle / dispatch produces, and must round-trip.  Files produced here carry the reference's header, so the
second half of each test doubles as "the stock reference could untransform this"."""
import struct

import numpy as np
import pytest

import dxt_lossless_transform_b200 as dlt
import oracle
from dxt_lossless_transform_b200 import file_formats as ff
from dxt_lossless_transform_b200 import synth
from dxt_lossless_transform_b200.file_formats import TransformFormat, TransformHeader

from dds_fixtures import FO, make_dds, real_fixture

pytestmark = pytest.mark.gpu


def o_transform(n, payload, variant, split):
    return oracle.transform(n, np.frombuffer(payload, np.uint8), variant, False, split).tobytes()


def o_untransform(n, payload, variant, split):
    return oracle.untransform(n, np.frombuffer(payload, np.uint8), variant, False, split).tobytes()


def o_auto(use_all):
    def run(n, payload):
        out, (v, _sa, sc) = oracle.auto(n, np.frombuffer(payload, np.uint8).copy(), use_all)
        return out.tobytes(), (v, sc)
    return run


def manual_bundle(variant: dlt.YCoCgVariant, split: bool) -> ff.TransformBundle:
    return (ff.TransformBundle.new()
            .with_bc1_manual(dlt.Bc1ManualTransformBuilder().decorrelation_mode(variant).split_colour_endpoints(split))
            .with_bc2_manual(dlt.Bc2ManualTransformBuilder().decorrelation_mode(variant).split_colour_endpoints(split)))


@pytest.mark.parametrize("name", ["bc1", "bc2"])
@pytest.mark.parametrize("variant", list(dlt.YCoCgVariant))
@pytest.mark.parametrize("split", [False, True])
def test_reference_fixtures_all_settings(name, variant, split):
    """assets/tests/r2-256-bc{1,2}.dds over all settings (what the reference's CLI tests round-trip)."""
    h = ff.DdsHandler()
    dds = real_fixture(name)
    out, back = np.zeros_like(dds), np.zeros_like(dds)
    h.transform_bundle(dds, out, manual_bundle(variant, split))
    want = FO.dds_transform_bundle(dds.tobytes(), dds.size, {FO.BC1: (int(variant), split), FO.BC2: (int(variant), split)}, o_transform)
    assert out.tobytes() == want
    hdr = TransformHeader.read_from(out)
    assert hdr.format() == (TransformFormat.Bc1 if name == "bc1" else TransformFormat.Bc2)
    got = hdr.to_bc1_settings() if name == "bc1" else hdr.to_bc2_settings()
    assert (got.decorrelation_mode, got.split_colour_endpoints) == (variant, split)
    h.untransform(out, back)
    assert np.array_equal(back, dds)
    assert back.tobytes() == FO.dds_untransform(want, len(want), o_untransform)


def test_bc3_fixture_is_format_not_implemented_like_the_reference():
    dds = real_fixture("bc3")
    with pytest.raises(ff.FormatNotImplemented) as e:
        ff.DdsHandler().transform_bundle(dds, np.zeros_like(dds), ff.TransformBundle.default_all())
    assert e.value.format == TransformFormat.Bc3


@pytest.mark.parametrize("fmt,w,h,mips", [(FO.DDS_BC1, 4, 4, 1), (FO.DDS_BC1, 17, 13, 1), (FO.DDS_BC1, 256, 256, 9),
                                          (FO.DDS_BC2, 4, 4, 1), (FO.DDS_BC2, 100, 36, 7), (FO.DDS_BC1, 2048, 1024, 12),
                                          (FO.DDS_BC2, 1024, 1024, 11), (FO.DDS_BC1, 1, 1, 1)])
def test_synthetic_dds_with_mips_and_leftover_data(fmt, w, h, mips):
    """Mip chains give odd block counts (unaligned streams inside the payload); trailing bytes survive verbatim
    (handler test `transform_and_untransform_preserves_leftover_data_roundtrip`)."""
    handler, bundle = ff.DdsHandler(), ff.TransformBundle.default_all()
    n = 1 if fmt == FO.DDS_BC1 else 2
    length = FO.parse_dds_ignore_magic(bytes(make_dds(fmt, w, h, mips)[:128]))[2]
    payload = synth.texture_blocks(n, length // (8 if n == 1 else 16), seed=w * 31 + h).tobytes()
    for leftover in (b"", b"Roundtrip preservation test data 123456!"):
        dds = make_dds(fmt, w, h, mips, payload=payload, leftover=leftover)
        out = np.full(dds.size + 5, 0xEE, np.uint8)  # a larger output buffer is fine; the tail stays untouched
        handler.transform_bundle(dds, out, bundle)
        want = FO.dds_transform_bundle(dds.tobytes(), out.size, {FO.BC1: (1, True), FO.BC2: (1, True)}, o_transform)
        assert out[:dds.size].tobytes() == want and bytes(out[dds.size:]) == b"\xEE" * 5
        back = np.zeros(dds.size, np.uint8)
        handler.untransform(out[:dds.size].copy(), back)
        assert np.array_equal(back, dds)


def test_unaligned_host_buffers():
    """The reference accepts any pointer alignment; misalign the whole file by 1 and 3 bytes."""
    handler, bundle = ff.DdsHandler(), ff.TransformBundle.default_all()
    dds = real_fixture("bc1")
    want = FO.dds_transform_bundle(dds.tobytes(), dds.size, {FO.BC1: (1, True)}, o_transform)
    for shift in (1, 3):
        src = np.zeros(dds.size + shift, np.uint8)
        src[shift:] = dds
        dst = np.zeros(dds.size + shift, np.uint8)
        handler.transform_bundle(src[shift:], dst[shift:], bundle)
        assert dst[shift:].tobytes() == want
        back = np.zeros(dds.size + shift, np.uint8)
        handler.untransform(dst[shift:], back[shift:])
        assert np.array_equal(back[shift:], dds)


@pytest.mark.parametrize("use_all", [False, True])
def test_auto_builder_bundle_embeds_the_chosen_settings(use_all):
    """Bundle with Bc{1,2}AutoTransformBuilder + the LTU estimator: the whole search runs on the GPU; the header
    records the winner (bundle/bc1.rs:57-70) and the bytes equal the oracle's search."""
    est = dlt.LosslessTransformUtilsSizeEstimation()
    bundle = (ff.TransformBundle.new()
              .with_bc1_auto(dlt.Bc1AutoTransformBuilder(est).use_all_decorrelation_modes(use_all))
              .with_bc2_auto(dlt.Bc2AutoTransformBuilder(est).use_all_decorrelation_modes(use_all)))
    handler = ff.DdsHandler()
    for name in ("bc1", "bc2"):
        dds = real_fixture(name)
        out, back = np.zeros_like(dds), np.zeros_like(dds)
        handler.transform_bundle(dds, out, bundle)
        want = FO.dds_transform_bundle(dds.tobytes(), dds.size, {FO.BC1: o_auto(use_all), FO.BC2: o_auto(use_all)}, o_transform)
        assert out.tobytes() == want
        handler.untransform(out, back)
        assert np.array_equal(back, dds)


def test_dispatch_functions_on_raw_texture_data():
    bundle = ff.TransformBundle.default_all()
    for n, fmt in ((1, TransformFormat.Bc1), (2, TransformFormat.Bc2)):
        data = synth.texture_blocks(n, 5463, seed=n)  # a 256x256 mip chain's block count: odd
        out, back = np.zeros_like(data), np.zeros_like(data)
        hdr = ff.dispatch_transform(fmt, data, out, bundle)
        assert hdr.value == FO.header_new(int(fmt), 0b00100)
        assert np.array_equal(out, oracle.transform(n, data, 1, False, True))
        ff.dispatch_untransform(hdr, out, back)
        assert np.array_equal(back, data)
        # a header written by "someone else" (every settings combination) untransforms to the oracle's answer
        for variant in range(4):
            for split in (False, True):
                t = oracle.transform(n, data, variant, False, split)
                ff.dispatch_untransform(TransformHeader(FO.header_new(int(fmt), FO.pack_bc12(variant, split))), t, back)
                assert np.array_equal(back, data)
    # empty payloads are legal
    e = np.zeros(0, np.uint8)
    assert ff.dispatch_transform(TransformFormat.Bc1, e, e.copy(), bundle).value == FO.header_new(0, 0b00100)


def test_batch_equals_single_file_calls():
    """A directory's worth of mixed files (BC1/BC2, mips, leftovers, bad files) through the batch entry points."""
    handler, bundle = ff.DdsHandler(), manual_bundle(dlt.YCoCgVariant.Variant2, False)
    rng = np.random.default_rng(5)
    files = []
    for i in range(24):
        fmt = FO.DDS_BC1 if i % 3 else FO.DDS_BC2
        w, h, mips = int(rng.choice([4, 20, 64, 256, 512])), int(rng.choice([4, 12, 128, 256])), int(rng.choice([1, 1, 3, 6]))
        n = 1 if fmt == FO.DDS_BC1 else 2
        length = FO.parse_dds_ignore_magic(bytes(make_dds(fmt, w, h, mips)[:128]))[2]
        payload = synth.texture_blocks(n, length // (8 if n == 1 else 16), seed=100 + i).tobytes()
        files.append(make_dds(fmt, w, h, mips, payload=payload, leftover=bytes(range(i))))
    files.insert(3, np.zeros(128, np.uint8))                    # not a DDS
    files.insert(9, real_fixture("bc3"))                        # FormatNotImplemented
    files.insert(15, make_dds(FO.DDS_BC1, 64, 64)[:-7].copy())  # truncated
    outs = [np.zeros_like(f) for f in files]
    res = handler.transform_bundle_batch(list(zip(files, outs)), bundle)
    good = []
    for i, (f, o, r) in enumerate(zip(files, outs, res)):
        single = np.zeros_like(f)
        try:
            handler.transform_bundle(f, single, bundle)
        except ff.TransformError as e:
            assert type(r) is type(e), (i, r, e)
            continue
        assert r is None, (i, r)
        assert np.array_equal(o, single), i
        assert o.tobytes() == FO.dds_transform_bundle(f.tobytes(), f.size, {FO.BC1: (2, False), FO.BC2: (2, False)}, o_transform)
        good.append(i)
    assert len(good) == 24
    backs = [np.zeros_like(f) for f in files]
    res = handler.untransform_batch([(outs[i], backs[i]) for i in good])
    assert res == [None] * len(good)
    for i in good:
        assert np.array_equal(backs[i], files[i]), i


def test_batch_over_every_visible_gpu():
    import torch

    devices = list(range(torch.cuda.device_count()))
    handler, bundle = ff.DdsHandler(), ff.TransformBundle.default_all()
    files = [make_dds(FO.DDS_BC1, 512, 512, 10, payload=None) for _ in range(6)] + [real_fixture("bc2")] * 3
    outs = [np.zeros_like(f) for f in files]
    assert handler.transform_bundle_batch(list(zip(files, outs)), bundle, devices=devices) == [None] * len(files)
    backs = [np.zeros_like(f) for f in files]
    assert handler.untransform_batch(list(zip(outs, backs)), devices=devices) == [None] * len(files)
    for f, o, b in zip(files, outs, backs):
        assert np.array_equal(b, f)
        assert o.tobytes() == FO.dds_transform_bundle(f.tobytes(), f.size, {FO.BC1: (1, True), FO.BC2: (1, True)}, o_transform)
