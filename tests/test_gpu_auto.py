"""GPU tests of the LTU-semantics estimator and the best-settings search against the oracle.

The estimator restates lossless-transform-utils 0.1.3 (not in the reference tree): these tests prove
GPU == oracle restatement bit for bit; parity with the real crate is UNPINNED."""
import ctypes as C
import zlib
from pathlib import Path

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def dlt():
    import dxt_lossless_transform_b200 as m

    return m


@pytest.fixture(scope="module")
def torch():
    import torch as t

    return t


def payload(fmt):
    return np.frombuffer(zlib.decompress((GOLDEN / f"r2-256-bc{fmt}.payload.zlib").read_bytes()), np.uint8).copy()


def estimator_inputs():
    rng = np.random.default_rng(42)
    yield "zeros64", np.zeros(64, np.uint8)
    for n in range(0, 41):
        yield f"zeros{n}", np.zeros(n, np.uint8)
        yield f"rand{n}", rng.integers(0, 4, n, dtype=np.uint8)
    yield "period4", np.tile(np.array([1, 2, 3, 4], np.uint8), 64)
    yield "period3", np.tile(np.array([9, 8, 7], np.uint8), 1000)
    yield "two_symbols", rng.integers(0, 2, 50_000, dtype=np.uint8)
    yield "low_entropy", rng.integers(0, 8, 300_001, dtype=np.uint8)
    yield "random", rng.integers(0, 256, 1 << 20, dtype=np.uint8)
    runs = np.repeat(rng.integers(0, 256, 4000, dtype=np.uint8), rng.integers(1, 40, 4000))
    yield "runs", runs
    # many keys colliding in one bucket: keys k * 2^16 have equal low product bits patterns
    yield "bc1_payload_colours", oracle.transform(1, payload(1), 1, False, True)[: 4096 * 4]
    yield "bc3_payload_alpha", oracle.transform(3, payload(3), 0, True, True)[: 4096 * 2]


def test_estimator_equals_oracle_on_host_buffers(dlt):
    est = dlt.LosslessTransformUtilsSizeEstimation()
    assert est.max_compressed_size(12345) == 0
    for name, data in estimator_inputs():
        assert est.estimate_compressed_size(data) == oracle.ltu_estimate(data), name


def test_estimator_reference_inequalities(dlt):
    # extensions/estimators/dxt-lossless-transform-ltu/src/lib.rs:125-225
    est = dlt.LosslessTransformUtilsSizeEstimation()
    assert est.estimate_compressed_size(np.zeros(64, np.uint8)) < 64
    rng = np.random.default_rng(0)
    rnd = rng.integers(0, 256, 4096, dtype=np.uint8)
    rep = np.tile(np.arange(16, dtype=np.uint8), 256)
    a, b = est.estimate_compressed_size(rep), est.estimate_compressed_size(rnd)
    assert a <= b <= rnd.size
    assert est.estimate_compressed_size(rnd) == b


def test_estimator_device_pointer_any_alignment(dlt, torch):
    rng = np.random.default_rng(5)
    data = rng.integers(0, 6, 200_000, dtype=np.uint8)
    d = torch.from_numpy(data).cuda()
    for off in (0, 1, 2, 3, 7):
        assert dlt.ltu_estimate_device(d.data_ptr() + off, data.size - off) == oracle.ltu_estimate(data[off:])


def adversarial_streams():
    """Inputs aimed at the worst cases of any parallel formulation of the sequential table: everything in one bucket, rare
    outliers in a flat stream (long look-backs in the hand-over between chunks), same-group bucket collisions across
    window / batch / chunk boundaries, short periods, low-entropy streams of many sizes."""
    rng = np.random.default_rng(77)
    yield "zeros_1M", np.zeros(1 << 20, np.uint8)
    yield "ff_3M+5", np.full((3 << 20) + 5, 0xFF, np.uint8)
    flat = np.full(6_000_001, 0x55, np.uint8)
    flat[::100_003] = 0xAA
    flat[1_234_567:1_234_600] = rng.integers(0, 256, 33, dtype=np.uint8)
    yield "flat_with_outliers", flat
    for period in (2, 3, 4, 5, 7, 8, 12):
        yield f"period{period}", np.tile(rng.integers(0, 256, period, dtype=np.uint8), 700_001 // period + 1)
    yield "two_symbols_2M", rng.integers(0, 2, 2_000_003, dtype=np.uint8)
    yield "three_symbols_9M", rng.integers(0, 3, 9_000_000, dtype=np.uint8)
    yield "random_20M", rng.integers(0, 256, 20_000_000, dtype=np.uint8)
    runs = np.repeat(rng.integers(0, 256, 200_000, dtype=np.uint8), rng.integers(1, 60, 200_000))
    yield "runs_6M", runs
    for n in (4100, 4104, 5000, 40_000, 131_072 + 7, 1_048_576 + 11):
        yield f"low_entropy_{n}", rng.integers(0, 5, n, dtype=np.uint8)
    # a smooth "texture" stream: few distinct neighbouring keys, heavy partitions
    t = (np.cumsum(rng.integers(-1, 2, 5_000_000)) // 7 % 256).astype(np.uint8)
    yield "random_walk_5M", t


def test_estimator_large_path_adversarial(dlt, torch):
    for name, data in adversarial_streams():
        d = torch.from_numpy(data).cuda()
        assert dlt.ltu_estimate_device(d.data_ptr(), data.size) == oracle.ltu_estimate(data), name
        if data.size > 100_000:  # an unaligned view of the same stream
            assert dlt.ltu_estimate_device(d.data_ptr() + 3, data.size - 3) == oracle.ltu_estimate(data[3:]), name


def test_estimator_is_deterministic_under_repetition(dlt, torch):
    """The table machine hands batches over through flags in shared memory (no barriers on the hot path): the same stream must
    give the same count every time, and the oracle's — 40 repetitions per kind, enough chunks to fill every SM."""
    rng = np.random.default_rng(31)
    n = 24_000_001
    kinds = {"texture_like": np.repeat(rng.integers(0, 256, n // 3 + 1, dtype=np.uint8), rng.integers(1, 6, n // 3 + 1))[:n],
             "random": rng.integers(0, 256, n, dtype=np.uint8), "two": rng.integers(0, 2, n, dtype=np.uint8)}
    for name, data in kinds.items():
        d = torch.from_numpy(data).cuda()
        want = oracle.ltu_estimate(data)
        got = {dlt.ltu_estimate_device(d.data_ptr(), data.size) for _ in range(40)}
        assert got == {want}, (name, got, want)


def test_estimator_batch_row_and_ring_boundaries(dlt, torch):
    """Sizes around every boundary of the table machine (csrc/estimator.cu): a row is 32 positions, a batch 256, the ring has
    20 slots (16 in an earlier version), a chunk starts at 128 batches, a call aims at one chunk per SM; the loop visits len - 7 positions rounded up to
    the group.  Three kinds of data: almost every row contested (two symbols), almost none (random), one bucket (flat);
    several segments per call go through the batched search tests."""
    rng = np.random.default_rng(2026)
    sizes = set()
    for positions in (1, 4, 31, 32, 33, 255, 256, 257, 511, 512, 513, 15 * 256, 16 * 256, 17 * 256, 19 * 256, 20 * 256, 21 * 256, 32 * 256 + 4,
                      39 * 256, 40 * 256, 41 * 256, 60 * 256, 127 * 256, 128 * 256,
                      128 * 256 + 4, 129 * 256, 2 * 128 * 256, 2 * 128 * 256 + 260, 148 * 128 * 256 - 256, 148 * 128 * 256, 148 * 128 * 256 + 512):
        for delta in (-4, -1, 0, 3):
            if positions + 7 + delta > 7:
                sizes.add(positions + 7 + delta)
    for n in sorted(sizes):
        kinds = {"two": rng.integers(0, 2, n, dtype=np.uint8), "flat": np.full(n, 9, np.uint8)}
        if n < 2_000_000:
            kinds["random"] = rng.integers(0, 256, n, dtype=np.uint8)
        for name, data in kinds.items():
            d = torch.from_numpy(data).cuda()
            assert dlt.ltu_estimate_device(d.data_ptr(), n) == oracle.ltu_estimate(data), (name, n)
            if n > 64:   # the same stream from an odd address
                assert dlt.ltu_estimate_device(d.data_ptr() + 1, n - 1) == oracle.ltu_estimate(data[1:]), (name, n, "+1")


@pytest.mark.parametrize("fmt", [1, 2, 3])
@pytest.mark.parametrize("use_all", [False, True])
def test_auto_matches_oracle_choice_and_bytes(dlt, torch, fmt, use_all):
    from dxt_lossless_transform_b200 import synth

    cases = [payload(fmt), oracle.generate_test_data(fmt, 3000)]
    for seed, smooth in ((1, 1.0), (2, 0.2), (3, 5.0)):
        cases.append(synth.texture_blocks(fmt, 20_000 + seed, seed=seed, smooth=smooth))
    cases.append(synth.random_blocks(fmt, 5000, seed=9))
    Est = dlt.Bc1EstimateSettings
    for data in cases:
        want_out, want = oracle.auto(fmt, data, use_all)
        # host path, core ABI
        out = np.zeros_like(data)
        best = {1: dlt.transform_bc1_auto, 2: dlt.transform_bc2_auto, 3: dlt.transform_bc3_auto}[fmt](
            data, out, Est(dlt.LosslessTransformUtilsSizeEstimation(), use_all))
        got = (int(best.decorrelation_mode), bool(getattr(best, "split_alpha_endpoints", False)), bool(best.split_colour_endpoints))
        assert got == want
        assert np.array_equal(out, want_out)
        # device path + per-candidate estimates
        d_in = torch.from_numpy(data).cuda()
        d_out = torch.zeros_like(d_in)
        torch.cuda.synchronize()
        best_d, sizes = dlt.transform_auto_device(fmt, d_in.data_ptr(), d_out.data_ptr(), data.size, use_all)
        assert sizes == oracle.auto_estimates(fmt, data, use_all)
        assert best_d == best
        assert np.array_equal(d_out.cpu().numpy(), want_out)


@pytest.mark.parametrize("fmt,use_all,expect", [
    (1, False, (0, False, False)), (1, True, (2, False, False)),
    (2, False, (0, False, False)), (2, True, (2, False, False)),
    (3, False, (1, True, False)), (3, True, (2, True, False)),
])
def test_auto_with_dummy_estimator_ties_pick_first_in_order(dlt, fmt, use_all, expect):
    """DummyEstimator (core/dxt-lossless-transform-bc1/src/test_prelude.rs:44-62) returns len_bytes."""
    data = oracle.generate_test_data(fmt, 333)
    dummy = dlt.CallbackSizeEstimator(lambda a: a.size)
    out = np.zeros_like(data)
    best = {1: dlt.transform_bc1_auto, 2: dlt.transform_bc2_auto, 3: dlt.transform_bc3_auto}[fmt](
        data, out, dlt.Bc1EstimateSettings(dummy, use_all))
    got = (int(best.decorrelation_mode), bool(getattr(best, "split_alpha_endpoints", False)), bool(best.split_colour_endpoints))
    assert got == expect
    assert np.array_equal(out, oracle.transform(fmt, data, *expect))
    n = data.size // 16
    per = {1: [data.size // 2], 2: [data.size // 4], 3: [2 * n, 4 * n]}[fmt]
    k = (16 if use_all else 8) if fmt == 3 else (8 if use_all else 4)
    assert dummy.calls == per * k


def test_auto_callback_sees_the_transformed_endpoint_streams(dlt):
    """A caller-supplied estimator must see exactly the bytes the reference would show it."""
    data = payload(1)
    seen = []
    est = dlt.CallbackSizeEstimator(lambda a: (seen.append(a.copy()), oracle.ltu_estimate(a.copy()))[1])
    out = np.zeros_like(data)
    best = dlt.transform_bc1_auto(data, out, dlt.Bc1EstimateSettings(est, True))
    order = dlt.api.auto_candidates(1, True)
    for s, view in zip(order, seen):
        t = oracle.transform(1, data, int(s.decorrelation_mode), False, s.split_colour_endpoints)
        assert np.array_equal(view, t[: data.size // 2])
    want_out, want = oracle.auto(1, data, True)
    assert (int(best.decorrelation_mode), False, best.split_colour_endpoints) == want
    assert np.array_equal(out, want_out)


def test_auto_failing_estimator_reports_size_estimation_error(dlt):
    """FailingEstimator (core/dxt-lossless-transform-bc1/src/transform/mod.rs:119-138)."""
    def boom(_a):
        raise RuntimeError("nope")

    data = oracle.generate_test_data(1, 64)
    with pytest.raises(dlt.api.SizeEstimationError):
        dlt.transform_bc1_auto(data, np.zeros_like(data), dlt.Bc1EstimateSettings(dlt.CallbackSizeEstimator(boom)))
    with pytest.raises(dlt.api.BcnError) as e:
        dlt.Bc2AutoTransformBuilder(dlt.CallbackSizeEstimator(boom)).transform(
            oracle.generate_test_data(2, 8), np.zeros(128, np.uint8))
    assert e.value.code == 4  # SizeEstimationFailed
    with pytest.raises(dlt.api.SizeEstimationError):
        dlt.transform_bc3_auto(oracle.generate_test_data(3, 8), np.zeros(128, np.uint8),
                               dlt.Bc1EstimateSettings(dlt.CallbackSizeEstimator(lambda a: 1, lambda n: 1 // 0)))


def test_stable_auto_builder_returns_configured_manual_builder(dlt):
    for Auto, fmt in ((dlt.Bc1AutoTransformBuilder, 1), (dlt.Bc2AutoTransformBuilder, 2)):
        data = payload(fmt)
        for ultra in (False, True):
            b = Auto.new_ultra(dlt.LosslessTransformUtilsSizeEstimation()) if ultra else Auto(dlt.LosslessTransformUtilsSizeEstimation())
            out = np.zeros_like(data)
            manual = b.transform(data, out)
            want_out, want = oracle.auto(fmt, data, ultra)
            s = manual.get_settings()
            assert (int(s.decorrelation_mode), False, s.split_colour_endpoints) == want
            assert np.array_equal(out, want_out)
            back = np.zeros_like(data)
            manual.untransform(out, back)
            assert np.array_equal(back, data)


def test_auto_empty_input(dlt):
    best = dlt.transform_bc1_auto(np.zeros(0, np.uint8), np.zeros(0, np.uint8),
                                  dlt.Bc1EstimateSettings(dlt.LosslessTransformUtilsSizeEstimation()))
    assert (best.decorrelation_mode, best.split_colour_endpoints) == (dlt.YCoCgVariant.NONE, False)


@pytest.mark.parametrize("use_all", [False, True])
def test_auto_batch_equals_single_calls(dlt, use_all):
    """dltcuda_transform_auto_batch: a mixed batch (formats, sizes incl. empty and tiny, > 64 estimator segments so the
    launch sets are split) picks the oracle's settings and writes the oracle's bytes for every payload."""
    from dxt_lossless_transform_b200 import synth

    rng = np.random.default_rng(11)
    items, datas = [], []
    for i in range(40):
        fmt = 1 + i % 3
        nb = int(rng.choice([0, 1, 7, 300, 4096, 5463, 20_000]))
        data = synth.texture_blocks(fmt, nb, seed=i, smooth=float(rng.choice([0.2, 1.0, 5.0]))) if nb else np.zeros(0, np.uint8)
        datas.append(data)
        items.append((fmt, data, np.zeros_like(data)))
    best = dlt.transform_auto_batch(items, use_all)
    for i, ((fmt, data, out), b) in enumerate(zip(items, best)):
        want_out, want = oracle.auto(fmt, data, use_all) if data.size else (data, None)
        got = (int(b.decorrelation_mode), bool(getattr(b, "split_alpha_endpoints", False)), bool(b.split_colour_endpoints))
        if data.size:
            assert got == want, (i, fmt, data.size)
            assert np.array_equal(out, want_out), (i, fmt, data.size)
        else:  # empty payload: every estimate is 0, the first candidate in test order wins
            first = dlt.auto_candidates(fmt, use_all)[0]
            assert b == first, (i, b, first)
    with pytest.raises(dlt.InvalidLength):
        dlt.transform_auto_batch([(1, np.zeros(12, np.uint8), np.zeros(12, np.uint8))])


def test_auto_batch_rounds_overlap_without_mixing_payloads(dlt):
    """A batch large enough for several upload / search / download rounds (two device slots reused round after
    round): every payload still gets the oracle's choice and bytes."""
    from dxt_lossless_transform_b200 import synth

    items = []
    for i in range(11):
        fmt = 1 if i % 2 == 0 else 3
        nb = (4 << 20) // (8 if fmt == 1 else 16) + (3 if i % 3 == 0 else 0)   # some odd block counts
        data = synth.texture_blocks(fmt, nb, seed=100 + i, smooth=[0.2, 1.0, 5.0][i % 3])
        items.append((fmt, data, np.zeros_like(data)))
    best = dlt.transform_auto_batch(items, False)
    for i, ((fmt, data, out), b) in enumerate(zip(items, best)):
        want_out, want = oracle.auto(fmt, data, False)
        got = (int(b.decorrelation_mode), bool(getattr(b, "split_alpha_endpoints", False)), bool(b.split_colour_endpoints))
        assert got == want, (i, fmt)
        assert np.array_equal(out, want_out), (i, fmt)


def test_auto_batch_with_page_locked_payloads_uses_mapped_memory_and_still_matches(dlt):
    """Small payloads in ONE page-locked pool: inputs are gathered by a kernel over the mapped memory and the winners are
    written straight into the mapped outputs (no per-payload copies) — same choices and bytes as the oracle.  Some
    buffers sit at 8-byte offsets (they take the copy path) and some payloads are above the mapped-memory threshold."""
    from dxt_lossless_transform_b200 import synth

    rng = np.random.default_rng(5)
    pool = 40 << 20
    pin_in, pin_out = dlt.alloc_pinned(pool), dlt.alloc_pinned(pool)
    pin_out.array[:] = 0xCD
    items, spans, off = [], [], 0
    for i in range(48):
        fmt = 1 + i % 3
        nb = int(rng.choice([5, 300, 2048, 4097, 16384, 40_000]))
        nbytes = nb * (8 if fmt == 1 else 16)
        off = (off + 255) // 256 * 256 + (8 if i % 7 == 3 else 0)
        if off + nbytes + 64 > pool:
            break
        pin_in.array[off:off + nbytes] = synth.texture_blocks(fmt, nb, seed=300 + i, smooth=[0.2, 1.0, 5.0][i % 3])
        items.append((fmt, pin_in.array[off:off + nbytes], pin_out.array[off:off + nbytes]))
        spans.append((off, nbytes))
        off += nbytes + 32
    best = dlt.transform_auto_batch(items, False)
    covered = np.zeros(pool, bool)
    for (fmt, data, out), b, (o, nbytes) in zip(items, best, spans):
        covered[o:o + nbytes] = True
        want_out, want = oracle.auto(fmt, np.array(data), False)
        got = (int(b.decorrelation_mode), bool(getattr(b, "split_alpha_endpoints", False)), bool(b.split_colour_endpoints))
        assert got == want, (fmt, o, nbytes)
        assert np.array_equal(out, want_out), (fmt, o, nbytes)
    assert (pin_out.array[~covered] == 0xCD).all()
    pin_in.free(), pin_out.free()


def test_dds_batch_with_auto_bundle_shares_one_search(dlt):
    """DdsHandler.transform_bundle_batch with auto builders + the LTU estimator: per-file results equal the single-file
    calls (which equal the oracle), through the batched search."""
    from dxt_lossless_transform_b200 import file_formats as ff
    from dxt_lossless_transform_b200 import synth
    from dds_fixtures import FO, make_dds, real_fixture

    est = dlt.LosslessTransformUtilsSizeEstimation()
    bundle = (ff.TransformBundle.new().with_bc1_auto(dlt.Bc1AutoTransformBuilder(est))
              .with_bc2_auto(dlt.Bc2AutoTransformBuilder(est).use_all_decorrelation_modes(True)))
    files = [real_fixture("bc1"), real_fixture("bc2")]
    for i in range(10):
        fmt = FO.DDS_BC1 if i % 2 else FO.DDS_BC2
        n = 1 if fmt == FO.DDS_BC1 else 2
        w, h, mips = [(64, 64, 1), (256, 128, 8), (20, 12, 2)][i % 3]
        length = FO.parse_dds_ignore_magic(bytes(make_dds(fmt, w, h, mips)[:128]))[2]
        files.append(make_dds(fmt, w, h, mips, payload=synth.texture_blocks(n, length // (8 * n), seed=i).tobytes(), leftover=b"xyz" * i))
    files.insert(4, real_fixture("bc3"))   # FormatNotImplemented stays a per-file result
    handler = ff.DdsHandler()
    outs = [np.zeros_like(f) for f in files]
    res = handler.transform_bundle_batch(list(zip(files, outs)), bundle)
    for i, (f, o, r) in enumerate(zip(files, outs, res)):
        single = np.zeros_like(f)
        try:
            handler.transform_bundle(f, single, bundle)
        except ff.TransformError as e:
            assert type(r) is type(e), i
            continue
        assert r is None, (i, r)
        assert np.array_equal(o, single), i
        back = np.zeros_like(f)
        handler.untransform(o, back)
        assert np.array_equal(back, f), i


def test_auto_batch_over_every_visible_gpu(dlt, torch):
    """Payload-granular sharding of the batched search: whole payloads per device, nothing exchanged."""
    from dxt_lossless_transform_b200 import file_formats as ff
    from dxt_lossless_transform_b200 import synth
    from dds_fixtures import real_fixture

    devices = list(range(torch.cuda.device_count()))
    items = []
    for i in range(12):
        fmt = 1 + i % 3
        data = synth.texture_blocks(fmt, 3000 + 211 * i, seed=50 + i)
        items.append((fmt, data, np.zeros_like(data)))
    best = dlt.transform_auto_batch(items, False, devices=devices)
    for (fmt, data, out), b in zip(items, best):
        want_out, want = oracle.auto(fmt, data, False)
        assert (int(b.decorrelation_mode), bool(getattr(b, "split_alpha_endpoints", False)), bool(b.split_colour_endpoints)) == want
        assert np.array_equal(out, want_out)
    # and through the DDS batch entry point with an auto bundle
    est = dlt.LosslessTransformUtilsSizeEstimation()
    bundle = ff.TransformBundle.new().with_bc1_auto(dlt.Bc1AutoTransformBuilder(est)).with_bc2_auto(dlt.Bc2AutoTransformBuilder(est))
    files = [real_fixture("bc1"), real_fixture("bc2")] * 3
    outs = [np.zeros_like(f) for f in files]
    handler = ff.DdsHandler()
    assert handler.transform_bundle_batch(list(zip(files, outs)), bundle, devices=devices) == [None] * len(files)
    for f, o in zip(files, outs):
        single = np.zeros_like(f)
        handler.transform_bundle(f, single, bundle)
        assert np.array_equal(o, single)
