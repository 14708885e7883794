"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle (bit-exact)."""
import json
import os
import zlib
from pathlib import Path

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

GOLDEN = Path(__file__).resolve().parent / "golden"
KNOWN = json.loads((GOLDEN / "known_answers.json").read_text())


@pytest.fixture(scope="module")
def dlt():
    import dxt_lossless_transform_b200 as m

    return m


@pytest.fixture(scope="module")
def torch():
    import torch as t

    assert t.cuda.is_available()
    return t


def bpb(fmt):
    return 8 if fmt == 1 else 16


def settings_list(dlt, fmt):
    return list({1: dlt.Bc1TransformSettings, 2: dlt.Bc2TransformSettings, 3: dlt.Bc3TransformSettings}[fmt].all_combinations())


def orc_args(s):
    return int(s.decorrelation_mode), bool(getattr(s, "split_alpha_endpoints", False)), bool(s.split_colour_endpoints)


def payload(fmt):
    return np.frombuffer(zlib.decompress((GOLDEN / f"r2-256-bc{fmt}.payload.zlib").read_bytes()), np.uint8).copy()


def rand_blocks(fmt, n, seed):
    return np.random.default_rng(seed).integers(0, 256, n * bpb(fmt), dtype=np.uint8)


# ---- host-pointer path (the reference's entry points) ------------------------------------------------
def test_known_answer_vectors_through_the_cabi(dlt):
    d1 = np.frombuffer(bytes.fromhex(KNOWN["generators"]["bc1"]), np.uint8).copy()
    for key, expect in KNOWN["bc1_3blocks"].items():
        v, s = map(int, key.split("/"))
        out = np.zeros_like(d1)
        dlt.transform_bc1_with_settings(d1, out, dlt.Bc1TransformSettings(dlt.YCoCgVariant(v), bool(s)))
        assert out.tobytes().hex() == expect
    d3 = oracle.generate_test_data(3, 2)
    for key, expect in KNOWN["bc3_2blocks"].items():
        v, sa, sc = map(int, key.split("/"))
        out = np.zeros_like(d3)
        dlt.transform_bc3_with_settings(d3, out, dlt.Bc3TransformSettings(dlt.YCoCgVariant(v), bool(sa), bool(sc)))
        assert out.tobytes().hex() == expect
    d2 = oracle.generate_test_data(2, 2)
    out = np.zeros_like(d2)
    dlt.transform_bc2_with_settings(d2, out, dlt.Bc2TransformSettings(dlt.YCoCgVariant.Variant1, True))
    assert out.tobytes().hex() == KNOWN["bc2_2blocks"]["1/1"]


@pytest.mark.parametrize("fmt", [1, 2, 3])
def test_reference_generators_1_to_130_blocks_all_settings(dlt, fmt):
    """The reference's per-ISA tests run 1..=2x the SIMD width; here 1..=130 blocks."""
    for nb in list(range(1, 131)) + [2047, 2048, 2049, 4095, 4097]:
        data = oracle.generate_test_data(fmt, nb)
        for s in settings_list(dlt, fmt):
            out = np.full(data.size + 8, 0xAA, np.uint8)  # output_len >= input_len: only the prefix is written
            dlt.transform_with_settings(fmt, data, out, s)
            assert np.array_equal(out[:data.size], oracle.transform(fmt, data, *orc_args(s))), (fmt, nb, s)
            assert (out[data.size:] == 0xAA).all()
            back = np.zeros_like(data)
            dlt.untransform_with_settings(fmt, out[:data.size].copy(), back, s)
            assert np.array_equal(back, data), (fmt, nb, s)


@pytest.mark.parametrize("fmt", [1, 2, 3])
def test_real_texture_payloads_all_settings(dlt, fmt):
    data = payload(fmt)
    for s in settings_list(dlt, fmt):
        out = np.zeros_like(data)
        dlt.transform_with_settings(fmt, data, out, s)
        assert np.array_equal(out, oracle.transform(fmt, data, *orc_args(s)))
        back = np.zeros_like(data)
        dlt.untransform_with_settings(fmt, out, back, s)
        assert np.array_equal(back, data)


@pytest.mark.parametrize("fmt", [1, 2, 3])
def test_misaligned_host_buffers(dlt, fmt):
    """The reference tests every untransform with 1-byte misaligned buffers (test_prelude.rs:340-536)."""
    nb = 777
    data = rand_blocks(fmt, nb, 3)
    for off_in, off_out in ((1, 0), (0, 1), (1, 1), (3, 5)):
        src = np.zeros(data.size + 16, np.uint8)
        src[off_in:off_in + data.size] = data
        for s in settings_list(dlt, fmt)[::3]:
            dst = np.zeros(data.size + 16, np.uint8)
            dlt.transform_with_settings(fmt, src[off_in:off_in + data.size], dst[off_out:off_out + data.size], s)
            t = oracle.transform(fmt, data, *orc_args(s))
            assert np.array_equal(dst[off_out:off_out + data.size], t)
            back = np.zeros(data.size + 16, np.uint8)
            dlt.untransform_with_settings(fmt, dst[off_out:off_out + data.size], back[off_in:off_in + data.size], s)
            assert np.array_equal(back[off_in:off_in + data.size], data)


@pytest.mark.parametrize("fmt", [1, 2, 3])
def test_multi_chunk_host_pipeline_pageable_and_pinned(dlt, fmt):
    """Several chunks for pageable buffers (16 MiB staging slots), ragged last chunk; pinned and pageable buffers."""
    nb = (28 << 20) // bpb(fmt) + 12345
    data = rand_blocks(fmt, nb, 11)
    s = settings_list(dlt, fmt)[0]
    expect = oracle.transform(fmt, data, *orc_args(s), threads=8)
    out = np.zeros_like(data)
    dlt.transform_with_settings(fmt, data, out, s)
    assert np.array_equal(out, expect)
    pin_in, pin_out = dlt.alloc_pinned(data.size), dlt.alloc_pinned(data.size)
    pin_in.array[:] = data
    dlt.transform_with_settings(fmt, pin_in.array, pin_out.array, s)
    assert np.array_equal(pin_out.array, expect)
    dlt.untransform_with_settings(fmt, pin_out.array, pin_in.array, s)
    assert np.array_equal(pin_in.array, data)
    back = np.zeros_like(data)
    dlt.untransform_with_settings(fmt, out, back, s)
    assert np.array_equal(back, data)
    pin_in.free(), pin_out.free()


_SMALL_CHUNK_SCRIPT = r"""
import sys
import numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
import dxt_lossless_transform_b200 as dlt
import oracle
for fmt, S in ((1, dlt.Bc1TransformSettings), (2, dlt.Bc2TransformSettings), (3, dlt.Bc3TransformSettings)):
    bpb = 8 if fmt == 1 else 16
    nb = (13 << 20) // bpb + 4321          # 13 MiB and a bit: 13+ chunks of 1 MiB, odd block count
    rng = np.random.default_rng(fmt)
    data = rng.integers(0, 256, nb * bpb, dtype=np.uint8)
    s = S()
    args = (int(s.decorrelation_mode), bool(getattr(s, "split_alpha_endpoints", False)), bool(s.split_colour_endpoints))
    expect = oracle.transform(fmt, data, *args, threads=8)
    pin_in, pin_out = dlt.alloc_pinned(data.size), dlt.alloc_pinned(data.size)
    for src, dst in ((data, np.zeros_like(data)), (pin_in.array, pin_out.array)):
        src[:] = data
        dlt.transform_with_settings(fmt, src, dst, s)
        assert np.array_equal(dst, expect), (fmt, "transform")
        src[:] = 0
        dlt.untransform_with_settings(fmt, dst, src, s)
        assert np.array_equal(src, data), (fmt, "untransform")
print("ok")
"""


@pytest.mark.parametrize("ramp,strided", [("0", "1"), ("1", "1"), ("1", "0")])
def test_host_pipeline_with_many_small_chunks(ramp, strided):
    """The copy pipeline's chunk schedule (ring of slots wrapping many times, short-chunk ramp at both ends, ragged
    last chunk, strided copies of equal-width neighbouring streams) — forced by 1 MiB chunks in a child process, because
    the host-path configuration is read once."""
    import os
    import subprocess
    import sys

    root = str(Path(__file__).resolve().parent.parent)
    env = dict(os.environ, DLTCUDA_CHUNK_MIB="1", DLTCUDA_RAMP=ramp, DLTCUDA_STRIDED=strided)
    r = subprocess.run([sys.executable, "-c", _SMALL_CHUNK_SCRIPT, root], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout + r.stderr


def test_stable_builders_roundtrip(dlt):
    for Builder, fmt in ((dlt.Bc1ManualTransformBuilder, 1), (dlt.Bc2ManualTransformBuilder, 2)):
        data = rand_blocks(fmt, 1000, 5)
        for v in dlt.YCoCgVariant:
            for split in (False, True):
                b = Builder().decorrelation_mode(v).split_colour_endpoints(split)
                out, back = np.zeros_like(data), np.zeros_like(data)
                b.transform(data, out)
                assert np.array_equal(out, oracle.transform(fmt, data, int(v), False, split))
                b.clone().untransform(out, back)
                assert np.array_equal(back, data)


# ---- device-resident path ---------------------------------------------------------------------------
def dev(torch, a):
    return torch.from_numpy(a).cuda()


@pytest.mark.parametrize("fmt", [1, 2, 3])
def test_device_resident_all_settings_odd_block_counts(dlt, torch, fmt):
    """Odd N puts the stream bases of the reference layout at odd offsets: shifted staging path."""
    for nb in (1, 2, 3, 15, 16, 17, 1023, 1025, 5463, 2048 * 3 + 1, 65537):
        data = rand_blocks(fmt, nb, nb)
        d_in = dev(torch, data)
        for s in settings_list(dlt, fmt):
            d_out = torch.zeros_like(d_in)
            dlt.transform_device(fmt, d_in.data_ptr(), d_out.data_ptr(), data.size, s)
            torch.cuda.synchronize()
            assert np.array_equal(d_out.cpu().numpy(), oracle.transform(fmt, data, *orc_args(s))), (fmt, nb, s)
            d_back = torch.zeros_like(d_in)
            dlt.untransform_device(fmt, d_out.data_ptr(), d_back.data_ptr(), data.size, s)
            torch.cuda.synchronize()
            assert np.array_equal(d_back.cpu().numpy(), data), (fmt, nb, s)


@pytest.mark.parametrize("fmt", [1, 2, 3])
def test_device_pointers_of_any_alignment(dlt, torch, fmt):
    """Block or stream pointers the tiled kernels cannot address take the byte-granular kernel."""
    nb = 3001
    data = rand_blocks(fmt, nb, 9)
    for off_in, off_out in ((1, 0), (0, 1), (8, 0), (2, 6)):
        d_in = torch.zeros(data.size + 32, dtype=torch.uint8, device="cuda")
        d_in[off_in:off_in + data.size] = dev(torch, data)
        for s in settings_list(dlt, fmt)[1::3]:
            d_out = torch.zeros(data.size + 32, dtype=torch.uint8, device="cuda")
            dlt.transform_device(fmt, d_in.data_ptr() + off_in, d_out.data_ptr() + off_out, data.size, s)
            torch.cuda.synchronize()
            got = d_out.cpu().numpy()
            assert np.array_equal(got[off_out:off_out + data.size], oracle.transform(fmt, data, *orc_args(s)))
            assert (got[:off_out] == 0).all() and (got[off_out + data.size:] == 0).all()
            d_back = torch.zeros(data.size + 32, dtype=torch.uint8, device="cuda")
            dlt.untransform_device(fmt, d_out.data_ptr() + off_out, d_back.data_ptr() + off_in, data.size, s)
            torch.cuda.synchronize()
            assert np.array_equal(d_back.cpu().numpy()[off_in:off_in + data.size], data)


@pytest.mark.parametrize("fmt", [1, 2, 3])
def test_no_write_outside_the_output(dlt, torch, fmt):
    nb = 2048 * 2 + 37
    data = rand_blocks(fmt, nb, 1)
    d_in = dev(torch, data)
    guard = 4096
    for s in settings_list(dlt, fmt)[::5]:
        d_out = torch.full((data.size + 2 * guard,), 0x5A, dtype=torch.uint8, device="cuda")
        dlt.transform_device(fmt, d_in.data_ptr(), d_out.data_ptr() + guard, data.size, s)
        torch.cuda.synchronize()
        got = d_out.cpu().numpy()
        assert (got[:guard] == 0x5A).all() and (got[guard + data.size:] == 0x5A).all()
        assert np.array_equal(got[guard:guard + data.size], oracle.transform(fmt, data, *orc_args(s)))


@pytest.mark.parametrize("fmt", [1, 2, 3])
def test_block_range_shards_compose_to_the_whole(dlt, torch, fmt):
    """shards share only (N, first_block); any partition gives the same bytes."""
    from dxt_lossless_transform_b200 import sharding

    nb = 100_003
    data = rand_blocks(fmt, nb, 21)
    d_in = dev(torch, data)
    for s in settings_list(dlt, fmt)[::3]:
        expect = oracle.transform(fmt, data, *orc_args(s))
        for shards in (2, 3, 8):
            d_out = torch.zeros_like(d_in)
            for first, count in sharding.shard_ranges(fmt, nb, shards):
                dlt.transform_device_range(fmt, d_in.data_ptr() + first * bpb(fmt), d_out.data_ptr(), nb, first, count, s)
            torch.cuda.synchronize()
            assert np.array_equal(d_out.cpu().numpy(), expect)
            d_back = torch.zeros_like(d_in)
            for first, count in sharding.shard_ranges(fmt, nb, shards):
                dlt.untransform_device_range(fmt, d_out.data_ptr(), d_back.data_ptr() + first * bpb(fmt), nb, first, count, s)
            torch.cuda.synchronize()
            assert np.array_equal(d_back.cpu().numpy(), data)
        # an arbitrary (odd) cut is legal too
        d_out = torch.zeros_like(d_in)
        for first, count in ((0, 777), (777, nb - 777)):
            dlt.transform_device_range(fmt, d_in.data_ptr() + first * bpb(fmt), d_out.data_ptr(), nb, first, count, s)
        torch.cuda.synchronize()
        assert np.array_equal(d_out.cpu().numpy(), expect)


@pytest.mark.parametrize("variant", [1, 2, 3])
def test_exhaustive_colour_table_on_the_gpu(dlt, torch, variant):
    """All 65,536 RGB565 values through the kernels' packed arithmetic vs the scalar oracle."""
    colours = np.arange(65536, dtype=np.uint16)
    blocks = np.zeros((65536, 4), np.uint16)
    blocks[:, 0] = colours
    blocks[:, 1] = colours[::-1]
    data = blocks.view(np.uint8).reshape(-1).copy()
    d_in = dev(torch, data)
    d_out = torch.zeros_like(d_in)
    s = dlt.Bc1TransformSettings(dlt.YCoCgVariant(variant), True)
    dlt.transform_device(1, d_in.data_ptr(), d_out.data_ptr(), data.size, s)
    torch.cuda.synchronize()
    got = d_out.cpu().numpy()
    c0 = got[:131072].view(np.uint16)
    expect = np.array([oracle.decorrelate(int(c), variant) for c in colours], np.uint16)
    assert np.array_equal(c0, expect)
    assert np.array_equal(got[131072:262144].view(np.uint16), expect[::-1])


def byte_histogram(torch, t):
    h = torch.zeros(256, dtype=torch.int64, device=t.device)
    step = 64 << 20
    for i in range(0, t.numel(), step):
        h += torch.bincount(t[i:i + step].int(), minlength=256)
    return h


# ---- BASELINE.json full sizes: size-independent properties + a full oracle comparison ---------------
@pytest.mark.parametrize("fmt", [1, 2, 3])
def test_one_gib_device_resident(dlt, torch, fmt):
    nbytes = 1 << 30
    g = torch.Generator(device="cuda").manual_seed(1234 + fmt)
    d_in = torch.randint(0, 256, (nbytes,), dtype=torch.uint8, device="cuda", generator=g)
    d_out, d_back = torch.empty_like(d_in), torch.empty_like(d_in)
    all_s = settings_list(dlt, fmt)
    hist_in = byte_histogram(torch, d_in)
    host = d_in.cpu().numpy()
    d_expect = torch.empty_like(d_in)
    threads = min(32, os.cpu_count() or 1)
    for s in all_s:
        dlt.transform_device(fmt, d_in.data_ptr(), d_out.data_ptr(), nbytes, s)
        dlt.untransform_device(fmt, d_out.data_ptr(), d_back.data_ptr(), nbytes, s)
        torch.cuda.synchronize()
        assert torch.equal(d_back, d_in), s
        if int(s.decorrelation_mode) == 0:
            # a pure permutation keeps the byte histogram
            assert torch.equal(byte_histogram(torch, d_out), hist_in)
        # EVERY setting, all 2^30 bytes, against the oracle (BASELINE configs[1] / [2]: "bit-exact vs reference")
        d_expect.copy_(torch.from_numpy(oracle.transform(fmt, host, *orc_args(s), threads=threads)))
        assert torch.equal(d_out, d_expect), s


def test_batch_of_mixed_payloads(dlt, torch):
    """A directory's worth of independent payloads (mixed BC1/BC2/BC3, mixed settings, pinned and pageable,
    tiny to multi-chunk) through one pipelined pass; and dealt out over the visible GPUs."""
    rng = np.random.default_rng(99)
    items, expect, keep = [], [], []
    sizes = [1, 7, 100, 4096, 70_001, 300_000, (20 << 20) // 16 + 5, 0, 2048, 33]
    for i, nb in enumerate(sizes):
        fmt = (1, 3, 2)[i % 3]
        s = settings_list(dlt, fmt)[(5 * i) % len(settings_list(dlt, fmt))]
        data = rng.integers(0, 256, nb * bpb(fmt), dtype=np.uint8)
        if i % 2:
            pin_in, pin_out = dlt.alloc_pinned(data.size), dlt.alloc_pinned(data.size)
            pin_in.array[:] = data
            src, dst = pin_in.array, pin_out.array
            keep += [pin_in, pin_out]
        else:
            src, dst = data, np.zeros_like(data)
        items.append((fmt, src, dst, s))
        expect.append(oracle.transform(fmt, data, *orc_args(s)) if nb else data)
    dlt.transform_batch(items)
    for (fmt, src, dst, s), want in zip(items, expect):
        assert np.array_equal(dst, want), (fmt, s, dst.size)
    # the way back, over every visible GPU
    back_items = [(fmt, dst.copy() if not isinstance(dst, np.ndarray) else dst, np.zeros(dst.size, np.uint8), s)
                  for fmt, src, dst, s in items]
    dlt.transform_batch(back_items, untransform=True, devices=list(range(torch.cuda.device_count())))
    for (fmt, src, dst, s), (_, _, back, _) in zip(items, back_items):
        assert np.array_equal(back, np.asarray(src)), (fmt, s)
    with pytest.raises(dlt.api.InvalidLength):
        dlt.transform_batch([(1, np.zeros(12, np.uint8), np.zeros(12, np.uint8), dlt.Bc1TransformSettings())])


def test_split_color_endpoints(dlt, torch):
    """common/src/transforms/split_565_color_endpoints/tests.rs: golden vector, 1..=512 pairs aligned and
    unaligned against the portable reference implementation."""
    src = np.frombuffer(bytes.fromhex(KNOWN["split_565"]["in"]), np.uint8).copy()
    dst = np.zeros_like(src)
    dlt.split_color_endpoints(src, dst)
    assert dst.tobytes().hex() == KNOWN["split_565"]["out"]
    rng = np.random.default_rng(4)
    for pairs in list(range(1, 513)) + [100_001]:
        data = rng.integers(0, 256, pairs * 4, dtype=np.uint8)
        want = np.zeros_like(data)
        oracle.lib().orc_split_color_endpoints(data.ctypes.data, want.ctypes.data, data.size)
        got = np.zeros_like(data)
        dlt.split_color_endpoints(data, got)
        assert np.array_equal(got, want), pairs
        if pairs % 37 == 0 or pairs > 512:
            for off in (0, 1):
                d_in = torch.zeros(data.size + 16, dtype=torch.uint8, device="cuda")
                d_in[off:off + data.size] = torch.from_numpy(data).cuda()
                d_out = torch.zeros(data.size + 16, dtype=torch.uint8, device="cuda")
                dlt.split_color_endpoints_device(d_in.data_ptr() + off, d_out.data_ptr() + off, data.size)
                torch.cuda.synchronize()
                assert np.array_equal(d_out.cpu().numpy()[off:off + data.size], want)
    with pytest.raises(dlt.api.InvalidLength):
        dlt.split_color_endpoints(np.zeros(6, np.uint8), np.zeros(6, np.uint8))


def test_concurrent_host_threads(dlt):
    """SURVEY §8b threading contract: the entry points are synchronous and re-entrant; distinct host threads (each with
    its own builders) may call them concurrently.  Contexts (streams, slots, scratch) are pooled per device."""
    import threading

    from dxt_lossless_transform_b200 import file_formats as ff
    from dxt_lossless_transform_b200 import synth
    from dds_fixtures import FO, make_dds

    SETTINGS = {fmt: settings_list(dlt, fmt) for fmt in (1, 2, 3)}
    errors = []

    def worker(k: int):
        try:
            rng = np.random.default_rng(k)
            for it in range(6):
                fmt = 1 + (k + it) % 3
                nblocks = int(rng.integers(1, 60_000))
                data = synth.texture_blocks(fmt, nblocks, seed=100 * k + it)
                s = list(SETTINGS[fmt])[(k * 7 + it) % len(SETTINGS[fmt])]
                out, back = np.zeros_like(data), np.zeros_like(data)
                dlt.transform_with_settings(fmt, data, out, s)
                assert np.array_equal(out, oracle.transform(fmt, data, *orc_args(s))), (k, it, "transform")
                dlt.untransform_with_settings(fmt, out, back, s)
                assert np.array_equal(back, data), (k, it, "roundtrip")
                if it % 3 == 0:  # a best-settings search (estimator scratch, several launches) in the mix
                    fa = 1 + k % 2
                    d2 = synth.texture_blocks(fa, 9_000 + 37 * k, seed=k)
                    o2 = np.zeros_like(d2)
                    auto = dlt.transform_bc1_auto if fa == 1 else dlt.transform_bc2_auto
                    best = auto(d2, o2, dlt.Bc1EstimateSettings(dlt.LosslessTransformUtilsSizeEstimation(), False))
                    want_out, want = oracle.auto(fa, d2, False)
                    assert (int(best.decorrelation_mode), False, bool(best.split_colour_endpoints)) == want, (k, it, "auto")
                    assert np.array_equal(o2, want_out), (k, it, "auto bytes")
                if it % 3 == 1:  # and a DDS file through the bundle
                    dds = make_dds(FO.DDS_BC1, 64, 32, 3)
                    t, b2 = np.zeros_like(dds), np.zeros_like(dds)
                    h = ff.DdsHandler()
                    h.transform_bundle(dds, t, ff.TransformBundle.default_all())
                    h.untransform(t, b2)
                    assert np.array_equal(b2, dds), (k, it, "dds")
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_batch_of_small_pinned_payloads_is_grouped_by_settings(dlt):
    """Many small payloads carved out of ONE page-locked pool (a directory of textures): the host path launches one
    kernel per settings combination over all of them (mapped memory, no copies).  Mixed formats, settings, sizes (odd
    block counts) and a few payloads at offsets the tiled kernels cannot take; both directions; guard gaps untouched."""
    rng = np.random.default_rng(77)
    pool_bytes = 24 << 20
    pin_in, pin_out, pin_back = dlt.alloc_pinned(pool_bytes), dlt.alloc_pinned(pool_bytes), dlt.alloc_pinned(pool_bytes)
    pin_in.array[:] = rng.integers(0, 256, pool_bytes, dtype=np.uint8)
    pin_out.array[:] = 0xEE
    pin_back.array[:] = 0xEE
    items, back_items, spans = [], [], []
    off = 0
    for i in range(300):
        fmt = 1 + i % 3
        all_s = settings_list(dlt, fmt)
        s = all_s[int(rng.integers(0, len(all_s)))]
        nb = int(rng.choice([1, 2, 31, 257, 1024, 2049, 4097]))
        nbytes = nb * bpb(fmt)
        off = (off + 15) // 16 * 16 + (4 if i % 37 == 0 else 0)    # now and then a 4-byte aligned payload
        if off + nbytes + 64 > pool_bytes:
            break
        src, dst, back = (a.array[off:off + nbytes] for a in (pin_in, pin_out, pin_back))
        items.append((fmt, src, dst, s))
        back_items.append((fmt, dst, back, s))
        spans.append((off, nbytes, fmt, s))
        off += nbytes + 48                                          # a guard gap after every payload
    dlt.transform_batch(items)
    dlt.transform_batch(back_items, untransform=True)
    covered = np.zeros(pool_bytes, bool)
    for o, nbytes, fmt, s in spans:
        covered[o:o + nbytes] = True
        expect = oracle.transform(fmt, pin_in.array[o:o + nbytes], *orc_args(s))
        assert np.array_equal(pin_out.array[o:o + nbytes], expect), (o, nbytes, fmt, s)
        assert np.array_equal(pin_back.array[o:o + nbytes], pin_in.array[o:o + nbytes]), (o, nbytes, fmt, s)
    assert (pin_out.array[~covered] == 0xEE).all() and (pin_back.array[~covered] == 0xEE).all()
    for a in (pin_in, pin_out, pin_back):
        a.free()


def test_repeated_calls_do_not_leak(dlt, torch):
    """Contexts, staging buffers and scratch are pooled: thousands of calls through every kind of entry point must leave
    device memory and host RSS where they were after the first few hundred."""
    import psutil

    from dxt_lossless_transform_b200 import synth

    data1, data3 = synth.texture_blocks(1, 9001, seed=1), synth.texture_blocks(3, 4097, seed=3)
    out1, out3 = np.zeros_like(data1), np.zeros_like(data3)
    est = dlt.Bc1EstimateSettings(dlt.LosslessTransformUtilsSizeEstimation(), False)
    cb = dlt.Bc1EstimateSettings(dlt.CallbackSizeEstimator(lambda a: int(a[::7].sum())), False)

    def burst(n):
        for i in range(n):
            dlt.transform_bc1_with_settings(data1, out1, dlt.Bc1TransformSettings())
            dlt.untransform_bc3_with_settings(data3, out3, dlt.Bc3TransformSettings())
            if i % 4 == 0:
                dlt.transform_bc1_auto(data1, out1, est)
                dlt.transform_bc3_auto(data3, out3, cb)
                dlt.transform_auto_batch([(1, data1, out1), (3, data3, out3)], False)
                dlt.transform_batch([(1, data1, out1, dlt.Bc1TransformSettings())])

    burst(200)
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    rss0 = psutil.Process().memory_info().rss
    burst(1500)
    torch.cuda.synchronize()
    assert free0 - torch.cuda.mem_get_info()[0] < (8 << 20)
    assert psutil.Process().memory_info().rss - rss0 < (64 << 20)
