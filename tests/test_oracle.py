"""CPU tests: the oracle against the reference's golden vectors / known answers, round trips in the
style of the reference's own tests, and the search-loop selection rules."""
import ctypes as C
import json
import zlib
from pathlib import Path

import numpy as np
import pytest

import oracle

GOLDEN = Path(__file__).resolve().parent / "golden"
KNOWN = json.loads((GOLDEN / "known_answers.json").read_text())


def hexarr(h: str) -> np.ndarray:
    return np.frombuffer(bytes.fromhex(h), np.uint8).copy()


def payload(fmt: int) -> np.ndarray:
    return np.frombuffer(zlib.decompress((GOLDEN / f"r2-256-bc{fmt}.payload.zlib").read_bytes()), np.uint8).copy()


def all_settings(fmt: int):
    for v in range(4):
        for sa in ((False, True) if fmt == 3 else (False,)):
            for sc in (False, True):
                yield v, sa, sc


# ---- golden vectors ---------------------------------------------------------------------------------
def test_decorrelate_known_answers():
    for colour, expect in KNOWN["decorrelate"].items():
        for k, e in enumerate(expect, start=1):
            assert oracle.decorrelate(int(colour, 16), k) == int(e, 16)


@pytest.mark.parametrize("variant", [1, 2, 3])
def test_decorrelate_is_a_bijection_and_recorrelate_inverts_it(variant):
    seen = set()
    for v in range(65536):
        d = oracle.decorrelate(v, variant)
        assert oracle.recorrelate(d, variant) == v
        seen.add(d)
    assert len(seen) == 65536
    assert oracle.decorrelate(0x1234, 0) == 0x1234 and oracle.recorrelate(0x1234, 0) == 0x1234


def test_reference_colour_list_roundtrip():
    # common/src/color_565/decorrelate.rs:408-446 — 13 named colours x 4 variants
    rgb = [(255, 0, 0), (0, 255, 0), (0, 0, 255), (255, 255, 0), (0, 255, 255), (255, 0, 255), (128, 128, 128),
           (255, 255, 255), (0, 0, 0), (255, 128, 64), (128, 0, 255), (0, 128, 64), (31, 79, 83)]
    for r, g, b in rgb:
        c = ((r >> 3) << 11) | ((g >> 2) << 5) | (b >> 3)
        for variant in range(4):
            assert oracle.recorrelate(oracle.decorrelate(c, variant), variant) == c


@pytest.mark.parametrize("fmt", [1, 2, 3])
def test_generators_match_reference_pinned_bytes(fmt):
    assert oracle.generate_test_data(fmt, 3).tobytes().hex() == KNOWN["generators"][f"bc{fmt}"]


def test_split_565_golden_vector():
    src = hexarr(KNOWN["split_565"]["in"])
    dst = np.zeros_like(src)
    oracle.lib().orc_split_color_endpoints(src.ctypes.data, dst.ctypes.data, src.size)
    assert dst.tobytes().hex() == KNOWN["split_565"]["out"]
    # and it is the colour part of the BC1 split layout
    blocks = np.zeros(3 * 8, np.uint8)
    blocks.reshape(3, 8)[:, :4] = src.reshape(3, 4)
    assert oracle.transform(1, blocks, 0, False, True)[:12].tobytes().hex() == KNOWN["split_565"]["out"]


def test_transform_known_answers():
    d1 = hexarr(KNOWN["generators"]["bc1"])
    for key, expect in KNOWN["bc1_3blocks"].items():
        v, s = map(int, key.split("/"))
        assert oracle.transform(1, d1, v, False, bool(s)).tobytes().hex() == expect
    d2 = oracle.generate_test_data(2, 2)
    for key, expect in KNOWN["bc2_2blocks"].items():
        v, s = map(int, key.split("/"))
        assert oracle.transform(2, d2, v, False, bool(s)).tobytes().hex() == expect
    d3 = oracle.generate_test_data(3, 2)
    for key, expect in KNOWN["bc3_2blocks"].items():
        v, sa, sc = map(int, key.split("/"))
        assert oracle.transform(3, d3, v, bool(sa), bool(sc)).tobytes().hex() == expect


# ---- round trips, as the reference tests them -----------------------------------------------------
@pytest.mark.parametrize("fmt", [1, 2, 3])
def test_roundtrip_all_settings_1_to_130_blocks(fmt):
    for nb in list(range(1, 131)) + [512, 1000]:
        data = oracle.generate_test_data(fmt, nb)
        for v, sa, sc in all_settings(fmt):
            t = oracle.transform(fmt, data, v, sa, sc)
            assert np.array_equal(oracle.untransform(fmt, t, v, sa, sc), data), (fmt, nb, v, sa, sc)
            # verbatim fields: a transform is a permutation of the non-colour bytes
            if v == 0:
                assert sorted(t.tobytes()) == sorted(data.tobytes())


@pytest.mark.parametrize("fmt", [1, 2, 3])
def test_real_texture_payload_roundtrip(fmt):
    data = payload(fmt)
    assert data.size == 4096 * (8 if fmt == 1 else 16)
    for v, sa, sc in all_settings(fmt):
        t = oracle.transform(fmt, data, v, sa, sc)
        assert np.array_equal(oracle.untransform(fmt, t, v, sa, sc), data)


@pytest.mark.parametrize("fmt", [1, 2, 3])
def test_layout_sections(fmt):
    """Section offsets of SURVEY.md §8a rows a5/a8/a10, checked field by field on random blocks."""
    rng = np.random.default_rng(7)
    n = 37
    bpb = 8 if fmt == 1 else 16
    data = rng.integers(0, 256, n * bpb, dtype=np.uint8)
    blk = data.reshape(n, bpb)
    col = blk[:, bpb - 8:bpb - 4]
    idx = blk[:, bpb - 4:]
    for sa in ((False, True) if fmt == 3 else (False,)):
        for sc in (False, True):
            t = oracle.transform(fmt, data, 0, sa, sc)
            cbase = {1: 0, 2: 8 * n, 3: 8 * n}[fmt]
            if sc:
                assert np.array_equal(t[cbase:cbase + 2 * n].reshape(n, 2), col[:, :2])
                assert np.array_equal(t[cbase + 2 * n:cbase + 4 * n].reshape(n, 2), col[:, 2:])
            else:
                assert np.array_equal(t[cbase:cbase + 4 * n].reshape(n, 4), col)
            assert np.array_equal(t[cbase + 4 * n:cbase + 8 * n].reshape(n, 4), idx)
            if fmt == 2:
                assert np.array_equal(t[:8 * n].reshape(n, 8), blk[:, :8])
            if fmt == 3:
                if sa:
                    assert np.array_equal(t[:n], blk[:, 0]) and np.array_equal(t[n:2 * n], blk[:, 1])
                else:
                    assert np.array_equal(t[:2 * n].reshape(n, 2), blk[:, :2])
                assert np.array_equal(t[2 * n:8 * n].reshape(n, 6), blk[:, 2:8])


@pytest.mark.parametrize("fmt", [1, 2, 3])
def test_multithreaded_and_ranged_drivers_equal_single_call(fmt):
    rng = np.random.default_rng(fmt)
    data = rng.integers(0, 256, 1237 * (8 if fmt == 1 else 16), dtype=np.uint8)
    for v, sa, sc in all_settings(fmt):
        t = oracle.transform(fmt, data, v, sa, sc)
        assert np.array_equal(oracle.transform(fmt, data, v, sa, sc, threads=5), t)
        assert np.array_equal(oracle.untransform(fmt, t, v, sa, sc, threads=3), data)
        parts = np.zeros_like(data)
        for b0, b1 in ((0, 400), (400, 401), (401, 1237)):
            oracle.run_range(fmt, False, data, parts, v, sa, sc, b0, b1)
        assert np.array_equal(parts, t)


@pytest.mark.parametrize("fmt", [1, 2, 3])
def test_vector_baseline_paths_equal_the_scalar_restatement(fmt):
    """The CPU baseline (orc_bcn_run_range / _mt: AVX-512 where the host has it, as the reference would pick) against
    the scalar restatement: every settings combination, both directions, block counts around the vector width, and
    buffers misaligned by one byte (the reference's own unaligned tests)."""
    import ctypes as C

    L = oracle.lib()
    L.orc_cpu_baseline_isa.restype = C.c_int
    assert L.orc_cpu_baseline_isa() in (0, 2, 5)
    bpb = 8 if fmt == 1 else 16
    rng = np.random.default_rng(100 + fmt)
    for nblocks in (1, 15, 16, 17, 31, 32, 33, 63, 64, 65, 333):
        raw = rng.integers(0, 256, nblocks * bpb + 1, dtype=np.uint8)
        for mis in (0, 1):
            data = raw[mis:mis + nblocks * bpb]
            for v, sa, sc in all_settings(fmt):
                want = oracle.transform(fmt, data.copy(), v, sa, sc)
                buf = np.zeros(nblocks * bpb + 1, np.uint8)
                got = buf[mis:mis + nblocks * bpb]
                oracle.run_range(fmt, False, data, got, v, sa, sc, 0, nblocks)
                assert np.array_equal(got, want), (nblocks, mis, v, sa, sc)
                back = np.zeros(nblocks * bpb + 1, np.uint8)[mis:mis + nblocks * bpb]
                oracle.run_range(fmt, True, got, back, v, sa, sc, 0, nblocks)
                assert np.array_equal(back, data), (nblocks, mis, v, sa, sc)


# ---- LTU estimator: the reference's own (inequality) tests -----------------------------------------
def test_ltu_reference_inequalities():
    # extensions/estimators/dxt-lossless-transform-ltu/src/lib.rs:125-225, tests/integration_test.rs
    assert oracle.ltu_estimate(np.zeros(0, np.uint8)) == 0
    zeros = np.zeros(64, np.uint8)
    assert oracle.ltu_estimate(zeros) < 64
    rng = np.random.default_rng(0)
    rnd = rng.integers(0, 256, 4096, dtype=np.uint8)
    rep = np.tile(np.arange(16, dtype=np.uint8), 256)
    assert oracle.ltu_estimate(rep) <= oracle.ltu_estimate(rnd) <= rnd.size
    assert oracle.ltu_estimate(rnd) == oracle.ltu_estimate(rnd.copy())


def test_ltu_small_cases_by_hand():
    # len <= 7: the loop body never runs
    for n in range(0, 8):
        assert oracle.ltu_matches(np.zeros(max(n, 1), np.uint8)[:n]) == 0 if n else True
    # 8 zero bytes: one group of 4 positions, all keys 0 == zero-initialised table -> 4 matches
    assert oracle.ltu_matches(np.zeros(8, np.uint8)) == 4
    # 12 bytes: end = 5 -> groups at 0 and 4
    assert oracle.ltu_matches(np.zeros(12, np.uint8)) == 8
    # distinct keys never match; the same 4-byte-periodic pattern matches from the second group on
    pat = np.tile(np.array([1, 2, 3, 4], np.uint8), 8)  # 32 bytes, end = 25 -> 7 groups
    assert oracle.ltu_matches(pat) == 6 * 4


# ---- search loop: order, strict '<', tie -> first, final re-transform ------------------------------
EST = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t))


def run_auto(fmt, data, use_all, fn):
    out = np.zeros_like(data)
    cb = EST(fn)
    v, sa, sc = C.c_int(), C.c_int(0), C.c_int()
    L = oracle.lib()
    if fmt == 3:
        rc = L.orc_bc3_transform_auto(data.ctypes.data, out.ctypes.data, data.size, int(use_all), cb, None,
                                      C.byref(v), C.byref(sa), C.byref(sc))
    else:
        rc = getattr(L, f"orc_bc{fmt}_transform_auto")(data.ctypes.data, out.ctypes.data, data.size, int(use_all),
                                                       cb, None, C.byref(v), C.byref(sc))
    return rc, out, (v.value, bool(sa.value), bool(sc.value))


@pytest.mark.parametrize("fmt,use_all,expect", [
    (1, False, (0, False, False)), (1, True, (2, False, False)),
    (2, False, (0, False, False)), (2, True, (2, False, False)),
    (3, False, (1, True, False)), (3, True, (2, True, False)),
])
def test_auto_all_tie_picks_first_in_order(fmt, use_all, expect):
    # DummyEstimator of the reference returns len_bytes: every candidate ties
    def dummy(_ctx, _p, n, out):
        out[0] = n
        return 0

    data = oracle.generate_test_data(fmt, 33)
    rc, out, best = run_auto(fmt, data, use_all, dummy)
    assert rc == 0 and best == expect
    assert np.array_equal(out, oracle.transform(fmt, data, *best))  # final re-transform happened


def test_auto_strict_less_and_last_candidate_wins_without_retransform():
    calls = []

    def decreasing(_ctx, _p, n, out):
        calls.append(n)
        out[0] = 1000 - len(calls)
        return 0

    data = oracle.generate_test_data(1, 16)
    rc, out, best = run_auto(1, data, False, decreasing)
    assert rc == 0 and best == (1, False, True)  # last of FAST_TEST_ORDER
    assert calls == [data.size // 2] * 4           # BC1 estimates out[0, len/2) once per candidate
    assert np.array_equal(out, oracle.transform(1, data, 1, False, True))


def test_auto_estimator_failure_propagates():
    def failing(_ctx, _p, n, out):
        return 7

    rc, _, _ = run_auto(2, oracle.generate_test_data(2, 4), True, failing)
    assert rc == 7


def test_auto_bc3_sums_alpha_and_colour_estimates():
    seen = []

    def rec(_ctx, _p, n, out):
        seen.append(n)
        out[0] = n
        return 0

    data = oracle.generate_test_data(3, 10)
    run_auto(3, data, False, rec)
    assert seen == [20, 40] * 8  # (2N, 4N) per candidate — bc3 transform_auto.rs:253-281


@pytest.mark.parametrize("fmt", [1, 2, 3])
def test_auto_with_ltu_matches_its_own_estimates(fmt):
    data = payload(fmt)
    for use_all in (False, True):
        sizes = oracle.auto_estimates(fmt, data, use_all)
        out, best = oracle.auto(fmt, data, use_all)
        assert np.array_equal(out, oracle.transform(fmt, data, *best))
        assert min(sizes) == sizes[[i for i, s in enumerate(sizes) if s == min(sizes)][0]]
