"""The restated LTU match estimator with its unverified parts as parameters (hash bits, index from the top / low bits
of the product, positions per loop iteration).

CPU part: (i) the exhaustive facts the CUDA estimator's 16-bit table entries rest on (csrc/estimator.cu): for every
supported parameter set, (bucket, 16-bit tag) identifies the 24-bit key, and storing a tag of 0xFFFF as 0xFFFE never
collides with another key of the same bucket; (ii) the parametrised C oracle against a pure-Python restatement.
GPU part: the CUDA estimator equals the oracle for every parameter set, and the search follows the parameters."""
import numpy as np
import pytest

import oracle

GOLD = np.uint32(0x9E3779B1)
PARAM_SETS = [(h, top, g) for h in range(12, 18) for top in (True, False) for g in (4, 1)]


def packets(hash_bits: int, top: bool):
    """The packet the producers of csrc/estimator.cu build for every 24-bit key: bucket << sb | tag."""
    k = np.arange(1 << 24, dtype=np.uint32)
    p = (k * GOLD).astype(np.uint32)
    sb = 32 - max(hash_bits, 16)
    tag_mask = np.uint32((1 << sb) - 1)
    if top:
        bucket, tag = p >> np.uint32(32 - hash_bits), p & tag_mask
    else:
        bucket, tag = p & np.uint32((1 << hash_bits) - 1), (p >> np.uint32(hash_bits)) & tag_mask
    return (bucket << np.uint32(sb)) | tag, bucket   # hash_bits + sb <= 32: the packet is a 32-bit word


@pytest.mark.parametrize("hash_bits", range(12, 18))
@pytest.mark.parametrize("top", [True, False])
def test_sixteen_bit_entries_identify_the_key(hash_bits, top):
    pkt, bucket = packets(hash_bits, top)
    ordered = np.sort(pkt)
    assert np.all(ordered[1:] != ordered[:-1]), "bucket + tag must identify the key"
    stored = pkt & np.uint32(0xFFFF)   # what a table entry holds
    with_marker = np.unique(bucket[stored == 0xFFFF])      # buckets in which some key's entry would read "untouched"
    with_alias = np.unique(bucket[stored == 0xFFFE])
    assert np.intersect1d(with_marker, with_alias).size == 0, "0xFFFF -> 0xFFFE must stay unique inside every bucket"
    assert pkt[0] == 0, "key 0 lives in entry 0 with tag 0 (the reference's zeroed table)"


def py_matches(data: bytes, hash_bits: int, top: bool, group: int) -> int:
    table = {}
    n = len(data)
    end = n - 7 if n > 7 else 0
    matches = 0
    for i in range(0, end, group):
        ks = []
        for j in range(group):
            d = data[i + j] | (data[i + j + 1] << 8) | (data[i + j + 2] << 16)
            prod = (d * 0x9E3779B1) & 0xFFFFFFFF
            idx = prod >> (32 - hash_bits) if top else prod & ((1 << hash_bits) - 1)
            ks.append((idx, d))
        for idx, d in ks:
            matches += table.get(idx, 0) == d
        for idx, d in ks:
            table[idx] = d
    return matches


@pytest.mark.parametrize("hash_bits,top,group", PARAM_SETS)
def test_parametrised_oracle_equals_python_restatement(hash_bits, top, group):
    rng = np.random.default_rng(hash_bits * 10 + group + top)
    for data in (rng.integers(0, 3, 3000, dtype=np.uint8), rng.integers(0, 256, 5000, dtype=np.uint8),
                 np.zeros(777, np.uint8), np.tile(np.array([7, 7, 9], np.uint8), 500),
                 np.zeros(5, np.uint8), np.zeros(8, np.uint8), np.zeros(11, np.uint8)):
        assert oracle.ltu_matches_params(data, hash_bits, top, group) == py_matches(data.tobytes(), hash_bits, top, group)


def test_default_parameters_are_the_restatement():
    rng = np.random.default_rng(3)
    data = rng.integers(0, 5, 100_000, dtype=np.uint8)
    assert oracle.ltu_matches(data) == oracle.ltu_matches_params(data, 16, True, 4)


# ------------------------------------------------------------------------------------------------ GPU
def gpu_streams():
    rng = np.random.default_rng(11)
    yield "zeros", np.zeros(300_000, np.uint8)
    yield "random_3M", rng.integers(0, 256, 3_000_000, dtype=np.uint8)
    yield "two_symbols", rng.integers(0, 2, 1_500_001, dtype=np.uint8)
    yield "low_entropy", rng.integers(0, 6, 2_000_003, dtype=np.uint8)
    yield "period3", np.tile(np.array([9, 8, 7], np.uint8), 400_000)
    yield "period6", np.tile(np.array([1, 2, 3, 1, 2, 4], np.uint8), 200_000)
    walk = (np.cumsum(rng.integers(-1, 2, 2_500_000)) // 5 % 256).astype(np.uint8)
    yield "random_walk", walk
    for n in (0, 1, 7, 8, 9, 11, 12, 135, 136, 137, 4103):
        yield f"small{n}", rng.integers(0, 4, n, dtype=np.uint8)


@pytest.mark.gpu
@pytest.mark.parametrize("hash_bits,top,group", PARAM_SETS)
def test_gpu_estimator_equals_oracle_for_every_parameter_set(hash_bits, top, group):
    import torch

    import dxt_lossless_transform_b200 as dlt

    dlt.ltu_set_params(hash_bits, top, group)
    try:
        assert dlt.ltu_get_params() == (hash_bits, top, group)
        for name, data in gpu_streams():
            want = oracle.ltu_matches_params(data, hash_bits, top, group) if data.size else 0
            want = 0 if data.size == 0 else max(data.size - want, 0)
            d = torch.from_numpy(data).cuda() if data.size else torch.zeros(1, dtype=torch.uint8).cuda()
            assert dlt.ltu_estimate_device(d.data_ptr(), data.size) == want, (name, hash_bits, top, group)
            if data.size > 1000:
                w1 = oracle.ltu_matches_params(data[1:], hash_bits, top, group)
                assert dlt.ltu_estimate_device(d.data_ptr() + 1, data.size - 1) == data.size - 1 - w1, name
    finally:
        dlt.ltu_set_params()


@pytest.mark.gpu
def test_gpu_search_follows_the_parameters():
    """Choice, output bytes and per-candidate estimates of the device search equal the oracle's under a non-default
    parameter set too (the selection logic does not depend on the estimator's constants)."""
    import torch

    import dxt_lossless_transform_b200 as dlt
    from dxt_lossless_transform_b200 import synth

    data = synth.texture_blocks(1, 50_001, seed=9)
    d_in = torch.from_numpy(data).cuda()
    d_out = torch.empty_like(d_in)
    for params in ((15, False, 1), (17, True, 4)):
        dlt.ltu_set_params(*params)
        oracle.ltu_set_params(*params)
        try:
            best, est = dlt.transform_auto_device(1, d_in.data_ptr(), d_out.data_ptr(), data.size, True)
            want_out, want = oracle.auto(1, data, True)
            assert (int(best.decorrelation_mode), False, best.split_colour_endpoints) == want
            assert np.array_equal(d_out.cpu().numpy(), want_out)
            assert est == oracle.auto_estimates(1, data, True)
        finally:
            dlt.ltu_set_params()
            oracle.ltu_set_params()


def test_unsupported_parameters_are_refused():
    import dxt_lossless_transform_b200 as dlt

    before = dlt.ltu_get_params()
    for bad in ((11, True, 4), (18, True, 4), (16, True, 2), (16, True, 0)):
        with pytest.raises(ValueError):
            dlt.ltu_set_params(*bad)
    assert dlt.ltu_get_params() == before
