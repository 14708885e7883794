"""Property-based GPU parity (hypothesis): random format, settings, block count, pointer misalignment and shard cut —
the CUDA path through the C ABI must equal the CPU oracle bit for bit, round-trip, and never write outside its output.
Complements the enumerated cases of test_gpu_parity.py (which mirror the reference's own test matrix)."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dlt():
    import dxt_lossless_transform_b200 as m

    return m


@pytest.fixture(scope="module")
def torch():
    import torch as t

    assert t.cuda.is_available()
    return t


def make_settings(dlt, fmt, variant, split_alpha, split_colour):
    v = dlt.YCoCgVariant(variant)
    if fmt == 3:
        return dlt.Bc3TransformSettings(v, split_alpha, split_colour)
    return (dlt.Bc1TransformSettings if fmt == 1 else dlt.Bc2TransformSettings)(v, split_colour)


case = st.tuples(
    st.integers(1, 3),                      # format
    st.integers(0, 3),                      # YCoCg variant (internal numbering)
    st.booleans(), st.booleans(),           # split_alpha (BC3 only), split_colour
    st.one_of(st.integers(0, 70), st.integers(2040, 2060), st.integers(0, 9000)),   # blocks: tiny, around a tile, anything
    st.integers(0, 15), st.integers(0, 15),  # byte misalignment of the device input / output pointers
    st.integers(0, 2**32 - 1),              # data seed
)


@settings(max_examples=120, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(case)
def test_device_transform_equals_oracle_and_round_trips(dlt, torch, c):
    fmt, variant, sa, sc, nb, mis_in, mis_out, seed = c
    sa = sa and fmt == 3
    s = make_settings(dlt, fmt, variant, sa, sc)
    nbytes = nb * (8 if fmt == 1 else 16)
    data = np.random.default_rng(seed).integers(0, 256, nbytes, dtype=np.uint8)
    guard = 64
    d_in = torch.zeros(nbytes + 16 + guard, dtype=torch.uint8, device="cuda")
    d_out = torch.full((nbytes + 16 + 2 * guard,), 0xA5, dtype=torch.uint8, device="cuda")
    d_back = torch.zeros(nbytes + 16, dtype=torch.uint8, device="cuda")
    d_in[mis_in:mis_in + nbytes] = torch.from_numpy(data).cuda()
    out_off = guard + mis_out
    dlt.transform_device(fmt, d_in.data_ptr() + mis_in, d_out.data_ptr() + out_off, nbytes, s)
    dlt.untransform_device(fmt, d_out.data_ptr() + out_off, d_back.data_ptr() + mis_in, nbytes, s)
    torch.cuda.synchronize()
    got = d_out.cpu().numpy()
    expect = oracle.transform(fmt, data, variant, sa, sc)
    assert np.array_equal(got[out_off:out_off + nbytes], expect)
    assert (got[:out_off] == 0xA5).all() and (got[out_off + nbytes:] == 0xA5).all()   # guard bands untouched
    assert np.array_equal(d_back.cpu().numpy()[mis_in:mis_in + nbytes], data)


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(case, st.integers(1, 5))
def test_block_range_shards_compose(dlt, torch, c, nshards):
    """Any cut of the block range into shards (the multi-GPU partition) writes exactly the single-call image."""
    fmt, variant, sa, sc, nb, _mi, _mo, seed = c
    sa = sa and fmt == 3
    s = make_settings(dlt, fmt, variant, sa, sc)
    bpb = 8 if fmt == 1 else 16
    data = np.random.default_rng(seed).integers(0, 256, nb * bpb, dtype=np.uint8)
    d_in = torch.from_numpy(data).cuda() if nb else torch.zeros(0, dtype=torch.uint8, device="cuda")
    d_out = torch.zeros(max(nb * bpb, 1), dtype=torch.uint8, device="cuda")
    cuts = sorted(np.random.default_rng(seed ^ 0x5EED).integers(0, nb + 1, nshards - 1).tolist())
    bounds = [0, *cuts, nb]
    for a, b in zip(bounds[:-1], bounds[1:]):
        dlt.transform_device_range(fmt, d_in.data_ptr() + a * bpb, d_out.data_ptr(), nb, a, b - a, s)
    torch.cuda.synchronize()
    assert np.array_equal(d_out.cpu().numpy()[:nb * bpb], oracle.transform(fmt, data, variant, sa, sc))


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(st.integers(0, 70_000), st.integers(0, 255), st.integers(0, 2**32 - 1), st.sampled_from(["random", "few", "runs"]))
def test_estimator_equals_oracle(dlt, torch, n, alphabet, seed, kind):
    """The LTU-semantics estimate of a device-resident byte range equals the oracle restatement (any length, any
    alphabet: flat, skewed and random streams stress different parts of the partition / runs / resolve pipeline)."""
    rng = np.random.default_rng(seed)
    if kind == "random":
        data = rng.integers(0, alphabet + 1, n, dtype=np.uint8)
    elif kind == "few":
        data = rng.choice(np.array([0, alphabet, 255 - alphabet], np.uint8), n, p=[0.9, 0.07, 0.03]) if n else np.zeros(0, np.uint8)
    else:
        data = np.repeat(rng.integers(0, alphabet + 1, n // 8 + 1, dtype=np.uint8), rng.integers(1, 16, n // 8 + 1))[:n]
    d = torch.from_numpy(np.ascontiguousarray(data)).cuda() if n else torch.zeros(1, dtype=torch.uint8, device="cuda")
    assert dlt.ltu_estimate_device(d.data_ptr(), len(data)) == oracle.ltu_estimate(data)
