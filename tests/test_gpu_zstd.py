"""transform_bcN_auto with the zstd estimator: the candidates are transformed on the GPU and
compressed concurrently on host threads; the choice must be what the reference's serial loop would pick — the first
minimum, in test order, of real zstd sizes over the candidates' endpoint streams — and the output the winner's bytes."""
import zlib
from pathlib import Path

import numpy as np
import pytest

import oracle
import zstd_ref

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not zstd_ref.available(), reason="no libzstd on this machine")]
GOLDEN = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def dlt():
    import dxt_lossless_transform_b200 as m

    return m


def payload(fmt):
    return np.frombuffer(zlib.decompress((GOLDEN / f"r2-256-bc{fmt}.payload.zlib").read_bytes()), np.uint8).copy()


def expected(dlt, fmt, data, use_all, level):
    """The reference's loop (bc1 transform_auto.rs:230-262, bc2 :226-264, bc3 :226-285) with libzstd as the estimator."""
    best, best_size, sizes = None, None, []
    n = len(data)
    for cand in dlt.auto_candidates(fmt, use_all):
        variant, split_colour = cand.decorrelation_mode, cand.split_colour_endpoints
        split_alpha = cand.split_alpha_endpoints if fmt == 3 else False
        t = oracle.transform(fmt, data, int(variant), split_alpha, split_colour)
        if fmt == 1:
            total = zstd_ref.compressed_size(t[: n // 2], level)
        elif fmt == 2:
            total = zstd_ref.compressed_size(t[n // 2: n // 2 + n // 4], level)
        else:
            nb = n // 16
            total = zstd_ref.compressed_size(t[: 2 * nb], level) + zstd_ref.compressed_size(t[n // 2: n // 2 + 4 * nb], level)
        sizes.append(total)
        if best_size is None or total < best_size:
            best, best_size = (cand, t), total
    return best[0], best[1], sizes


@pytest.mark.parametrize("fmt", [1, 2, 3])
@pytest.mark.parametrize("use_all", [False, True])
def test_auto_with_zstd_picks_the_reference_choice(dlt, fmt, use_all):
    from dxt_lossless_transform_b200 import synth

    est = dlt.ZStandardSizeEstimation.new_fast()
    fn = {1: dlt.transform_bc1_auto, 2: dlt.transform_bc2_auto, 3: dlt.transform_bc3_auto}[fmt]
    for data in (payload(fmt), synth.texture_blocks(fmt, 20_001, seed=7 + fmt), synth.random_blocks(fmt, 333, seed=5)):
        out = np.empty_like(data)
        got = fn(data, out, dlt.Bc1EstimateSettings(est, use_all))
        want_settings, want, _sizes = expected(dlt, fmt, data, use_all, 1)
        assert got == want_settings
        assert np.array_equal(out, want)


def test_stable_auto_builder_with_zstd_level_3(dlt):
    data = payload(1)
    out = np.empty_like(data)
    b = dlt.Bc1AutoTransformBuilder(dlt.ZStandardSizeEstimation.new_default()).use_all_decorrelation_modes(True)
    manual = b.transform(data, out)
    want_settings, want, _ = expected(dlt, 1, data, True, 3)
    assert manual.get_settings() == want_settings
    assert np.array_equal(out, want)
    back = np.empty_like(data)
    manual.untransform(out, back)
    assert np.array_equal(back, data)


def test_zstd_and_generic_callback_paths_agree(dlt):
    """The concurrent path and the one-candidate-at-a-time callback path (a Python estimator calling the same libzstd)
    must make the same choice."""
    data = payload(3)
    seen = []

    def estimate(arr):
        seen.append(len(arr))
        return zstd_ref.compressed_size(arr, 1)

    out_a, out_b = np.empty_like(data), np.empty_like(data)
    a = dlt.transform_bc3_auto(data, out_a, dlt.Bc1EstimateSettings(dlt.ZStandardSizeEstimation.new_fast(), True))
    b = dlt.transform_bc3_auto(data, out_b, dlt.Bc1EstimateSettings(
        dlt.CallbackSizeEstimator(estimate, max_compressed_size=zstd_ref.compress_bound), True))
    assert a == b and np.array_equal(out_a, out_b)
    assert len(seen) == 32   # 16 candidates x 2 ranges
