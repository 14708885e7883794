"""ZStandardSizeEstimation behind the DltSizeEstimator vtable: CPU tests — the callbacks are host
code, so everything except the search itself runs without a GPU.  Mirrors the reference crate's tests
(extensions/compressors/dxt-lossless-transform-zstd/src/lib.rs tests: level validation, empty / null input -> 0,
max_compressed_size, compressible data shrinks) and pins the sizes against libzstd called directly."""
import ctypes as C

import numpy as np
import pytest

import dxt_lossless_transform_b200 as dlt
from dxt_lossless_transform_b200 import _native as N

import zstd_ref

pytestmark = pytest.mark.skipif(not zstd_ref.available(), reason="no libzstd on this machine")


def test_level_validation_matches_the_reference():
    for bad in (0, -1, 23, 100):
        with pytest.raises(dlt.InvalidLevel):
            dlt.ZStandardSizeEstimation(bad)
        assert not N.lib().dltzstd_new_size_estimator(bad)
    for ok in (1, 3, 22):
        assert dlt.ZStandardSizeEstimation(ok).compression_level == ok
    assert dlt.ZStandardSizeEstimation.new_fast().compression_level == 1
    assert dlt.ZStandardSizeEstimation.new_default().compression_level == 3
    assert dlt.ZStandardSizeEstimation.new_best().compression_level == 22
    N.lib().dltzstd_free_size_estimator(None)   # null-safe


def test_library_version_is_reported():
    assert dlt.ZStandardSizeEstimation.library_version() == zstd_ref.version() >= 10400


def test_max_compressed_size_is_compress_bound():
    e = dlt.ZStandardSizeEstimation.new_fast()
    assert e.max_compressed_size(0) == 0
    for n in (1, 100, 4096, 1 << 20, (1 << 30) + 7):
        assert e.max_compressed_size(n) == zstd_ref.compress_bound(n)


def test_empty_and_null_inputs_estimate_zero():
    e = dlt.ZStandardSizeEstimation.new_fast()
    assert e.estimate_compressed_size(np.zeros(0, np.uint8)) == 0
    c = e.c_estimator().contents
    out = C.c_size_t(123)
    scratch = (C.c_uint8 * 64)()
    assert c.estimate_compressed_size(c.context, None, 100, scratch, 64, C.byref(out)) == 0 and out.value == 0
    assert c.estimate_compressed_size(None, None, 0, scratch, 64, C.byref(out)) == 1       # null context
    assert c.estimate_compressed_size(c.context, None, 0, scratch, 64, None) == 1          # null out
    assert c.max_compressed_size(None, 10, C.byref(out)) == 1


@pytest.mark.parametrize("level", [1, 3, 9])
def test_sizes_equal_libzstd_called_with_the_reference_parameters(level):
    rng = np.random.default_rng(level)
    e = dlt.ZStandardSizeEstimation(level)
    cases = [
        np.zeros(1000, np.uint8),
        rng.integers(0, 256, 50_000, dtype=np.uint8),
        np.tile(rng.integers(0, 256, 97, dtype=np.uint8), 700),
        (np.arange(300_000) // 7 % 251).astype(np.uint8),
        np.frombuffer(b"a", np.uint8),
    ]
    for data in cases:
        assert e.estimate_compressed_size(data) == zstd_ref.compressed_size(data, level)
    assert e.estimate_compressed_size(cases[0]) < 100   # compressible data shrinks (lib.rs tests)


def test_a_too_small_output_buffer_is_an_estimator_error_not_a_crash():
    e = dlt.ZStandardSizeEstimation.new_fast()
    c = e.c_estimator().contents
    data = np.random.default_rng(0).integers(0, 256, 10_000, dtype=np.uint8)
    scratch = (C.c_uint8 * 16)()
    out = C.c_size_t(0)
    assert c.estimate_compressed_size(c.context, data.ctypes.data, data.size, scratch, 16, C.byref(out)) == 3
