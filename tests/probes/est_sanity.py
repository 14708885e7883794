#!/usr/bin/env python
"""Small estimator / search workload for compute-sanitizer runs (memcheck, racecheck, initcheck): every kernel of the
large-input estimator path on a few awkward sizes, checked against the oracle."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import torch  # noqa: E402

import dxt_lossless_transform_b200 as dlt  # noqa: E402
import oracle  # noqa: E402
from dxt_lossless_transform_b200 import synth  # noqa: E402


def main():
    rng = np.random.default_rng(3)
    cases = [np.zeros(70_001, np.uint8), rng.integers(0, 256, 300_007, dtype=np.uint8), rng.integers(0, 3, 41_000, dtype=np.uint8),
             np.tile(np.array([5, 5, 5, 9], np.uint8), 30_000), rng.integers(0, 256, 4_104, dtype=np.uint8)]
    for data in cases:
        d = torch.from_numpy(data).cuda()
        for off in (0, 1):
            got = dlt.ltu_estimate_device(d.data_ptr() + off, data.size - off)
            assert got == oracle.ltu_estimate(data[off:]), (data.size, off, got)
    for fmt in (1, 3):
        data = synth.texture_blocks(fmt, 9_001, seed=fmt)
        d_in = torch.from_numpy(data).cuda()
        d_out = torch.zeros_like(d_in)
        best, sizes = dlt.transform_auto_device(fmt, d_in.data_ptr(), d_out.data_ptr(), data.size, True)
        assert sizes == oracle.auto_estimates(fmt, data, True)
    print("sanity ok")


if __name__ == "__main__":
    main()
