#!/usr/bin/env python
"""Measures the host-pointer (reference-facing) path: GB/s per direction of the host link for one
transform + one untransform of a BC1 payload, with pinned and with pageable caller buffers, next to
the raw pinned H2D / D2H / bidirectional copy bandwidth of the box.  Tuning knobs come from the
environment (DLTCUDA_CHUNK_MIB, DLTCUDA_STAGES, DLTCUDA_ZEROCOPY)."""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import dxt_lossless_transform_b200 as dlt  # noqa: E402
from dxt_lossless_transform_b200 import synth  # noqa: E402


def link_bandwidth(nbytes: int) -> dict:
    h_a = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_b = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    d_b = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    out = {}

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return nbytes * reps / (time.perf_counter() - t0) / 1e9

    out["h2d_gbs"] = timed(lambda: d_a.copy_(h_a, non_blocking=True))
    out["d2h_gbs"] = timed(lambda: h_b.copy_(d_b, non_blocking=True))

    def both():
        with torch.cuda.stream(s1):
            d_a.copy_(h_a, non_blocking=True)
        with torch.cuda.stream(s2):
            h_b.copy_(d_b, non_blocking=True)

    out["bidir_gbs_per_direction"] = timed(both)
    return out


def main():
    nbytes = int(float(os.environ.get("PROBE_GIB", "1")) * (1 << 30))
    torch.cuda.set_device(0)
    res = {"env": {k: v for k, v in os.environ.items() if k.startswith("DLTCUDA_")}, "bytes": nbytes}
    res["link"] = link_bandwidth(nbytes)
    data = synth.random_blocks(1, nbytes // 8, seed=5)
    s = dlt.Bc1TransformSettings()
    for kind in ("pinned", "pageable"):
        if kind == "pinned":
            bi, bt, bb = dlt.alloc_pinned(nbytes), dlt.alloc_pinned(nbytes), dlt.alloc_pinned(nbytes)
            a_in, a_t, a_back = bi.array, bt.array, bb.array
        else:
            a_in, a_t, a_back = np.empty(nbytes, np.uint8), np.empty(nbytes, np.uint8), np.empty(nbytes, np.uint8)
        a_in[:] = data
        dlt.transform_bc1_with_settings(a_in, a_t, s)  # warm-up
        reps = 3
        t0 = time.perf_counter()
        for _ in range(reps):
            dlt.transform_bc1_with_settings(a_in, a_t, s)
        t_fwd = (time.perf_counter() - t0) / reps
        t0 = time.perf_counter()
        for _ in range(reps):
            dlt.untransform_bc1_with_settings(a_t, a_back, s)
        t_inv = (time.perf_counter() - t0) / reps
        assert np.array_equal(a_back, a_in)
        res[kind] = {"transform_gbs_per_direction": nbytes / t_fwd / 1e9, "untransform_gbs_per_direction": nbytes / t_inv / 1e9}
    # small-payload latency (64 KiB texture), pinned
    small = 64 << 10
    bi, bt = dlt.alloc_pinned(small), dlt.alloc_pinned(small)
    bi.array[:] = data[:small]
    dlt.transform_bc1_with_settings(bi.array, bt.array, s)
    t0 = time.perf_counter()
    for _ in range(200):
        dlt.transform_bc1_with_settings(bi.array, bt.array, s)
    res["latency_64KiB_us"] = (time.perf_counter() - t0) / 200 * 1e6
    print(json.dumps(res))


if __name__ == "__main__":
    main()
