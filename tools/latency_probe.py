#!/usr/bin/env python
"""Host-buffer latency of one transform / untransform call by payload size (page-locked buffers):
    python tools/latency_probe.py [fmt=1]      (knobs from the environment: DLTCUDA_ZEROCOPY_MAX_KIB, DLTCUDA_CHUNK_MIB, DLTCUDA_RAMP)"""
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import dxt_lossless_transform_b200 as dlt  # noqa: E402
from dxt_lossless_transform_b200 import synth  # noqa: E402

fmt = int(sys.argv[1]) if len(sys.argv) > 1 else 1
torch.cuda.set_device(0)
top = 256 << 20
src, dst, back = dlt.alloc_pinned(top), dlt.alloc_pinned(top), dlt.alloc_pinned(top)
src.array[:] = synth.random_blocks(fmt, top // (8 if fmt == 1 else 16), seed=3)
s = {1: dlt.Bc1TransformSettings, 2: dlt.Bc2TransformSettings, 3: dlt.Bc3TransformSettings}[fmt]()
fwd = {1: dlt.transform_bc1_with_settings, 2: dlt.transform_bc2_with_settings, 3: dlt.transform_bc3_with_settings}[fmt]
inv = {1: dlt.untransform_bc1_with_settings, 2: dlt.untransform_bc2_with_settings, 3: dlt.untransform_bc3_with_settings}[fmt]
res = {"env": {k: v for k, v in os.environ.items() if k.startswith("DLTCUDA_")}, "fmt": fmt, "sizes": {}}
for mib in (0.0625, 0.25, 1, 2, 4, 8, 16, 32, 64, 128, 256):
    n = int(mib * (1 << 20))
    a, b, c = src.array[:n], dst.array[:n], back.array[:n]
    fwd(a, b, s), inv(b, c, s)
    assert np.array_equal(a, c)
    reps = 50 if mib <= 8 else 10
    t0 = time.perf_counter()
    for _ in range(reps):
        fwd(a, b, s)
    t_f = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    for _ in range(reps):
        inv(b, c, s)
    t_i = (time.perf_counter() - t0) / reps
    res["sizes"][str(mib)] = {"transform_us": round(t_f * 1e6, 1), "untransform_us": round(t_i * 1e6, 1),
                              "transform_gbs_per_direction": round(n / t_f / 1e9, 2), "untransform_gbs_per_direction": round(n / t_i / 1e9, 2)}
print(json.dumps(res))
