#!/usr/bin/env python
"""One transform configuration at 1 GiB (minus `delta` blocks) for ncu: python tools/ragged_one.py fmt variant sa sc delta"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import dxt_lossless_transform_b200 as dlt  # noqa: E402

fmt, variant, sa, sc, delta = (int(x) for x in sys.argv[1:6])
v = dlt.YCoCgVariant(variant)
s = dlt.Bc3TransformSettings(v, bool(sa), bool(sc)) if fmt == 3 else (dlt.Bc1TransformSettings if fmt == 1 else dlt.Bc2TransformSettings)(v, bool(sc))
bpb = 8 if fmt == 1 else 16
nbytes = (1 << 30) - delta * bpb
torch.cuda.set_device(0)
d_in = torch.randint(0, 256, (nbytes,), dtype=torch.uint8, device="cuda")
d_out = torch.empty_like(d_in)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for _ in range(2):
    dlt.transform_device(fmt, d_in.data_ptr(), d_out.data_ptr(), nbytes, s)
ev[0].record()
for _ in range(5):
    dlt.transform_device(fmt, d_in.data_ptr(), d_out.data_ptr(), nbytes, s)
ev[1].record()
torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / 5
print(fmt, s, delta, "ms", ms, "GB/s", 2 * nbytes / ms / 1e6)
