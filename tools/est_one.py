#!/usr/bin/env python
"""One device-resident best-settings search for ncu: python tools/est_one.py [MiB=64] [fmt=1] [use_all=0] [reps=2]"""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import dxt_lossless_transform_b200 as dlt  # noqa: E402
from dxt_lossless_transform_b200 import synth  # noqa: E402

mib = float(sys.argv[1]) if len(sys.argv) > 1 else 64
fmt = int(sys.argv[2]) if len(sys.argv) > 2 else 1
use_all = bool(int(sys.argv[3])) if len(sys.argv) > 3 else False
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
torch.cuda.set_device(0)
nbytes = int(mib * (1 << 20))
data = synth.texture_blocks(fmt, nbytes // (8 if fmt == 1 else 16), seed=11)
d_in = torch.from_numpy(data).cuda()
d_out = torch.empty_like(d_in)
dlt.transform_auto_device(fmt, d_in.data_ptr(), d_out.data_ptr(), data.size, use_all)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(reps):
    best, est = dlt.transform_auto_device(fmt, d_in.data_ptr(), d_out.data_ptr(), data.size, use_all)
print("ms", (time.perf_counter() - t0) / reps * 1e3, best, est)
