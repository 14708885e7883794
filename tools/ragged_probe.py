#!/usr/bin/env python
"""Odd block counts (unaligned streams -> RAGGED transform kernel) at 1 GiB scale, BC1/BC2/BC3: GB/s and, under ncu,
the DRAM traffic of the ragged kernels."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import dxt_lossless_transform_b200 as dlt  # noqa: E402


def main():
    torch.cuda.set_device(0)
    stream = torch.cuda.current_stream().cuda_stream
    for fmt, s in ((1, dlt.Bc1TransformSettings()), (2, dlt.Bc2TransformSettings()), (3, dlt.Bc3TransformSettings())):
        bpb = 8 if fmt == 1 else 16
        for delta in (0, 3):
            nbytes = (1 << 30) - delta * bpb
            d_in = torch.randint(0, 256, (nbytes,), dtype=torch.uint8, device="cuda")
            d_out = torch.empty_like(d_in)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            for _ in range(2):
                dlt.transform_device(fmt, d_in.data_ptr(), d_out.data_ptr(), nbytes, s, stream)
            ev[0].record()
            for _ in range(5):
                dlt.transform_device(fmt, d_in.data_ptr(), d_out.data_ptr(), nbytes, s, stream)
            ev[1].record()
            torch.cuda.synchronize()
            ms = ev[0].elapsed_time(ev[1]) / 5
            print(json.dumps({"format": fmt, "odd": bool(delta), "ms": ms, "gbs": 2 * nbytes / ms / 1e6}), flush=True)


if __name__ == "__main__":
    main()
