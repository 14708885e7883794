#!/usr/bin/env python
"""A directory of SMALL textures through dltcuda_transform_batch (manual settings): payloads of 16 KiB .. 1 MiB in pinned
memory, GB/s of payload bytes in (an equal amount comes back) and microseconds per payload."""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import dxt_lossless_transform_b200 as dlt  # noqa: E402
from dxt_lossless_transform_b200 import _native as N  # noqa: E402


def main():
    torch.cuda.set_device(0)
    total = 64 << 20
    pin_in, pin_out = dlt.alloc_pinned(total), dlt.alloc_pinned(total)
    pin_in.array[:] = np.random.default_rng(0).integers(0, 256, total, dtype=np.uint8)
    for fmt, S in ((1, dlt.Bc1TransformSettings), (3, dlt.Bc3TransformSettings)):
        for size in (16 << 10, 64 << 10, 256 << 10, 1 << 20):
            n = total // size
            # the C entry point is timed directly (the ctypes marshalling of thousands of payloads is Python's cost)
            arr = (N.DltcudaPayload * n)()
            base_in, base_out = pin_in.array.ctypes.data, pin_out.array.ctypes.data
            st = dlt.api._dsettings(fmt, S())
            for i in range(n):
                arr[i] = N.DltcudaPayload(base_in + i * size, base_out + i * size, size, st)
            for inverse in (False, True):
                assert N.lib().dltcuda_transform_batch(arr, n, inverse) == 0
                t0 = time.perf_counter()
                for _ in range(3):
                    N.lib().dltcuda_transform_batch(arr, n, inverse)
                dt = (time.perf_counter() - t0) / 3
                print(json.dumps({"format": fmt, "payload_bytes": size, "payloads": n, "untransform": inverse,
                                  "ms": dt * 1e3, "us_per_payload": dt * 1e6 / n, "input_gbs": total / dt / 1e9}), flush=True)
    pin_in.free(), pin_out.free()


if __name__ == "__main__":
    main()
