#!/usr/bin/env python
"""A/B helper: times transform/untransform for a handful of settings (1 GiB, aligned) with whatever
library DLT_LIB_PATH points to.  Prints one compact JSON line."""
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import dxt_lossless_transform_b200 as dlt  # noqa: E402


def avg_ms(fn, reps=10):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    torch.cuda.set_device(0)
    stream = torch.cuda.current_stream().cuda_stream
    nbytes = 1 << 30
    out = {"lib": os.environ.get("DLT_LIB_PATH", "default")}
    V = dlt.YCoCgVariant
    cases = {
        1: [dlt.Bc1TransformSettings(V.Variant1, True), dlt.Bc1TransformSettings(V.NONE, True), dlt.Bc1TransformSettings(V.Variant3, False)],
        2: [dlt.Bc2TransformSettings(V.Variant1, True), dlt.Bc2TransformSettings(V.NONE, False)],
        3: [dlt.Bc3TransformSettings(V.Variant1, True, True), dlt.Bc3TransformSettings(V.Variant3, True, False),
            dlt.Bc3TransformSettings(V.Variant1, False, False), dlt.Bc3TransformSettings(V.NONE, False, True)],
    }
    for fmt, settings in cases.items():
        d_in = torch.randint(0, 256, (nbytes,), dtype=torch.uint8, device="cuda")
        d_t, d_b = torch.empty_like(d_in), torch.empty_like(d_in)
        for s in settings:
            t = avg_ms(lambda: dlt.transform_device(fmt, d_in.data_ptr(), d_t.data_ptr(), nbytes, s, stream))
            u = avg_ms(lambda: dlt.untransform_device(fmt, d_t.data_ptr(), d_b.data_ptr(), nbytes, s, stream))
            key = f"bc{fmt}/{s.decorrelation_mode.name}/{'a' if getattr(s, 'split_alpha_endpoints', False) else '-'}{'c' if s.split_colour_endpoints else '-'}"
            out[key] = [round(2 * nbytes / t / 1e6), round(2 * nbytes / u / 1e6)]
        del d_in, d_t, d_b
    print(json.dumps(out))


if __name__ == "__main__":
    main()
