#!/usr/bin/env python
"""Odd block counts (1 GiB - 3 blocks: every stream base only 2-byte aligned -> the RAGGED transform kernel), every
settings combination of BC2 and BC3 (+ BC1), for whichever build DLT_LIB_PATH points to.  One JSON line:
min / max GB/s (read + written) per format and the per-setting figures.

    DLT_LIB_PATH=build/variants/libB.so python tools/ragged_variants.py [label]
"""
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import dxt_lossless_transform_b200 as dlt  # noqa: E402


def avg_ms(fn, reps=5):
    fn(), fn()
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
    out = {"label": sys.argv[1] if len(sys.argv) > 1 else os.environ.get("DLT_LIB_PATH", "default")}
    d_src = torch.randint(0, 256, (1 << 30,), dtype=torch.uint8, device="cuda")
    d_dst = torch.empty_like(d_src)
    for fmt, cls in ((3, dlt.Bc3TransformSettings), (2, dlt.Bc2TransformSettings), (1, dlt.Bc1TransformSettings)):
        bpb = 8 if fmt == 1 else 16
        nbytes = (1 << 30) - 3 * bpb
        per = {}
        for s in cls.all_combinations():
            ms = avg_ms(lambda: dlt.transform_device(fmt, d_src.data_ptr(), d_dst.data_ptr(), nbytes, s, stream))
            key = f"{s.decorrelation_mode.name}/{'a' if getattr(s, 'split_alpha_endpoints', False) else '-'}{'c' if s.split_colour_endpoints else '-'}"
            per[key] = round(2 * nbytes / ms / 1e6)
        out[f"bc{fmt}_min"], out[f"bc{fmt}_max"] = min(per.values()), max(per.values())
        out[f"bc{fmt}"] = per
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
