#!/usr/bin/env python
"""BASELINE.json configs[1..2]: device-resident transform / untransform GB/s (read+written bytes) for
every settings combination of BC1, BC2 and BC3 on 1 GiB of synthetic blocks, with round-trip check.
Also times odd block counts (unaligned stream bases in the reference layout)."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import dxt_lossless_transform_b200 as dlt  # noqa: E402

PEAK = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0


def time_ms(fn, reps=5):
    for _ in range(3):
        fn()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    return min(ev[i].elapsed_time(ev[i + 1]) for i in range(reps)), ev[0].elapsed_time(ev[-1]) / reps


def main():
    torch.cuda.set_device(0)
    stream = torch.cuda.current_stream().cuda_stream
    for fmt, cls in ((1, dlt.Bc1TransformSettings), (2, dlt.Bc2TransformSettings), (3, dlt.Bc3TransformSettings)):
        bpb = 8 if fmt == 1 else 16
        for nbytes, tag in ((1 << 30, "1GiB"), ((1 << 30) - 3 * bpb, "1GiB-3blocks(odd N, unaligned streams)")):
            g = torch.Generator(device="cuda").manual_seed(fmt)
            d_in = torch.randint(0, 256, (nbytes,), dtype=torch.uint8, device="cuda", generator=g)
            d_t, d_back = torch.empty_like(d_in), torch.empty_like(d_in)
            settings = list(cls.all_combinations())
            if "odd" in tag:
                settings = settings[:2] + settings[-2:]
            for s in settings:
                best_f, avg_f = time_ms(lambda: dlt.transform_device(fmt, d_in.data_ptr(), d_t.data_ptr(), nbytes, s, stream))
                best_i, avg_i = time_ms(lambda: dlt.untransform_device(fmt, d_t.data_ptr(), d_back.data_ptr(), nbytes, s, stream))
                ok = bool(torch.equal(d_back, d_in))
                rec = {"format": fmt, "size": tag, "settings": str(s), "roundtrip_ok": ok,
                       "transform_gbs": 2 * nbytes / (avg_f * 1e-3) / 1e9, "untransform_gbs": 2 * nbytes / (avg_i * 1e-3) / 1e9}
                rec["transform_frac_of_measured"] = rec["transform_gbs"] / PEAK
                rec["untransform_frac_of_measured"] = rec["untransform_gbs"] / PEAK
                print(json.dumps(rec), flush=True)
            del d_in, d_t, d_back


if __name__ == "__main__":
    main()
