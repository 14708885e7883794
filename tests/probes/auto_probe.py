#!/usr/bin/env python
"""BASELINE.json configs[3]: determine-best-settings throughput (GPU LTU estimator + search) on
device-resident payloads, next to the CPU oracle port, and a check that both pick the same settings.
Prints one JSON line per (format, payload size, mode)."""
import json
import sys
import time
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
    torch.cuda.set_device(0)
    sizes = [64 << 10, 1 << 20, 8 << 20, 64 << 20, 1 << 30]
    if len(sys.argv) > 1:
        sizes = [int(float(x) * (1 << 20)) for x in sys.argv[1:]]
    for fmt in (1, 3):
        bpb = 8 if fmt == 1 else 16
        for nbytes in sizes:
            nb = nbytes // bpb
            data = synth.texture_blocks(fmt, nb, seed=nbytes % 1000 + fmt)
            d_in = torch.from_numpy(data).cuda()
            d_out = torch.empty_like(d_in)
            for use_all in (False, True):
                torch.cuda.synchronize()
                best, est = dlt.transform_auto_device(fmt, d_in.data_ptr(), d_out.data_ptr(), data.size, use_all)  # warm
                reps = 3 if nbytes <= (64 << 20) else 1
                t0 = time.perf_counter()
                for _ in range(reps):
                    best, est = dlt.transform_auto_device(fmt, d_in.data_ptr(), d_out.data_ptr(), data.size, use_all)
                dt = (time.perf_counter() - t0) / reps
                rec = {"format": fmt, "bytes": nbytes, "use_all": use_all, "gpu_ms": dt * 1e3,
                       "gpu_input_gbs": nbytes / dt / 1e9, "best": str(best)}
                if nbytes <= (64 << 20):
                    t0 = time.perf_counter()
                    want_out, want = oracle.auto(fmt, data, use_all)
                    cdt = time.perf_counter() - t0
                    got = (int(best.decorrelation_mode), bool(getattr(best, "split_alpha_endpoints", False)),
                           bool(best.split_colour_endpoints))
                    rec.update({"cpu_ms": cdt * 1e3, "cpu_input_gbs": nbytes / cdt / 1e9, "same_choice": got == want,
                                "same_bytes": bool(np.array_equal(d_out.cpu().numpy(), want_out)),
                                "same_estimates": est == oracle.auto_estimates(fmt, data, use_all)})
                print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
