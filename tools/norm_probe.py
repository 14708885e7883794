#!/usr/bin/env python
"""experimental::normalize_blocks at 1 GiB: stand-alone normalization pass and the transform with fused normalization,
GB/s of read+written bytes next to the plain transform; plus transform_bc1_auto_with_normalization on 64 MiB."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import torch  # noqa: E402

import dxt_lossless_transform_b200 as dlt  # noqa: E402
from dxt_lossless_transform_b200 import experimental as ex  # noqa: E402
from dxt_lossless_transform_b200 import synth  # noqa: E402
from norm_cases import crafted_blocks  # noqa: E402


def timed(fn, reps=5):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    fn(), fn()
    ev[0].record()
    for _ in range(reps):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / reps


def main():
    torch.cuda.set_device(0)
    stream = torch.cuda.current_stream().cuda_stream
    n = (1 << 30) // 8
    for name, tile in (("crafted (half the blocks normalizable)", crafted_blocks(1 << 20, seed=1)),
                       ("texture-like (few normalizable blocks)", synth.texture_blocks(1, 1 << 20, seed=1))):
        d_in = torch.from_numpy(tile).cuda().repeat(n // (1 << 20))
        d_out = torch.empty_like(d_in)
        nbytes = d_in.numel()
        rec = {"data": name, "bytes": nbytes}
        ms = timed(lambda: dlt.transform_device(1, d_in.data_ptr(), d_out.data_ptr(), nbytes, dlt.Bc1TransformSettings(), stream))
        rec["transform_plain_gbs"] = 2 * nbytes / ms / 1e6
        for mode in (ex.ColorNormalizationMode.Color0Only, ex.ColorNormalizationMode.ReplicateColor):
            ms = timed(lambda: ex.normalize_blocks_device(d_in.data_ptr(), d_out.data_ptr(), nbytes, mode, stream))
            rec[f"normalize_{mode.name}_gbs"] = 2 * nbytes / ms / 1e6
            det = ex.Bc1TransformDetailsWithNormalization(mode, dlt.YCoCgVariant.Variant1, True)
            ms = timed(lambda: ex.transform_bc1_with_normalize_blocks_device(d_in.data_ptr(), d_out.data_ptr(), nbytes, det, stream))
            rec[f"transform_fused_{mode.name}_gbs"] = 2 * nbytes / ms / 1e6
        print(json.dumps(rec), flush=True)
    data = crafted_blocks((64 << 20) // 8, seed=2)
    out = np.zeros_like(data)
    for use_all in (False, True):
        ex.transform_bc1_auto_with_normalization(data, out, use_all)
        t0 = time.perf_counter()
        best = ex.transform_bc1_auto_with_normalization(data, out, use_all)
        print(json.dumps({"auto_with_normalization_64MiB_host_ms": (time.perf_counter() - t0) * 1e3, "use_all": use_all,
                          "best": str(best)}), flush=True)


if __name__ == "__main__":
    main()
