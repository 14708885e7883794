#!/usr/bin/env python
"""BASELINE.json configs[3] as a BATCH: 64 payloads ("chunks") per size class through dltcuda_transform_auto_batch
(one upload per payload, one set of estimator launches for all candidates of all payloads) next to 64 single
transform_bc1_auto calls on the same host buffers.  Prints one JSON line per (size, mode)."""
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
    sizes = [64 << 10, 1 << 20, 8 << 20]
    args = [a for a in sys.argv[1:] if not a.startswith("bc")]
    fmt = 3 if "bc3" in sys.argv[1:] else 1   # `bc3` anywhere on the command line: BC3 payloads (K = 8 / 16)
    if args:
        sizes = [int(float(x) * (1 << 20)) for x in args]
    count = 64
    est = dlt.Bc1EstimateSettings(dlt.LosslessTransformUtilsSizeEstimation(), False)
    single = dlt.transform_bc3_auto if fmt == 3 else dlt.transform_bc1_auto
    for nbytes in sizes:
        nb = nbytes // (16 if fmt == 3 else 8)
        pin_in, pin_out = dlt.alloc_pinned(nbytes * count), dlt.alloc_pinned(nbytes * count)
        items = []
        for i in range(count):
            pin_in.array[i * nbytes:(i + 1) * nbytes] = synth.texture_blocks(fmt, nb, seed=i, smooth=[0.2, 1.0, 5.0][i % 3])
            items.append((fmt, pin_in.array[i * nbytes:(i + 1) * nbytes], pin_out.array[i * nbytes:(i + 1) * nbytes]))
        for use_all in (False, True):
            est.use_all_decorrelation_modes = use_all
            dlt.transform_auto_batch(items, use_all)  # warm
            t0 = time.perf_counter()
            best = dlt.transform_auto_batch(items, use_all)
            dt_batch = time.perf_counter() - t0
            batch_out = pin_out.array.copy()
            t0 = time.perf_counter()
            singles = [single(i_, o_, est) for _f, i_, o_ in items]
            dt_single = time.perf_counter() - t0
            same = all(b == s for b, s in zip(best, singles)) and np.array_equal(batch_out, pin_out.array)
            rec = {"format": f"BC{fmt}", "payload_bytes": nbytes, "payloads": count, "use_all": use_all, "batch_ms": dt_batch * 1e3,
                   "batch_input_gbs": nbytes * count / dt_batch / 1e9, "single_calls_ms": dt_single * 1e3,
                   "single_calls_input_gbs": nbytes * count / dt_single / 1e9, "batch_equals_single_calls": bool(same),
                   "distinct_winners": len({str(b) for b in best})}
            if nbytes <= (1 << 20):
                t0 = time.perf_counter()
                want = [oracle.auto(fmt, np.asarray(i_), use_all)[1] for _f, i_, _o in items[:8]]
                rec["cpu_oracle_ms_extrapolated"] = (time.perf_counter() - t0) * 1e3 * count / 8
                rec["same_choice_as_oracle_first8"] = all(
                    (int(b.decorrelation_mode), bool(getattr(b, "split_alpha_endpoints", False)), bool(b.split_colour_endpoints)) == w
                    for b, w in zip(best[:8], want))
            print(json.dumps(rec), flush=True)
        pin_in.free()
        pin_out.free()


if __name__ == "__main__":
    main()
