#!/usr/bin/env python
"""How well does the LTU-semantics estimate rank transform candidates compared with real zstd?

For every payload of tools/texture_corpus.py and every BC1 candidate (comprehensive order, 8 candidates) this takes the
GPU LTU estimate (dltcuda_transform_auto_device) and the real zstd level 1 / level 3 sizes of the candidate's endpoint
streams (ZStandardSizeEstimation, the system libzstd), and reports the regret of choosing by each estimator, measured in
zstd-1 bytes: size(chosen) / size(best) - 1.  Also times transform_bc1_auto with the zstd estimator: concurrent path vs a
one-candidate-at-a-time callback calling the same libzstd.  Needs a GPU; writes one JSON object to stdout."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))

import torch  # noqa: E402

import dxt_lossless_transform_b200 as dlt  # noqa: E402
from dxt_lossless_transform_b200 import synth  # noqa: E402
from texture_corpus import corpus  # noqa: E402


def main():
    torch.cuda.set_device(0)
    z1, z3 = dlt.ZStandardSizeEstimation(1), dlt.ZStandardSizeEstimation(3)
    cands = dlt.auto_candidates(1, True)
    rows = []
    for name, data in corpus():
        n = len(data)
        d_in = torch.from_numpy(data).cuda()
        d_out = torch.empty_like(d_in)
        _best, ltu = dlt.transform_auto_device(1, d_in.data_ptr(), d_out.data_ptr(), n, True)
        s1, s3 = [], []
        out = np.empty_like(data)
        for c in cands:
            dlt.transform_bc1_with_settings(data, out, c)
            s1.append(z1.estimate_compressed_size(out[: n // 2]))
            s3.append(z3.estimate_compressed_size(out[: n // 2]))
        rows.append({"payload": name, "stream_bytes": n // 2, "ltu": [int(x) for x in ltu], "zstd1": s1, "zstd3": s3})

    def regret(key):
        r, exact = [], 0
        for row in rows:
            truth = np.array(row["zstd1"], float)
            k = int(np.argmin(np.array(row[key], float))) if key else 0
            r.append(truth[k] / truth.min() - 1)
            exact += truth[k] == truth.min()
        return {"mean_regret_pct": 100 * float(np.mean(r)), "max_regret_pct": 100 * float(np.max(r)), "exact_choices": int(exact),
                "payloads": len(rows)}

    result = {
        "zstd_version": dlt.ZStandardSizeEstimation.library_version(),
        "candidates": [repr(c) for c in cands],
        "regret_in_zstd1_bytes": {"no search (first candidate)": regret(None), "ltu (GPU estimator)": regret("ltu"),
                                  "zstd level 3": regret("zstd3"), "zstd level 1": regret("zstd1")},
        "rows": rows,
    }

    # timing: transform_bc1_auto with the zstd estimator, host buffers
    timing = []
    for mib in (8, 64):
        data = synth.texture_blocks(1, (mib << 20) // 8, seed=3)
        out = np.empty_like(data)
        for use_all in (False, True):
            rec = {"payload_mib": mib, "use_all": use_all}
            for label, est in (("concurrent_ms", z1),
                               ("serial_callback_ms", dlt.CallbackSizeEstimator(lambda a: z1.estimate_compressed_size(a),
                                                                                max_compressed_size=z1.max_compressed_size))):
                opts = dlt.Bc1EstimateSettings(est, use_all)
                dlt.transform_bc1_auto(data, out, opts)
                t0 = time.perf_counter()
                got = dlt.transform_bc1_auto(data, out, opts)
                rec[label] = (time.perf_counter() - t0) * 1e3
                rec[label.replace("_ms", "_choice")] = repr(got)
            timing.append(rec)
    result["auto_with_zstd1_timing"] = timing
    print(json.dumps(result))


if __name__ == "__main__":
    main()
