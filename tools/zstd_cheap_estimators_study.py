#!/usr/bin/env python
"""CPU-only companion of tools/zstd_rank_study.py: could a CHEAP GPU-friendly formula — the LTU match
count combined with an order-0 byte histogram of the stream — rank BC1 transform candidates closer to real zstd-1 than the
LTU estimate alone?  Corpus: tools/texture_corpus.py (43 payloads), 8 candidates each, truth = real zstd level 1 size of the
endpoint stream (system libzstd), regret = size(chosen) / size(best) - 1.  Uses the oracle for transforms and match counts
(no GPU).      python tools/zstd_cheap_estimators_study.py"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "tests", ROOT / "tools"):
    sys.path.insert(0, str(p))

import oracle  # noqa: E402
import zstd_ref  # noqa: E402
from texture_corpus import corpus  # noqa: E402

CANDS = [(2, 0), (0, 0), (0, 1), (3, 0), (3, 1), (2, 1), (1, 0), (1, 1)]   # COMPREHENSIVE_TEST_ORDER: (variant, split_colour)


def main():
    assert zstd_ref.available(), "needs libzstd"
    rows = []
    for name, data in corpus():
        n = len(data)
        row = {"n": n // 2, "z1": [], "ltu": [], "h0": []}
        for v, sc in CANDS:
            s = oracle.transform(1, data, v, False, bool(sc))[: n // 2]
            row["z1"].append(zstd_ref.compressed_size(s, 1))
            row["ltu"].append(len(s) - oracle.ltu_matches_params(s, 16, True, 4))
            h = np.bincount(s, minlength=256).astype(float)
            p = h[h > 0] / len(s)
            row["h0"].append(float(-(p * np.log2(p)).sum()))
        rows.append(row)

    def regret(f):
        r = []
        for row in rows:
            truth = np.array(row["z1"], float)
            k = int(np.argmin([f(row, i) for i in range(len(CANDS))]))
            r.append(truth[k] / truth.min() - 1)
        return {"mean_regret_pct": round(100 * float(np.mean(r)), 3), "max_regret_pct": round(100 * float(np.max(r)), 3)}

    out = {"payloads": len(rows), "zstd_version": zstd_ref.version(),
           "no search (first candidate)": regret(lambda r, i: i),
           "ltu": regret(lambda r, i: r["ltu"][i]),
           "order-0 entropy x length": regret(lambda r, i: r["h0"][i] * r["n"]),
           "ltu x order-0 entropy": regret(lambda r, i: r["ltu"][i] * r["h0"][i]),
           "ltu x entropy / 8 + 0.05 x matches": regret(lambda r, i: r["ltu"][i] * r["h0"][i] / 8 + 0.05 * (r["n"] - r["ltu"][i])),
           "ltu x entropy^2": regret(lambda r, i: r["ltu"][i] * r["h0"][i] ** 2)}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
