#!/usr/bin/env python
"""BASELINE.json configs[4]: a 64 GiB batch of 16 MiB payloads alternating BC1 / BC3 (default settings),
streamed end to end (pinned host -> GPU(s) -> pinned host) with dltcuda_transform_batch[_multi_gpu].

Host RAM need not hold 64 GiB: the batch cycles a pinned pool (default 4 GiB in + 4 GiB out), i.e. the
same pool payloads are submitted repeatedly until the requested total has gone through; every pool
payload is verified against the oracle once (sampled subset) and by a round trip.

    python tests/probes/batch_probe.py [total_GiB=64] [pool_GiB=4] [num_gpus=all]
"""
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

PAYLOAD = 16 << 20


def main():
    total_gib = float(sys.argv[1]) if len(sys.argv) > 1 else 64.0
    pool_gib = float(sys.argv[2]) if len(sys.argv) > 2 else 4.0
    ngpu = int(sys.argv[3]) if len(sys.argv) > 3 else torch.cuda.device_count()
    npool = max(2, int(pool_gib * (1 << 30)) // PAYLOAD)
    pin_in, pin_out = dlt.alloc_pinned(npool * PAYLOAD), dlt.alloc_pinned(npool * PAYLOAD)
    items = []
    for i in range(npool):
        fmt = 1 if i % 2 == 0 else 3
        nb = PAYLOAD // (8 if fmt == 1 else 16)
        src = pin_in.array[i * PAYLOAD:(i + 1) * PAYLOAD]
        src[:] = synth.random_blocks(fmt, nb, seed=synth.BASE_SEED + 5, first_block=i * nb)
        s = dlt.Bc1TransformSettings() if fmt == 1 else dlt.Bc3TransformSettings()
        items.append((fmt, src, pin_out.array[i * PAYLOAD:(i + 1) * PAYLOAD], s))
    devices = list(range(ngpu))
    dlt.transform_batch(items, devices=devices)  # warm-up + verification pass
    for i in list(range(0, npool, max(1, npool // 8)))[:8]:
        fmt, src, dst, s = items[i]
        args = (int(s.decorrelation_mode), bool(getattr(s, "split_alpha_endpoints", False)), bool(s.split_colour_endpoints))
        assert np.array_equal(dst, oracle.transform(fmt, np.asarray(src), *args, threads=8)), i
    rounds = max(1, int(round(total_gib * (1 << 30) / (npool * PAYLOAD))))
    t0 = time.perf_counter()
    for _ in range(rounds):
        dlt.transform_batch(items, devices=devices)
    dt = time.perf_counter() - t0
    moved = rounds * npool * PAYLOAD
    # and the way back for one pool pass (round trip identity)
    back = dlt.alloc_pinned(npool * PAYLOAD)
    inv = [(fmt, dst, back.array[i * PAYLOAD:(i + 1) * PAYLOAD], s) for i, (fmt, src, dst, s) in enumerate(items)]
    t1 = time.perf_counter()
    dlt.transform_batch(inv, untransform=True, devices=devices)
    dt_inv = time.perf_counter() - t1
    assert np.array_equal(back.array, pin_in.array)
    print(json.dumps({
        "workload": f"{moved / (1 << 30):.0f} GiB batch: {rounds} x {npool} payloads of 16 MiB alternating BC1/BC3, default settings, "
                    f"pinned pool of {npool * PAYLOAD >> 20} MiB cycled",
        "gpus": ngpu, "transform_input_gbs": moved / dt / 1e9, "seconds": dt,
        "untransform_input_gbs_one_pool_pass": npool * PAYLOAD / dt_inv / 1e9,
        "bytes_counted": "payload bytes in (an equal amount comes back)"}))


if __name__ == "__main__":
    main()
