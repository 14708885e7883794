"""World-size-2 `gloo` test of the multi-GPU path's host logic, on CPU.

The data path has no collective: every rank transforms its own contiguous block
range and owns the matching slice of every output stream.  Here each rank computes its range with
the library's shard helper, produces its slices (with the oracle standing in for the GPU kernels,
which need a device), and rank 0 assembles the payload from the (offset, bytes) slices alone — the
"host-side prefix of shard offsets" — and checks it against the single-call result."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, fmt: int, nblocks: int, q):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import torch
    import torch.distributed as dist

    import oracle
    import dxt_lossless_transform_b200 as dlt
    from dxt_lossless_transform_b200 import sharding, synth

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        bpb = 8 if fmt == 1 else 16
        settings = dlt.Bc3TransformSettings() if fmt == 3 else dlt.Bc1TransformSettings()
        args = (int(settings.decorrelation_mode), bool(getattr(settings, "split_alpha_endpoints", False)),
                bool(settings.split_colour_endpoints))
        first, count = sharding.shard_ranges(fmt, nblocks, world)[rank]
        # each rank generates only its own blocks (counter-based generator)
        mine = synth.random_blocks(fmt, count, seed=77, first_block=first)
        full_in = np.zeros(nblocks * bpb, np.uint8)
        full_in[first * bpb:(first + count) * bpb] = mine
        full_out = np.zeros_like(full_in)
        oracle.run_range(fmt, False, full_in, full_out, *args, first, first + count)
        slices = sharding.stream_slices(fmt, settings, nblocks, first, count)
        payload = np.concatenate([full_out[o:o + l] for o, l in slices]) if count else np.zeros(0, np.uint8)
        # rank 0 gathers (slices, bytes): plumbing only, not part of the data path
        gathered = [None] * world
        dist.gather_object((slices, payload.tobytes()), gathered if rank == 0 else None, dst=0)
        if rank == 0:
            out = np.zeros(nblocks * bpb, np.uint8)
            covered = 0
            for sl, raw in gathered:
                buf = np.frombuffer(raw, np.uint8)
                pos = 0
                for o, l in sl:
                    out[o:o + l] = buf[pos:pos + l]
                    pos += l
                    covered += l
            whole = synth.random_blocks(fmt, nblocks, seed=77)
            expect = oracle.transform(fmt, whole, *args)
            q.put((covered == nblocks * bpb, bool(np.array_equal(out, expect))))
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert t.item() == world
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("fmt,nblocks", [(1, 50_001), (3, 9_999), (2, 1_500)])
def test_two_ranks_compose_the_payload_without_a_collective(fmt, nblocks):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, fmt, nblocks, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    covered, equal = q.get(timeout=10)
    assert covered and equal
