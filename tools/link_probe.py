#!/usr/bin/env python
"""Host link under different copy shapes: one or several concurrent copies per direction, large or chunked.
    python tools/link_probe.py"""
import json
import time

import torch

torch.cuda.set_device(0)
N = 1 << 30
h_a = torch.empty(N, dtype=torch.uint8).pin_memory()
h_b = torch.empty(N, dtype=torch.uint8).pin_memory()
d_a = torch.empty(N, dtype=torch.uint8, device="cuda")
d_b = torch.empty(N, dtype=torch.uint8, device="cuda")
streams = [torch.cuda.Stream() for _ in range(8)]


def run(h2d_parts, d2h_parts, chunk=None, reps=3):
    def once():
        k = 0
        for direction, parts in (("h2d", h2d_parts), ("d2h", d2h_parts)):
            if parts == 0:
                continue
            per = N // parts
            for p in range(parts):
                s = streams[k % len(streams)]
                k += 1
                with torch.cuda.stream(s):
                    lo, hi = p * per, (p + 1) * per
                    step = chunk or per
                    for c in range(lo, hi, step):
                        e = min(hi, c + step)
                        if direction == "h2d":
                            d_a[c:e].copy_(h_a[c:e], non_blocking=True)
                        else:
                            h_b[c:e].copy_(d_b[c:e], non_blocking=True)
    once()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    return round(N / dt / 1e9, 2)


out = {}
out["h2d x1"] = run(1, 0)
out["d2h x1"] = run(0, 1)
out["h2d x2"] = run(2, 0)
out["d2h x2"] = run(0, 2)
for parts in (1, 2, 4):
    out[f"both x{parts} (per direction)"] = run(parts, parts)
for chunk_mib in (64, 16, 4):
    out[f"both x1, {chunk_mib} MiB chunks"] = run(1, 1, chunk_mib << 20)
    out[f"both x2, {chunk_mib} MiB chunks"] = run(2, 2, chunk_mib << 20)
print(json.dumps(out))
