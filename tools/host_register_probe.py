#!/usr/bin/env python
"""How fast is cudaHostRegister on this box?  (Would pinning a caller's pageable buffers in place beat the staged copies
of the host path?  Measured: 5-10 GB/s to register plus 17-50 ms per GiB to unregister, against 16-20 GB/s for the
staged path end to end - no.)"""
import time, numpy as np, torch
torch.cuda.init(); torch.zeros(1, device="cuda")
rt = torch.cuda.cudart()
for mib in (64, 1024):
    a = np.ones(mib << 20, np.uint8)
    for rep in range(2):
        t0 = time.perf_counter(); r = rt.cudaHostRegister(a.ctypes.data, a.nbytes, 0); t1 = time.perf_counter()
        u = rt.cudaHostUnregister(a.ctypes.data); t2 = time.perf_counter()
        print(mib, "MiB register %.1f ms (%.1f GB/s) unregister %.1f ms" % ((t1-t0)*1e3, a.nbytes/(t1-t0)/1e9, (t2-t1)*1e3), r, u)
