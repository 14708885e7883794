#!/usr/bin/env python
"""Summarise an ncu report of one kernel: key raw metrics + per-instruction stall samples (needs -lineinfo / --import-source on).
    python tools/ncu_src.py report.ncu-rep [min_share=0.006]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
share = float(sys.argv[2]) if len(sys.argv) > 2 else 0.006
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "smsp__inst_executed.min", "smsp__inst_executed.max",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "sm__inst_executed_pipe_lsu.sum", "l1tex__lsu_writeback_active.sum", "smsp__inst_executed_pipe_lsu.sum"]
for r in rows[2:]:
    print(r[hdr.index("Kernel Name")][:90])
    for w in want:
        if w in hdr:
            print("  ", w, r[hdr.index(w)])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, body = rows[hi], rows[hi + 1:]
ia, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[isamp]) for r in body)
print("total samples", tot)
agg = collections.Counter()
for k, r in enumerate(body):
    s = int(r[isamp])
    for i, h in stall:
        if r[i] not in ("", "0"):
            agg[h] += int(r[i])
    if s > tot * share:
        st = sorted([(int(r[i]), h) for i, h in stall if r[i] not in ("", "0")], reverse=True)[:2]
        print(f"{k:5d} {r[ia].strip()[:58]:58s} {s:6d} {r[iex]:>10s} {st}")
print(agg.most_common(8))
