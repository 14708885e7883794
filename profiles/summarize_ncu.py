#!/usr/bin/env python
"""Turns an ncu report (gpurun_out/*.ncu-rep) into the small text/JSON summaries kept under profiles/.

    python profiles/summarize_ncu.py gpurun_out/prof_r01.ncu-rep r01 [outdir=profiles/]
(run it on the GPU box with outdir=gpurun_out when the .ncu-rep is too big to bring back)
"""
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__block_size",
    "launch__grid_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__waves_per_multiprocessor", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__cycles_elapsed.max", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "smsp__cycles_active.avg", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum",
]


def to_bytes(value: str, unit: str) -> float:
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(value.replace(",", "")) * scale.get(unit, 1)


def main(rep: str, tag: str, outdir: Path = HERE) -> None:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    lines, traffic = [], {}
    for r in data:
        name = r[idx["Kernel Name"]]
        lines.append(name)
        for k in KEEP:
            if k in idx:
                lines.append(f"    {k:70s} {r[idx[k]]} {units[idx[k]]}")
        rd = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
        wr = to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
        traffic[name] = {"dram_read_bytes": rd, "dram_write_bytes": wr, "dram_total_bytes": rd + wr}
        lines.append(f"    {'dram read+write bytes per launch':70s} {rd + wr:.0f}")
    (outdir / f"{tag}_ncu_full_summary.txt").write_text("\n".join(lines) + "\n")
    (outdir / f"{tag}_dram_traffic_per_kernel.json").write_text(json.dumps(traffic, indent=1) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], Path(sys.argv[3]) if len(sys.argv) > 3 else HERE)
