#!/usr/bin/env python
"""bench.py — BCn transform+untransform throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline workload (BASELINE.json configs[1]): 1 GiB of device-resident synthetic BC1 blocks per GPU; one STEP
is transform -> untransform for every one of the 8 (decorrelation mode x split_colour_endpoints)
settings = 16 kernel launches.  GB/s counts the bytes the path reads plus the bytes it writes
(2*len per transform, 2*len per untransform), the convention of MEASURED_PEAKS.json; the reference's
own convention (input bytes per second) is half of it and is printed as `input_gbs`.

With N > 1 every rank owns one 1 GiB block-range shard of an N GiB payload (weak scaling, no
collective on the data path; torch.distributed is only used for the barrier and the max over ranks).

The same JSON line also carries the other BASELINE configs, measured in the same run (extra keys, never the headline):
  formats           configs[2]: BC2 / BC3, every settings combination, 1 GiB, per-kernel GB/s
  odd_n             1 GiB - 3 blocks (odd block count: unaligned stream bases, the ragged kernel), BC1 / BC2 / BC3
  cfg1_latency      configs[0]: 8 MiB BC1 (Variant1, split): device-resident and host-buffer latency
  determine_best_settings   configs[3]: transform_bc1_auto with the LTU-semantics estimator
  mixed_batch       configs[4]: 64 GiB of 16 MiB payloads alternating BC1 / BC3 through dltcuda_transform_batch, split over the N ranks
  e2e / e2e_pageable        the reference-facing C ABI on page-locked / ordinary host buffers
  host              host-side ceilings measured in the same run: synchronised aggregate link figures, all-core memcpy

`--impl reference` times the CPU restatement of the reference (oracle/, all host threads, AVX-512 where the host
has it — what the reference itself would run) on the same workload and prints the same `config` — the reference
is Rust and cannot be built in this image.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

GIB = 1 << 30
FMT = 1
BPB = 8
METRIC = "BCn transform+untransform GB/s per B200 (% HBM roofline); 1/2/4/8-GPU GB/s"
WORKLOAD = ("BC1 all decorrelation modes x split_colour_endpoints, transform->untransform, "
            "1 GiB device-resident synthetic BC1 blocks per GPU")
# Bc1TransformSettings::all_combinations order (bc1 settings.rs:68-77): variant-major, split first
SETTING_LABELS = [f"{v}/{s}" for v in ("NONE", "Variant1", "Variant2", "Variant3") for s in ("split", "nosplit")]


def config_dict(shard_bytes: int) -> dict:
    """The workload description BOTH arms print (same keys, same values: the driver compares them)."""
    return {"workload": WORKLOAD, "bytes_per_gpu": shard_bytes, "settings": SETTING_LABELS,
            "bytes_counted": "read+written (4*len per round trip)",
            "l2": "inputs (1 GiB) are larger than L2 (126 MB); no flush needed",
            "sharding": "contiguous block range per rank, host-side offset prefix, no collective"}


def shard_bytes_of(args) -> int:
    b = args.gib_per_gpu * GIB if args.gib_per_gpu >= 1 else int(args.gib_per_gpu * GIB)
    return int(b) // (2048 * BPB) * (2048 * BPB)


def measured_peak_hbm() -> tuple[float, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])), mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def label(s) -> str:
    return f"{s.decorrelation_mode.name}/{'split' if s.split_colour_endpoints else 'nosplit'}"


def label3(s) -> str:
    return (f"{s.decorrelation_mode.name}/{'split_a' if getattr(s, 'split_alpha_endpoints', False) else 'a'}/"
            f"{'split_c' if s.split_colour_endpoints else 'c'}")


# --------------------------------------------------------------------------------------------------
# CPU arm (oracle port) — used by `--impl reference` and by the cpu_baseline object of the GPU arm
# --------------------------------------------------------------------------------------------------
def cpu_isa() -> str:
    import oracle

    L = oracle.lib()
    L.orc_cpu_baseline_isa.restype = C.c_int
    isa = L.orc_cpu_baseline_isa()
    return {5: "AVX-512 (vpermt2d / vpermt2w / vpermt2b) BC1, BC2 and BC3 paths",
            2: "explicit AVX2 BC1 path"}.get(isa, "word-wise scalar paths")


def cpu_roundtrip(sample_bytes: int, threads: int, steps: int, warmup: int) -> dict:
    """The 8-settings round trip of the headline workload on the host cores: GB/s (read+written bytes), ms / step.
    Outputs are touched before the timed region (no first-touch page faults inside it) and `warmup` full steps run first."""
    import oracle
    from dxt_lossless_transform_b200 import synth

    data = synth.random_blocks(FMT, sample_bytes // BPB, seed=synth.BASE_SEED + 2)
    t, back = np.zeros_like(data), np.zeros_like(data)
    L = oracle.lib()
    combos = [(v, sc) for v in (0, 1, 2, 3) for sc in (1, 0)]

    def step():
        for v, sc in combos:
            L.orc_bcn_run_mt(FMT, 0, data.ctypes.data, t.ctypes.data, data.size, v, 0, sc, threads)
            L.orc_bcn_run_mt(FMT, 1, t.ctypes.data, back.ctypes.data, data.size, v, 0, sc, threads)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    assert np.array_equal(back, data)
    traffic = 4 * data.size * len(combos)
    return {"gbs": traffic / dt / 1e9, "ms_per_step": dt * 1e3, "warmup_steps": warmup, "steps": steps}


def cpu_auto(payload: np.ndarray, use_all: bool, threads: int) -> dict:
    """transform_bc1_auto of the oracle: one payload on one thread (what the reference does for ONE texture) and
    `threads` payloads at once, one per thread (what its CLI does for a directory: rayon, one file per task,
    tools/dxt-lossless-transform-cli/src/commands/transform/mod.rs:154-176)."""
    import oracle

    oracle.auto(FMT, payload[: 1 << 20], use_all)   # warm the library
    t0 = time.perf_counter()
    _, choice = oracle.auto(FMT, payload, use_all)
    one = time.perf_counter() - t0
    copies = [payload.copy() for _ in range(threads)]
    with ThreadPoolExecutor(threads) as pool:
        list(pool.map(lambda p: oracle.auto(FMT, p[: 1 << 20], use_all), copies))   # thread start-up, scratch first touch
        t0 = time.perf_counter()
        list(pool.map(lambda p: oracle.auto(FMT, p, use_all), copies))
        many = time.perf_counter() - t0
    return {"one_payload_one_thread_ms": one * 1e3, "one_thread_input_mibs": payload.size / one / (1 << 20),
            "all_core_payloads": threads, "all_core_ms": many * 1e3,
            "all_core_input_mibs": threads * payload.size / many / (1 << 20), "choice": choice}


def host_memcpy_gbs(nbytes: int, threads: int) -> float:
    """All-core memcpy of this box (GB/s of bytes copied; DRAM traffic is twice that, three times with write-allocate)."""
    src, dst = np.ones(nbytes, np.uint8), np.zeros(nbytes, np.uint8)
    cuts = [nbytes * i // threads for i in range(threads + 1)]

    def part(i):
        np.copyto(dst[cuts[i]:cuts[i + 1]], src[cuts[i]:cuts[i + 1]])

    with ThreadPoolExecutor(threads) as pool:
        list(pool.map(part, range(threads)))
        t0 = time.perf_counter()
        for _ in range(3):
            list(pool.map(part, range(threads)))
        return 3 * nbytes / (time.perf_counter() - t0) / 1e9


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    shard_bytes = shard_bytes_of(args)
    steps, warmup = max(1, args.steps), max(1, args.warmup)
    r = cpu_roundtrip(shard_bytes, cores, steps, warmup)
    one = cpu_roundtrip(min(shard_bytes, 256 << 20), 1, 2, 1)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": r["gbs"], "unit": "GB/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u16", "data": "synthetic", "input_gbs": r["gbs"] / 2,
        "config": config_dict(shard_bytes),
        "note": "CPU restatement (oracle/) of the reference path on the host cores of this box; the Rust reference cannot be "
                "built here.  One step = the full 1 GiB, 8 settings x (transform + untransform); with N > 1 the rate is "
                "the same (the job is N such shards).",
        "cpu_baseline": {"value": r["gbs"], "unit": "GB/s", "cores": cores, "kind": "port",
                         "single_thread_value": one["gbs"],
                         "sample": f"{shard_bytes >> 20} MiB BC1, 8 settings x (transform+untransform) per step, {warmup} warm-up + "
                                   f"{steps} timed steps, C port of the reference ({cpu_isa()}), {cores} threads by block range; "
                                   f"single thread: 256 MiB, 1 warm-up + 2 timed steps"},
        "e2e": {"value": r["gbs"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def pin_rank_to_gpu_numa_node(torch, local_rank: int) -> dict:
    """With several ranks on one box, keep each rank's threads (and therefore the first touch of its pinned pools and the
    library's staging threads) on the CPUs next to its GPU.  Reads the PCI device's local_cpulist from sysfs; a VM
    that hides the topology (numa_node = -1) leaves the affinity alone."""
    info = {"applied": False}
    try:
        p = torch.cuda.get_device_properties(local_rank)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        base = Path("/sys/bus/pci/devices") / bdf
        node = int((base / "numa_node").read_text().strip())
        info.update({"pci": bdf, "numa_node": node})
        if node < 0:
            return info
        cpus = set()
        for part in (base / "local_cpulist").read_text().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus |= set(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            info.update({"applied": True, "cpus": len(cpus)})
    except Exception as e:  # noqa: BLE001 — topology is best effort
        info["error"] = str(e)[:80]
    return info


def run_gpu_arm(args) -> None:
    import torch
    import torch.distributed as dist

    import dxt_lossless_transform_b200 as dlt
    from dxt_lossless_transform_b200 import sharding, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    all_cpus = os.sched_getaffinity(0)
    affinity = pin_rank_to_gpu_numa_node(torch, local_rank) if world > 1 else {"applied": False}
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    shard_bytes = shard_bytes_of(args)
    shard_blocks = shard_bytes // BPB
    first_block = shard_blocks * rank  # the host-side prefix of shard offsets: all the "exchange" there is
    settings = list(dlt.Bc1TransformSettings.all_combinations())
    assert [label(s) for s in settings] == SETTING_LABELS

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce(value: float, op) -> float:
        t = torch.tensor([value], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=op)
        return float(t.item())

    # Synthetic input: counter-based SplitMix64, so each rank generates exactly its own block range.
    host_in = dlt.alloc_pinned(shard_bytes)
    gen_chunk = 64 << 20
    for off in range(0, shard_bytes, gen_chunk):
        nb = min(gen_chunk, shard_bytes - off) // BPB
        host_in.array[off:off + nb * BPB] = synth.random_blocks(FMT, nb, seed=synth.BASE_SEED + 2,
                                                                first_block=first_block + off // BPB)
    d_in = torch.empty(shard_bytes, dtype=torch.uint8, device="cuda")
    d_in.copy_(torch.from_numpy(host_in.array), non_blocking=False)
    d_t = torch.empty_like(d_in)
    d_back = torch.empty_like(d_in)
    stream = torch.cuda.current_stream().cuda_stream

    def stream_ptrs(s):
        # this rank's compact shard image: stream k at shard_blocks * prefix_k
        ptrs, prefix = [], 0
        for w in sharding.stream_widths(FMT, s):
            ptrs.append(d_t.data_ptr() + shard_blocks * prefix)
            prefix += w
        return ptrs + [0] * (6 - len(ptrs))

    def launch_pair(s):
        dlt.transform_device_streams(FMT, d_in.data_ptr(), stream_ptrs(s), shard_blocks, s, stream)
        dlt.untransform_device_streams(FMT, stream_ptrs(s), d_back.data_ptr(), shard_blocks, s, stream)

    # ---- warm-up + correctness of what is about to be timed
    for _ in range(max(args.warmup, 3)):
        for s in settings:
            launch_pair(s)
    torch.cuda.synchronize()
    assert torch.equal(d_back, d_in), "round trip mismatch"

    # ---- timed region: K steps, CUDA events on the launching stream, one event per kernel boundary
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = dlt.kernel_launch_count()
    n_kernels = 2 * len(settings)
    events = [[torch.cuda.Event(enable_timing=True) for _ in range(n_kernels + 1)] for _ in range(args.steps)]
    barrier()
    for k in range(args.steps):
        events[k][0].record()
        for i, s in enumerate(settings):
            dlt.transform_device_streams(FMT, d_in.data_ptr(), stream_ptrs(s), shard_blocks, s, stream)
            events[k][2 * i + 1].record()
            dlt.untransform_device_streams(FMT, stream_ptrs(s), d_back.data_ptr(), shard_blocks, s, stream)
            events[k][2 * i + 2].record()
    barrier()
    gpu_launches = dlt.kernel_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    total_ms = reduce(events[0][0].elapsed_time(events[-1][-1]), dist.ReduceOp.MAX)
    ms_per_step = total_ms / args.steps
    traffic_per_step = 4 * shard_bytes * len(settings) * world  # whole job
    value = traffic_per_step / (ms_per_step * 1e-3) / 1e9

    # per-kernel durations (this rank), dominant kernel -> roofline
    per_kernel: dict[str, list[float]] = {}
    for k in range(args.steps):
        for i, s in enumerate(settings):
            per_kernel.setdefault("transform " + label(s), []).append(events[k][2 * i].elapsed_time(events[k][2 * i + 1]))
            per_kernel.setdefault("untransform " + label(s), []).append(events[k][2 * i + 1].elapsed_time(events[k][2 * i + 2]))
    avg = {name: sum(v) / len(v) for name, v in per_kernel.items()}
    med = {name: statistics.median(v) for name, v in per_kernel.items()}
    # The dominant kernel is elected by its MEDIAN launch (one preempted launch — the clock sampler's NVML queries land inside
    # the timed region by design — must not elect an arbitrary kernel); its figure is the plain average over the K launches, and
    # the launches far off the median are counted beside it.
    dominant = max(med, key=med.get)
    outliers = sum(1 for name, v in per_kernel.items() for x in v if x > 1.5 * med[name])
    peak, peak_src = measured_peak_hbm()
    algo_bytes = 2 * shard_bytes  # 16 B per BC1 block: 8 read + 8 written
    achieved = algo_bytes / (avg[dominant] * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": None, "peak_source": peak_src + ", burst copy figure",
        "algorithmic_bytes_per_launch": algo_bytes, "avg_launch_ms": avg[dominant], "median_launch_ms": med[dominant],
        "launches_timed": args.steps * n_kernels, "launches_over_1p5x_median": outliers,
        "frac_of_nominal_8TBs": achieved / 8000.0,
        "per_kernel_gbs": {name: algo_bytes / (ms * 1e-3) / 1e9 for name, ms in sorted(avg.items())},
    }
    ncu_traffic = ROOT / "profiles" / "dram_traffic.json"
    if ncu_traffic.exists():
        try:
            doc = json.loads(ncu_traffic.read_text())
            per_gib = doc.get("per_kernel", {}).get(dominant)   # only a measurement of THIS kernel counts (else null)
            # ncu capture of the same kernels on 1 GiB (profiles/README.md); traffic is linear in the payload
            roofline["traffic"] = per_gib * shard_bytes / GIB if per_gib else None
            roofline["traffic_source"] = doc.get("source") if per_gib else "no ncu capture of this kernel (profiles/dram_traffic.json)"
        except Exception:
            pass

    if args.profile_mode:
        if rank == 0:
            print(json.dumps({"profile_mode": True, "ms_per_step": ms_per_step, "value": value, "roofline": roofline}))
        return

    def time_launch(fn, reps=3) -> float:
        """Average device time (ms) of one launch of fn() on the current stream, after two warm-up launches."""
        fn(), fn()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(reps):
            fn()
        ev[1].record()
        torch.cuda.synchronize()
        return ev[0].elapsed_time(ev[1]) / reps

    # ---- BASELINE configs[2] and the odd-block-count case, device resident, rank 0 (kernel figures do not depend on N)
    formats, odd_n, cfg1 = None, None, None
    if rank == 0 and not args.no_extras:
        formats, odd_n = {}, {}
        classes = {1: dlt.Bc1TransformSettings, 2: dlt.Bc2TransformSettings, 3: dlt.Bc3TransformSettings}
        for fmt in (2, 3, 1):
            bpb = 8 if fmt == 1 else 16
            combos = list(classes[fmt].all_combinations())
            for tag, nbytes in (("full", GIB), ("odd", GIB - 3 * bpb)):
                if tag == "full" and fmt == 1:
                    continue   # the headline itself
                # d_in's random bytes are valid blocks of any format; the odd case re-uses the first nbytes
                fwd, inv = {}, {}
                for s in combos:
                    name = label3(s) if fmt == 3 else label(s)
                    ms = time_launch(lambda: dlt.transform_device(fmt, d_in.data_ptr(), d_t.data_ptr(), nbytes, s, stream))
                    fwd[name] = 2 * nbytes / (ms * 1e-3) / 1e9
                    if tag == "full":
                        ms = time_launch(lambda: dlt.untransform_device(fmt, d_t.data_ptr(), d_back.data_ptr(), nbytes, s, stream))
                        inv[name] = 2 * nbytes / (ms * 1e-3) / 1e9
                        assert torch.equal(d_back[:nbytes], d_in[:nbytes]), (fmt, name)
                rec = {"transform_gbs_min": min(fwd.values()), "transform_gbs_max": max(fwd.values()),
                       "transform_frac_of_measured_min": min(fwd.values()) / peak, "transform_gbs": fwd}
                if inv:
                    rec.update({"untransform_gbs_min": min(inv.values()), "untransform_gbs_max": max(inv.values()), "untransform_gbs": inv})
                (formats if tag == "full" else odd_n)[f"bc{fmt}"] = rec
        formats["workload"] = "1 GiB device-resident, every settings combination, reference layout, 3 timed launches each (GB/s read+written)"
        odd_n["workload"] = "1 GiB - 3 blocks (odd block count: stream bases only 2-byte aligned), transform, every settings combination"
        # ---- configs[0]: 8 MiB BC1, Variant1 + split — a latency config (L2 resident once warm)
        n8 = 8 << 20
        s0 = dlt.Bc1TransformSettings()
        cfg1 = {"workload": "8 MiB BC1 (Variant1, split), the reference's criterion size; L2-resident once warm: latency, not HBM",
                "transform_device_us": 1e3 * time_launch(lambda: dlt.transform_device(1, d_in.data_ptr(), d_t.data_ptr(), n8, s0, stream), 20),
                "untransform_device_us": 1e3 * time_launch(lambda: dlt.untransform_device(1, d_t.data_ptr(), d_back.data_ptr(), n8, s0, stream), 20)}
        h8, o8 = dlt.alloc_pinned(n8), dlt.alloc_pinned(n8)
        h8.array[:] = host_in.array[:n8]
        dlt.transform_bc1_with_settings(h8.array, o8.array, s0)
        t0 = time.perf_counter()
        for _ in range(10):
            dlt.transform_bc1_with_settings(h8.array, o8.array, s0)
        cfg1["transform_host_pinned_us"] = (time.perf_counter() - t0) / 10 * 1e6
        del h8, o8

    # ---- the host side of this box, measured in the same run with ALL ranks active at the same time
    def link_gbs() -> dict:
        nbytes = min(shard_bytes, GIB)
        h_a, h_b = torch.from_numpy(host_in.array[:nbytes]), torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

        def timed(fn, reps=3):
            fn()
            barrier()           # every rank starts its copies together: the figure is what the link gives under full load
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            barrier()
            return nbytes * reps / dt / 1e9

        def both():
            with torch.cuda.stream(s1):
                d_t[:nbytes].copy_(h_a, non_blocking=True)
            with torch.cuda.stream(s2):
                h_b.copy_(d_back[:nbytes], non_blocking=True)

        mine = {"h2d_gbs": timed(lambda: d_t[:nbytes].copy_(h_a, non_blocking=True)),
                "d2h_gbs": timed(lambda: h_b.copy_(d_back[:nbytes], non_blocking=True)),
                "bidir_gbs_per_direction": timed(both)}
        agg = {k: reduce(v, dist.ReduceOp.SUM) for k, v in mine.items()}
        return {"this_rank": mine, "all_ranks_sum": agg, "note": "ranks measure simultaneously (barrier before every phase)"}

    host_link = link_gbs()

    # ---- end to end through the reference-facing C ABI: pinned HOST buffers, H2D + D2H in the timed region
    host_t, host_back = dlt.alloc_pinned(shard_bytes), dlt.alloc_pinned(shard_bytes)
    dlt.set_device(local_rank)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))

    def e2e_step():
        for s in settings:
            dlt.transform_bc1_with_settings(host_in.array, host_t.array, s)
            dlt.untransform_bc1_with_settings(host_t.array, host_back.array, s)

    e2e_step()  # warm-up (allocates the context's slots)
    assert np.array_equal(host_back.array[:1 << 20], host_in.array[:1 << 20])
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s_per_step = reduce(time.perf_counter() - t0, dist.ReduceOp.MAX) / e2e_steps
    assert np.array_equal(host_back.array, host_in.array), "e2e round trip mismatch"
    per_dir = 2 * shard_bytes * len(settings) * world / e2e_s_per_step / 1e9
    e2e = {
        "value": traffic_per_step / e2e_s_per_step / 1e9, "unit": "GB/s",
        "h2d_bytes_per_step": 2 * shard_bytes * len(settings) * world,
        "d2h_bytes_per_step": 2 * shard_bytes * len(settings) * world,
        "steps": e2e_steps, "ms_per_step": e2e_s_per_step * 1e3,
        "host_link_gbs_per_direction": per_dir,
        "frac_of_bidirectional_link": per_dir / host_link["all_ranks_sum"]["bidir_gbs_per_direction"],
        "api": "dltbc1core_transform / dltbc1core_untransform on pinned host buffers",
    }

    # ---- the same call on ORDINARY (pageable) caller memory — what the reference's callers pass (Vec<u8>, mmap)
    pg_in = np.empty(shard_bytes, np.uint8)
    pg_in[:] = host_in.array
    pg_t, pg_back = np.full(shard_bytes, 1, np.uint8), np.full(shard_bytes, 1, np.uint8)   # touched: no first-touch page faults in the timed region
    s0 = dlt.Bc1TransformSettings()
    dlt.transform_bc1_with_settings(pg_in[: 64 << 20], pg_t[: 64 << 20], s0)   # allocates the staging slots
    barrier()
    t0 = time.perf_counter()
    dlt.transform_bc1_with_settings(pg_in, pg_t, s0)
    t_fwd = reduce(time.perf_counter() - t0, dist.ReduceOp.MAX)
    barrier()
    t0 = time.perf_counter()
    dlt.untransform_bc1_with_settings(pg_t, pg_back, s0)
    t_inv = reduce(time.perf_counter() - t0, dist.ReduceOp.MAX)
    assert np.array_equal(pg_back, pg_in), "pageable round trip mismatch"
    e2e_pageable = {"value": 4 * shard_bytes * world / (t_fwd + t_inv) / 1e9, "unit": "GB/s (read+written, as `e2e`)",
                    "transform_gbs_per_direction": shard_bytes * world / t_fwd / 1e9,
                    "untransform_gbs_per_direction": shard_bytes * world / t_inv / 1e9,
                    "workload": "1 GiB BC1 per rank, Variant1/split, transform then untransform, pageable numpy buffers"}
    del pg_in, pg_t, pg_back

    # ---- BASELINE configs[4]: 64 GiB of 16 MiB payloads alternating BC1 / BC3, default settings, through
    # dltcuda_transform_batch, split over the ranks; the pinned pools of the e2e leg are cycled
    mixed = None
    if not args.no_extras:
        payload = 16 << 20
        npool = shard_bytes // payload
    if not args.no_extras and npool >= 2:
        items = []
        for i in range(npool):
            fmt = 1 if i % 2 == 0 else 3
            st = dlt.Bc1TransformSettings() if fmt == 1 else dlt.Bc3TransformSettings()
            items.append((fmt, host_in.array[i * payload:(i + 1) * payload], host_t.array[i * payload:(i + 1) * payload], st))
        dlt.transform_batch(items)   # warm-up
        total = int(args.batch_gib * GIB)
        rounds = max(1, total // world // (npool * payload))
        barrier()
        t0 = time.perf_counter()
        for _ in range(rounds):
            dlt.transform_batch(items)
        dt = reduce(time.perf_counter() - t0, dist.ReduceOp.MAX)
        moved = rounds * npool * payload * world
        inv_items = [(fmt, dst, host_back.array[i * payload:(i + 1) * payload], st) for i, (fmt, src, dst, st) in enumerate(items)]
        dlt.transform_batch(inv_items, untransform=True)
        assert np.array_equal(host_back.array[:npool * payload], host_in.array[:npool * payload]), "mixed batch round trip"
        mixed = {"workload": f"{moved / GIB:.0f} GiB: {rounds} x {npool} payloads of 16 MiB per rank, alternating BC1 / BC3, default settings, "
                             f"page-locked pool of {npool * payload >> 20} MiB per rank cycled, dltcuda_transform_batch",
                 "input_gbs": moved / dt / 1e9, "seconds": dt, "n_gpus": world,
                 "bytes_counted": "payload bytes in (an equal amount comes back)",
                 "frac_of_bidirectional_link": moved / dt / 1e9 / host_link["all_ranks_sum"]["bidir_gbs_per_direction"]}

    # ---- BASELINE configs[3] beside the headline: determine-best-settings (GPU LTU estimator + search) on a 64 MiB
    # BC1 payload, device resident and through host buffers, with the CPU oracle's search: one payload on one thread
    # (the reference's own unit of work) and one payload per core (its CLI's unit of parallelism)
    auto = None
    if rank == 0 and not args.no_auto:
        import oracle

        os.sched_setaffinity(0, all_cpus)
        cores = len(all_cpus)
        nb = (64 << 20) // BPB
        a_host = synth.texture_blocks(FMT, nb, seed=synth.BASE_SEED + 4)
        a_in = torch.from_numpy(a_host).cuda()
        a_out = torch.empty_like(a_in)
        pin_in, pin_out = dlt.alloc_pinned(a_host.size), dlt.alloc_pinned(a_host.size)
        pin_in.array[:] = a_host
        est = dlt.Bc1EstimateSettings(dlt.LosslessTransformUtilsSizeEstimation(), False)
        auto = {"workload": "transform_bc1_auto, LTU-semantics estimator (PARITY UNPINNED: restated third-party crate), "
                            "64 MiB texture-like BC1 payload", "unit": "ms"}
        for name, use_all in (("fast_k4", False), ("comprehensive_k8", True)):
            dlt.transform_auto_device(FMT, a_in.data_ptr(), a_out.data_ptr(), a_host.size, use_all)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                best, sizes = dlt.transform_auto_device(FMT, a_in.data_ptr(), a_out.data_ptr(), a_host.size, use_all)
            auto[name + "_device_ms"] = (time.perf_counter() - t0) / 5 * 1e3
            auto[name + "_best"] = f"{best.decorrelation_mode.name}/{'split' if best.split_colour_endpoints else 'nosplit'}"
            if not use_all:
                k4_out = a_out.cpu().numpy()
                k4_sizes = list(sizes)
        # the same search on 1 GiB (BASELINE: the estimator at HBM scale)
        big = d_in[:GIB] if shard_bytes >= GIB else None
        if big is not None:
            dlt.transform_auto_device(FMT, big.data_ptr(), d_t.data_ptr(), GIB, False)
            t0 = time.perf_counter()
            dlt.transform_auto_device(FMT, big.data_ptr(), d_t.data_ptr(), GIB, False)
            auto["fast_k4_device_1gib_ms"] = (time.perf_counter() - t0) * 1e3
        est.use_all_decorrelation_modes = False
        dlt.transform_bc1_auto(pin_in.array, pin_out.array, est)
        t0 = time.perf_counter()
        for _ in range(3):
            dlt.transform_bc1_auto(pin_in.array, pin_out.array, est)
        auto["fast_k4_host_e2e_ms"] = (time.perf_counter() - t0) / 3 * 1e3
        cpu = cpu_auto(a_host, False, cores)
        want_out, want = oracle.auto(FMT, a_host, False)
        got = dlt.transform_auto_device(FMT, a_in.data_ptr(), a_out.data_ptr(), a_host.size, False)[0]
        assert (int(got.decorrelation_mode), False, bool(got.split_colour_endpoints)) == want, "GPU and oracle searches differ"
        assert np.array_equal(k4_out, want_out) and k4_sizes == oracle.auto_estimates(FMT, a_host, False)
        auto["cpu_oracle_fast_k4"] = cpu
        auto["cpu_cores"] = cores
        auto["note"] = ("GPU choice, output bytes and every per-candidate estimate equal the oracle's on the full 64 MiB payload; the CPU "
                        "figures are the oracle's search: one payload on one thread (the reference's unit of work) and one payload "
                        "per core at once (its CLI's rayon loop)")

    if rank == 0:
        os.sched_setaffinity(0, all_cpus)
        cores = len(all_cpus)
        sample = 256 << 20
        cpu_all = cpu_roundtrip(sample, cores, 2, 1)
        cpu_one = cpu_roundtrip(sample, 1, 1, 1)
        memcpy_gbs = host_memcpy_gbs(512 << 20, cores)
        # a pageable call moves every payload byte through host DRAM six times (staging copy in: read + write, DMA read, DMA
        # write, staging copy out: read + write); an all-core memcpy moves two bytes per byte copied — three if its stores
        # allocate the destination lines first (numpy / glibc decide by size), hence a range for the ceiling
        e2e_pageable["host_dram_ceiling_gbs_per_direction"] = [memcpy_gbs * 2 / 6, memcpy_gbs * 3 / 6]
        e2e_pageable["frac_of_host_dram_ceiling"] = ([min(e2e_pageable["transform_gbs_per_direction"], e2e_pageable["untransform_gbs_per_direction"]) / c
                                                      for c in (memcpy_gbs * 3 / 6, memcpy_gbs * 2 / 6)] if world == 1 else None)
        out = {
            "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u16", "data": "synthetic", "input_gbs": value / 2,
            "config": config_dict(shard_bytes),
            "roofline": roofline,
            "cpu_baseline": {"value": cpu_all["gbs"], "unit": "GB/s", "cores": cores, "kind": "port",
                             "single_thread_value": cpu_one["gbs"],
                             "sample": f"{sample >> 20} MiB BC1, the same 8-settings round trip, 1 warm-up + 2 timed steps (single thread: "
                                       f"1 + 1), outputs pre-touched, C port of the reference (oracle/, {cpu_isa()}), {cores} threads by block range"},
            "e2e": e2e, "e2e_pageable": e2e_pageable, "gpu_launches": int(gpu_launches), "clocks": clocks,
            "host": {"link": host_link, "all_core_memcpy_gbs": memcpy_gbs, "cpus": cores,
                     "rank_affinity": affinity,
                     "note": "e2e moves every payload byte across the link once per direction AND through host DRAM (DMA read of the "
                             "source, DMA write of the destination); with several GPUs on one box the sum of the links exceeds what host "
                             "DRAM sustains — compare all_ranks_sum with all_core_memcpy_gbs (a copy reads and writes: DRAM traffic is 2-3x)"},
        }
        if formats:
            out["formats"], out["odd_n"], out["cfg1_latency"] = formats, odd_n, cfg1
        if mixed:
            out["mixed_batch"] = mixed
        if auto:
            out["determine_best_settings"] = auto
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--gib-per-gpu", type=float, default=1.0)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--batch-gib", type=float, default=64.0, help="size of the mixed BC1/BC3 batch (BASELINE configs[4]), whole job")
    ap.add_argument("--no-auto", action="store_true", help="skip the determine-best-settings side measurement")
    ap.add_argument("--no-extras", action="store_true", help="skip the BC2/BC3, odd-N, 8 MiB and mixed-batch side measurements")
    ap.add_argument("--profile-mode", action="store_true",
                    help="kernel-only run for ncu: skips the e2e and cpu_baseline legs (never a bench value)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
