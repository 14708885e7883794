#!/usr/bin/env python
"""bench.py — BCn transform+untransform throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): 1 GiB of device-resident synthetic BC1 blocks per GPU; one STEP
is transform -> untransform for every one of the 8 (decorrelation mode x split_colour_endpoints)
settings = 16 kernel launches.  GB/s counts the bytes the path reads plus the bytes it writes
(2*len per transform, 2*len per untransform), the convention of MEASURED_PEAKS.json; the reference's
own convention (input bytes per second) is half of it and is printed as `input_gbs`.

With N > 1 every rank owns one 1 GiB block-range shard of an N GiB payload (weak scaling, no
collective on the data path; torch.distributed is only used for the barrier and the max over ranks).

`--impl reference` times the CPU restatement of the reference (oracle/, all host threads) on a
bounded sample of the same workload — the reference is Rust and cannot be built in this image.
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


def all_settings(dlt):
    return list(dlt.Bc1TransformSettings.all_combinations())


def label(s) -> str:
    return f"{s.decorrelation_mode.name}/{'split' if s.split_colour_endpoints else 'nosplit'}"


# --------------------------------------------------------------------------------------------------
# CPU arm (oracle port) — used by `--impl reference` and by the cpu_baseline object of the GPU arm
# --------------------------------------------------------------------------------------------------
def cpu_roundtrip_gbs(sample_bytes: int, threads: int, steps: int, warmup: int) -> tuple[float, float]:
    """GB/s (read+written bytes) and ms/step of the oracle over the 8-settings round trip."""
    import oracle
    from dxt_lossless_transform_b200 import synth

    data = synth.random_blocks(FMT, sample_bytes // BPB, seed=synth.BASE_SEED + 2)
    t, back = np.empty_like(data), np.empty_like(data)
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
    return traffic / dt / 1e9, dt * 1e3


def cpu_isa() -> str:
    import oracle

    L = oracle.lib()
    L.orc_cpu_baseline_uses_avx2.restype = C.c_int
    return "explicit AVX2 BC1 path" if L.orc_cpu_baseline_uses_avx2() else "word-wise scalar BC1 path"


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = 256 << 20
    gbs, ms = cpu_roundtrip_gbs(sample, cores, max(1, args.steps), max(1, min(args.warmup, 2)))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": gbs, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u16", "data": "synthetic", "input_gbs": gbs / 2,
        "config": {"workload": WORKLOAD, "bytes_counted": "read+written (4*len per round trip)",
                   "note": "CPU restatement (oracle/) of the reference path; the Rust reference cannot be built here"},
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": cores, "kind": "port",
                         "sample": f"{sample >> 20} MiB BC1, 8 settings x (transform+untransform) per step, "
                                   f"C port of the reference ({cpu_isa()}), {cores} threads by block range"},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
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
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    shard_bytes = args.gib_per_gpu * GIB if args.gib_per_gpu >= 1 else int(args.gib_per_gpu * GIB)
    shard_bytes = int(shard_bytes) // (2048 * BPB) * (2048 * BPB)
    shard_blocks = shard_bytes // BPB
    total_blocks = shard_blocks * world
    first_block = shard_blocks * rank  # the host-side prefix of shard offsets: all the "exchange" there is
    settings = all_settings(dlt)

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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

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
    total_ms = events[0][0].elapsed_time(events[-1][-1])
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
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
    dominant = max(avg, key=avg.get)
    peak, peak_src = measured_peak_hbm()
    algo_bytes = 2 * shard_bytes  # 16 B per BC1 block: 8 read + 8 written
    achieved = algo_bytes / (avg[dominant] * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": None, "peak_source": peak_src + ", burst copy figure",
        "algorithmic_bytes_per_launch": algo_bytes, "avg_launch_ms": avg[dominant],
        "frac_of_nominal_8TBs": achieved / 8000.0,
        "per_kernel_gbs": {name: algo_bytes / (ms * 1e-3) / 1e9 for name, ms in sorted(avg.items())},
    }
    ncu_traffic = ROOT / "profiles" / "dram_traffic.json"
    if ncu_traffic.exists():
        try:
            doc = json.loads(ncu_traffic.read_text())
            per_gib = doc.get("per_kernel", {}).get(dominant, doc.get("bytes_per_launch_1gib"))
            # ncu capture of the same kernels on 1 GiB (profiles/README.md); traffic is linear in the payload
            roofline["traffic"] = per_gib * shard_bytes / GIB if per_gib else None
            roofline["traffic_source"] = doc.get("source")
        except Exception:
            pass

    if args.profile_mode:
        if rank == 0:
            print(json.dumps({"profile_mode": True, "ms_per_step": ms_per_step, "value": value, "roofline": roofline}))
        return

    # ---- the host link of this box, measured in the same run (denominator of the e2e number)
    def link_gbs():
        nbytes = min(shard_bytes, GIB)
        h_a, h_b = torch.from_numpy(host_in.array[:nbytes]), torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

        def timed(fn, reps=3):
            fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            torch.cuda.synchronize()
            return nbytes * reps / (time.perf_counter() - t0) / 1e9

        def both():
            with torch.cuda.stream(s1):
                d_t[:nbytes].copy_(h_a, non_blocking=True)
            with torch.cuda.stream(s2):
                h_b.copy_(d_back[:nbytes], non_blocking=True)

        return {"h2d_gbs": timed(lambda: d_t[:nbytes].copy_(h_a, non_blocking=True)),
                "d2h_gbs": timed(lambda: h_b.copy_(d_back[:nbytes], non_blocking=True)),
                "bidir_gbs_per_direction": timed(both)}

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
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_s_per_step = float(dt.item()) / e2e_steps
    assert np.array_equal(host_back.array, host_in.array), "e2e round trip mismatch"
    e2e = {
        "value": traffic_per_step / e2e_s_per_step / 1e9, "unit": "GB/s",
        "h2d_bytes_per_step": 2 * shard_bytes * len(settings) * world,
        "d2h_bytes_per_step": 2 * shard_bytes * len(settings) * world,
        "steps": e2e_steps, "ms_per_step": e2e_s_per_step * 1e3,
        "host_link_gbs_per_direction": 2 * shard_bytes * len(settings) * world / e2e_s_per_step / 1e9,
        "host_link_measured": host_link,
        "frac_of_bidirectional_link": (2 * shard_bytes * len(settings) / e2e_s_per_step / 1e9)
        / host_link["bidir_gbs_per_direction"],
        "api": "dltbc1core_transform / dltbc1core_untransform on pinned host buffers",
    }

    # ---- BASELINE configs[3] beside the headline: determine-best-settings (GPU LTU estimator + search) on a 64 MiB
    # BC1 payload, device resident and through host buffers, with the CPU oracle's search on a bounded sample
    auto = None
    if rank == 0 and not args.no_auto:
        import oracle

        nb = (64 << 20) // BPB
        a_host = synth.texture_blocks(FMT, nb, seed=synth.BASE_SEED + 4)
        a_in = torch.from_numpy(a_host).cuda()
        a_out = torch.empty_like(a_in)
        pin_in, pin_out = dlt.alloc_pinned(a_host.size), dlt.alloc_pinned(a_host.size)
        pin_in.array[:] = a_host
        est = dlt.Bc1EstimateSettings(dlt.LosslessTransformUtilsSizeEstimation(), False)
        auto = {"workload": "transform_bc1_auto, LTU-semantics estimator, 64 MiB texture-like BC1 payload", "unit": "ms"}
        for name, use_all in (("fast_k4", False), ("comprehensive_k8", True)):
            dlt.transform_auto_device(FMT, a_in.data_ptr(), a_out.data_ptr(), a_host.size, use_all)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                best, _sizes = dlt.transform_auto_device(FMT, a_in.data_ptr(), a_out.data_ptr(), a_host.size, use_all)
            auto[name + "_device_ms"] = (time.perf_counter() - t0) / 5 * 1e3
            auto[name + "_best"] = f"{best.decorrelation_mode.name}/{'split' if best.split_colour_endpoints else 'nosplit'}"
        est.use_all_decorrelation_modes = False
        dlt.transform_bc1_auto(pin_in.array, pin_out.array, est)
        t0 = time.perf_counter()
        for _ in range(3):
            dlt.transform_bc1_auto(pin_in.array, pin_out.array, est)
        auto["fast_k4_host_e2e_ms"] = (time.perf_counter() - t0) / 3 * 1e3
        small = a_host[: 8 << 20].copy()
        t0 = time.perf_counter()
        want_out, want = oracle.auto(FMT, small, False)
        auto["cpu_oracle_fast_k4_8MiB_sample_ms"] = (time.perf_counter() - t0) * 1e3
        chk = np.zeros_like(small)
        got = dlt.transform_bc1_auto(small, chk, est)
        assert (int(got.decorrelation_mode), False, bool(got.split_colour_endpoints)) == want and np.array_equal(chk, want_out)
        auto["cpu_oracle_fast_k4_64MiB_extrapolated_ms"] = auto["cpu_oracle_fast_k4_8MiB_sample_ms"] * 8
        auto["note"] = "GPU choice and bytes checked against the oracle on the 8 MiB sample; CPU figure is single-thread"

    if rank == 0:
        cores = os.cpu_count() or 1
        sample = 128 << 20
        cpu_all, _ = cpu_roundtrip_gbs(sample, cores, 1, 1)
        cpu_one, _ = cpu_roundtrip_gbs(sample, 1, 1, 0)
        out = {
            "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u16", "data": "synthetic", "input_gbs": value / 2,
            "config": {"workload": WORKLOAD, "bytes_per_gpu": shard_bytes, "settings": [label(s) for s in settings],
                       "bytes_counted": "read+written (4*len per round trip)",
                       "l2": "inputs (1 GiB) are larger than L2 (126 MB); no flush needed",
                       "sharding": "contiguous block range per rank, host-side offset prefix, no collective"},
            "roofline": roofline,
            "cpu_baseline": {"value": cpu_all, "unit": "GB/s", "cores": cores, "kind": "port",
                             "single_thread_value": cpu_one,
                             "sample": f"{sample >> 20} MiB BC1, the same 8-settings round trip, C port of the "
                                       f"reference (oracle/, {cpu_isa()}), {cores} threads by block range"},
            "e2e": e2e, "gpu_launches": int(gpu_launches), "clocks": clocks,
        }
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
    ap.add_argument("--no-auto", action="store_true", help="skip the determine-best-settings side measurement")
    ap.add_argument("--profile-mode", action="store_true",
                    help="kernel-only run for ncu: skips the e2e and cpu_baseline legs (never a bench value)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
