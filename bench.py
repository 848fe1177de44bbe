#!/usr/bin/env python
"""bench.py -- POA windows/sec of the B200-native engine (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (config.workload): BASELINE.json configs[1] -- "single B200: 10k windows, 500 bp x 32 reads,
15% ONT error, banded width 256" -- synthetic windows (racon_gpu_b200/windows.py), per
GPU (weak scaling).  One step = one pass of the hot path over the whole batch of windows.

  value      windows/s, whole job, inputs ALREADY RESIDENT IN HBM: K back-to-back launches of the POA
             kernel timed with CUDA events on the launching stream, max over ranks.
  e2e        the same metric through the public host API (api.Polisher.polish == racon's GPU window
             scheduler) with HOST buffers: pinned staging + H2D + kernel + D2H + coverage trim every
             step (+ the NCCL gather of the consensus to rank 0 when N > 1).
  roofline   the DP-fill kernel against the measured HBM peak: algorithmic bytes = DP cells of the
             256-column band x 2 B x 2 (one write + one read), SURVEY.md 8(d) / BASELINE.md 4.
  cpu_baseline  the reference's own CPU path (oracle/_ref: racon::Window + spoa AVX2) timed on this
             box's host cores over a bounded sample of the same windows (rank 0, N=1).

`--impl reference` times that CPU path as the job itself (all host threads, bounded sample per step).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

M, X, G = 3, -5, -4
WORKLOADS = {
    # name: (windows per GPU, backbone length, reads per window, error rate, banded)
    "A_banded": (10000, 500, 32, 0.15, True),   # BASELINE configs[1]  (headline)
    "A_full": (10000, 500, 32, 0.15, False),    # BASELINE configs[2]
    "C_small": (100, 500, 8, 0.05, False),      # BASELINE configs[0] shape
}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.thr = index, [], None, None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        self.thr = threading.Thread(target=self._read, daemon=True)
        self.thr.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.thr.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic(workload: str):
    """dram bytes per launch from the committed `ncu --set full` capture, if one matches."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        rec = json.load(open(path))
        if workload in rec:
            r = rec[workload]
            return float(r["dram_bytes_per_launch"]) if isinstance(r, dict) else float(r)
    return None


def cpu_reference_rate(batch, L, seconds_target, threads):
    """Times the reference's CPU path on a bounded sample; returns the cpu_baseline dict."""
    from oracle_lib import Oracle, Ref, processing_order
    ref = Ref()
    est = 40.0 * threads  # windows/s guess for config A (~50 windows/s/thread)
    n = int(max(threads * 2, min(batch.n_windows, seconds_target * est)))
    sample = batch.slice(0, n)
    t0 = time.perf_counter()
    if ref.available:
        ref.polish(sample, M, X, G, tgs=True, trim=True, threads=threads, window_length=L)
        kind = "reference"
    else:  # the restatement, when oracle/_ref did not travel
        from racon_gpu_b200 import api
        Oracle().polish(sample, api.processing_order(sample), M, X, G, tgs=True, trim=True, threads=threads)
        kind = "port"
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "windows/s", "cores": threads, "kind": kind,
            "sample": f"first {n} windows of the workload, {dt:.1f} s, racon::Window::generate_consensus + spoa "
                      f"(kNW {M}/{X}/{G}), one engine per thread"}


def algorithmic_bytes_per_window(batch, banded, sample=48):
    """sum over reads of rows x cols DP cells (rows from the oracle's trace) x 2 B x 2."""
    from oracle_lib import Oracle
    from racon_gpu_b200 import api
    sub = batch.slice(0, min(sample, batch.n_windows))
    _, _, _, st = Oracle().polish(sub, api.processing_order(sub), M, X, G, tgs=False, trim=False,
                                  threads=min(16, os.cpu_count() or 1), want_stats=True)
    cells = st[:, 4].mean() if banded else st[:, 2].mean()
    return float(cells) * 4.0, float(cells)


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path as the job itself."""
    if rank != 0:
        return
    from racon_gpu_b200.windows import synth_windows
    nwin, L, D, err, banded = WORKLOADS[args.workload]
    threads = os.cpu_count() or 1
    per_step = int(max(threads * 2, min(nwin, 8.0 * 40.0 * threads)))  # ~8 s of CPU work per step
    batch = synth_windows(per_step, L, D, err, seed=args.seed)
    from oracle_lib import Ref, Oracle
    ref = Ref()
    def step():
        if ref.available:
            ref.polish(batch, M, X, G, tgs=True, trim=True, threads=threads, window_length=L)
        else:
            from racon_gpu_b200 import api
            Oracle().polish(batch, api.processing_order(batch), M, X, G, tgs=True, trim=True, threads=threads)
    for _ in range(min(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    print(json.dumps({
        "impl": "reference", "metric": "POA windows/sec", "value": value, "unit": "windows/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {L} bp x {D} reads, {err:.0%} error, m{M}/x{X}/g{G}; CPU path is unbanded "
                               f"(spoa has no band); bounded sample of {per_step} windows per step"},
        "cpu_baseline": {"value": value, "unit": "windows/s", "cores": threads,
                         "kind": "reference" if ref.available else "port",
                         "sample": f"{per_step} windows per step x {args.steps} steps"},
        "e2e": {"value": value, "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="A_banded", choices=sorted(WORKLOADS))
    ap.add_argument("--windows", type=int, default=0, help="override windows per GPU")
    ap.add_argument("--seed", type=int, default=12345)
    ap.add_argument("--batches", type=int, default=4, help="batch processors per GPU for the e2e leg (racon -c)")
    ap.add_argument("--rounds", type=float, default=1.0, help="e2e leg: launches per batch processor per step (chunk = windows / (batches x rounds))")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from racon_gpu_b200 import api
    from racon_gpu_b200.shard import gather_consensus
    from racon_gpu_b200.windows import synth_windows

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible (the engine has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    warmup = max(args.warmup, 3)

    nwin, L, D, err, banded = WORKLOADS[args.workload]
    if args.windows:
        nwin = args.windows
    batch = synth_windows(nwin, L, D, err, seed=args.seed + 1000003 * rank)  # each rank its own windows

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    free_b, _ = torch.cuda.mem_get_info()
    mem = int(min(0.6 * free_b, 64 << 30))

    # ---------------- value: kernel with inputs resident in HBM -------------------------------------
    stream = torch.cuda.Stream(device=device)
    pb = api.PoaBatch(device=local_rank, stream=stream.cuda_stream, max_gpu_mem=mem, banded=banded,
                      gap=G, mismatch=X, match=M)
    n_added, _ = pb.add_windows(batch)
    if n_added != nwin:
        raise SystemExit(f"bench.py: batch accepted {n_added}/{nwin} windows; raise the memory budget")
    pb.upload()
    info = pb.info()
    sampler = ClockSampler(local_rank)
    sampler.start()  # started before the warm-up: nvidia-smi needs ~1 s to deliver its first sample
    with torch.cuda.stream(stream):
        for _ in range(warmup):
            pb.launch()
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        t0 = time.perf_counter()
        for a, b in ev:
            a.record(stream)
            pb.launch()
            b.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        clocks = sampler.stop()
    kernel_ms = [a.elapsed_time(b) for a, b in ev]
    dev_ms = float(sum(kernel_ms))
    pb.download()
    cons_k, _, status_k = pb.get_consensus()
    n_fail = int((status_k != 0).sum())
    pb.close()

    # ---------------- e2e: host buffers through the public API -------------------------------------
    pol = api.Polisher(devices=[local_rank], batches_per_device=args.batches, mem_per_batch=mem // args.batches,
                       banded=banded, match=M, mismatch=X, gap=G)
    chunk = max(256, int(-(-nwin // (args.batches * args.rounds))))  # windows per launch of a batch processor
    out = None
    for _ in range(warmup):
        out_t = pol.polish(batch, tgs=True, trim=True, max_windows_per_round=chunk)
        out = (out_t[0], out_t[1], out_t[2].astype(np.uint8), out_t[3])
        if world > 1:
            gather_consensus(out_t[0], out_t[1], device)
    barrier()
    t0 = time.perf_counter()
    e2e_launches = 0
    for _ in range(args.steps):
        cons, clen, polished, status = pol.polish(batch, tgs=True, trim=True, max_windows_per_round=chunk, out=out)
        e2e_launches += pol.last["kernel_launches"]
        if world > 1:
            gather_consensus(cons, clen, device)
    barrier()
    e2e_s = time.perf_counter() - t0
    h2d, d2h = pol.last["h2d_bytes"], pol.last["d2h_bytes"]
    n_unpolished = int((~polished).sum())
    pol.close()

    # ---------------- reduce over ranks ---------------------------------------------------------------
    t = torch.tensor([dev_ms, e2e_s, wall, float(n_fail + n_unpolished)], dtype=torch.float64, device=device)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dev_ms, e2e_s, wall = float(tmax[0]), float(tmax[1]), float(tmax[2])
        failures = int(tsum[3])
    else:
        failures = n_fail + n_unpolished
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    total_windows = nwin * world
    value = total_windows * args.steps / (dev_ms / 1e3)
    e2e_value = total_windows * args.steps / e2e_s
    bytes_per_window, cells = algorithmic_bytes_per_window(batch, banded)
    peak, peak_src = measured_hbm_peak()
    launch_s = (dev_ms / 1e3) / args.steps
    achieved = bytes_per_window * nwin / launch_s / 1e9
    result = {
        "metric": "POA windows/sec", "value": value, "unit": "windows/s", "n_gpus": world, "steps": args.steps,
        "warmup": warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int16", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {nwin} windows per GPU, {L} bp x {D} reads, {err:.0%} error, "
                               f"{'static band 256' if banded else 'full band'}, m{M}/x{X}/g{G} (BASELINE.json configs[1])",
                   "windows_per_gpu": nwin, "l2": f"inputs larger than L2: {info['n_slots']} resident window workspaces x "
                                                  f"{info['slot_bytes'] / 1048576:.1f} MB are rewritten every step",
                   "resident_warps": info["n_slots"], "blocks_per_sm": info["blocks_per_sm"],
                   "e2e_batches_per_gpu": args.batches, "failed_windows": failures},
        "e2e": {"value": e2e_value, "unit": "windows/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "api.Polisher.polish (b200poa_polisher_polish): staging + H2D + kernel + D2H + trim"
                       + (" + NCCL gather" if world > 1 else "")},
        "gpu_launches": args.steps + e2e_launches,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": recorded_traffic(args.workload), "kernel": "poa_window_kernel",
                     "algorithmic_bytes_per_window": bytes_per_window, "dp_cells_per_window": cells,
                     "launch_ms": launch_s * 1e3, "peak_source": peak_src},
        "wall_s_timed_region": wall,
    }
    if world == 1 and not args.no_cpu_baseline:
        result["cpu_baseline"] = cpu_reference_rate(batch, L, 15.0, os.cpu_count() or 1)
    print(json.dumps(result))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
