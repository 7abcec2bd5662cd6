#!/usr/bin/env python
"""Benchmark of the deephisto_b200 hot path (contract: one JSON line on stdout from rank 0).

Default workload = BASELINE.json configs[1]: `examples.sample_annotated_rnd --torch` -- random 224x224 patches inside
50 synthetic annotation polygons on a 32768 x 32768 synthetic slide, batch 256, fp32 NHWC features in [0,1] + int64
labels + (y,x) coords, exactly what AnnoRegionRndSampler.torch_generator yields. One step = one batch:
    dh_region_sample (Philox draws + exact clip-area acceptance, 256 slots)  ->  dh_gather_normalize (256 patches).

  value     patches/s over all ranks, inputs (slide, polygon tables) resident in HBM, CUDA-event timed, max over ranks
  e2e       the same metric through the public Python API (AnnoRegionRndSampler.torch_generator), every step ending
            with the device->host read of the step's labels and coordinates into pinned memory (the features stay
            in HBM for the consumer CNN; `features_to_host` also reports the PCIe-bound variant)
  roofline  dominant kernel (gather+normalise): algorithmic bytes / CUDA-event time of that kernel vs measured HBM peak
  cpu_baseline / --impl reference: the reference's CPU path (oracle/cpu_pipeline.py restates
            AnnoRegionRndSampler.torch_generator; the reference itself needs psimage + shapely, which do not exist)

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload annotated_rnd|dense|predict]
"""

from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

PS = 224
PATCH_IN = PS * PS * 3                 # 150 528 B uint8 read per patch
SLIDE_HW = (32768, 32768)
N_POLY = 50
BATCH = 256
K_PER_REGION = 4
RI = 0.75
L2_BYTES = 126 * 1024 * 1024


# --------------------------------------------------------------------------------------------------------------------
def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons through NVML while the GPU is under load."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.ok = index, [], False, False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.max_sm = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                util = nv.nvmlDeviceGetUtilizationRates(self.h).gpu
                self.samples.append((sm, reasons, util))
            except Exception:
                pass
            time.sleep(0.02)

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_sm, "reasons": [], "samples": 0}
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
        }
        busy = [s for s in self.samples if s[2] > 0] or self.samples
        clocks = sorted(s[0] for s in busy)
        seen = sorted({n for s in busy for n, bit in names.items() if s[1] & bit})
        return {"sm_mhz": clocks[len(clocks) // 2], "sm_max_mhz": self.max_sm, "reasons": seen, "samples": len(busy)}


# --------------------------------------------------------------------------------------------------------------------
# CPU legs (run in a process that never touched CUDA: they fork worker pools)
def cpu_leg_annotated(steps: int, warmup: int, workers: int | None, budget_s: float | None):
    from oracle import cpu_pipeline, synth
    from deephisto_b200.synthetic import synth_polygons

    H, W = SLIDE_HW
    cores = workers or os.cpu_count()
    path = synth.synth_slide_shared(H, W, 0, workers=cores, return_path=True)
    images = [((H, W), synth_polygons(N_POLY, H, W, seed=0))]
    pipe = cpu_pipeline.AnnotatedRndCPU(path, (H, W, 3), images, layer=1, one_image_for_batch=True, max_workers=cores)

    def run(n_batches, seed):
        t0 = time.perf_counter()
        n = 0
        for f, l, c in pipe.batches(PS, BATCH, n_batches, batches_per_worker=2, k=K_PER_REGION, ri=RI, seed=seed):
            n += f.shape[0]
        return n, time.perf_counter() - t0

    try:
        return _cpu_leg_run(run, steps, warmup, cores, budget_s, H, W)
    finally:
        pipe.close()


def _cpu_leg_run(run, steps, warmup, cores, budget_s, H, W):
    if warmup > 0:
        run(max(2, min(warmup, 2 * cores)), seed=1)        # page-cache / import warm-up, untimed
    if budget_s is not None:                                # bounded sample: grow until ~budget seconds of CPU work
        n_batches = 2 * cores
        while True:
            n, dt = run(n_batches, seed=2)
            if dt >= budget_s / 2 or n_batches >= 4096:
                break
            n_batches = int(min(4096, max(n_batches * 2, n_batches * budget_s / max(dt, 1e-3))))
    else:
        n_batches = steps
        n, dt = run(n_batches, seed=2)
    return {"value": n / dt, "unit": "patches/s", "cores": cores, "kind": "port",
            "sample": f"{n_batches} batches x {BATCH} patches of the same workload ({H}x{W} slide in host RAM, {N_POLY} polygons), "
                      f"{cores} worker processes x 2 batches per job like the reference's spawn ProcessPoolExecutor (pool start-up excluded); {dt:.2f} s",
            "seconds": dt, "steps": n_batches}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_leg_annotated(args.steps, args.warmup, None, None)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "patches/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds"] / max(r["steps"], 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": CONFIG,
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "oracle port of AnnoRegionRndSampler.torch_generator (region_samplers.py:685-738): the reference cannot run here "
                "(psimage and shapely are not installable); CPU tensors are left on the host as the reference yields them",
    }
    print(json.dumps(line), flush=True)


METRIC = "patches/sec sampled+normalised"
CONFIG = {
    "workload": "examples.sample_annotated_rnd --torch (BASELINE configs[1]): 224x224 random patches inside 50 synthetic polygons, "
                "32768x32768 uint8 RGB slide, batch 256, patches_from_one_region 4, region_intersection 0.75, fp32 NHWC /255",
    "slide": list(SLIDE_HW), "patch": PS, "batch": BATCH, "polygons": N_POLY,
    "l2_policy": "inputs larger than L2: random patches of a 3.2 GB slide; each step writes a 193 MB batch (> 126 MB L2)",
}


# --------------------------------------------------------------------------------------------------------------------
def ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from deephisto_b200 import _lib, ops
    from deephisto_b200.patch_samplers.region_samplers import AnnoRegionRndSampler, build_tables
    from deephisto_b200.slide import SyntheticSlide
    from deephisto_b200.synthetic import synth_polygons

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.require_device()
    K, Wm = args.steps, args.warmup
    H, W = SLIDE_HW
    sampler_src = SyntheticSlide(H, W, seed=0)
    slide = sampler_src.device_slide(dev)
    polys = synth_polygons(N_POLY, H, W, seed=0)
    tables, _, classes = build_tables([((H, W), polys)], layer=1, area_influence=0.5, classes=None, one_image_for_batch=True, device=dev)
    thr = PS * PS * RI
    stream = torch.cuda.current_stream().cuda_stream

    # ---- device-resident loop through the C-ABI, outputs preallocated ---------------------------------------------------
    # Coordinates are counter-based (Philox keyed by the global slot index): one dh_region_sample launch draws the next
    # AHEAD batches (identical results to per-batch launches), then each step gathers one batch of 256 patches.
    AHEAD = 16
    coords = torch.empty((2, AHEAD * BATCH, 2), dtype=torch.int32, device=dev)
    labels = torch.empty((2, AHEAD * BATCH), dtype=torch.int64, device=dev)
    images = torch.empty((2, AHEAD * BATCH), dtype=torch.int32, device=dev)
    status = torch.zeros((2, AHEAD * BATCH), dtype=torch.uint8, device=dev)
    nbuf = 3
    feats = [torch.empty((BATCH, PS, PS, 3), dtype=torch.float32, device=dev) for _ in range(nbuf)]
    tstruct = C.byref(tables.struct)
    sp, gp = lib.dh_region_sample, lib.dh_gather_normalize
    fps = [f.data_ptr() for f in feats]
    cptr = [coords[0].data_ptr(), coords[1].data_ptr()]
    sl_ptr, pitch = slide.storage.data_ptr(), slide.pitch
    fail = torch.zeros(1, dtype=torch.uint8, device=dev)
    launches = [0]

    def step(i, ev=None):
        # rank r draws from its own Philox slot range (rank << 40): disjoint streams, no data-path collective
        buf = (i // AHEAD) & 1
        rc = 0
        if i % AHEAD == 0:
            off = (rank << 40) + i * BATCH
            rc = sp(tstruct, AHEAD * BATCH, K_PER_REGION, PS, thr, 500, 64, -1, 2 * BATCH, 0, off, coords[buf].data_ptr(),
                    labels[buf].data_ptr(), images[buf].data_ptr(), status[buf].data_ptr(), stream)
            launches[0] += 1
        if ev is not None:
            ev[0].record()
        rc |= gp(sl_ptr, H, W, pitch, cptr[buf] + (i % AHEAD) * BATCH * 8, None, BATCH, PS, fps[i % nbuf], 0, 0, 1, None, None, None, stream)
        launches[0] += 1
        if ev is not None:
            ev[1].record()
        if rc:
            raise RuntimeError(_lib.last_error())

    sampler_thread = ClockSampler(local)
    sampler_thread.start()
    Wm = (Wm + AHEAD - 1) // AHEAD * AHEAD          # keep the sampling launches aligned with the timed region
    for i in range(Wm):
        step(i)
    torch.maximum(fail, status.max().reshape(1), out=fail)
    torch.cuda.synchronize()
    launches[0] = 0
    if world > 1:
        dist.barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t_start.record()
    for i in range(K):
        step(Wm + i, evs[i])
    t_end.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms_total = t_start.elapsed_time(t_end)
    gather_ms = sorted(a.elapsed_time(b) for a, b in evs)
    gather_avg_ms = sum(gather_ms) / len(gather_ms)
    if int(fail.item()) != 0 or int(status.max().item()) != 0:
        raise RuntimeError("region sampling reported failed slots")
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * K * BATCH / (ms_total / 1e3)

    # ---- end to end through the public API ---------------------------------------------------------------------------
    api = AnnoRegionRndSampler([(sampler_src, polys)], layer=1, patch_size=PS, patches_from_one_region=K_PER_REGION,
                               one_image_for_batch=True, seed=1 + rank, device=dev, verbose=False)
    h_labels = torch.empty(BATCH, dtype=torch.int64).pin_memory()
    h_coords = torch.empty((BATCH, 2), dtype=torch.float32).pin_memory()
    d2h = h_labels.numel() * 8 + h_coords.numel() * 4

    def run_api(n_batches, to_host=False, h_feats=None):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for f, l, c in api.torch_generator(batch_size=BATCH, n_batches=n_batches, batches_per_worker=2):
            h_labels.copy_(l, non_blocking=True)
            h_coords.copy_(c, non_blocking=True)
            if to_host:
                h_feats.copy_(f, non_blocking=True)
            torch.cuda.current_stream().synchronize()      # the consumer reads this step's result
        return time.perf_counter() - t0

    run_api(max(Wm, 4))
    if world > 1:
        dist.barrier()
    e2e_s = run_api(K)
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * K * BATCH / float(t.item())
    h_feats = torch.empty((BATCH, PS, PS, 3), dtype=torch.float32).pin_memory()
    kh = max(4, min(K, 32))
    run_api(2, True, h_feats)
    e2e_host_s = run_api(kh, True, h_feats)
    sampler_thread.stop_flag = True
    sampler_thread.join(timeout=1)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peaks()
    alg_bytes = BATCH * (PATCH_IN + PATCH_IN * 4)
    achieved = alg_bytes / (gather_avg_ms / 1e3) / 1e9
    traffic = None
    tf = ROOT / "profiles" / "gather_traffic.json"
    if tf.exists():
        traffic = json.loads(tf.read_text()).get("gather_nhwc_f32_bytes_per_launch")
    line = {
        "metric": METRIC, "value": value, "unit": "patches/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": dict(CONFIG, parallelism=f"replicated slide, batches sharded by rank (x{world})"),
        "gigapixels_per_s": value * PS * PS / 1e9,
        "roofline": {"bound": "hbm", "kernel": "gather_nhwc_vec<float> (dh_gather_normalize)", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms_avg": gather_avg_ms, "kernel_ms_median": gather_ms[len(gather_ms) // 2],
                     "frac_of_nominal_8TBs": achieved / 8000.0},
        "e2e": {"value": e2e_value, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": d2h,
                "api": "AnnoRegionRndSampler.torch_generator(batch_size=256) -> CUDA features; labels+coords read back per step",
                "features_to_host": {"value": kh * BATCH / e2e_host_s, "unit": "patches/s", "d2h_bytes_per_step": d2h + h_feats.numel() * 4}},
        "gpu_launches": launches[0],
        "clocks": sampler_thread.summary(),
    }
    if world == 1 and not args.no_cpu_baseline:
        out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--cpu-leg", "--cpu-budget", str(args.cpu_budget)],
                             capture_output=True, text=True)
        try:
            line["cpu_baseline"] = json.loads(out.stdout.strip().splitlines()[-1])
        except Exception:
            line["cpu_baseline"] = {"error": (out.stderr or out.stdout)[-400:]}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work for the bounded cpu_baseline sample")
    ap.add_argument("--cpu-leg", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.cpu_leg:
        r = cpu_leg_annotated(0, 1, None, args.cpu_budget)
        print(json.dumps({k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}), flush=True)
        return
    if args.impl == "reference":
        reference_arm(args)
        return
    ours(args)


if __name__ == "__main__":
    main()
