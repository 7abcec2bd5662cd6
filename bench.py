#!/usr/bin/env python
"""Benchmark of the deephisto_b200 hot path (contract: one JSON line on stdout from rank 0).

Default workload = BASELINE.json configs[1]: `examples.sample_annotated_rnd --torch` -- random 224x224 patches inside
50 synthetic annotation polygons on a 32768 x 32768 synthetic slide, batch 256, fp32 NHWC features in [0,1] + int64
labels + (y,x) coords, exactly what AnnoRegionRndSampler.torch_generator yields. One step = one batch of 256 patches.
The device pipeline prefetches CHUNK (32) batches per pair of launches, like AnnoRegionRndSampler.torch_generator:
    dh_region_sample (Philox draws + exact clip-area acceptance, 32 x 256 slots)  ->  dh_gather_normalize (32 x 256 patches)
so a K-step run is ceil(K/32) chunk pairs (the last one partial) and every batch is a contiguous slice of the chunk buffer.

  value     patches/s over all ranks, inputs (slide, polygon tables) resident in HBM, CUDA-event timed, max over ranks
  e2e       the same metric through the public Python API (AnnoRegionRndSampler.torch_generator) starting from a slide in
            PINNED HOST memory: its upload to HBM is inside the timed region (once per run -- the slide then stays resident,
            which is the product's design), and every step ends with the device->host read of the step's labels and
            coordinates; features stay in HBM for the consumer CNN (`features_to_host` also reports the PCIe-bound variant)
  roofline  dominant kernel (gather+normalise): algorithmic bytes / CUDA-event time of that kernel vs measured HBM peak
  cpu_baseline / --impl reference: the reference's CPU path (oracle/cpu_pipeline.py restates
            AnnoRegionRndSampler.torch_generator; the reference itself needs psimage + shapely, which do not exist)

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload annotated_rnd|predict]

`--workload predict` (BASELINE configs[2]/[3]): whole-slide patched prediction, ResNet18 (torch/cuDNN) on a synthetic
40k x 40k slide (1 GPU) or 100k x 100k slide row-band sharded over N GPUs; metric = gigapixels/s; one step = one slide.
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
CHUNK = 32                             # batches prefetched per (sample, gather) launch pair
K_PER_REGION = 4
RI = 0.75


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
            time.sleep(0.01)

    def finish(self):
        self.stop_flag = True
        self.join(timeout=1)
        return self.summary()

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


# stdout carries exactly ONE JSON line: everything else that writes to fd 1 (NCCL's version banner, the samplers' "Image H x W"
# prints, worker processes) is pointed at stderr for the life of the process.
sys.stdout.flush()
_JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line: dict):
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


def dist_env():
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))


def init_nccl(dev):
    """init_process_group + first collective with fd 1 pointed at stderr: NCCL prints its version banner on stdout, which
    must carry exactly one JSON line."""
    import torch
    import torch.distributed as dist

    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        dist.init_process_group("nccl", device_id=dev)
        dist.all_reduce(torch.zeros(1, device=dev))
        torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)


# --------------------------------------------------------------------------------------------------------------------
# CPU legs (run in a process that never touched CUDA: they start worker pools)
METRIC = "patches/sec sampled+normalised"
CONFIG = {
    "workload": "examples.sample_annotated_rnd --torch (BASELINE configs[1]): 224x224 random patches inside 50 synthetic polygons, "
                "32768x32768 uint8 RGB slide, batch 256, patches_from_one_region 4, region_intersection 0.75, fp32 NHWC /255",
    "slide": list(SLIDE_HW), "patch": PS, "batch": BATCH, "polygons": N_POLY, "chunk_batches": CHUNK,
    "l2_policy": "inputs larger than L2: random patches of a 3.2 GB slide; every gather launch writes a 4.9 GB chunk (32 batches of "
                 "154 MB) into one of two alternating buffers (126 MB L2)",
}


def cpu_pipe(cores):
    from oracle import cpu_pipeline, synth
    from deephisto_b200.synthetic import synth_polygons

    H, W = SLIDE_HW
    path = synth.synth_slide_shared(H, W, 0, workers=cores, return_path=True)
    images = [((H, W), synth_polygons(N_POLY, H, W, seed=0))]
    return cpu_pipeline.AnnotatedRndCPU(path, (H, W, 3), images, layer=1, one_image_for_batch=True, max_workers=cores)


def cpu_run(pipe, batch, n_batches, seed, batches_per_worker=2):
    t0 = time.perf_counter()
    n = 0
    for f, l, c in pipe.batches(PS, batch, n_batches, batches_per_worker=batches_per_worker, k=K_PER_REGION, ri=RI, seed=seed):
        n += f.shape[0]
    return n, time.perf_counter() - t0


def cpu_leg_bounded(budget_s: float):
    """cpu_baseline of our arm: ~budget_s seconds of the reference's CPU path on the same workload (batch 256)."""
    cores = os.cpu_count()
    pipe = cpu_pipe(cores)
    try:
        cpu_run(pipe, BATCH, 2 * cores, seed=1)                     # page-cache / import warm-up, untimed
        n_batches = 2 * cores
        while True:
            n, dt = cpu_run(pipe, BATCH, n_batches, seed=2)
            if dt >= budget_s / 2 or n_batches >= 4096:
                break
            n_batches = int(min(4096, max(n_batches * 2, n_batches * budget_s / max(dt, 1e-3))))
    finally:
        pipe.close()
    H, W = SLIDE_HW
    return {"value": n / dt, "unit": "patches/s", "cores": cores, "kind": "port",
            "sample": f"{n_batches} batches x {BATCH} patches of the same workload ({H}x{W} slide in host RAM, {N_POLY} polygons), {cores} "
                      f"worker processes x 2 batches per job like the reference's spawn ProcessPoolExecutor (pool start-up excluded); {dt:.2f} s"}


def reference_arm(args):
    """The reference's CPU implementation of the path on the host cores, K steps; a step is a bounded sample (a batch of
    `sample` <= 256 patches) so that the whole run ends within a few minutes."""
    world, rank, _ = dist_env()
    if rank != 0:
        return
    if args.workload == "predict":
        return reference_arm_predict(args)
    cores = os.cpu_count()
    pipe = cpu_pipe(cores)
    try:
        n, dt = cpu_run(pipe, BATCH, max(2 * cores, min(args.warmup, 4 * cores)), seed=1)      # warm-up, also calibrates the sample
        rate = n / dt
        sample = int(max(K_PER_REGION, min(BATCH, rate * args.ref_budget / max(args.steps, 1)) // K_PER_REGION * K_PER_REGION))
        # keep the reference's worker-job size (2 batches of 256 = 512 patches per job, region_samplers.py:722-728): the result pipe of
        # the process pool is part of the path, and smaller jobs would make it look ~2x faster than it is at batch 256
        bpw = max(2, (2 * BATCH) // sample)
        n, dt = cpu_run(pipe, sample, args.steps, seed=2, batches_per_worker=bpw)
    finally:
        pipe.close()
    H, W = SLIDE_HW
    value = n / dt
    desc = (f"{args.steps} steps x {sample} patches (bounded sample of the 256-patch batch) of the same workload ({H}x{W} slide in host RAM, "
            f"{N_POLY} polygons), {cores} worker processes, {bpw} sampled batches (= {bpw * sample} patches, the reference's 2 x 256) per job like the "
            f"reference's spawn ProcessPoolExecutor (pool start-up excluded); {dt:.2f} s")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "patches/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": CONFIG,
        "cpu_baseline": {"value": value, "unit": "patches/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "oracle port of AnnoRegionRndSampler.torch_generator (region_samplers.py:685-738): the reference itself cannot run "
                "(psimage and shapely are neither vendored nor installable); CPU tensors are left on the host as the reference yields them",
    }
    emit(line)


# --------------------------------------------------------------------------------------------------------------------
MODES = {
    # BASELINE configs[1]
    "annotated_rnd": dict(batch=BATCH, chunk=CHUNK, dtype="f32", torch_dtype="float32", layout="NHWC", dcode=0, lcode=0, esize=4, flips=False,
                          config=CONFIG, kernel="gather_tma_kernel<float, NHWC, /255> (dh_gather_normalize), one launch per 32-batch chunk"),
    # BASELINE configs[4]: the input pipeline of models.patch_cls_simple.train at 8k patches/step, bf16 NCHW with the batch-level
    # random H/V flips of train.py:71-81 fused into the gather
    "train_input": dict(batch=8192, chunk=1, dtype="bf16", torch_dtype="bfloat16", layout="NCHW", dcode=1, lcode=1, esize=2, flips=True,
                        config=dict(CONFIG, workload="models.patch_cls_simple.train input pipeline (BASELINE configs[4]): on-device annotated random sampling, "
                                    "8192 patches per step, bf16 NCHW /255 with batch-level random H/V flips (train.py:71-81), same slide and polygons "
                                    "as configs[1]", batch=8192, chunk_batches=1,
                                    l2_policy="inputs larger than L2: random patches of a 3.2 GB slide; every step writes a 2.5 GB batch"),
                        kernel="gather_tma_kernel<bf16, NCHW, /255> with flips (dh_gather_normalize), one launch per 8192-patch step"),
}


def ours(args):
    import torch
    import torch.distributed as dist

    from deephisto_b200 import _lib
    from deephisto_b200.patch_samplers.region_samplers import AnnoRegionRndSampler, build_tables
    from deephisto_b200.slide import PinnedSlide, SyntheticSlide
    from deephisto_b200.synthetic import synth_polygons

    world, rank, local = dist_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        init_nccl(dev)
    lib = _lib.require_device()
    K, Wm = args.steps, args.warmup
    mode = MODES[args.workload]
    BATCH, CHUNK = mode["batch"], mode["chunk"]        # shadow the module-level defaults
    out_dtype = getattr(torch, mode["torch_dtype"])
    H, W = SLIDE_HW
    source = SyntheticSlide(H, W, seed=0)
    slide = source.device_slide(dev)
    polys = synth_polygons(N_POLY, H, W, seed=0)
    tables, _, classes = build_tables([((H, W), polys)], layer=1, area_influence=0.5, classes=None, one_image_for_batch=True, device=dev)
    thr = PS * PS * RI
    stream = torch.cuda.current_stream().cuda_stream

    # ---- device-resident loop through the C-ABI, outputs preallocated ---------------------------------------------------
    # Coordinates are counter-based (Philox keyed by the global slot index): one dh_region_sample launch draws CHUNK batches
    # (identical results to per-batch launches), one dh_gather_normalize launch writes their features.
    n_slots = CHUNK * BATCH
    coords = torch.empty((2, n_slots, 2), dtype=torch.int32, device=dev)
    labels = torch.empty((2, n_slots), dtype=torch.int64, device=dev)
    images = torch.empty((2, n_slots), dtype=torch.int32, device=dev)
    status = torch.zeros((2, n_slots), dtype=torch.uint8, device=dev)
    feats = torch.empty((2, n_slots, PS, PS, 3) if mode["layout"] == "NHWC" else (2, n_slots, 3, PS, PS), dtype=out_dtype, device=dev)
    # one H and one V coin per batch, like torchvision's flips of the whole [B,3,H,W] tensor (train.py:71-81)
    flip_bits = None
    if mode["flips"]:
        n_coins = (max(Wm, 3) + CHUNK + K) * 2 + 4 * CHUNK
        coins = torch.randint(0, 4, (n_coins,), generator=torch.Generator().manual_seed(7 + rank), dtype=torch.uint8)
        flip_bits = coins.repeat_interleave(BATCH).to(dev)
    tstruct = C.byref(tables.struct)
    sp, gp = lib.dh_region_sample, lib.dh_gather_normalize
    sl_ptr, pitch = slide.storage.data_ptr(), slide.pitch
    fail = torch.zeros(1, dtype=torch.uint8, device=dev)
    launches = [0]

    # the coordinate launch of chunk i+1 runs on a second stream next to the gather of chunk i, as in AnnoRegionRndSampler.torch_generator
    # (DH_BENCH_OVERLAP=0 puts both on one stream: A/B in profiles/r01_gather.md)
    overlap = os.environ.get("DH_BENCH_OVERLAP", "1") == "1"
    side = torch.cuda.Stream(dev) if overlap else None
    drawn = [torch.cuda.Event(), torch.cuda.Event()]
    gathered = [torch.cuda.Event(), torch.cuda.Event()]

    def chunk(first_step, n_batches, ev=None):
        """Steps [first_step, first_step + n_batches): rank r draws from its own Philox slot range (rank << 40) -- disjoint
        streams, no data-path collective."""
        buf = (first_step // CHUNK) & 1
        n = n_batches * BATCH
        off = (rank << 40) + first_step * BATCH
        if overlap:
            side.wait_event(gathered[buf])                            # the coords buffer is free once its previous gather is done
        rc = sp(tstruct, n, K_PER_REGION, PS, thr, 500, 64, -1, 2 * BATCH, 0, off, coords[buf].data_ptr(), labels[buf].data_ptr(),
                images[buf].data_ptr(), status[buf].data_ptr(), side.cuda_stream if overlap else stream)
        if overlap:
            drawn[buf].record(side)
            torch.cuda.current_stream().wait_event(drawn[buf])
        if ev is not None:
            ev[0].record()
        fl = None if flip_bits is None else flip_bits.data_ptr() + first_step * BATCH
        rc |= gp(sl_ptr, H, W, pitch, coords[buf].data_ptr(), None, n, PS, feats[buf].data_ptr(), mode["dcode"], mode["lcode"], 1, None, None, fl,
                 stream)
        if ev is not None:
            ev[1].record()
        if overlap:
            gathered[buf].record()
        launches[0] += 2
        if rc:
            raise RuntimeError(_lib.last_error())

    def run_steps(first, n_steps, evs=None):
        done = 0
        while done < n_steps:
            nb = min(CHUNK, n_steps - done)
            e = None
            if evs is not None:
                e = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), nb)
                evs.append(e)
            chunk(first + done, nb, e)
            done += nb
        return done

    clocks = ClockSampler(local)
    clocks.start()
    Wm = max(Wm, 3)
    run_steps(0, Wm)
    torch.maximum(fail, status.max().reshape(1), out=fail)
    torch.cuda.synchronize()
    launches[0] = 0
    first_timed = (Wm + CHUNK - 1) // CHUNK * CHUNK                 # keep chunk boundaries aligned with the buffers
    if world > 1:
        dist.barrier()
    evs = []
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t_start.record()
    run_steps(first_timed, K, evs)
    t_end.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms_total = t_start.elapsed_time(t_end)
    gather_ms = [(a.elapsed_time(b), nb) for a, b, nb in evs]
    gather_total_ms = sum(m for m, _ in gather_ms)
    torch.maximum(fail, status.max().reshape(1), out=fail)
    if int(fail.item()) != 0:
        raise RuntimeError("region sampling reported failed slots")
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * K * BATCH / (ms_total / 1e3)
    del feats, coords, labels, images, status
    torch.cuda.empty_cache()

    # ---- end to end through the public API, slide starting in pinned host memory -----------------------------------------
    host_slide = PinnedSlide.from_device(slide)                       # setup (not timed): the slide as it sits in host RAM
    h_labels = torch.empty(BATCH, dtype=torch.int64).pin_memory()
    h_coords = torch.empty((BATCH, 2), dtype=torch.float32).pin_memory()
    d2h = h_labels.numel() * 8 + h_coords.numel() * 4

    def run_api(api, n_batches, to_host=False, h_feats=None):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for f, l, c in api.torch_generator(batch_size=BATCH, n_batches=n_batches, batches_per_worker=2):
            h_labels.copy_(l, non_blocking=True)
            h_coords.copy_(c, non_blocking=True)
            if to_host:
                h_feats.copy_(f, non_blocking=True)
            torch.cuda.current_stream().synchronize()                 # the consumer reads this step's result
        return time.perf_counter() - t0

    def new_api(src, seed):
        return AnnoRegionRndSampler([(src, polys)], layer=1, patch_size=PS, patches_from_one_region=K_PER_REGION, one_image_for_batch=True,
                                    seed=seed, device=dev, verbose=False, out_dtype=out_dtype, out_layout=mode["layout"], flips=mode["flips"],
                                    shard_upload=True if world > 1 and isinstance(src, PinnedSlide) else None)

    warm = new_api(source, 1 + rank)
    run_api(warm, max(Wm, 3 * CHUNK))                                # warm-up of the API path (kernels, and the caching allocator's pool: a
    del warm, slide                                                  # long-lived process re-uses its slide and feature buffers instead of paying
    source._dev.clear()                                              # cudaMalloc -- measured 10-40 ms of jitter on the first call otherwise)
    api = new_api(host_slide, 101 + rank)                              # polygon parsing / table build: constructor, not timed (as in the reference)
    if world > 1:
        dist.barrier()
    e2e_s = run_api(api, K)                                          # includes the one-time H2D upload of the 3.2 GB slide
    steady_s = run_api(api, K)                                       # same call again: slide already resident
    t = torch.tensor([e2e_s, steady_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value, steady_value = (world * K * BATCH / float(x) for x in t.tolist())
    h_feats = torch.empty((BATCH, PS, PS, 3) if mode["layout"] == "NHWC" else (BATCH, 3, PS, PS), dtype=out_dtype).pin_memory()
    kh = max(2, min(K, 32 if BATCH <= 256 else 4))
    run_api(api, 2, True, h_feats)
    e2e_host_s = run_api(api, kh, True, h_feats)
    clk = clocks.finish()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peaks()
    alg_bytes_batch = BATCH * (PATCH_IN + PATCH_IN * mode["esize"])
    alg_total = alg_bytes_batch * sum(nb for _, nb in gather_ms)
    achieved = alg_total / (gather_total_ms / 1e3) / 1e9
    full = [m for m, nb in gather_ms if nb == CHUNK]
    traffic = None
    tf = ROOT / "profiles" / "gather_traffic.json"
    if tf.exists() and args.workload == "annotated_rnd":               # ncu capture of exactly this kernel / launch shape
        cap = json.loads(tf.read_text())
        if int(cap.get("algorithmic_bytes_per_launch", 0)) == alg_bytes_batch * CHUNK:      # same launch shape as the timed kernel
            traffic = cap.get("dram_bytes_per_launch")
    slide_bytes = host_slide.nbytes
    uploaded = int(api.uploaded_bytes)                                # what actually travelled (the whole layer unless < 1/5 of it is annotated)
    line = {
        "metric": METRIC, "value": value, "unit": "patches/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": mode["dtype"],
        "data": "synthetic", "config": dict(mode["config"], parallelism=f"replicated slide, batches sharded by rank (x{world}), no data-path collective"),
        "gigapixels_per_s": value * PS * PS / 1e9,
        "roofline": {"bound": "hbm", "kernel": mode["kernel"],
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes_batch * CHUNK, "algorithmic_bytes_per_patch": PATCH_IN * (1 + mode["esize"]),
                     "kernel_ms_avg_full_chunk": (sum(full) / len(full)) if full else None, "launches_timed": len(gather_ms),
                     "kernel_ms_per_batch": gather_total_ms / max(sum(nb for _, nb in gather_ms), 1), "frac_of_nominal_8TBs": achieved / 8000.0},
        "e2e": {"value": e2e_value, "unit": "patches/s", "h2d_bytes_per_step": uploaded / K, "d2h_bytes_per_step": d2h,
                "api": f"AnnoRegionRndSampler.torch_generator(batch_size={BATCH}, n_batches=K) over a slide in pinned host memory: the timed region "
                       "contains the one-time H2D upload of the slide (h2d_bytes_total per rank), K batches, and per step the D2H read of labels+coords"
                       + ("; the ranks hold the same slide, so each uploads 1/world of its rows over its own PCIe link and one NCCL all-gather over "
                          "NVLink replicates them (slide.sharded_upload)" if world > 1 else ""),
                "h2d_bytes_total": uploaded, "slide_bytes": slide_bytes, "host_memory_pinned": bool(host_slide.pinned), "seconds": e2e_s,
                "steady_state": {"value": steady_value, "unit": "patches/s", "note": "the same call repeated with the slide already resident"},
                "features_to_host": {"value": kh * BATCH / e2e_host_s, "unit": "patches/s", "d2h_bytes_per_step": d2h + h_feats.numel() * mode["esize"]}},
        "gpu_launches": launches[0],
        "clocks": clk,
    }
    if args.with_training and args.workload == "train_input":
        line["training_consumer"] = training_consumer(api, dev, BATCH, torch)
    if world == 1 and not args.no_cpu_baseline and args.workload == "annotated_rnd":
        out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--cpu-leg", "--cpu-budget", str(args.cpu_budget)],
                             capture_output=True, text=True)
        try:
            line["cpu_baseline"] = json.loads(out.stdout.strip().splitlines()[-1])
        except Exception:
            line["cpu_baseline"] = {"error": (out.stderr or out.stdout)[-400:]}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def training_consumer(api, dev, batch, torch, steps=3, micro=1024):
    """BASELINE configs[4] in full: the sampler feeding a bf16 ResNet18 training step (models/patch_cls_simple/train.py:153-172:
    forward, cross-entropy, backward, Adam 1e-4) at `batch` patches per step, micro-batched. Two measurements, same model:
    (a) every step reuses one resident batch (no input pipeline at all), (b) every step consumes a fresh batch from
    AnnoRegionRndSampler.torch_generator (sampling + gather of the next batch run on the producer stream). (b) / (a) is the
    cost of the input pipeline as the training loop sees it."""
    from deephisto_b200.examples.predict_full_patched import get_model

    torch.manual_seed(0)
    model = get_model(5).to(dev).to(memory_format=torch.channels_last).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    loss_fn = torch.nn.CrossEntropyLoss()

    def train_step(f, l):
        opt.zero_grad(set_to_none=True)
        for a in range(0, f.shape[0], micro):
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out = model(f[a : a + micro].contiguous(memory_format=torch.channels_last))
                loss = loss_fn(out.float(), l[a : a + micro]) * (min(micro, f.shape[0] - a) / f.shape[0])
            loss.backward()
        opt.step()
        return loss

    gen = api.torch_generator(batch_size=batch, n_batches=2 * steps + 2, batches_per_worker=2)
    f0, l0, _ = next(gen)
    train_step(f0, l0)                                               # warm-up (cuDNN algorithm selection, allocator)
    torch.cuda.synchronize()
    res = {}
    for name in ("resident_batch", "fresh_batch_per_step"):
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(steps):
            if name == "fresh_batch_per_step":
                f0, l0, _ = next(gen)
            loss = train_step(f0, l0)
        t1.record()
        torch.cuda.synchronize()
        res[name] = {"patches_per_s": steps * batch / (t0.elapsed_time(t1) / 1e3), "ms_per_step": t0.elapsed_time(t1) / steps}
    res["input_pipeline_overhead"] = res["resident_batch"]["patches_per_s"] / res["fresh_batch_per_step"]["patches_per_s"] - 1.0
    res["final_loss"] = float(loss.item()) * 1.0
    res["note"] = f"ResNet18 bf16 autocast channels_last, Adam, {batch} patches per step in micro-batches of {micro}, {steps} steps each"
    return res


# --------------------------------------------------------------------------------------------------------------------
# whole-slide patched prediction (BASELINE configs[2] / [3])
def predict_config(world, args):
    hw = tuple(args.slide) if args.slide else ((40000, 40000) if world == 1 else (100000, 100000))
    return hw, {
        "workload": f"examples.predict_full_patched (BASELINE configs[{2 if world == 1 else 3}]): patch_cls_simple ResNet18 (random init, seed 0, "
                    f"{'bf16 channels_last' if args.bf16 else 'fp32, torch defaults (TF32 convolutions)'}{', BatchNorm folded' if args.fold_bn else ''}"
                    f"{', cudnn.benchmark' if args.cudnn_benchmark else ''}) on a synthetic {hw[0]}x{hw[1]} slide, "
                    f"224x224 patches at stride 112, dense sampler batch 64 (CNN batch {args.cnn_batch}), stitch downscale 16, argmax map",
        "slide": list(hw), "patch": PS, "stride": 112, "downscale": 16, "cnn_batch": args.cnn_batch,
        "l2_policy": "inputs larger than L2: each step re-reads the whole slide band (GBs) from HBM",
    }


def ours_predict(args):
    import torch
    import torch.distributed as dist

    from deephisto_b200 import _lib, bands
    from deephisto_b200.anno.utils import AnnoDescription
    from deephisto_b200.examples import predict_full_patched as pfp
    from deephisto_b200.patch_samplers import full_samplers as fs
    from deephisto_b200.slide import SyntheticSlide

    world, rank, local = dist_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        init_nccl(dev)
    _lib.require_device()
    (H, W), cfg = predict_config(world, args)
    torch.manual_seed(0)
    model = pfp.get_model(5)
    torch.backends.cudnn.benchmark = bool(args.cudnn_benchmark)
    pred = pfp.DeviceBatchPredictor(model, dev, torch.bfloat16 if args.bf16 else torch.float32, fold_bn=bool(args.fold_bn))
    anno = AnnoDescription.with_auto_colors(["AT", "BG", "LP", "MM", "TUM"])
    src = SyntheticSlide(H, W, seed=0)
    mode = fs.SamplerExecutionMode.INMEMORY_SINGLEPROC
    sampler = fs.FullImageDenseSampler(src, 1, PS, 64, mode, stride=112, device=dev, lazy_slide=True)
    plan = bands.plan_band(H, W, PS, 112, 16, 64, rank, world)
    # the band of the slide (with its patch-size halo) is made resident once, like the reference loads the layer in its constructor
    band, y_off = sampler.band_slide(plan.slide_y0, plan.slide_y1)
    sampler.band_slide = lambda y0, y1: (band, y_off)
    ipp = pfp.ImagePredictorPatched(src, sampler, pred, anno, layer=1, downscale=16, device=dev, cnn_batch=args.cnn_batch,
                                    stream_bands=False)           # `value`: the band stays resident in HBM across steps
    K, Wm = args.steps, max(1, args.warmup)

    def step():
        return ipp.process_device(rank=rank, world=world)["argmax"] if world > 1 else ipp.dense_band_local(0, 1)["argmax_band"]

    clocks = ClockSampler(local)
    clocks.start()
    for _ in range(Wm):
        out = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ipp.stage_events = {}                                            # CUDA events per stage: our kernels vs torch's CNN
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(K):
        out = step()
    t1.record()
    torch.cuda.synchronize()
    stage_ms = {k: v / K for k, v in ipp.stage_ms().items()}
    ipp.stage_events = None
    if world > 1:
        dist.barrier()
    t = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    # e2e: every step starts from this rank's slide band in PINNED HOST memory: the public call streams it through HBM in row chunks
    # on a copy stream, one chunk ahead of the CNN (ImagePredictorPatched._logits_streamed), and the class map is read back (what
    # process() returns) -- the reference likewise reads the layer from storage in its constructor (full_samplers.py:53-55)
    from deephisto_b200.slide import PinnedSlide

    host_band = PinnedSlide.from_device(band, y_origin=y_off, full_height=H)     # setup, not timed
    h_map = torch.empty(out.shape, dtype=torch.uint8).pin_memory()
    del band, sampler, ipp
    torch.cuda.empty_cache()
    sampler2 = fs.FullImageDenseSampler(host_band, 1, PS, 64, mode, stride=112, device=dev, lazy_slide=True)
    ipp2 = pfp.ImagePredictorPatched(host_band, sampler2, pred, anno, layer=1, downscale=16, device=dev, cnn_batch=args.cnn_batch)

    def step2():
        return ipp2.process_device(rank=rank, world=world)["argmax"] if world > 1 else ipp2.dense_band_local(0, 1)["argmax_band"]

    h_map.copy_(step2())                                               # warm-up of the streamed path (allocator)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    w0 = time.perf_counter()
    for _ in range(K):
        h_map.copy_(step2(), non_blocking=True)
        torch.cuda.synchronize()
    e2e_s = time.perf_counter() - w0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    clk = clocks.finish()
    if rank == 0:
        gpx = H * W / 1e9
        g = bands.dense_grid(H, W, PS, 112, 64)
        line = {
            "metric": "WSI gigapixels/sec patched predict", "value": K * gpx / (ms_total / 1e3), "unit": "Gpx/s", "n_gpus": world, "steps": K,
            "warmup": Wm, "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16" if args.bf16 else "f32", "data": "synthetic",
            "config": dict(cfg, parallelism=f"row bands x{world} with patch-size halo (recomputed halo patch rows), NCCL all-gather of the u8 class-map bands"),
            "patches_per_s": K * g.n_padded / (ms_total / 1e3), "patches_per_slide": g.n_padded, "patches_this_rank": plan.n_patches,
            "e2e": {"value": K * gpx / e2e_s, "unit": "Gpx/s", "h2d_bytes_per_step": int(host_band.nbytes), "d2h_bytes_per_step": int(h_map.numel()),
                    "api": "per step: ImagePredictorPatched.process_device on a lazy sampler over this rank's slide band in pinned host memory (row "
                           "chunks of <= 1 GiB uploaded on a copy stream one chunk ahead of the CNN), class map copied to pinned host memory "
                           "(h2d/d2h bytes are per rank)"},
            "stage_ms_per_step_rank0": stage_ms,
            "roofline": None, "gpu_launches": None, "clocks": clk,
            "note": "CNN-bound (torch/cuDNN ResNet18, not part of the rebuilt path): stage_ms_per_step_rank0 separates this repo's kernels "
                    "(coords+gather, stitch, assemble) from the CNN",
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def reference_arm_predict(args):
    """ImagePredictorPatched.process + batch_predictor restated on the CPU (oracle), 2048 x 2048 crop, model on the host cores."""
    import numpy as np
    import torch

    from deephisto_b200.examples.predict_full_patched import get_model
    from oracle import cpu_pipeline, stitch as ostitch, synth

    H = W = 2048
    slide = synth.synth_slide(H, W, 0)
    torch.manual_seed(0)
    model = get_model(5).eval()
    t0 = time.perf_counter()
    n = 0
    for _ in range(max(1, min(args.steps, 2))):
        logits, coords = [], []
        for feats, c, _ in cpu_pipeline.dense_batches(slide, PS, 112, 64):
            with torch.no_grad():
                logits.append(model(feats.permute(0, 3, 1, 2).contiguous()).numpy())
            coords.append(c.numpy().astype(np.int64))
        ostitch.stitch(np.concatenate(logits), np.concatenate(coords), H, W, PS, 16)
        n += 1
    dt = time.perf_counter() - t0
    val = n * H * W / 1e9 / dt
    emit({"impl": "reference", "metric": "WSI gigapixels/sec patched predict", "value": val, "unit": "Gpx/s", "n_gpus": args.gpus,
                      "steps": n, "warmup": 0, "ms_per_step": 1e3 * dt / n, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                      "dtype": "f32", "data": "synthetic", "config": {"workload": "oracle port of examples.predict_full_patched on a 2048x2048 crop, CPU ResNet18"},
                      "cpu_baseline": {"value": val, "unit": "Gpx/s", "cores": torch.get_num_threads(), "kind": "port", "sample": f"{n} x 2048x2048 crop"},
                      "e2e": {"value": val, "unit": "Gpx/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="annotated_rnd", choices=["annotated_rnd", "train_input", "predict"])
    ap.add_argument("--slide", type=int, nargs=2, default=None, help="predict workload: slide H W")
    ap.add_argument("--bf16", action="store_true", help="predict workload: run the CNN in bf16 channels_last")
    ap.add_argument("--cnn-batch", type=int, default=1024)
    ap.add_argument("--fold-bn", action="store_true", help="predict workload: fold eval-mode BatchNorm into the convolutions")
    ap.add_argument("--cudnn-benchmark", action="store_true", help="predict workload: torch.backends.cudnn.benchmark = True")
    ap.add_argument("--with-training", action="store_true", help="train_input workload: also time a ResNet18 bf16 training step fed by the sampler")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work for the bounded cpu_baseline sample")
    ap.add_argument("--ref-budget", type=float, default=60.0, help="--impl reference: target seconds for the K timed steps")
    ap.add_argument("--cpu-leg", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.steps is None:
        # default job = 2500 batches of 256 = 640 000 patches, the size of the reference's own training run (config.yaml: 50 epochs,
        # batch 64; train.py:142: 200 steps per epoch): the slide is uploaded once and stays resident for the whole job
        args.steps = {"predict": 3, "train_input": 40}.get(args.workload, 2500)
    if args.warmup is None:
        args.warmup = {"predict": 3, "train_input": 4}.get(args.workload, 32)
    args.warmup = max(args.warmup, 3) if args.workload != "predict" else args.warmup
    if args.cpu_leg:
        emit(cpu_leg_bounded(args.cpu_budget))
        return
    if args.impl == "reference":
        reference_arm(args)
        return
    if args.workload == "predict":
        ours_predict(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
