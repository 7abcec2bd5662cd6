#!/usr/bin/env python
"""Benchmark of the deephisto_b200 hot path (contract: one JSON line on stdout from rank 0).

Primary workload = BASELINE.json configs[1]: `examples.sample_annotated_rnd --torch` -- random 224x224 patches inside
50 synthetic annotation polygons on a 32768 x 32768 synthetic slide, batch 256, fp32 NHWC features in [0,1] + int64
labels + (y,x) coords, exactly what AnnoRegionRndSampler.torch_generator yields. One step = one batch of 256 patches.
The device pipeline prefetches CHUNK (32) batches per pair of launches, like AnnoRegionRndSampler.torch_generator:
    dh_region_sample (Philox draws + exact clip-area acceptance, 32 x 256 slots)  ->  dh_gather_normalize (32 x 256 patches)
so a K-step region is ceil(K/32) chunk pairs (the last one partial) and every batch is a contiguous slice of the chunk buffer.
The K-step region is timed REGIONS (5) times back to back, each bracketed by barrier + synchronize; `value` is the median region.

  value     patches/s over all ranks, inputs (slide, polygon tables) resident in HBM, CUDA-event timed, max over ranks
  e2e       the same metric through the public Python API (AnnoRegionRndSampler.torch_generator) starting from a slide in
            PINNED HOST memory, every step ending with the device->host read of the step's labels and coordinates. A K-step job
            that needs fewer pixel bytes than half the slide gathers them IN PLACE from the pinned buffer (zero-copy: the gather's
            bulk row copies read host memory over PCIe; the layer becomes resident in the background, after the job's last
            gather); a longer job uploads the layer first (once) and then runs from HBM. e2e.* flat keys attribute the time.
  roofline  dominant kernel (gather+normalise): algorithmic bytes of the launches ACTUALLY TIMED / their CUDA-event time vs the
            measured HBM peak; at N = 1 also roofline.stitch_* = the stitch kernels on the 40k x 40k case (BASELINE configs[2])
  e2e.predict_*  BASELINE's second metric in the same line: whole-slide patched prediction of a 100 000 x 100 000 synthetic
            slide (configs[3]; ResNet18 bf16 channels_last via torch/cuDNN), row bands over the N ranks + NCCL all-gather,
            Gpx/s with the band resident and end to end from pinned host memory, our kernels' and the CNN's share, band parity
  cpu_baseline / --impl reference: the reference's CPU path (oracle/cpu_pipeline.py restates
            AnnoRegionRndSampler.torch_generator; the reference itself needs psimage + shapely, which do not exist)

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload annotated_rnd|train_input|predict]

`--workload predict` runs only the prediction part (any slide size / CNN options) as its own line; one step = one slide.
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

    def __init__(self, index: int, enabled: bool = True):
        """enabled=False (ranks other than 0: only rank 0 reports clocks): no NVML polling thread next to the rank's launch thread."""
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.ok = index, [], False, False
        self.max_sm = None
        if not enabled:
            return
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
REGIONS = 5                            # the K-step region is timed this many times; `value` is the median
PREDICT_HW = (100000, 100000)          # BASELINE configs[3]; the same slide for every N (strong scaling)
PREDICT_STEPS = 2                      # timed slides of the prediction part (1 warm-up slide before them)
STITCH_HW = (40000, 40000)             # BASELINE configs[2]: the stitch roofline case


def make_config(world: int) -> dict:
    """`config` of BOTH arms (identical dicts, so the driver's same_config check holds)."""
    return {
        "workload": "examples.sample_annotated_rnd --torch (BASELINE configs[1]): 224x224 random patches inside 50 synthetic polygons, "
                    "32768x32768 uint8 RGB slide, batch 256, patches_from_one_region 4, region_intersection 0.75, fp32 NHWC /255",
        "slide": list(SLIDE_HW), "patch": PS, "batch": BATCH, "polygons": N_POLY, "chunk_batches": CHUNK, "timed_regions": REGIONS,
        "l2_policy": "inputs larger than L2: random patches of a 3.2 GB slide; every gather launch writes >= 3.8 GB of features into one "
                     "of two alternating buffers (126 MB L2)",
        "parallelism": f"replicated slide, batches sharded by rank (x{world}), no data-path collective",
        "predict_workload": f"examples.predict_full_patched (BASELINE configs[3]) in the same line as e2e.predict_*: patch_cls_simple ResNet18 "
                            f"(random init, seed 0; bf16 channels_last, FusedResNetForward) on a synthetic {PREDICT_HW[0]}x{PREDICT_HW[1]} slide, 224x224 patches at stride 112, dense "
                            f"sampler batch 64, stitch downscale 16, argmax map; row bands x{world} with patch-size halo + NCCL all-gather of "
                            f"the u8 class-map bands; 1 warm-up + {PREDICT_STEPS} timed slides (strong scaling: the slide is the same for every N); "
                            f"predict_e2e_*: 1 warm-up + 1 timed slide streamed from pinned host memory",
        "stitch_workload": f"roofline.stitch_*: dh_stitch_dense / dh_stitch_binned sum maps of the {STITCH_HW[0]}x{STITCH_HW[1]} stride-112 case "
                           "(BASELINE configs[2]), n = 5 classes; binned = the coverage sampler's own coordinate list; unaligned = 39999x39999 "
                           "(map rows not 16-byte aligned); N = 1 only; CUDA-event median of 5 timed regions of 8 / 4 / 2 / 1 back-to-back calls at "
                           "d = 16 / 4 / 2 / 1 (per-call time = region / calls: the ~25 us python launch path is not timed as kernel time); d = 16 "
                           "(125 MB map ~ L2) writes a ring of 3 maps, larger maps exceed the L2 by themselves",
    }


CONFIG = make_config(1)


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
    e2e = {"value": value, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    if not args.no_predict:
        # BASELINE's second metric on the host cores: the oracle port of examples.predict_full_patched on a 2048 x 2048 crop (324 patches,
        # CPU ResNet18 fp32); seconds per 100k x 100k slide = crop time x (795 712 / 384 padded patches), linear in the patch count
        crop = cpu_predict_crop(steps=1)
        from deephisto_b200 import bands

        full = bands.dense_grid(PREDICT_HW[0], PREDICT_HW[1], PS, 112, 64).n_padded
        s_full = crop["s_per_crop"] * full / crop["patches"]
        e2e.update({"predict_gpx_per_s": PREDICT_HW[0] * PREDICT_HW[1] / 1e9 / s_full, "predict_e2e_gpx_per_s": PREDICT_HW[0] * PREDICT_HW[1] / 1e9 / s_full,
                    "predict_s_per_slide": s_full, "predict_ms_ours": None, "predict_ms_cnn": 1e3 * crop["cnn_s"] * full / crop["patches"],
                    "predict_halo_frac": 0.0, "predict_band_parity": None, "predict_patches_per_s": crop["patches"] / crop["s_per_crop"],
                    "predict_crop_gpx_per_s": crop["gpx_per_s"], "predict_extrapolated_from": "2048x2048 crop, linear in padded patch count"})
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "patches/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": make_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "patches/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": e2e,
        "gpu_launches": 0,
        "note": "oracle port of AnnoRegionRndSampler.torch_generator (region_samplers.py:685-738): the reference itself cannot run "
                "(psimage and shapely are neither vendored nor installable); CPU tensors are left on the host as the reference yields them",
    }
    emit(line)


def cpu_predict_crop(steps=1, hw=2048):
    """ImagePredictorPatched.process + batch_predictor restated on the CPU (oracle), hw x hw crop, ResNet18 on the host cores."""
    import numpy as np
    import torch

    from deephisto_b200.examples.predict_full_patched import get_model
    from oracle import cpu_pipeline, stitch as ostitch, synth

    H = W = hw
    slide = synth.synth_slide(H, W, 0)
    torch.manual_seed(0)
    model = get_model(5).eval()
    n, cnn_s, patches = 0, 0.0, 0
    t0 = time.perf_counter()
    for _ in range(max(1, steps)):
        logits, coords = [], []
        for feats, c, _ in cpu_pipeline.dense_batches(slide, PS, 112, 64):
            c0 = time.perf_counter()
            with torch.no_grad():
                logits.append(model(feats.permute(0, 3, 1, 2).contiguous()).numpy())
            cnn_s += time.perf_counter() - c0
            coords.append(c.numpy().astype(np.int64))
        ostitch.stitch(np.concatenate(logits), np.concatenate(coords), H, W, PS, 16)
        patches = sum(len(c) for c in coords)
        n += 1
    dt = time.perf_counter() - t0
    return {"s_per_crop": dt / n, "cnn_s": cnn_s / n, "patches": patches, "gpx_per_s": n * H * W / 1e9 / dt, "n": n, "threads": torch.get_num_threads()}


# --------------------------------------------------------------------------------------------------------------------
MODES = {
    # BASELINE configs[1]
    "annotated_rnd": dict(batch=BATCH, chunk=CHUNK, dtype="f32", torch_dtype="float32", layout="NHWC", dcode=0, lcode=0, esize=4, flips=False,
                          kernel="gather_tma_kernel<float, NHWC, /255> (dh_gather_normalize), one launch per chunk of <= 32 batches"),
    # BASELINE configs[4]: the input pipeline of models.patch_cls_simple.train at 8k patches/step, bf16 NCHW with the batch-level
    # random H/V flips of train.py:71-81 fused into the gather
    "train_input": dict(batch=8192, chunk=1, dtype="bf16", torch_dtype="bfloat16", layout="NCHW", dcode=1, lcode=1, esize=2, flips=True,
                        kernel="gather_tma_kernel<bf16, NCHW, /255> with flips (dh_gather_normalize), one launch per 8192-patch step"),
}


def mode_config(workload: str, world: int) -> dict:
    cfg = make_config(world)
    if workload == "train_input":
        cfg.update(workload="models.patch_cls_simple.train input pipeline (BASELINE configs[4]): on-device annotated random sampling, "
                            "8192 patches per step, bf16 NCHW /255 with batch-level random H/V flips (train.py:71-81), same slide and polygons "
                            "as configs[1]", batch=8192, chunk_batches=1,
                   l2_policy="inputs larger than L2: random patches of a 3.2 GB slide; every step writes a 2.5 GB batch")
    return cfg


def median(xs):
    xs = sorted(xs)
    return xs[len(xs) // 2]


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
    Wm = max(Wm, 3)
    chunks_per_region = (K + CHUNK - 1) // CHUNK
    first_timed = (Wm + CHUNK - 1) // CHUNK * CHUNK                 # keep chunk boundaries aligned with the buffers
    region_stride = chunks_per_region * CHUNK                       # steps between the starts of consecutive timed regions
    # one H and one V coin per batch, like torchvision's flips of the whole [B,3,H,W] tensor (train.py:71-81)
    flip_bits = None
    if mode["flips"]:
        n_coins = first_timed + REGIONS * region_stride + CHUNK
        coins = torch.randint(0, 4, (n_coins,), generator=torch.Generator().manual_seed(7 + rank), dtype=torch.uint8)
        flip_bits = coins.repeat_interleave(BATCH).to(dev)
    tstruct = C.byref(tables.struct)
    sp, gp = lib.dh_region_sample, lib.dh_gather_normalize
    sl_ptr, pitch = slide.ptr, slide.pitch
    fail = torch.zeros(1, dtype=torch.uint8, device=dev)
    launches = [0]

    # the coordinate launch of chunk i+1 runs on a second stream next to the gather of chunk i, as in AnnoRegionRndSampler.torch_generator
    # (DH_BENCH_OVERLAP=0 puts both on one stream: A/B in profiles/r01_gather.md). A region of a single chunk (the driver's --steps 20)
    # has nothing to overlap with: one stream then -- the cross-stream event hops only delay the gather (measured 7.20-7.22 M vs
    # 6.97-7.14 M patches/s on one box, profiles/r02_e2e.md)
    overlap = os.environ.get("DH_BENCH_OVERLAP", "1" if K > CHUNK else "0") == "1"
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

    clocks = ClockSampler(local, enabled=rank == 0)
    clocks.start()
    run_steps(0, Wm)
    torch.maximum(fail, status.max().reshape(1), out=fail)
    torch.cuda.synchronize()
    # ---- REGIONS timed regions of exactly K steps each, every one bracketed by barrier + synchronize, time = max over ranks ------------
    evs, region_ms = [], []
    for r in range(REGIONS):
        launches[0] = 0
        if world > 1:
            dist.barrier()
        t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t_start.record()
        run_steps(first_timed + r * region_stride, K, evs)
        t_end.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t = torch.tensor([t_start.elapsed_time(t_end)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        region_ms.append(float(t.item()))
    launches_per_region = launches[0]
    ms_total = median(region_ms)
    gather_ms = [(a.elapsed_time(b), nb) for a, b, nb in evs]
    torch.maximum(fail, status.max().reshape(1), out=fail)
    if int(fail.item()) != 0:
        raise RuntimeError("region sampling reported failed slots")
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
        first = None
        for f, l, c in api.torch_generator(batch_size=BATCH, n_batches=n_batches, batches_per_worker=2):
            h_labels.copy_(l, non_blocking=True)
            h_coords.copy_(c, non_blocking=True)
            if to_host:
                h_feats.copy_(f, non_blocking=True)
            torch.cuda.current_stream().synchronize()                 # the consumer reads this step's result
            if first is None:
                first = time.perf_counter() - t0
        return time.perf_counter() - t0, first

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
    e2e_s, first_s = run_api(api, K)                                 # K batches from the pinned host slide (in place, or upload first: cost rule)
    in_flight = int(api.upload_in_flight_bytes())                    # background upload not finished when the timed region ended
    zero_copy = int(api.zero_copy_bytes)
    uploaded_in_region = int(api.uploaded_bytes) if zero_copy == 0 else 0
    steady_s, _ = run_api(api, K)                                    # same call again: slide already resident (run_api synchronises first)
    ingest = api.ingest_stats()
    t = torch.tensor([e2e_s, steady_s, first_s or 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s, steady_s, first_s = (float(x) for x in t.tolist())
    e2e_value, steady_value = world * K * BATCH / e2e_s, world * K * BATCH / steady_s
    h_feats = torch.empty((BATCH, PS, PS, 3) if mode["layout"] == "NHWC" else (BATCH, 3, PS, PS), dtype=out_dtype).pin_memory()
    kh = max(2, min(K, 32 if BATCH <= 256 else 4))
    run_api(api, 2, True, h_feats)
    e2e_host_s, _ = run_api(api, kh, True, h_feats)
    slide_bytes = int(host_slide.nbytes)
    del api, host_slide, h_feats
    torch.cuda.empty_cache()

    # ---- the stitch kernels against the same roofline (N = 1), and BASELINE's second metric: whole-slide prediction -----------------
    stitch = stitch_rooflines(torch) if (world == 1 and args.workload == "annotated_rnd" and not args.no_stitch) else {}
    predict = {}
    if args.workload == "annotated_rnd" and not args.no_predict:
        pargs = argparse.Namespace(slide=list(PREDICT_HW), bf16=True, fold_bn=True, cudnn_benchmark=True, cnn_batch=1024, fused=True,
                                   steps=PREDICT_STEPS, warmup=1)
        predict = predict_measure(pargs, dev, world, rank, e2e_steps=1)["flat"]
    clk = clocks.finish()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peaks()
    per_patch = PATCH_IN * (1 + mode["esize"])
    patches_timed = sum(nb for _, nb in gather_ms) * BATCH
    gather_total_ms = sum(m for m, _ in gather_ms)
    achieved = per_patch * patches_timed / (gather_total_ms / 1e3) / 1e9
    launch_patches = sorted({nb * BATCH for _, nb in gather_ms})
    traffic, traffic_src = None, None
    tf = ROOT / "profiles" / "gather_traffic.json"
    if tf.exists() and args.workload == "annotated_rnd":
        doc = json.loads(tf.read_text())
        caps = doc.get("captures", [doc])                                # one ncu --set full capture per launch shape
        by_patches = {int(c.get("algorithmic_bytes_per_launch", 0)) // per_patch: c for c in caps}
        if len(launch_patches) == 1 and launch_patches[0] in by_patches:  # a capture of exactly the launch shape timed here
            cap = by_patches[launch_patches[0]]
            traffic = cap.get("dram_bytes_per_launch")
            traffic_src = f"profiles/gather_traffic.json: {cap.get('source', 'ncu --set full')} ({launch_patches[0]} patches per launch, as timed here)"
        else:
            traffic_src = (f"not reported: profiles/gather_traffic.json holds ncu captures of {sorted(by_patches)}-patch launches, the launches timed "
                           f"here hold {launch_patches} patches")
    line = {
        "metric": METRIC, "value": value, "unit": "patches/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": mode["dtype"],
        "data": "synthetic", "config": mode_config(args.workload, world),
        "gigapixels_per_s": value * PS * PS / 1e9,
        "timed_regions": REGIONS, "region_ms_min": min(region_ms), "region_ms_median": ms_total, "region_ms_max": max(region_ms),
        "roofline": dict({"bound": "hbm", "kernel": mode["kernel"],
                          "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                          "peak_source": peak_src, "algorithmic_bytes_per_launch": per_patch * patches_timed / len(gather_ms),
                          "algorithmic_bytes_per_patch": per_patch, "patches_per_launch": patches_timed / len(gather_ms),
                          "launches_timed": len(gather_ms), "kernel_ms_avg": gather_total_ms / len(gather_ms),
                          "kernel_ms_median": median([m for m, _ in gather_ms]),
                          "kernel_ms_per_batch": gather_total_ms / max(sum(nb for _, nb in gather_ms), 1), "frac_of_nominal_8TBs": achieved / 8000.0},
                         **stitch),
        "e2e": dict({"value": e2e_value, "unit": "patches/s",
                     "h2d_bytes_per_step": (zero_copy + uploaded_in_region) / K, "d2h_bytes_per_step": d2h,
                     "api": f"AnnoRegionRndSampler.torch_generator(batch_size={BATCH}, n_batches=K) over a slide in pinned host memory, per step the D2H "
                            "read of labels+coords. Cost rule of the sampler: a job whose patches need fewer bytes than half the slide reads them IN "
                            "PLACE from the pinned buffer (the gather kernel's bulk row copies go over PCIe: zero_copy_bytes, counted as patch "
                            "bytes ps*ps*3; row alignment adds <= 5 %) and the layer is uploaded in the background AFTER the job's last gather "
                            "(upload_bytes_in_flight_at_end); a longer job uploads the layer first"
                            + ("; the ranks hold the same slide, so a rank uploads 1/world of its rows over its own PCIe link and one NCCL "
                               "all-gather over NVLink replicates them (slide.sharded_upload)" if world > 1 else ""),
                     "seconds": e2e_s, "ms_first_batch": 1e3 * first_s, "zero_copy_bytes": zero_copy, "uploaded_bytes_in_region": uploaded_in_region,
                     "upload_bytes_in_flight_at_end": in_flight, "ms_alloc": ingest["alloc_ms"], "ms_upload": ingest["upload_ms"],
                     "ms_allgather": ingest["allgather_ms"], "upload_bytes": ingest["bytes"], "slide_bytes": slide_bytes,
                     "steady_value": steady_value, "steady_seconds": steady_s,
                     "features_to_host_value": kh * BATCH / e2e_host_s, "features_to_host_d2h_bytes_per_step": d2h + BATCH * PS * PS * 3 * mode["esize"]},
                    **predict),
        "gpu_launches": launches_per_region,
        "clocks": clk,
    }
    if args.with_training and args.workload == "train_input":
        api2 = new_api(source, 55 + rank)
        line["training_consumer"] = training_consumer(api2, dev, BATCH, torch)
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


def stitch_rooflines(torch, reps=5) -> dict:
    """roofline.stitch_*: CUDA-event median of `reps` launches of the stitch kernels on the 40k x 40k stride-112 case, sum map, n = 5.
    Algorithmic bytes = dh*dw*n*4 (every output written once) + P*n*4 (logits read once); frac = that / time / hbm peak."""
    from deephisto_b200 import ops

    peak, _ = measured_peaks()
    out = {}
    n = 5

    def timed(fn, group=1):
        # `group` back-to-back calls per timed region: the python + ctypes launch path (~25 us per call) would otherwise sit between the
        # first event and the kernel and be timed as kernel time -- 40 % of a d = 16 launch, 6 % at d = 4
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(group):
                fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) / group)
        return median(ts)

    def ring_for(d):
        # d = 16: the 125 MB map is as large as the L2 (126 MB) -- the calls of a group write a ring of 3 maps, so every launch's lines
        # have been evicted before they are written again; larger maps (>= 2 GB) need no ring
        return 3 if d >= 8 else 1

    def group_for(d):
        return {16: 8, 4: 4, 2: 2}.get(d, 1)

    def cover_list(H, W):
        st = ops.CoverState(H, W, PS, 16, 2, 1024, seed=0)
        cells = (H // 16) * (W // 16)
        parts = []
        while True:
            c, counts = st.next_group(16)
            parts.append(c.reshape(-1, 2))
            if int(counts[-1].item()) >= cells:
                keep = int((counts < cells).sum().item()) + 1           # batches up to and including the one that completes coverage
                parts[-1] = parts[-1][: keep * 1024]
                break
        return torch.cat(parts).contiguous()

    keep = {}
    H, W = STITCH_HW
    npad = ops.dense_count(H, W, PS, 112, 64)[1]
    g = torch.Generator(device="cuda").manual_seed(0)
    dense_lg = torch.randn((npad, n), generator=g, device="cuda")
    for d in (1, 2, 4, 16):
        ring = [None] * ring_for(d)
        pos = [0]

        def run():
            ring[pos[0]] = None                                         # free the oldest map first (d = 1: 32 GB)
            ring[pos[0]] = ops.stitch_dense(dense_lg, H, W, PS, 112, d, 64, want_sum=True)
            pos[0] = (pos[0] + 1) % len(ring)
        ms = timed(run, group_for(d))
        del ring
        alg = (H // d) * (W // d) * n * 4 + npad * n * 4
        out[f"stitch_dense_d{d}_frac"] = alg / ms / 1e6 / peak
        out[f"stitch_dense_d{d}_ms"] = ms
    keep.clear()
    for tag, (h, w), ds in (("", (H, W), (1, 2, 4, 16)), ("unaligned_", (H - 1, W - 1), (4,))):
        coords = cover_list(h, w)
        lg = torch.randn((coords.shape[0], n), generator=g, device="cuda")
        for d in ds:
            ring = [None] * ring_for(d)
            pos = [0]

            def run():
                ring[pos[0]] = None
                ring[pos[0]] = ops.stitch_binned(lg, coords, PS, d, h // d, w // d, want_sum=True)
                pos[0] = (pos[0] + 1) % len(ring)
            ms = timed(run, group_for(d))
            del ring
            alg = (h // d) * (w // d) * n * 4 + coords.shape[0] * n * 4
            out[f"stitch_binned_{tag}d{d}_frac"] = alg / ms / 1e6 / peak
            out[f"stitch_binned_{tag}d{d}_ms"] = ms
        out[f"stitch_binned_{tag}patches"] = int(coords.shape[0])
        keep.clear()
    del dense_lg
    torch.cuda.empty_cache()
    return out


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
        "workload": f"examples.predict_full_patched (BASELINE configs[{2 if hw[0] <= 40000 else 3}]): patch_cls_simple ResNet18 (random init, seed 0, "
                    f"{'bf16 channels_last' if args.bf16 else 'fp32, torch defaults (TF32 convolutions)'}{', BatchNorm folded' if args.fold_bn else ''}"
                    f"{', cudnn.benchmark' if args.cudnn_benchmark else ''}{', fused forward' if getattr(args, 'fused', False) else ''}) on a synthetic {hw[0]}x{hw[1]} slide, "
                    f"224x224 patches at stride 112, dense sampler batch 64 (CNN batch {args.cnn_batch}), stitch downscale 16, argmax map",
        "slide": list(hw), "patch": PS, "stride": 112, "downscale": 16, "cnn_batch": args.cnn_batch,
        "l2_policy": "inputs larger than L2: each step re-reads the whole slide band (GBs) from HBM",
    }


def band_parity_check(torch, dist, dev, rank, world) -> int:
    """Small banded-vs-single self-check inside the bench run (N > 1): with logits that are a fixed function of the patch pixels the
    NCCL-assembled sum map and class map must equal the single-GPU ones bit for bit (the FixedLogits case of tests/test_multigpu.py)."""
    from deephisto_b200.anno.utils import AnnoDescription
    from deephisto_b200.examples import predict_full_patched as pfp
    from deephisto_b200.patch_samplers import full_samplers as fs
    from deephisto_b200.slide import SyntheticSlide

    class FixedLogits(pfp.DeviceBatchPredictor):
        def logits(self, features):
            f = features.float()
            return torch.stack([f[:, 0, 3, 7], f[:, 1, 100, 50], f[:, 2, 223, 223], f[:, 0, 0, 0] * 2, f[:, 1, 17, 200] - f[:, 2, 5, 5]], 1).contiguous()

    anno = AnnoDescription.with_auto_colors([f"c{i}" for i in range(5)])
    mode = fs.SamplerExecutionMode.INMEMORY_SINGLEPROC
    model = torch.nn.Identity()
    H, W, stride, d = 3000, 2100, 112, 16
    ok = 1
    try:
        lazy = fs.FullImageDenseSampler(SyntheticSlide(H, W, seed=5), 1, PS, 64, mode, stride=stride, device=dev, lazy_slide=True)
        out = pfp.ImagePredictorPatched(None, lazy, FixedLogits(model, dev), anno, layer=1, downscale=d, device=dev, cnn_batch=96).process_device(
            want_sum=True, rank=rank, world=world)
        full_s = fs.FullImageDenseSampler(SyntheticSlide(H, W, seed=5), 1, PS, 64, mode, stride=stride, device=dev)
        full = pfp.ImagePredictorPatched(None, full_s, FixedLogits(model, dev), anno, layer=1, downscale=d, device=dev, cnn_batch=128).process_device(want_sum=True)
        ok = int(torch.equal(out["sum"], full["sum"]) and torch.equal(out["argmax"], full["argmax"]) and out["argmax"].shape == (H // d, W // d))
    except Exception as e:                                              # a failed self-check must not hide the measured numbers
        print(f"band parity self-check raised: {e!r}", file=sys.stderr)
        ok = 0
    t = torch.tensor([ok], dtype=torch.int32, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return int(t.item())


def predict_measure(args, dev, world, rank, e2e_steps=None) -> dict:
    """Whole-slide patched prediction on `world` ranks (process group already initialised when world > 1): args.warmup warm-up slides,
    args.steps timed slides with this rank's slide band resident (CUDA events, max over ranks), then e2e slides starting from the band in
    pinned host memory. Returns {"flat": the e2e.predict_* keys, "line": pieces of the standalone --workload predict line}."""
    import torch
    import torch.distributed as dist

    from deephisto_b200 import bands
    from deephisto_b200.anno.utils import AnnoDescription
    from deephisto_b200.examples import predict_full_patched as pfp
    from deephisto_b200.patch_samplers import full_samplers as fs
    from deephisto_b200.slide import PinnedSlide, SyntheticSlide

    (H, W), cfg = predict_config(world, args)
    torch.manual_seed(0)
    model = pfp.get_model(5)
    torch.backends.cudnn.benchmark = bool(args.cudnn_benchmark)
    pred = pfp.DeviceBatchPredictor(model, dev, torch.bfloat16 if args.bf16 else torch.float32, fold_bn=bool(args.fold_bn),
                                    fused=bool(getattr(args, "fused", False)))
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
    host_band = PinnedSlide.from_device(band, y_origin=y_off, full_height=H)     # setup, not timed
    h_map = torch.empty(out.shape, dtype=torch.uint8).pin_memory()
    del band, sampler, ipp
    torch.cuda.empty_cache()
    sampler2 = fs.FullImageDenseSampler(host_band, 1, PS, 64, mode, stride=112, device=dev, lazy_slide=True)
    ipp2 = pfp.ImagePredictorPatched(host_band, sampler2, pred, anno, layer=1, downscale=16, device=dev, cnn_batch=args.cnn_batch)

    def step2():
        return ipp2.process_device(rank=rank, world=world)["argmax"] if world > 1 else ipp2.dense_band_local(0, 1)["argmax_band"]

    Ke = K if e2e_steps is None else e2e_steps
    h_map.copy_(step2())                                               # warm-up slide of the streamed path: a long-lived process re-uses its
                                                                       # chunk buffers (the allocator pool was emptied above: the first streamed
                                                                       # slide would pay cudaMalloc for every chunk, measured +0.1-0.2 s)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    w0 = time.perf_counter()
    for _ in range(Ke):
        h_map.copy_(step2(), non_blocking=True)
        torch.cuda.synchronize()
    e2e_s = time.perf_counter() - w0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    h2d, d2h = int(host_band.nbytes), int(h_map.numel())
    del host_band, sampler2, ipp2, h_map
    torch.cuda.empty_cache()
    parity = band_parity_check(torch, dist, dev, rank, world) if world > 1 else None
    gpx = H * W / 1e9
    g = bands.dense_grid(H, W, PS, 112, 64)
    halo = sum(bands.plan_band(H, W, PS, 112, 16, 64, r, world).n_patches for r in range(world)) / g.n_padded - 1.0
    ours_ms = sum(v for k, v in stage_ms.items() if k != "cnn")
    flat = {
        "predict_gpx_per_s": K * gpx / (ms_total / 1e3), "predict_e2e_gpx_per_s": Ke * gpx / e2e_s, "predict_s_per_slide": ms_total / K / 1e3,
        "predict_ms_ours": ours_ms, "predict_ms_cnn": stage_ms.get("cnn", 0.0), "predict_ms_gather": stage_ms.get("coords+gather", 0.0),
        "predict_ms_stitch": stage_ms.get("stitch", 0.0), "predict_ms_assemble": stage_ms.get("assemble (NCCL all-gather)", 0.0),
        "predict_halo_frac": halo, "predict_band_parity": parity, "predict_patches_per_s": K * g.n_padded / (ms_total / 1e3),
        "predict_slide_px": H * W, "predict_steps": K, "predict_e2e_steps": Ke, "predict_h2d_bytes_per_slide_per_rank": h2d,
        "predict_d2h_bytes_per_slide": d2h,
        "predict_cnn": ("bf16 channels_last" if args.bf16 else "fp32 (TF32 convolutions)") + (", BatchNorm folded" if args.fold_bn else "")
                       + (", cudnn.benchmark" if args.cudnn_benchmark else "")
                       + (", FusedResNetForward (space-to-depth stem written by the gather, dh_maxpool3x3s2_nhwc, cuDNN fused conv+bias(+add)+relu calls)"
                          if getattr(args, "fused", False) else "")
                       + f", CNN batch {args.cnn_batch} (convolutions = torch/cuDNN: not part of the rebuilt path)",
    }
    return {"flat": flat, "cfg": cfg, "ms_total": ms_total, "e2e_s": e2e_s, "Ke": Ke, "stage_ms": stage_ms, "grid": g, "plan": plan, "hw": (H, W),
            "h2d": h2d, "d2h": d2h}


def ours_predict(args):
    import torch
    import torch.distributed as dist

    from deephisto_b200 import _lib

    world, rank, local = dist_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        init_nccl(dev)
    _lib.require_device()
    clocks = ClockSampler(local, enabled=rank == 0)
    clocks.start()
    m = predict_measure(args, dev, world, rank)
    clk = clocks.finish()
    if rank == 0:
        (H, W), K, g = m["hw"], args.steps, m["grid"]
        gpx = H * W / 1e9
        line = {
            "metric": "WSI gigapixels/sec patched predict", "value": K * gpx / (m["ms_total"] / 1e3), "unit": "Gpx/s", "n_gpus": world, "steps": K,
            "warmup": max(1, args.warmup), "ms_per_step": m["ms_total"] / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16" if args.bf16 else "f32", "data": "synthetic",
            "config": dict(m["cfg"], parallelism=f"row bands x{world} with patch-size halo (recomputed halo patch rows), NCCL all-gather of the u8 class-map bands"),
            "patches_per_s": K * g.n_padded / (m["ms_total"] / 1e3), "patches_per_slide": g.n_padded, "patches_this_rank": m["plan"].n_patches,
            "e2e": dict({"value": m["Ke"] * gpx / m["e2e_s"], "unit": "Gpx/s", "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": m["d2h"],
                         "api": "per step: ImagePredictorPatched.process_device on a lazy sampler over this rank's slide band in pinned host memory (row "
                                "chunks of <= 1 GiB uploaded on a copy stream one chunk ahead of the CNN), class map copied to pinned host memory "
                                "(h2d/d2h bytes are per rank)"}, **m["flat"]),
            "stage_ms_per_step_rank0": m["stage_ms"],
            "roofline": None, "gpu_launches": None, "clocks": clk,
            "note": "CNN-bound (torch/cuDNN ResNet18, not part of the rebuilt path): stage_ms_per_step_rank0 separates this repo's kernels "
                    "(coords+gather, stitch, assemble) from the CNN",
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def reference_arm_predict(args):
    """ImagePredictorPatched.process + batch_predictor restated on the CPU (oracle), 2048 x 2048 crop, model on the host cores."""
    crop = cpu_predict_crop(steps=max(1, min(args.steps, 2)))
    val = crop["gpx_per_s"]
    emit({"impl": "reference", "metric": "WSI gigapixels/sec patched predict", "value": val, "unit": "Gpx/s", "n_gpus": args.gpus,
          "steps": crop["n"], "warmup": 0, "ms_per_step": 1e3 * crop["s_per_crop"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
          "dtype": "f32", "data": "synthetic", "config": {"workload": "oracle port of examples.predict_full_patched on a 2048x2048 crop, CPU ResNet18"},
          "cpu_baseline": {"value": val, "unit": "Gpx/s", "cores": crop["threads"], "kind": "port", "sample": f"{crop['n']} x 2048x2048 crop"},
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
    ap.add_argument("--fused", action="store_true", help="predict workload: FusedResNetForward (bf16; space-to-depth stem, fused cuDNN epilogues)")
    ap.add_argument("--with-training", action="store_true", help="train_input workload: also time a ResNet18 bf16 training step fed by the sampler")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-predict", action="store_true", help="default workload: skip the e2e.predict_* part (100k x 100k whole-slide prediction)")
    ap.add_argument("--no-stitch", action="store_true", help="default workload: skip the roofline.stitch_* part")
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
