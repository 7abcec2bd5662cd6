"""Round-2 measurement probes (not a product path). CUDA-event medians, printed as text on stderr and JSON on stdout.
    python profiles/r02_probe.py cover      coverage sampler: persistent group kernel vs one launch per batch
    python profiles/r02_probe.py binned     dh_stitch_binned: segment kernel (variant 0) vs row-run kernels (variant 1), aligned / unaligned rows
    python profiles/r02_probe.py zerocopy   gather reading a pinned HOST slide in place (PCIe) vs ring depth
    python profiles/r02_probe.py gather     gather kernel, bf16 / fp32 output modes at 8 192 patches per launch
    python profiles/r02_probe.py cnn        where the ResNet18 forward (torch/cuDNN) spends its time, and library-level variants"""
import json
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from deephisto_b200 import _lib, ops  # noqa: E402

peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
lib = _lib.require_device()
PS, N = 224, 5
rows = []


def timeit(fn, reps=10, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def say(**kw):
    rows.append(kw)
    print("  ".join(f"{k}={v:.4g}" if isinstance(v, float) else f"{k}={v}" for k, v in kw.items()), file=sys.stderr)


def cover_list(H, W, B=1024):
    st = ops.CoverState(H, W, PS, 16, 2, B, seed=0)
    cells = (H // 16) * (W // 16)
    parts = []
    while True:
        c, counts = st.next_group(16)
        parts.append(c.reshape(-1, 2))
        if int(counts[-1].item()) >= cells:
            keep = int((counts < cells).sum().item()) + 1
            parts[-1] = parts[-1][: keep * B]
            break
    return torch.cat(parts).contiguous()


def cover():
    for variant in (0, 1):
        lib.dh_cover_set_variant(variant)
        for (h, w, B) in ((40000, 40000, 64), (40000, 40000, 1024), (100000, 100000, 64), (8192, 8192, 64)):
            for group in (1, 16, 64):
                st = ops.CoverState(h, w, PS, 16, 2, B, seed=0)
                st.next_group(4)                                      # state built, kernels loaded
                ms = timeit(lambda: st.next_group(group), reps=8, warm=1) / group
                say(kernel="cover", variant=variant, slide=h, B=B, group=group, us_per_batch=1e3 * ms, patches_per_s=B / ms * 1e3)
    lib.dh_cover_set_variant(0)
    # the whole sampler through the public API: 40k x 40k to full coverage at batch 64, gather included
    import time

    from deephisto_b200.patch_samplers import full_samplers as fs
    from deephisto_b200.slide import SyntheticSlide

    src = SyntheticSlide(40000, 40000, seed=0)
    for rep in range(2):
        s = fs.FullImageRndSampler(src, 1, PS, 64, fs.SamplerExecutionMode.INMEMORY_SINGLEPROC, seed=rep, quiet=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 0
        for f, c, r in s.generator_torch():
            n += f.shape[0]
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        say(kernel="FullImageRndSampler.generator_torch 40k^2 to coverage", rep=rep, patches=n, ms=1e3 * dt, patches_per_s=n / dt)


def binned():
    which = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [40000, 39999]
    codes = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0]
    for hw in which:
        H = W = hw
        coords = cover_list(H, W)
        P = coords.shape[0]
        logits = torch.randn((P, N), generator=torch.Generator(device="cuda").manual_seed(0), device="cuda")
        for d in (16, 4, 2, 1):
            dh, dw = H // d, W // d
            cells = dh * dw
            reps = 10 if d >= 4 else 4
            for label, kw, out_bytes in (("sum", dict(want_sum=True), cells * N * 4), ("argmax", dict(want_sum=False, want_argmax=True), cells)):
                res = {}
                for variant in ((3, 2, 1) if label == "sum" else (2, 1)):
                    for code in (codes if label == "sum" else [0]):
                        lib.dh_stitch_binned_set_variant(variant)
                        lib.dh_stitch_binned_set_tile_rows(code)
                        keep = {}

                        def run():
                            keep["o"] = None
                            keep["o"] = ops.stitch_binned(logits, coords, PS, d, dh, dw, **kw)

                        ms = timeit(run, reps)
                        res[variant] = keep["o"]
                        alg = out_bytes + P * N * 4
                        say(kernel="stitch_binned", hw=hw, P=P, d=d, out=label, variant=variant, code=code, ms=ms, GBs=alg / ms / 1e6, frac=alg / ms / 1e6 / peak)
                        keep.clear()
                lib.dh_stitch_binned_set_tile_rows(0)
                for va in sorted(v for v in res if v != 1):
                    a, b = res[va], res[1]
                    same = all((x is None and y is None) or torch.equal(x.view(torch.int32) if x.dtype == torch.float32 else x,
                                                                        y.view(torch.int32) if y.dtype == torch.float32 else y) for x, y in zip(a, b))
                    say(kernel="stitch_binned bit-identical (variant %d vs 1)" % va, hw=hw, d=d, out=label, same=bool(same))
                del res, a, b
                torch.cuda.empty_cache()
    lib.dh_stitch_binned_set_variant(0)


def zerocopy():
    from deephisto_b200.slide import PinnedSlide

    H = W = 32768
    dev = ops.DeviceSlide.synthetic(H, W, 0)
    host = PinnedSlide.from_device(dev)
    mapped = ops.MappedHostSlide(host.host, H, W, host.pitch)
    g = torch.Generator(device="cuda").manual_seed(0)
    for n in (5120, 1024):
        coords = torch.stack([torch.randint(0, H - PS, (n,), generator=g, device="cuda"), torch.randint(0, W - PS, (n,), generator=g, device="cuda")], 1).to(torch.int32).contiguous()
        out = torch.empty((n, PS, PS, 3), dtype=torch.float32, device="cuda")
        for stages in (2, 3, 4, 6, 8):
            os.environ["DH_GATHER_STAGES"] = str(stages)
            for occ in (0, 1, 2, 4):
                os.environ["DH_GATHER_OCC"] = str(occ)
                ms = timeit(lambda: ops.gather_normalize(mapped, coords, PS, out=out), reps=5, warm=1)
                say(kernel="gather zero-copy (pinned host slide)", patches=n, stages=stages, occ=occ, ms=ms, pcie_GBs=n * PS * PS * 3 / ms / 1e6, patches_per_s=n / ms * 1e3)
        os.environ.pop("DH_GATHER_STAGES")
        os.environ.pop("DH_GATHER_OCC")
        ref = ops.gather_normalize(dev, coords, PS)
        say(kernel="zero-copy == resident", patches=n, same=bool(torch.equal(ref, ops.gather_normalize(mapped, coords, PS))))
    up = torch.empty(H * host.pitch, dtype=torch.uint8, device="cuda")
    ms = timeit(lambda: up.copy_(host.host, non_blocking=True), reps=3, warm=1)
    say(kernel="contiguous upload of the slide (copy engine)", ms=ms, pcie_GBs=H * host.pitch / ms / 1e6)


def gather():
    H = W = 32768
    dev = ops.DeviceSlide.synthetic(H, W, 0)
    g = torch.Generator(device="cuda").manual_seed(0)
    n = 8192
    coords = torch.stack([torch.randint(0, H - PS, (n,), generator=g, device="cuda"), torch.randint(0, W - PS, (n,), generator=g, device="cuda")], 1).to(torch.int32).contiguous()
    for dtype, esz in ((torch.bfloat16, 2), (torch.float32, 4)):
        for layout in ("NHWC", "NCHW"):
            out = torch.empty((2, n, PS, PS, 3) if layout == "NHWC" else (2, n, 3, PS, PS), dtype=dtype, device="cuda")
            for stages in (2, 3, 4):
                os.environ["DH_GATHER_STAGES"] = str(stages)
                for occ in (0, 2, 3, 4, 5):
                    os.environ["DH_GATHER_OCC"] = str(occ)
                    i = [0]

                    def run():
                        i[0] ^= 1
                        ops.gather_normalize(dev, coords, PS, dtype=dtype, layout=layout, out=out[i[0]])

                    ms = timeit(run, reps=7, warm=2)
                    alg = n * PS * PS * 3 * (1 + esz)
                    say(kernel="gather", dtype=str(dtype).split(".")[-1], layout=layout, stages=stages, occ=occ, ms=ms, GBs=alg / ms / 1e6, frac=alg / ms / 1e6 / peak)
            os.environ.pop("DH_GATHER_STAGES")
            os.environ.pop("DH_GATHER_OCC")
            del out
            torch.cuda.empty_cache()


def gather2():
    """bf16 output modes, occupancy x ring depth, three interleaved repetitions (box noise is ~1-2 %)."""
    H = W = 32768
    dev = ops.DeviceSlide.synthetic(H, W, 0)
    g = torch.Generator(device="cuda").manual_seed(0)
    n = 8192
    coords = torch.stack([torch.randint(0, H - PS, (n,), generator=g, device="cuda"), torch.randint(0, W - PS, (n,), generator=g, device="cuda")], 1).to(torch.int32).contiguous()
    for layout in ("NHWC", "NCHW", "S2D48"):
        shape = {"NHWC": (2, n, PS, PS, 3), "NCHW": (2, n, 3, PS, PS), "S2D48": (2, n, PS // 4, PS // 4, 48)}[layout]
        out = torch.empty(shape, dtype=torch.bfloat16, device="cuda")
        res = {}
        for rep in range(3):
            for stages in (2, 3):
                for occ in (2, 3, 4):
                    os.environ["DH_GATHER_STAGES"], os.environ["DH_GATHER_OCC"] = str(stages), str(occ)
                    i = [0]

                    def run():
                        i[0] ^= 1
                        ops.gather_normalize(dev, coords, PS, dtype=torch.bfloat16, layout=layout, out=out[i[0]])

                    res.setdefault((stages, occ), []).append(timeit(run, reps=11, warm=2))
        os.environ.pop("DH_GATHER_STAGES")
        os.environ.pop("DH_GATHER_OCC")
        alg = n * PS * PS * 3 * 3
        for (stages, occ), ms in sorted(res.items()):
            say(kernel="gather bf16", layout=layout, stages=stages, occ=occ, ms_reps=" ".join(f"{m:.4f}" for m in ms), frac_best=alg / min(ms) / 1e6 / peak,
                frac_median=alg / sorted(ms)[1] / 1e6 / peak)
        del out
        torch.cuda.empty_cache()


def gather4():
    """fp32 NHWC (the bench's roofline kernel) at 5 120 and 8 192 patches per launch: occupancy x ring depth, three interleaved repetitions."""
    H = W = 32768
    dev = ops.DeviceSlide.synthetic(H, W, 0)
    g = torch.Generator(device="cuda").manual_seed(0)
    for n in (5120, 8192):
        coords = torch.stack([torch.randint(0, H - PS, (n,), generator=g, device="cuda"), torch.randint(0, W - PS, (n,), generator=g, device="cuda")], 1).to(torch.int32).contiguous()
        out = torch.empty((2, n, PS, PS, 3), dtype=torch.float32, device="cuda")
        res = {}
        for rep in range(3):
            for stages in (2, 3):
                for occ in (2, 3, 4):
                    os.environ["DH_GATHER_STAGES"], os.environ["DH_GATHER_OCC"] = str(stages), str(occ)
                    i = [0]

                    def run():
                        i[0] ^= 1
                        ops.gather_normalize(dev, coords, PS, out=out[i[0]])

                    res.setdefault((stages, occ), []).append(timeit(run, reps=11, warm=2))
        os.environ.pop("DH_GATHER_STAGES")
        os.environ.pop("DH_GATHER_OCC")
        alg = n * PS * PS * 3 * 5
        for (stages, occ), ms in sorted(res.items()):
            say(kernel="gather fp32 NHWC", patches=n, stages=stages, occ=occ, ms_reps=" ".join(f"{m:.4f}" for m in ms), frac_median=alg / sorted(ms)[1] / 1e6 / peak)
        del out
        torch.cuda.empty_cache()


def gather5():
    """bf16 with per-batch random H/V flips (BASELINE configs[4]): NCHW (mirrored units on the fast path) vs NHWC (H flips on the byte path)."""
    H = W = 32768
    dev = ops.DeviceSlide.synthetic(H, W, 0)
    g = torch.Generator(device="cuda").manual_seed(0)
    n = 8192
    coords = torch.stack([torch.randint(0, H - PS, (n,), generator=g, device="cuda"), torch.randint(0, W - PS, (n,), generator=g, device="cuda")], 1).to(torch.int32).contiguous()
    flips = {"none": None, "all four combinations": torch.randint(0, 4, (n,), generator=g, device="cuda").to(torch.uint8),
             "V only": (torch.randint(0, 2, (n,), generator=g, device="cuda") * 2).to(torch.uint8), "H only": torch.randint(0, 2, (n,), generator=g, device="cuda").to(torch.uint8)}
    for layout in ("NCHW", "NHWC"):
        out = torch.empty((2, n, PS, PS, 3) if layout == "NHWC" else (2, n, 3, PS, PS), dtype=torch.bfloat16, device="cuda")
        for name, fl in flips.items():
            i = [0]

            def run():
                i[0] ^= 1
                ops.gather_normalize(dev, coords, PS, dtype=torch.bfloat16, layout=layout, flip=fl, out=out[i[0]])

            ms = timeit(run, reps=11, warm=2)
            say(kernel="gather bf16 + flips", layout=layout, flips=name, ms=ms, frac=n * PS * PS * 9 / ms / 1e6 / peak)
        del out
        torch.cuda.empty_cache()


def gather3():
    """bf16 NCHW / NHWC / S2D48: rows per tile (DH_GATHER_ROWS) x CTAs per SM, 2-deep ring."""
    H = W = 32768
    dev = ops.DeviceSlide.synthetic(H, W, 0)
    g = torch.Generator(device="cuda").manual_seed(0)
    n = 8192
    coords = torch.stack([torch.randint(0, H - PS, (n,), generator=g, device="cuda"), torch.randint(0, W - PS, (n,), generator=g, device="cuda")], 1).to(torch.int32).contiguous()
    for layout in ("NCHW", "NHWC", "S2D48"):
        shape = {"NHWC": (2, n, PS, PS, 3), "NCHW": (2, n, 3, PS, PS), "S2D48": (2, n, PS // 4, PS // 4, 48)}[layout]
        out = torch.empty(shape, dtype=torch.bfloat16, device="cuda")
        for rows in (8, 16, 28, 32):
            for occ in (2, 3):
                os.environ["DH_GATHER_ROWS"], os.environ["DH_GATHER_OCC"] = str(rows), str(occ)
                i = [0]

                def run():
                    i[0] ^= 1
                    ops.gather_normalize(dev, coords, PS, dtype=torch.bfloat16, layout=layout, out=out[i[0]])

                try:
                    ms = timeit(run, reps=11, warm=2)
                    say(kernel="gather bf16", layout=layout, rows=rows, occ=occ, ms=ms, frac=n * PS * PS * 9 / ms / 1e6 / peak)
                except Exception as e:  # noqa: BLE001
                    say(kernel="gather bf16", layout=layout, rows=rows, occ=occ, error=repr(e)[:80])
        os.environ.pop("DH_GATHER_ROWS")
        os.environ.pop("DH_GATHER_OCC")
        del out
        torch.cuda.empty_cache()


def cnn():
    from deephisto_b200.examples import predict_full_patched as pfp

    torch.manual_seed(0)
    torch.backends.cudnn.benchmark = True
    model = pfp.get_model(5)
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    pred = pfp.DeviceBatchPredictor(model, "cuda", torch.bfloat16, fold_bn=True)
    m = pred.model
    x = torch.rand((B, PS, PS, 3), device="cuda").to(torch.bfloat16).permute(0, 3, 1, 2)        # NHWC storage viewed as NCHW
    with torch.no_grad():
        ms = timeit(lambda: m(x), reps=5, warm=3)
        say(kernel="ResNet18 forward (bf16 channels_last, BN folded, cudnn.benchmark)", batch=B, ms=ms, patches_per_s=B / ms * 1e3,
            tflops=B * 3.64e9 / ms / 1e9)
        # per stage
        stages = [("conv1", lambda t: m.conv1(t)), ("bn1+relu", lambda t: m.relu(m.bn1(t))), ("maxpool", lambda t: m.maxpool(t)),
                  ("layer1", lambda t: m.layer1(t)), ("layer2", lambda t: m.layer2(t)), ("layer3", lambda t: m.layer3(t)), ("layer4", lambda t: m.layer4(t)),
                  ("avgpool+fc", lambda t: m.fc(torch.flatten(m.avgpool(t), 1)))]
        t = x
        for name, fn in stages:
            ms = timeit(lambda: fn(t), reps=5, warm=2)
            out = fn(t)
            say(kernel=f"  stage {name}", ms=ms, in_shape=str(tuple(t.shape)), out_MB=out.numel() * out.element_size() / 1e6)
            t = out
        # conv1 with the input padded to 4 / 8 channels (zero weights: same function)
        for cpad in (4, 8):
            xp = torch.zeros((B, PS, PS, cpad), device="cuda", dtype=torch.bfloat16)
            xp[..., :3] = x.permute(0, 2, 3, 1)
            xp = xp.permute(0, 3, 1, 2)
            wpad = torch.zeros((64, cpad, 7, 7), device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
            wpad[:, :3] = m.conv1.weight
            bias = m.conv1.bias
            ms = timeit(lambda: torch.nn.functional.conv2d(xp, wpad, bias, stride=2, padding=3), reps=5, warm=2)
            err = (torch.nn.functional.conv2d(xp, wpad, bias, stride=2, padding=3).float() - m.conv1(x).float()).abs().max().item()
            say(kernel=f"  conv1 with input padded to {cpad} channels", ms=ms, max_abs_diff=err)
        # fused conv + bias + relu through cuDNN (torch.cudnn_convolution_relu) on a layer1-sized convolution
        t1 = torch.rand((B, 64, 56, 56), device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        conv = m.layer1[0].conv1
        ms_a = timeit(lambda: torch.relu_(torch.nn.functional.conv2d(t1, conv.weight, conv.bias, padding=1)), reps=5, warm=2)
        say(kernel="  layer1 conv3x3 + bias, then relu (2 kernels)", ms=ms_a)
        try:
            ms_b = timeit(lambda: torch.cudnn_convolution_relu(t1, conv.weight, conv.bias, (1, 1), (1, 1), (1, 1), 1), reps=5, warm=2)
            a = torch.relu(torch.nn.functional.conv2d(t1, conv.weight, conv.bias, padding=1))
            b = torch.cudnn_convolution_relu(t1, conv.weight, conv.bias, (1, 1), (1, 1), (1, 1), 1)
            say(kernel="  torch.cudnn_convolution_relu (1 kernel)", ms=ms_b, max_abs_diff=(a.float() - b.float()).abs().max().item(),
                channels_last_out=bool(b.is_contiguous(memory_format=torch.channels_last)))
            z = torch.rand_like(t1)
            ms_c = timeit(lambda: torch.relu_(torch.nn.functional.conv2d(t1, conv.weight, conv.bias, padding=1).add_(z)), reps=5, warm=2)
            ms_d = timeit(lambda: torch.cudnn_convolution_add_relu(t1, conv.weight, z, 1.0, conv.bias, (1, 1), (1, 1), (1, 1), 1), reps=5, warm=2)
            say(kernel="  conv + add + relu: 3 kernels vs torch.cudnn_convolution_add_relu", ms_3=ms_c, ms_fused=ms_d)
        except Exception as e:  # noqa: BLE001
            say(kernel="  torch.cudnn_convolution_relu failed", error=repr(e)[:200])
        # whole forward under a CUDA graph
        try:
            gph = torch.cuda.CUDAGraph()
            static_x = x.clone()
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                for _ in range(2):
                    m(static_x)
            torch.cuda.current_stream().wait_stream(s)
            with torch.cuda.graph(gph):
                static_y = m(static_x)
            ms = timeit(lambda: gph.replay(), reps=5, warm=2)
            say(kernel="ResNet18 forward under a CUDA graph", batch=B, ms=ms, patches_per_s=B / ms * 1e3)
        except Exception as e:  # noqa: BLE001
            say(kernel="CUDA graph capture failed", error=repr(e)[:200])


def cnn2():
    """FusedResNetForward vs the plain bf16 model: total and per stage."""
    from deephisto_b200.examples import predict_full_patched as pfp

    torch.manual_seed(0)
    torch.backends.cudnn.benchmark = True
    model = pfp.get_model(5)
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    plain = pfp.DeviceBatchPredictor(model, "cuda", torch.bfloat16, fold_bn=True)
    fused = pfp.DeviceBatchPredictor(model, "cuda", torch.bfloat16, fused=True)
    f = fused.fused
    x = torch.rand((B, PS, PS, 3), device="cuda").to(torch.bfloat16)
    with torch.no_grad():
        ms = timeit(lambda: plain.logits(x.permute(0, 3, 1, 2)), reps=5, warm=3)
        say(kernel="plain forward (bf16 channels_last, BN folded)", batch=B, ms=ms, patches_per_s=B / ms * 1e3)
        ms = timeit(lambda: f.space_to_depth(x), reps=5, warm=2)
        say(kernel="space_to_depth (torch ops)", ms=ms)
        s2d = f.space_to_depth(x)
        ms = timeit(lambda: f(s2d), reps=5, warm=3)
        say(kernel="FusedResNetForward", batch=B, ms=ms, patches_per_s=B / ms * 1e3, tflops=B * 3.64e9 / ms / 1e9)
        a, b = plain.logits(x.permute(0, 3, 1, 2)), f(s2d)
        say(kernel="fused vs plain logits", max_abs_diff=(a - b).abs().max().item(), max_abs=a.abs().max().item(),
            argmax_agree=(a.argmax(1) == b.argmax(1)).float().mean().item())
        one = (1, 1)
        pad = one if f.stem == "s2d4" else (0, 0)
        pool = ops.maxpool3x3s2_d2s if f.stem == "s2d4" else ops.maxpool3x3s2_nhwc
        ms = timeit(lambda: torch.cudnn_convolution_relu(s2d, f.stem_w, f.stem_b, one, pad, one, 1), reps=5, warm=2)
        say(kernel=f"  stem {f.stem}: conv + bias + relu (cuDNN fused)", ms=ms)
        y = torch.cudnn_convolution_relu(s2d, f.stem_w, f.stem_b, one, pad, one, 1)
        ms = timeit(lambda: pool(y), reps=5, warm=2)
        say(kernel=f"  {pool.__name__}", ms=ms, GBs=(y.numel() * 2 * 1.25) / ms / 1e6, frac=(y.numel() * 2 * 1.25) / ms / 1e6 / peak)
        t = pool(y)
        i = 0
        for w1, b1, stride, w2, b2, down in f.blocks:
            def blk(t=t):
                identity = t if down is None else torch.nn.functional.conv2d(t, down[0], down[1], stride=down[2])
                h = torch.cudnn_convolution_relu(t, w1, b1, stride, one, one, 1)
                return torch.cudnn_convolution_add_relu(h, w2, identity, 1.0, b2, one, one, one, 1)
            ms = timeit(blk, reps=5, warm=2)
            say(kernel=f"  block {i}", ms=ms, in_shape=str(tuple(t.shape)))
            t = blk()
            i += 1
        for bs in (256, 512, 2048):
            xs = torch.rand((bs, PS, PS, 3), device="cuda").to(torch.bfloat16)
            s2 = f.space_to_depth(xs)
            ms = timeit(lambda: f(s2), reps=5, warm=3)
            say(kernel="FusedResNetForward", batch=bs, ms=ms, patches_per_s=bs / ms * 1e3)


def cnn3():
    """Stem alternatives (timing only, random tensors): 2x2 space-to-depth (16 ch, 4x4 conv, 64 out) vs 4x4 space-to-depth (48 ch, 3x3 conv,
    256 out = 2x2 output pixels x 64)."""
    torch.backends.cudnn.benchmark = True
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    cl = torch.channels_last
    one = (1, 1)
    for name, cin, hw, cout, k, pad in (("s2d-2: [B,16,115,115] -> 64, 4x4", 16, 115, 64, 4, 0), ("s2d-4: [B,48,56,56] -> 256, 3x3 pad 1", 48, 56, 256, 3, 1),
                                        ("s2d-4 with 64 padded input channels", 64, 56, 256, 3, 1), ("s2d-2 with 32 channels", 32, 115, 64, 4, 0)):
        x = torch.rand((B, cin, hw, hw), device="cuda").to(torch.bfloat16).contiguous(memory_format=cl)
        w = torch.rand((cout, cin, k, k), device="cuda").to(torch.bfloat16).contiguous(memory_format=cl)
        b = torch.zeros(cout, device="cuda", dtype=torch.bfloat16)
        ms = timeit(lambda: torch.cudnn_convolution_relu(x, w, b, one, (pad, pad), one, 1), reps=5, warm=3)
        ms2 = timeit(lambda: torch.relu_(torch.nn.functional.conv2d(x, w, b, padding=pad)), reps=5, warm=3)
        oh = hw + 2 * pad - k + 1
        say(kernel=name, batch=B, ms_fused=ms, ms_conv_relu=ms2, tflops=2.0 * B * oh * oh * cout * cin * k * k / ms / 1e9)


def ncu_binned():
    """One dh_stitch_binned call per variant (for `ncu -k regex:bin_`): 40k x 40k coverage list, sum map at downscale argv[2]."""
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    hw = int(sys.argv[3]) if len(sys.argv) > 3 else 40000
    coords = cover_list(hw, hw)
    logits = torch.randn((coords.shape[0], N), generator=torch.Generator(device="cuda").manual_seed(0), device="cuda")
    for variant in ([int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else (2, 1, 2, 1)):
        lib.dh_stitch_binned_set_variant(variant)
        ops.stitch_binned(logits, coords, PS, d, hw // d, hw // d, want_sum=True)
        torch.cuda.synchronize()
    lib.dh_stitch_binned_set_variant(0)


def binned2():
    """What bounds the tile kernels at d = 4? (a) the same map with ONE patch (tiles only write zeros: the write pattern alone),
    (b) tile heights / groups, (c) dh_stitch_dense and a plain memset of the same map for reference."""
    hw, d = 40000, int(sys.argv[2]) if len(sys.argv) > 2 else 4
    dh = dw = hw // d
    coords = cover_list(hw, hw)
    logits = torch.randn((coords.shape[0], N), generator=torch.Generator(device="cuda").manual_seed(0), device="cuda")
    one_c, one_l = coords[:1].contiguous(), logits[:1].contiguous()
    alg = dh * dw * N * 4
    buf = torch.empty((dh, dw, N), device="cuda")
    ms = timeit(lambda: buf.zero_(), 10)
    say(kernel="memset of the map", d=d, ms=ms, GBs=alg / ms / 1e6, frac=alg / ms / 1e6 / peak)
    del buf
    for name, c, l in (("coverage list", coords, logits), ("one patch", one_c, one_l)):
        for variant in (4, 3, 1):
            for code in ([int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else (0, 1000000, 3000000)):
                lib.dh_stitch_binned_set_variant(variant)
                lib.dh_stitch_binned_set_tile_rows(code)
                keep = {}

                def run():
                    keep["o"] = None
                    keep["o"] = ops.stitch_binned(l, c, PS, d, dh, dw, want_sum=True)

                ms = timeit(run, 10)
                keep.clear()
                say(kernel="stitch_binned", list=name, d=d, variant=variant, code=code, ms=ms, GBs=alg / ms / 1e6, frac=alg / ms / 1e6 / peak)
    lib.dh_stitch_binned_set_variant(0)
    lib.dh_stitch_binned_set_tile_rows(0)


def binned3():
    """Resident CTAs per SM (extra dynamic shared memory: code = 100000 * KB) x tile kernel, every downscale, aligned and unaligned rows."""
    which = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [40000, 39999]
    for hw in which:
        coords = cover_list(hw, hw)
        logits = torch.randn((coords.shape[0], N), generator=torch.Generator(device="cuda").manual_seed(0), device="cuda")
        for d in (16, 4, 2, 1):
            dh = dw = hw // d
            alg = dh * dw * N * 4 + coords.shape[0] * N * 4
            for variant in (3, 2, 1):
                for code in (0, 1000000, 3000000):
                    lib.dh_stitch_binned_set_variant(variant)
                    lib.dh_stitch_binned_set_tile_rows(code)
                    keep = {}

                    def run():
                        keep["o"] = None
                        keep["o"] = ops.stitch_binned(logits, coords, PS, d, dh, dw, want_sum=True)

                    ms = timeit(run, 10 if d >= 4 else 4)
                    keep.clear()
                    say(kernel="stitch_binned", hw=hw, d=d, variant=variant, code=code, ms=ms, frac=alg / ms / 1e6 / peak)
    lib.dh_stitch_binned_set_variant(0)
    lib.dh_stitch_binned_set_tile_rows(0)


def ncu_dense():
    """A few dh_stitch_dense calls (for `ncu -k regex:stitch_dense`): 40k x 40k, sum map at downscale argv[2]."""
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    H = W = 40000
    npad = ops.dense_count(H, W, PS, 112, 64)[1]
    lg = torch.randn((npad, N), generator=torch.Generator(device="cuda").manual_seed(0), device="cuda")
    for _ in range(3):
        ops.stitch_dense(lg, H, W, PS, 112, d, 64, want_sum=True)
        torch.cuda.synchronize()
    ms = timeit(lambda: ops.stitch_dense(lg, H, W, PS, 112, d, 64, want_sum=True), reps=10)
    say(kernel="stitch_dense sum", d=d, ms=ms, frac=((H // d) * (W // d) * N * 4 + npad * N * 4) / ms / 1e6 / peak)
    ms = timeit(lambda: ops.stitch_dense(lg, H, W, PS, 112, d, 64, want_sum=False, want_argmax=True), reps=10)
    say(kernel="stitch_dense argmax only", d=d, ms=ms)


def dense2():
    """dh_stitch_dense: logits staged per block (variant 0) vs read per row class (variant 1), every downscale and output set,
    aligned (40000) and unaligned (39999) rows; bit-identical check."""
    for hw in (40000, 39999):
        H = W = hw
        npad = ops.dense_count(H, W, PS, 112, 64)[1]
        lg = torch.randn((npad, N), generator=torch.Generator(device="cuda").manual_seed(0), device="cuda")
        for d in (16, 4, 2, 1):
            alg = (H // d) * (W // d) * N * 4 + npad * N * 4
            for label, kw in (("sum", dict(want_sum=True)), ("sum+argmax+count", dict(want_sum=True, want_argmax=True, want_count=True)), ("argmax", dict(want_sum=False, want_argmax=True))):
                if d < 4 and label != "sum":
                    continue
                res = {}
                for variant in (0, 1):
                    lib.dh_stitch_dense_set_variant(variant)
                    keep = {}

                    def run():
                        keep["o"] = None
                        keep["o"] = ops.stitch_dense(lg, H, W, PS, 112, d, 64, **kw)

                    ms = timeit(run, 10 if d >= 4 else 4)
                    res[variant] = keep["o"]
                    keep.clear()
                    say(kernel="stitch_dense", hw=hw, d=d, out=label, variant=variant, ms=ms, frac=alg / ms / 1e6 / peak)
                same = all((x is None and y is None) or torch.equal(x.view(torch.int32) if x.dtype == torch.float32 else x,
                                                                    y.view(torch.int32) if y.dtype == torch.float32 else y) for x, y in zip(res[0], res[1]))
                say(kernel="stitch_dense bit-identical (variant 0 vs 1)", hw=hw, d=d, out=label, same=bool(same))
                del res
                torch.cuda.empty_cache()
    lib.dh_stitch_dense_set_variant(0)


def dense3():
    """dh_stitch_dense at d = argv[2]: rows per block (DH_STITCH_RPB) sweep, grouped launches into a ring of maps as in bench.py."""
    import os
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    H = W = 40000
    npad = ops.dense_count(H, W, PS, 112, 64)[1]
    lg = torch.randn((npad, N), generator=torch.Generator(device="cuda").manual_seed(0), device="cuda")
    alg = (H // d) * (W // d) * N * 4 + npad * N * 4
    group = 8 if d >= 8 else 4
    for label, kw in (("sum", dict(want_sum=True)), ("argmax", dict(want_sum=False, want_argmax=True))):
        for rpb in [int(x) for x in (sys.argv[3].split(",") if len(sys.argv) > 3 else "0,8,10,14,17,20,25,30,40,60".split(","))]:
            if rpb:
                os.environ["DH_STITCH_RPB"] = str(rpb)
            else:
                os.environ.pop("DH_STITCH_RPB", None)
            ring = [None, None, None]
            pos = [0]

            def run():
                for _ in range(group):
                    ring[pos[0]] = None
                    ring[pos[0]] = ops.stitch_dense(lg, H, W, PS, 112, d, 64, **kw)
                    pos[0] = (pos[0] + 1) % 3

            ms = timeit(run, 7) / group
            say(kernel="stitch_dense", d=d, out=label, rpb=rpb, ms=ms, frac=alg / ms / 1e6 / peak)
    os.environ.pop("DH_STITCH_RPB", None)


def ncu_predict_parts():
    """One S2D48 gather launch and one stem pooling launch at the predictor's batch size (for ncu)."""
    H = W = 32768
    dev = ops.DeviceSlide.synthetic(H, W, 0)
    g = torch.Generator(device="cuda").manual_seed(0)
    n = 1024
    coords = torch.stack([torch.randint(0, H - PS, (n,), generator=g, device="cuda"), torch.randint(0, W - PS, (n,), generator=g, device="cuda")], 1).to(torch.int32).contiguous()
    y = torch.randn((n, 256, 56, 56), generator=g, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    for _ in range(3):
        ops.gather_normalize(dev, coords, PS, dtype=torch.bfloat16, layout="S2D48")
        ops.maxpool3x3s2_d2s(y)
        torch.cuda.synchronize()


def ncu_cover():
    st = ops.CoverState(40000, 40000, PS, 16, 2, 64, seed=0)
    for _ in range(3):
        st.next_group(16)
        torch.cuda.synchronize()


if __name__ == "__main__":
    {"dense3": dense3, "dense2": dense2, "binned3": binned3, "binned2": binned2, "ncu_predict_parts": ncu_predict_parts, "gather5": gather5, "gather4": gather4, "gather3": gather3, "gather2": gather2, "cnn3": cnn3, "ncu_dense": ncu_dense, "ncu_binned": ncu_binned, "ncu_cover": ncu_cover, "cnn2": cnn2, "cover": cover, "binned": binned, "zerocopy": zerocopy, "gather": gather, "cnn": cnn}[sys.argv[1]]()
    print(json.dumps({"peak_gbs": peak, "rows": rows}, indent=1))
