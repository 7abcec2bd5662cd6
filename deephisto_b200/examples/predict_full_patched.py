"""Whole-slide patched prediction and stitching -- drop-in for the reference's `examples/predict_full_patched.py`.

Same names and call signatures as the reference: `ImagePredictorPatched(psim_path, patch_sampler, batch_predictor, anno,
layer, downscale=4).process()` (:22-63), `batch_predictor(patches, model, device)` (:66-78),
`perform_and_save_visualizations` (:81-113), `load_model` (:116-126). What changes is where the work happens:

  reference (per batch)                                     here
  np.stack(u8)/255 -> tensor -> .to(device) -> permute       dh_gather_normalize writes the NCHW batch from the HBM-resident slide
  model(features).detach().cpu().numpy()                     logits stay in HBM ([n_padded, n] float32 buffer)
  python loop  prediction[y//d:(y+ps)//d, ...] += logits_i   dh_stitch_dense (reference order, bit-exact sums) / dh_stitch_scatter
  np.argmax(prediction, axis=2)                              fused into the stitch epilogue

`process()` picks the path from what it is given:
  * a FullImageDenseSampler / FullImageRndSampler OBJECT of this package + a DeviceBatchPredictor -> device pipeline;
  * anything else (the reference's usage: an iterator of (list[Patch], progress) and a callable list[Patch] -> ndarray)
    -> the same control flow as the reference with the accumulation done by dh_stitch_scatter on the device.
Multi-GPU (BASELINE config 4): `process(rank=r, world=G)` computes the row band of deephisto_b200/bands.py and
assembles the bands with an NCCL all-gather (`torch.distributed` must be initialised)."""

from __future__ import annotations

import weakref
from pathlib import Path
from typing import Callable, Optional

import numpy as np
import torch

from .. import bands, ops
from ..anno.utils import AnnoDescription
from ..patch_samplers.full_samplers import FullImageDenseSampler, FullImageRndSampler, SamplerExecutionMode
from ..slide import Patch, open_slide


def get_model(n_classes: int, pretrained: bool = False) -> torch.nn.Module:
    """models/patch_cls_simple/model.py:5-11: torchvision ResNet18 with the final layer replaced. The reference asks for
    ImageNet weights (a download); offline the default is random initialisation (weights come from load_model)."""
    import torch.nn as nn
    from torchvision import models

    model = models.resnet18(weights=models.ResNet18_Weights.DEFAULT if pretrained else None)
    model.fc = nn.Linear(model.fc.in_features, n_classes)
    return model


def load_model(weights_path: Path, device, n_classes: int = 5) -> torch.nn.Module:
    """Reference :116-126."""
    model = get_model(n_classes=n_classes).to(device)
    model.load_state_dict(torch.load(weights_path, weights_only=True, map_location=device))
    model.to(device).eval()
    return model


class DeviceBatchPredictor:
    """features -> logits on the device. Called with a CUDA NCHW batch it returns CUDA float32 logits [B, n]; called with a
    list[Patch] (the reference's batch_predictor contract, :66-78) it uploads the uint8 pixels, normalises them with
    dh_gather_normalize and returns a numpy array like the reference.

    dtype float32 keeps torch's defaults (what the reference runs on a GPU); bfloat16 runs the CNN in bf16 channels_last."""

    def __init__(self, model: torch.nn.Module, device="cuda", dtype=torch.float32, channels_last: bool = True, fold_bn: bool = False,
                 fused: bool = False, stem: str = "s2d4", cuda_graph: bool = True):
        """fold_bn=True folds every eval-mode BatchNorm into the preceding convolution (torch.nn.utils.fusion): the same
        function with ~20 fewer memory-bound elementwise kernels per forward; logits change at rounding level only.
        fused=True (bfloat16, torchvision BasicBlock ResNets): the forward runs through FusedResNetForward -- space-to-depth
        stem, dh_maxpool3x3s2_nhwc, cuDNN's fused conv+bias(+residual)+ReLU calls -- over the same weights; `gather` then writes the
        stem's space-to-depth input directly. cuda_graph (fused predictors): the forward over the predictor's gather buffer always has
        the same shape and addresses, so after two eager runs (cuDNN picks its algorithms there) it is captured once and replayed --
        one launch per CNN batch instead of ~45, which keeps a rank's launch thread off the critical path when many ranks share the
        host's cores."""
        self.device = torch.device(device)
        self.dtype = dtype
        self.channels_last = channels_last and dtype != torch.float32
        model = model.to(self.device).eval()
        self.fused = None
        self._s2d = None                                       # space-to-depth stem input of the last gather() (fused predictors)
        self._graph_on, self._graph, self._graph_out, self._graph_buf, self._eager_runs = bool(cuda_graph), None, None, None, 0
        if fused:
            if dtype != torch.bfloat16:
                raise ValueError("fused=True needs dtype=torch.bfloat16 (the float32 predictor is the parity path; the gather writes bf16 stem inputs)")
            self.fused = FusedResNetForward(model, dtype, stem=stem)      # works on its own folded copy of the weights
            self.model = None
        else:
            if fold_bn:
                model = fold_batchnorm(model)                   # a copy
            elif dtype != torch.float32:
                import copy

                model = copy.deepcopy(model)                    # the caller's module keeps its precision
            if dtype != torch.float32:
                model = model.to(dtype)
            if self.channels_last:
                model = model.to(memory_format=torch.channels_last)
            self.model = model
        self._source_model = None                              # weak reference to the caller's module (batch_predictor's cache check)

    def source_model(self):
        return None if self._source_model is None else self._source_model()

    def gather(self, slide, coords: torch.Tensor, ps: int) -> torch.Tensor:
        """[B,3,ps,ps] model input for patches at `coords`, written by dh_gather_normalize in the memory format the model runs in:
        channels_last models get an NHWC buffer viewed as NCHW (no layout pass between the gather and the first convolution)."""
        if self.fused is not None:
            f = self.fused
            if f.stem == "s2d4" and ps % 4:
                raise ValueError(f"the s2d4 stem needs a patch size divisible by 4, got {ps}: build the predictor with stem='s2d2'")
            if self.dtype == torch.bfloat16 and ps % 2 == 0:
                # dh_gather_normalize writes the stem's space-to-depth input itself, into a buffer of this predictor (s2d2: its zero
                # border is written once, the kernel only touches the interior). The returned tensor is valid until the next gather().
                B = coords.shape[0]
                shape = f.s2d_shape(max(B, 1), ps)
                if self._s2d is None or self._s2d.shape[0] < B or tuple(self._s2d.shape[1:]) != (shape[2], shape[3], shape[1]):
                    self._s2d = torch.zeros((shape[0], shape[2], shape[3], shape[1]), dtype=self.dtype, device=coords.device)
                out = ops.gather_normalize(slide, coords, ps, dtype=self.dtype, layout=f.gather_layout, scale255=True, out=self._s2d[:B])
                return out.permute(0, 3, 1, 2)
            return f.space_to_depth(ops.gather_normalize(slide, coords, ps, dtype=self.dtype, layout="NHWC", scale255=True))
        if self.channels_last:
            return ops.gather_normalize(slide, coords, ps, dtype=self.dtype, layout="NHWC", scale255=True).permute(0, 3, 1, 2)
        return ops.gather_normalize(slide, coords, ps, dtype=self.dtype, layout="NCHW", scale255=True)

    def batch_buffer(self, B: int, ps: int):
        """The fused bf16 predictor's own stem-input buffer with room for B patches (None for every other predictor): callers that
        fill a CNN batch from several sources (row chunks of a streamed slide) gather into slices of it with gather_into() and run
        buffer_logits() when it is full."""
        f = self.fused
        if f is None or self.dtype != torch.bfloat16 or ps % 2 or (f.stem == "s2d4" and ps % 4):
            return None
        shape = f.s2d_shape(max(B, 1), ps)
        if self._s2d is None or self._s2d.shape[0] < B or tuple(self._s2d.shape[1:]) != (shape[2], shape[3], shape[1]):
            self._s2d = torch.zeros((shape[0], shape[2], shape[3], shape[1]), dtype=self.dtype, device=self.device)
        return self._s2d

    def gather_into(self, slide, coords: torch.Tensor, ps: int, offset: int) -> None:
        """Patches at `coords` into rows [offset, offset + len(coords)) of batch_buffer()."""
        ops.gather_normalize(slide, coords, ps, dtype=self.dtype, layout=self.fused.gather_layout, scale255=True,
                             out=self._s2d[offset : offset + coords.shape[0]])

    @torch.no_grad()
    def buffer_logits(self) -> torch.Tensor:
        """Logits of every row of batch_buffer() (rows that were not filled since the last call hold older patches)."""
        return self._forward_buffer()

    @torch.no_grad()
    def logits(self, features: torch.Tensor) -> torch.Tensor:
        if self.fused is not None:
            if features.shape[1] == 3:                       # a plain [B,3,H,W] batch (list[Patch] callers): fold it here
                features = self.fused.space_to_depth(features.permute(0, 2, 3, 1))
            B = features.shape[0]
            buf = self._s2d
            if buf is not None and features.data_ptr() == buf.data_ptr() and B < buf.shape[0]:
                # a short batch inside this predictor's gather buffer (the tail of a patch range): run the buffer's full batch -- the rows
                # behind B hold the previous batch, their logits are dropped -- so that cuDNN sees ONE input shape per predictor
                # (cudnn.benchmark re-tunes every new shape; a streamed slide has a different tail per row chunk)
                return self._forward_buffer()[:B]
            if buf is not None and features.data_ptr() == buf.data_ptr() and B == buf.shape[0]:
                return self._forward_buffer()
            return self.fused(features)
        if self.channels_last:
            features = features.contiguous(memory_format=torch.channels_last)
        return self.model(features).float()

    def _forward_buffer(self) -> torch.Tensor:
        """Logits of the whole gather buffer (fused predictors). Eager for the first two runs of a buffer, then a captured CUDA graph
        is replayed; the result is copied out of the graph's static output (callers may keep it across calls)."""
        buf = self._s2d
        x = buf.permute(0, 3, 1, 2)
        if not self._graph_on:
            return self.fused(x)
        if self._graph is not None and self._graph_buf is buf:
            self._graph.replay()
            return self._graph_out.clone()
        if self._graph_buf is not buf:                        # a new (larger) buffer: start over
            self._graph, self._graph_out, self._graph_buf, self._eager_runs = None, None, buf, 0
        self._eager_runs += 1
        if self._eager_runs <= 2:
            return self.fused(x)
        try:
            torch.cuda.current_stream(self.device).synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self.fused(x)
            self._graph, self._graph_out = g, out
            g.replay()
            return out.clone()
        except Exception as e:                                # capture is an optimisation: fall back to eager launches
            import warnings

            warnings.warn(f"CUDA graph capture of the fused forward failed ({e!r}); running eagerly", RuntimeWarning)
            self._graph_on, self._graph, self._graph_out = False, None, None
            return self.fused(x)

    def features_from_patches(self, patches: list[Patch]) -> torch.Tensor:
        """uint8 patch pixels -> [B,3,ps,ps] in [0,1] through the gather kernel: the stacked batch is a (B*ps) x ps 'slide'."""
        ps = patches[0].patch_size
        stack = np.stack([np.asarray(p.data) for p in patches]).reshape(len(patches) * ps, ps, 3)
        slide = ops.DeviceSlide.from_numpy(stack, self.device)
        coords = torch.arange(len(patches), dtype=torch.int32, device=self.device).mul_(ps).reshape(-1, 1)
        coords = torch.cat([coords, torch.zeros_like(coords)], 1).contiguous()
        return ops.gather_normalize(slide, coords, ps, dtype=self.dtype, layout="NCHW", scale255=True)

    def __call__(self, batch):
        if isinstance(batch, torch.Tensor):
            return self.logits(batch)
        return self.logits(self.features_from_patches(batch)).cpu().numpy()


class FusedResNetForward:
    """The eval-mode forward of a torchvision BasicBlock ResNet (ResNet18 / 34: models/patch_cls_simple/model.py:5-11) restated over
    the same weights with cuDNN's fused epilogues and a tensor-core friendly stem. Every convolution is still cuDNN through torch;
    what changes is how it is called (measured on a B200, batch 1024, bf16 channels_last, profiles/r02_predict.md):

      stem   conv1 is a 7x7 stride-2 convolution over THREE input channels: cuDNN has no tensor-core kernel for C = 3 (8.2 ms of the
             19.8 ms forward). The same function on a SPACE-TO-DEPTH image, the 7x7 kernel zero-extended and folded the same way:
             stem "s2d4" (patch size % 4 == 0): 4x4 blocks -> 48 input channels, a 3x3 stride-1 padding-1 convolution with 4 x 64
             output channels (the 2x2 output pixels of a block), 0.70 ms; stem "s2d2": 2x2 blocks -> 12 channels padded to 16, a 4x4
             convolution over an explicitly zero-bordered image, 2.05 ms. Bias + ReLU in the cuDNN epilogue either way.
      pool   3x3 stride-2 max pooling of the stem output by dh_maxpool3x3s2_d2s / _nhwc (HBM-bound; torch's kernel takes 3.0 ms).
      blocks conv + bias + ReLU and conv + bias + residual + ReLU are ONE cuDNN call each (torch.cudnn_convolution_relu /
             cudnn_convolution_add_relu) instead of three kernels; BatchNorm is folded into the convolutions first (eval mode).

    Logits equal the plain bf16 model's up to bf16 rounding / accumulation order (tests/test_predict_gpu.py); the float32 predictor
    (bit-level parity path against the reference's arithmetic) does not use this class."""

    def __init__(self, model: torch.nn.Module, dtype=torch.bfloat16, stem: str = "s2d4"):
        from torchvision.models.resnet import BasicBlock, ResNet

        if stem not in ("s2d4", "s2d2"):
            raise ValueError("stem must be 's2d4' or 's2d2'")
        self.stem = stem

        if not isinstance(model, ResNet) or not all(isinstance(b, BasicBlock) for layer in (model.layer1, model.layer2, model.layer3, model.layer4) for b in layer):
            raise TypeError("FusedResNetForward needs a torchvision ResNet made of BasicBlocks (ResNet18 / ResNet34)")
        m = fold_batchnorm(model).to(dtype)
        dev = next(m.parameters()).device
        cl = torch.channels_last

        def wb(conv):
            w = conv.weight.detach().to(dtype).contiguous(memory_format=cl)
            b = conv.bias.detach().to(dtype) if conv.bias is not None else torch.zeros(w.shape[0], dtype=dtype, device=w.device)
            return w, b

        if tuple(m.conv1.kernel_size) != (7, 7) or tuple(m.conv1.stride) != (2, 2) or tuple(m.conv1.padding) != (3, 3) or m.conv1.in_channels != 3:
            raise TypeError("FusedResNetForward: unexpected stem convolution")
        w7, self.stem_b = wb(m.conv1)
        # out[oy] = sum_ky w[ky] in[2 oy + ky - 3]; with in[2 (oy + a) + p] =: S[oy + a][p] and ky = 2 a + p + 3: taps a in {-2..1}
        # (a = -2, p = 0 would be ky = -1: zero). Channel order inside a space-to-depth pixel: (p, q, c) -> p * 8 + q * 3 + c, two pad
        # channels behind each input row's 6 values, so that a 16-byte half pixel holds bytes of ONE input row (dh_gather_normalize S2D mode).
        w4 = torch.zeros((w7.shape[0], 16, 4, 4), dtype=torch.float32, device=dev)
        w7f = m.conv1.weight.detach().float()
        for a in range(-2, 2):
            for p_ in range(2):
                ky = 2 * a + p_ + 3
                if not 0 <= ky < 7:
                    continue
                for b_ in range(-2, 2):
                    for q in range(2):
                        kx = 2 * b_ + q + 3
                        if 0 <= kx < 7:
                            w4[:, p_ * 8 + q * 3 : p_ * 8 + q * 3 + 3, a + 2, b_ + 2] = w7f[:, :, ky, kx]
        self.stem_w = w4.to(dtype).contiguous(memory_format=cl)
        if stem == "s2d4":
            # output row oy = 2Y + P reads input rows 2 oy + ky - 3 = 4 (Y + A) + p  =>  ky = 4A + p - 2P + 3, block taps A in {-1, 0, 1}:
            # a 3x3 convolution over the block image with padding 1 (a zero block = the original zero padding of 3 pixels and beyond).
            # Input channel p*12 + q*3 + c (DH_S2D48), output channel (P*2 + Q)*64 + o (depth-to-space, dh_maxpool3x3s2_d2s).
            co = w7f.shape[0]
            w48 = torch.zeros((4 * co, 48, 3, 3), dtype=torch.float32, device=dev)
            for P_ in range(2):
                for A in range(-1, 2):
                    for p_ in range(4):
                        ky = 4 * A + p_ - 2 * P_ + 3
                        if not 0 <= ky < 7:
                            continue
                        for Q_ in range(2):
                            for B_ in range(-1, 2):
                                for q in range(4):
                                    kx = 4 * B_ + q - 2 * Q_ + 3
                                    if 0 <= kx < 7:
                                        oc = (P_ * 2 + Q_) * co
                                        w48[oc : oc + co, p_ * 12 + q * 3 : p_ * 12 + q * 3 + 3, A + 1, B_ + 1] = w7f[:, :, ky, kx]
            self.stem_w = w48.to(dtype).contiguous(memory_format=cl)
            self.stem_b = self.stem_b.repeat(4)
        self.blocks = []
        for layer in (m.layer1, m.layer2, m.layer3, m.layer4):
            for blk in layer:
                w1, b1 = wb(blk.conv1)
                w2, b2 = wb(blk.conv2)
                down = None
                if blk.downsample is not None:
                    dconv = blk.downsample[0]
                    down = (*wb(dconv), tuple(dconv.stride))
                self.blocks.append((w1, b1, tuple(blk.conv1.stride), w2, b2, down))
        self.fc_w, self.fc_b = m.fc.weight.detach().to(dtype), m.fc.bias.detach().to(dtype)
        self.dtype = dtype

    def s2d_shape(self, batch: int, ps: int) -> tuple:
        """Logical NCHW shape of the stem input for a batch of ps x ps patches, stored channels_last. s2d4: [B, 48, ps/4, ps/4];
        s2d2: [B, 16, ps/2 + 3, ps/2 + 3] with a zero border of 2 (top / left) and 1 (bottom / right) space-to-depth pixels."""
        return (batch, 48, ps // 4, ps // 4) if self.stem == "s2d4" else (batch, 16, ps // 2 + 3, ps // 2 + 3)

    @property
    def gather_layout(self) -> str:
        """The dh_gather_normalize layout that writes this stem's input."""
        return "S2D48" if self.stem == "s2d4" else "S2D16"

    def space_to_depth(self, x_nhwc: torch.Tensor) -> torch.Tensor:
        """[B, H, W, 3] (values already normalised) -> the stem input (torch ops; the predictor lets dh_gather_normalize write this
        layout directly)."""
        B, H, W, _ = x_nhwc.shape
        if self.stem == "s2d4":
            v = x_nhwc.reshape(B, H // 4, 4, W // 4, 4, 3).permute(0, 1, 3, 2, 4, 5)        # [B, Y, X, p, q, c]
            return v.reshape(B, H // 4, W // 4, 48).contiguous().permute(0, 3, 1, 2)
        out = torch.zeros((B, H // 2 + 3, W // 2 + 3, 16), dtype=x_nhwc.dtype, device=x_nhwc.device)
        v = x_nhwc.reshape(B, H // 2, 2, W // 2, 2, 3).permute(0, 1, 3, 2, 4, 5)            # [B, y', x', p, q, c]
        inner = out[:, 2 : 2 + H // 2, 2 : 2 + W // 2]
        inner[..., 0:6] = v[:, :, :, 0].reshape(B, H // 2, W // 2, 6)
        inner[..., 8:14] = v[:, :, :, 1].reshape(B, H // 2, W // 2, 6)
        return out.permute(0, 3, 1, 2)

    @torch.no_grad()
    def __call__(self, s2d: torch.Tensor) -> torch.Tensor:
        """s2d: stem input as produced by space_to_depth / the gather's S2D mode -> float32 logits [B, n]."""
        one = (1, 1)
        if self.stem == "s2d4":
            x = torch.cudnn_convolution_relu(s2d, self.stem_w, self.stem_b, one, one, one, 1)           # [B, 4 * 64, ps/4, ps/4], depth-to-space
            x = ops.maxpool3x3s2_d2s(x)
        else:
            x = torch.cudnn_convolution_relu(s2d, self.stem_w, self.stem_b, one, (0, 0), one, 1)        # [B, 64, ps/2, ps/2]
            x = ops.maxpool3x3s2_nhwc(x)
        for w1, b1, stride, w2, b2, down in self.blocks:
            identity = x if down is None else torch.nn.functional.conv2d(x, down[0], down[1], stride=down[2])
            h = torch.cudnn_convolution_relu(x, w1, b1, stride, one, one, 1)
            x = torch.cudnn_convolution_add_relu(h, w2, identity, 1.0, b2, one, one, one, 1)
        x = x.float().mean(dim=(2, 3))
        return torch.nn.functional.linear(x, self.fc_w.float(), self.fc_b.float())


def fold_batchnorm(model: torch.nn.Module) -> torch.nn.Module:
    """A copy of an eval-mode torchvision ResNet with conv+bn pairs fused (conv1/bn1, every block's conv/bn and downsample)."""
    import copy

    from torch.nn.utils.fusion import fuse_conv_bn_eval

    m = copy.deepcopy(model).eval()

    def fuse_pairs(mod):
        names = [n for n, _ in mod.named_children()]
        for a, b in zip(names, names[1:]):
            ca, cb = getattr(mod, a), getattr(mod, b)
            if isinstance(ca, torch.nn.Conv2d) and isinstance(cb, torch.nn.BatchNorm2d):
                setattr(mod, a, fuse_conv_bn_eval(ca, cb))
                setattr(mod, b, torch.nn.Identity())
        for child in mod.children():
            fuse_pairs(child)

    fuse_pairs(m)
    return m


def batch_predictor(patches: list[Patch], model, device) -> np.ndarray:
    """Reference :66-78 (same signature and return type); the /255, NHWC->NCHW and float conversion run in the gather kernel."""
    per_model = _PREDICTORS.setdefault(model, {})        # keyed by the model OBJECT (weakly): a collected model takes its predictors along
    pred = per_model.get(str(device))
    if pred is None or pred.source_model() is not model:
        pred = per_model[str(device)] = DeviceBatchPredictor(model, device)
        pred._source_model = weakref.ref(model)
    return pred(patches)


_PREDICTORS: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()


class ImagePredictorPatched:
    def __init__(self, psim_path, patch_sampler, batch_predictor: Callable, anno: AnnoDescription, layer: int, downscale: int = 4,
                 *, device="cuda", cnn_batch: Optional[int] = None, progress: bool = False, stream_bands: bool = True,
                 stream_band_bytes: int = 1 << 30):
        self.patch_sampler = patch_sampler
        self.batch_predictor = batch_predictor
        self.anno = anno
        self.layer = layer
        self.downscale = downscale
        self._device = torch.device(device)
        self._cnn_batch = cnn_batch
        self._progress = progress
        self._stream_bands = stream_bands                      # lazy (non-resident) slides: upload row chunks behind the CNN
        self._stream_band_bytes = stream_band_bytes
        self._copy_stream = None
        if isinstance(patch_sampler, (FullImageDenseSampler, FullImageRndSampler)):
            self.h, self.w = patch_sampler.h, patch_sampler.w
        else:
            with open_slide(psim_path) as psim:
                self.h, self.w = psim.layer_size(self.layer)
        self.last_sum_map: Optional[torch.Tensor] = None      # kept when process_device(want_sum=True)
        self.stage_events: Optional[dict] = None              # set to {} to collect CUDA events per stage (bench breakdown)
        self.nvtx = False                                     # True: NVTX ranges "deephisto/<stage>" around the stages (Nsight timelines)

    def _mark(self, stage: str):
        """(start, end) CUDA events appended to stage_events[stage]; a no-op context when profiling is off."""
        import contextlib

        if self.stage_events is None:
            return torch.cuda.nvtx.range(f"deephisto/{stage}") if self.nvtx else contextlib.nullcontext()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.stage_events.setdefault(stage, []).append((a, b))

        @contextlib.contextmanager
        def ctx():
            a.record()
            yield
            b.record()

        return ctx()

    def _mark_on(self, stage: str, stream):
        """_mark for work enqueued on another stream (the events are recorded there)."""
        import contextlib

        if self.stage_events is None:
            return contextlib.nullcontext()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.stage_events.setdefault(stage, []).append((a, b))

        @contextlib.contextmanager
        def ctx():
            a.record(stream)
            yield
            b.record(stream)

        return ctx()

    def stage_ms(self) -> dict:
        """Total milliseconds per stage from the collected events (synchronises)."""
        torch.cuda.synchronize()
        return {k: sum(a.elapsed_time(b) for a, b in v) for k, v in (self.stage_events or {}).items()}

    # ---- reference entry point --------------------------------------------------------------------------------------
    def process(self, rank: Optional[int] = None, world: Optional[int] = None) -> np.ndarray:
        """int64 [h//d, w//d] class map, like the reference's np.argmax(prediction, axis=2) (:62)."""
        return self.process_device(rank=rank, world=world)["argmax"].cpu().numpy().astype(np.int64)

    def process_device(self, want_sum: bool = False, want_count: bool = False, rank: Optional[int] = None,
                       world: Optional[int] = None) -> dict:
        fast = isinstance(self.batch_predictor, DeviceBatchPredictor)
        if fast and isinstance(self.patch_sampler, FullImageDenseSampler):
            if world is not None and world > 1:
                return self._dense_banded(rank, world, want_sum)
            return self._dense_device(want_sum, want_count)
        if fast and isinstance(self.patch_sampler, FullImageRndSampler):
            if world is not None and world > 1:
                return self._rnd_banded(rank, world, want_sum, want_count)
            return self._scatter_device(self._rnd_batches(), want_sum, want_count)
        if world is not None and world > 1:
            raise ValueError("row-band sharding needs a FullImageDenseSampler / FullImageRndSampler object and a DeviceBatchPredictor")
        return self._scatter_device(self._host_batches(), want_sum, want_count)

    # ---- dense, device resident, bit-exact sums -----------------------------------------------------------------------
    def _logits_for(self, sampler: FullImageDenseSampler, slide, logits: torch.Tensor, first: int, count: int, y_off: int = 0):
        """Fill logits[first:first+count] for entries [first, first+count) of the padded dense enumeration."""
        self._logits_for_ranges(sampler, slide, logits, [(first, count)], y_off)

    def _logits_for_ranges(self, sampler: FullImageDenseSampler, slide, logits: torch.Tensor, ranges, y_off: int = 0):
        """Fill the rows of `logits` listed by `ranges` = [(first, count)] (entries of the padded dense enumeration). The ranges are
        concatenated into ONE coordinate list and cut into CNN batches, so a band pays for one short tail batch, not one per range
        (a band plan has three ranges: main-grid rows, their last-column patches, the last row / corner / padding copies)."""
        pred: DeviceBatchPredictor = self.batch_predictor
        step = self._cnn_batch or max(sampler.batch_size, 512)
        ps = sampler.patch_size
        ranges = [(int(f), int(c)) for f, c in ranges if c > 0]
        total = sum(c for _, c in ranges)
        if total == 0:
            return
        with self._mark("coords+gather"):
            parts = [ops.dense_coords(sampler.h, sampler.w, ps, sampler.stride, sampler.batch_size, first=f, count=c, device=self._device) for f, c in ranges]
            coords = parts[0] if len(parts) == 1 else torch.cat(parts)
            if y_off:
                coords[:, 0] -= y_off
        single = len(ranges) == 1
        out = logits[ranges[0][0] : ranges[0][0] + total] if single else torch.empty((total, logits.shape[1]), dtype=logits.dtype, device=self._device)
        for a in range(0, total, step):
            c = min(step, total - a)
            with self._mark("coords+gather"):
                feats = pred.gather(slide, coords[a : a + c], ps)
            with self._mark("cnn"):
                out[a : a + c] = pred.logits(feats)
        if not single:
            a = 0
            for f, c in ranges:
                logits[f : f + c] = out[a : a + c]
                a += c

    def _logits_streamed(self, sampler: FullImageDenseSampler, logits: torch.Tensor, patch_ranges, max_band_bytes: Optional[int] = None):
        """Logits of `patch_ranges` (as produced by bands.plan_band) for a slide that is NOT resident in HBM: the slide rows are
        uploaded in row chunks on a copy stream, one chunk ahead of the chunk the CNN is working on, so the host->device traffic
        hides behind the convolutions and at most three chunks (<= max_band_bytes each, default 1 GiB) live in HBM -- slides
        larger than HBM (or than host RAM, through a memory-mapped .npy) are predicted the same way. Consecutive chunks re-upload
        the ps - stride rows they share."""
        from ..slide import band_to_device

        g = bands.dense_grid(sampler.h, sampler.w, sampler.patch_size, sampler.stride, sampler.batch_size)
        budget = int(max_band_bytes or self._stream_band_bytes)
        jobs = bands.stream_jobs(g, patch_ranges, ops.DeviceSlide.pitch_for(sampler.w), budget, first_budget_bytes=max(budget // 8, 1))
        cur = torch.cuda.current_stream(self._device)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self._device)
        copy = self._copy_stream

        def upload(job):
            with torch.cuda.stream(copy), self._mark_on("upload", copy):
                with sampler._src as psim:
                    band = band_to_device(psim, sampler.layer, job[0], job[1], self._device)
                ready = torch.cuda.Event()
                ready.record(copy)
            return band, ready

        # CNN batches are filled ACROSS chunks (fused bf16 predictor: its stem-input buffer is the batch): the patches a chunk leaves
        # over wait in the buffer for the next chunk's first patches, so a slide pays for one short batch, not one per chunk
        # (1 GiB chunks of a 100k-wide slide hold ~26.7 batches of 1024: 1.9 % of the CNN time went into padding, profiles/r02_predict.md)
        pred: DeviceBatchPredictor = self.batch_predictor
        step = self._cnn_batch or max(sampler.batch_size, 512)
        ps = sampler.patch_size
        buf = pred.batch_buffer(step, ps) if isinstance(pred, DeviceBatchPredictor) else None
        fill = 0
        dst = torch.empty(step, dtype=torch.int64, device=self._device) if buf is not None else None

        def flush():
            nonlocal fill
            if fill:
                with self._mark("cnn"):
                    out = pred.buffer_logits()
                    logits.index_copy_(0, dst[:fill], out[:fill].to(logits.dtype))
                fill = 0

        copy.wait_stream(cur)
        nxt = upload(jobs[0]) if jobs else None
        for i, job in enumerate(jobs):
            band, ready = nxt
            nxt = upload(jobs[i + 1]) if i + 1 < len(jobs) else None
            cur.wait_event(ready)
            band.storage.record_stream(cur)
            if buf is None:
                self._logits_for_ranges(sampler, band, logits, job[2], y_off=job[0])
            else:
                with self._mark("coords+gather"):
                    ranges = [(int(f), int(c)) for f, c in job[2] if c > 0]
                    parts = [ops.dense_coords(sampler.h, sampler.w, ps, sampler.stride, sampler.batch_size, first=f, count=c, device=self._device)
                             for f, c in ranges]
                    if not parts:
                        continue
                    coords = parts[0] if len(parts) == 1 else torch.cat(parts)
                    if job[0]:
                        coords[:, 0] -= job[0]
                    where = torch.cat([torch.arange(f, f + c, device=self._device) for f, c in ranges])
                a, m = 0, int(coords.shape[0])
                while a < m:
                    c = min(step - fill, m - a)
                    with self._mark("coords+gather"):
                        pred.gather_into(band, coords[a : a + c], ps, fill)
                        dst[fill : fill + c] = where[a : a + c]
                    fill += c
                    a += c
                    if fill == step:
                        flush()
            del band
        if buf is not None:
            flush()

    def _dense_device(self, want_sum: bool, want_count: bool) -> dict:
        s: FullImageDenseSampler = self.patch_sampler
        n = len(self.anno.anno_classes)
        logits = torch.empty((s.n_padded, n), dtype=torch.float32, device=self._device)
        if s._slide_dev is None:                                               # lazy_slide=True: stream the slide through HBM
            g = bands.dense_grid(s.h, s.w, s.patch_size, s.stride, s.batch_size)
            self._logits_streamed(s, logits, [(0, g.main_n), (g.main_n, g.ny), (g.main_n + g.ny, s.n_padded - g.main_n - g.ny)])
        else:
            self._logits_for(s, s._slide, logits, 0, s.n_padded)
        sum_map, cnt, amax = ops.stitch_dense(logits, s.h, s.w, s.patch_size, s.stride, self.downscale, s.batch_size,
                                              want_sum=want_sum, want_count=want_count, want_argmax=True)
        self.last_sum_map = sum_map
        return {"argmax": amax, "sum": sum_map, "count": cnt, "logits": logits}

    def dense_band_local(self, rank: int, world: int, want_sum: bool = False) -> dict:
        """The work of ONE rank of a row-band sharded prediction, without the exchange step: logits of the band's patches
        (halo patch rows recomputed), then the band of the stitched map, padded to `plan.rows_max` rows."""
        s: FullImageDenseSampler = self.patch_sampler
        n = len(self.anno.anno_classes)
        d = self.downscale
        plan = bands.plan_band(s.h, s.w, s.patch_size, s.stride, d, s.batch_size, rank, world)
        logits = torch.zeros((s.n_padded, n), dtype=torch.float32, device=self._device)
        if plan.patch_ranges and s._slide_dev is None and self._stream_bands:
            self._logits_streamed(s, logits, plan.patch_ranges)               # band rows streamed through HBM, upload hidden behind the CNN
        elif plan.patch_ranges:
            slide, y_off = s.band_slide(plan.slide_y0, plan.slide_y1)
            self._logits_for_ranges(s, slide, logits, plan.patch_ranges, y_off)
        dw = s.w // d
        amax_band = torch.zeros((plan.rows_max, dw), dtype=torch.uint8, device=self._device)
        sum_band = torch.zeros((plan.rows_max, dw, n), dtype=torch.float32, device=self._device) if want_sum else None
        if plan.row_end > plan.row_begin:
            with self._mark("stitch"):
                sm, _, am = ops.stitch_dense(logits, s.h, s.w, s.patch_size, s.stride, d, s.batch_size, row_begin=plan.row_begin,
                                             row_end=plan.row_end, want_sum=want_sum, want_argmax=True)
                amax_band[: plan.row_end - plan.row_begin] = am
                if want_sum:
                    sum_band[: plan.row_end - plan.row_begin] = sm
        return {"argmax_band": amax_band, "sum_band": sum_band, "logits": logits, "plan": plan}

    def _dense_banded(self, rank: int, world: int, want_sum: bool) -> dict:
        import torch.distributed as dist

        s: FullImageDenseSampler = self.patch_sampler
        dh = s.h // self.downscale
        loc = self.dense_band_local(rank, world, want_sum)
        # the one exchange step: band maps are disjoint row ranges -> all-gather, then drop the padding rows
        with self._mark("assemble (NCCL all-gather)"):
            out = {"argmax": assemble_bands(loc["argmax_band"], dh, world, dist), "sum": None, "count": None, "logits": loc["logits"],
                   "plan": loc["plan"]}
            if want_sum:
                out["sum"] = assemble_bands(loc["sum_band"], dh, world, dist)
        self.last_sum_map = out["sum"]
        return out

    # ---- arbitrary coordinates: scatter-accumulate ------------------------------------------------------------------------
    def _rnd_batches(self, sampler: Optional[FullImageRndSampler] = None):
        """Coverage-driven random sampling (the reference's default, :156-163). The sampler's batches depend on each other through
        the coverage accumulator, the CNN does not feed back into them: coordinates of several sampler batches are collected
        and sent through gather + CNN together (the reference's batch of 64 leaves the tensor cores mostly idle)."""
        s: FullImageRndSampler = self.patch_sampler if sampler is None else sampler
        pred: DeviceBatchPredictor = self.batch_predictor
        target = self._cnn_batch or max(s.batch_size, 1024)
        pending, n_pending, progress = [], 0, 0.0

        def flush():
            coords = pending[0] if len(pending) == 1 else torch.cat(pending)
            with self._mark("coords+gather"):
                feats = pred.gather(s._slide, coords, s.patch_size)
            with self._mark("cnn"):
                lg = pred.logits(feats)
            return lg, coords, s.patch_size, progress

        for coords, progress in s.coords_generator():
            pending.append(coords)
            n_pending += len(coords)
            if n_pending >= target:
                yield flush()
                pending, n_pending = [], 0
        if pending:
            yield flush()

    def _rnd_banded(self, rank: int, world: int, want_sum: bool, want_count: bool = False) -> dict:
        """Row-band sharded prediction with the coverage-driven random sampler (the reference's default sampler, :156-163). Rank r
        owns the map rows of bands.rnd_band and runs ITS OWN coverage sampler (Philox substream r) over the slide rows of that band
        only, so the CNN work splits evenly and no rank holds more than its band of the slide. Exchange steps (NCCL): (1) the ranks
        all-gather their (coords, logits) lists -- 28 bytes per patch -- so that every rank stitches ALL patches that touch its rows,
        in rank-major list order, with dh_stitch_binned: the assembled map is bit-identical to a single-GPU stitch of the concatenated
        list; (2) the all-gather of the band maps, as for the dense sampler."""
        import torch.distributed as dist

        s: FullImageRndSampler = self.patch_sampler
        d, n, ps = self.downscale, len(self.anno.anno_classes), s.patch_size
        dh, dw = s.h // d, s.w // d
        plan = bands.rnd_band(s.h, ps, d, s._downscale, rank, world)
        lgs, cos = [], []
        if plan.row_end > plan.row_begin:
            sub = s.band_sampler(plan.slide_y0, plan.slide_y1, rank)
            for lg, coords, _, _ in self._rnd_batches(sub):
                lgs.append(lg.reshape(-1, n))
                c = coords.reshape(-1, 2).clone()
                c[:, 0] += plan.slide_y0                                           # band-relative -> layer coordinates
                cos.append(c)
        lg = torch.cat(lgs) if lgs else torch.zeros((0, n), dtype=torch.float32, device=self._device)
        co = torch.cat(cos) if cos else torch.zeros((0, 2), dtype=torch.int32, device=self._device)
        with self._mark("assemble (NCCL all-gather)"):
            lg_all, co_all, counts = gather_patch_lists(lg, co, world, dist)
        rows = plan.row_end - plan.row_begin
        amax_band = torch.zeros((plan.rows_max, dw), dtype=torch.uint8, device=self._device)
        sum_band = torch.zeros((plan.rows_max, dw, n), dtype=torch.float32, device=self._device) if want_sum else None
        cnt_band = torch.zeros((plan.rows_max, dw), dtype=torch.int32, device=self._device) if want_count else None
        if rows > 0:
            with self._mark("stitch"):
                sm, cn, am = ops.stitch_binned(lg_all, co_all, ps, d, rows, dw, row_offset=plan.row_begin, want_sum=want_sum or n > 8,
                                               want_count=want_count, want_argmax=True)
            amax_band[:rows] = am
            if want_sum:
                sum_band[:rows] = sm
            if want_count:
                cnt_band[:rows] = cn
        with self._mark("assemble (NCCL all-gather)"):
            out = {"argmax": assemble_bands(amax_band, dh, world, dist), "sum": None, "count": None, "logits": lg_all, "coords": co_all,
                   "plan": plan, "patches_per_rank": counts}
            if want_sum:
                out["sum"] = assemble_bands(sum_band, dh, world, dist)
            if want_count:
                out["count"] = assemble_bands(cnt_band, dh, world, dist)
        self.last_sum_map = out["sum"]
        return out

    def _host_batches(self):
        """The reference's loop (:47-54): any iterator of (list[Patch], progress) and any callable list[Patch] -> [B, n]."""
        for patches, progress in self.patch_sampler:
            preds = self.batch_predictor(patches)
            lg = torch.as_tensor(np.asarray(preds) if not isinstance(preds, torch.Tensor) else preds, dtype=torch.float32).to(self._device)
            coords = torch.tensor([[p.pos_y, p.pos_x] for p in patches], dtype=torch.int32, device=self._device)
            yield lg, coords, patches[0].patch_size, progress

    def _scatter_device(self, batches, want_sum: bool, want_count: bool) -> dict:
        """The reference's accumulation loop (:47-54) for arbitrary coordinates. The logits and coordinates of every batch stay
        on the device (28 bytes per patch) and ONE dh_stitch_binned call at the end adds them per map cell in sampler order:
        bit-identical to `prediction[...] += logits_i` patch by patch, no atomics, every map byte written once."""
        d = self.downscale
        dh, dw = self.h // d, self.w // d
        n = len(self.anno.anno_classes)
        bar = None
        if self._progress:
            from tqdm import tqdm

            bar = tqdm(total=100, desc="Predicting", unit="step")
        all_lg, all_coords, ps_seen = [], [], None
        for lg, coords, ps, progress in batches:
            if ps_seen is not None and ps != ps_seen:
                raise ValueError(f"patch size changed from {ps_seen} to {ps} between batches")
            ps_seen = ps
            all_lg.append(lg.reshape(-1, n))
            all_coords.append(coords.reshape(-1, 2))
            if bar is not None:
                bar.n = round(progress * 100, 2)
                bar.refresh()
        if all_lg:
            lg = all_lg[0].contiguous() if len(all_lg) == 1 else torch.cat(all_lg)
            coords = all_coords[0].contiguous() if len(all_coords) == 1 else torch.cat(all_coords)
        else:
            lg = torch.zeros((0, n), dtype=torch.float32, device=self._device)
            coords = torch.zeros((0, 2), dtype=torch.int32, device=self._device)
        with self._mark("stitch"):
            sum_map, cnt, amax = ops.stitch_binned(lg, coords, ps_seen or 1, d, dh, dw, want_sum=want_sum or n > 8, want_count=want_count,
                                                   want_argmax=True)
        self.last_sum_map = sum_map
        return {"argmax": amax, "sum": sum_map if want_sum else None, "count": cnt, "logits": lg, "coords": coords}


def gather_patch_lists(logits: torch.Tensor, coords: torch.Tensor, world: int, dist):
    """All-gather per-rank patch lists of different lengths: (logits [sum P_r, n], coords [sum P_r, 2], [P_0 .. P_{world-1}]),
    concatenated in rank order. Lists are padded to the longest one for the collective (NCCL or gloo)."""
    n = logits.shape[1]
    cnt = torch.tensor([logits.shape[0]], dtype=torch.int64, device=logits.device)
    cnts = torch.empty(world, dtype=torch.int64, device=logits.device)
    dist.all_gather_into_tensor(cnts, cnt)
    counts = [int(v) for v in cnts.tolist()]
    pmax = max(max(counts), 1)
    # one int32 buffer per rank: [pmax][n] logits bit patterns, then [pmax][2] coordinates
    mine = torch.zeros(pmax * (n + 2), dtype=torch.int32, device=logits.device)
    mine[: logits.shape[0] * n] = logits.contiguous().view(torch.int32).reshape(-1)
    mine[pmax * n : pmax * n + coords.shape[0] * 2] = coords.contiguous().reshape(-1)
    everything = torch.empty(world * pmax * (n + 2), dtype=torch.int32, device=logits.device)
    dist.all_gather_into_tensor(everything, mine)
    everything = everything.view(world, pmax * (n + 2))
    lg = torch.cat([everything[r, : counts[r] * n].view(torch.float32).reshape(counts[r], n) for r in range(world)])
    co = torch.cat([everything[r, pmax * n : pmax * n + counts[r] * 2].reshape(counts[r], 2) for r in range(world)])
    return lg.contiguous(), co.contiguous(), counts


def assemble_bands(band: torch.Tensor, dh: int, world: int, dist) -> torch.Tensor:
    """All-gather equal-height (padded) row bands and keep rows [r*dh//G, (r+1)*dh//G) of each: [dh, ...]. Works with
    NCCL (CUDA tensors) and gloo (CPU tensors; used by the world_size-2 CPU tests)."""
    rows_max = band.shape[0]
    gathered = torch.empty((world * rows_max,) + tuple(band.shape[1:]), dtype=band.dtype, device=band.device)
    dist.all_gather_into_tensor(gathered, band.contiguous())
    heights = [bands.band_rows(dh, r, world)[1] - bands.band_rows(dh, r, world)[0] for r in range(world)]
    if all(hh == rows_max for hh in heights):
        return gathered
    return torch.cat([gathered[r * rows_max : r * rows_max + heights[r]] for r in range(world)], 0)


def visualize_device(pred_u8: torch.Tensor, slide, anno_dsc: AnnoDescription, d: int, alpha: float = 0.6):
    """(mask, thumbnail, overlay) uint8 [h,w,3] CUDA tensors for a class map u8 [h,w] and the full-resolution DeviceSlide it was
    predicted from at total downscale d: one launch of dh_colorize_overlay (reference :89-110)."""
    lut = torch.zeros((256, 3), dtype=torch.uint8)
    for a in anno_dsc.anno_classes:
        lut[a.id] = torch.tensor(a.color, dtype=torch.uint8)
    return ops.colorize_overlay(pred_u8.contiguous(), lut.to(pred_u8.device), slide, d, alpha, want_mask=True, want_thumb=True, want_overlay=True)


def perform_and_save_visualizations(img_path, anno_dsc: AnnoDescription, pred: np.ndarray, out_dir: Path = Path("."), *, device="cuda"):
    """Reference :81-113: colourised mask, downscaled slide, 0.6/0.4 overlay, saved as JPEG. The three images are computed on
    the device from the full-resolution layer in one pass; only the JPEG encoding (PIL) is host work."""
    from PIL import Image

    from ..slide import layer_to_device

    out_dir.mkdir(exist_ok=True, parents=True)
    stem = Path(img_path).stem if isinstance(img_path, (str, Path)) else "slide"
    h, w = pred.shape[:2]
    with open_slide(img_path) as psim:
        full = layer_to_device(psim, 1, device)
    d = min(full.H // h, full.W // w)
    if d < 1:
        raise ValueError("the class map is larger than the slide")
    pred_u8 = torch.as_tensor(np.asarray(pred)).to(torch.uint8).to(device)
    mask, thumb, over = visualize_device(pred_u8, full, anno_dsc, d)
    Image.fromarray(mask.cpu().numpy()).save(out_dir / f"{stem}_mask.jpg", quality=95)
    Image.fromarray(thumb.cpu().numpy()).save(out_dir / f"{stem}.jpg", quality=95)
    Image.fromarray(over.cpu().numpy()).save(out_dir / f"{stem}_overlay.jpg", quality=95)


def main(argv=None):
    """python -m deephisto_b200.examples.predict_full_patched --synthetic H W [--weights best_model.pth] (reference :129-183)."""
    import argparse
    import os
    import time

    ap = argparse.ArgumentParser()
    ap.add_argument("--image", default=None, help=".npy slide (uint8 [H,W,3]) or .psi when psimage is installed")
    ap.add_argument("--synthetic", type=int, nargs=2, metavar=("H", "W"), default=None)
    ap.add_argument("--weights", default=None)
    ap.add_argument("--layer", type=int, default=1)
    ap.add_argument("--stride", type=int, default=112)
    ap.add_argument("--downscale", type=int, default=16)
    ap.add_argument("--random-sampler", action="store_true", help="the reference's default (coverage-driven random sampling)")
    ap.add_argument("--bf16", action="store_true")
    ap.add_argument("--fused", action="store_true", help="bf16 fast path: FusedResNetForward (space-to-depth stem, fused cuDNN epilogues); implies --bf16")
    ap.add_argument("--out", default="./output/")
    args = ap.parse_args(argv)

    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=device)
    anno_dsc = AnnoDescription.with_known_colors({"AT": (245, 119, 34), "BG": (153, 255, 255), "LP": (64, 170, 72), "MM": (255, 0, 0),
                                                  "TUM": (33, 67, 156)})
    if args.weights:
        model = load_model(args.weights, device)
    else:
        torch.manual_seed(0)
        model = get_model(5).to(device).eval()
    from ..slide import SyntheticSlide

    src = SyntheticSlide(*args.synthetic) if args.synthetic else args.image
    if src is None:
        ap.error("give --image or --synthetic H W")
    mode = SamplerExecutionMode.INMEMORY_SINGLEPROC
    if args.random_sampler:
        sampler = FullImageRndSampler(src, layer=args.layer, patch_size=224, batch_size=64, mode=mode, device=device, lazy_slide=world > 1)
    else:
        sampler = FullImageDenseSampler(src, layer=args.layer, patch_size=224, batch_size=64, mode=mode, stride=args.stride, device=device,
                                        lazy_slide=world > 1)
    bf16 = args.bf16 or args.fused
    predictor = ImagePredictorPatched(src, patch_sampler=sampler,
                                      batch_predictor=DeviceBatchPredictor(model, device, torch.bfloat16 if bf16 else torch.float32, fused=args.fused),
                                      anno=anno_dsc, layer=args.layer, downscale=args.downscale, device=device)
    t0 = time.perf_counter()
    pred = predictor.process(rank=rank, world=world) if world > 1 else predictor.process()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rank == 0:
        print(f"{sampler.h} x {sampler.w} slide -> {pred.shape} class map in {dt:.3f} s ({sampler.h * sampler.w / dt / 1e9:.3f} Gpx/s, {world} GPU)")
        if not args.synthetic:
            perform_and_save_visualizations(src, anno_dsc, pred, out_dir=Path(args.out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
