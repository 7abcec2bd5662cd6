"""Torch-tensor level wrappers over the C-ABI (device memory and streams come from torch; all
arithmetic happens in the hand-written sm_100a kernels of libdeephisto_b200.so)."""

from __future__ import annotations

import ctypes as C
import warnings
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import DH_BF16, DH_F32, DH_NCHW, DH_NHWC, DH_S2D16, DH_S2D48, DH_U8, check

_DTYPES = {torch.float32: DH_F32, torch.bfloat16: DH_BF16, torch.uint8: DH_U8}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _need_cuda(t: torch.Tensor, name: str, dtype=None) -> None:
    if not t.is_cuda:
        raise _lib.DeepHistoError(f"{name} must be a CUDA tensor (there is no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")


class DeviceSlide:
    """uint8 RGB slide layer resident in HBM: rows of `pitch` bytes (pitch % 16 == 0), 3*W used."""

    def __init__(self, storage: torch.Tensor, H: int, W: int, pitch: int):
        _need_cuda(storage, "slide storage", torch.uint8)
        if storage.numel() < H * pitch:
            raise ValueError("slide storage smaller than H * pitch")
        self.storage, self.H, self.W, self.pitch = storage, int(H), int(W), int(pitch)

    @staticmethod
    def pitch_for(W: int) -> int:
        return (3 * W + 15) // 16 * 16

    @classmethod
    def empty(cls, H: int, W: int, device="cuda") -> "DeviceSlide":
        pitch = cls.pitch_for(W)
        return cls(torch.empty(H * pitch, dtype=torch.uint8, device=device), H, W, pitch)

    @classmethod
    def from_numpy(cls, arr: np.ndarray, device="cuda") -> "DeviceSlide":
        if arr.ndim != 3 or arr.shape[2] != 3 or arr.dtype != np.uint8:
            raise ValueError("slide must be uint8 [H, W, 3]")
        H, W, _ = arr.shape
        s = cls.empty(H, W, device)
        with warnings.catch_warnings():                     # read-only sources (np.load(mmap_mode="r")) are only read from
            warnings.filterwarnings("ignore", message="The given NumPy array is not writable")
            src = torch.from_numpy(np.ascontiguousarray(arr)).view(H, 3 * W)
        s.storage.view(H, s.pitch)[:, : 3 * W].copy_(src, non_blocking=False)
        return s

    @classmethod
    def synthetic(cls, H: int, W: int, seed: int = 0, device="cuda", y0: int = 0, rows: Optional[int] = None) -> "DeviceSlide":
        """Counter-hash slide generated on the device (oracle/synth.py restates the bytes). With y0 / rows only that row band
        of the H x W slide is generated; the returned slide has `rows` rows and row 0 is slide row y0."""
        lib = _lib.require_device()
        rows = H - y0 if rows is None else rows
        s = cls.empty(rows, W, device)
        with torch.cuda.device(s.storage.device):
            check(lib.dh_synth_slide_rows(s.storage.data_ptr(), H, W, s.pitch, y0, rows, seed, _stream()), "dh_synth_slide_rows")
        return s

    def to_numpy(self) -> np.ndarray:
        return self.rows2d()[:, : 3 * self.W].cpu().numpy().reshape(self.H, self.W, 3)

    def rows2d(self) -> torch.Tensor:
        """uint8 [H, pitch] view of the rows (the storage may be longer than H * pitch: padded shards of a collective upload)."""
        return self.storage[: self.H * self.pitch].view(self.H, self.pitch)

    @property
    def device(self):
        return self.storage.device

    @property
    def ptr(self) -> int:
        """Address the kernels read the slide from."""
        return self.storage.data_ptr()

    def view_rows(self, y0: int, y1: int) -> "DeviceSlide":
        """Rows [y0, y1) as a slide of their own (a view: row 0 of the result is row y0)."""
        return DeviceSlide(self.storage[y0 * self.pitch : y1 * self.pitch], y1 - y0, self.W, self.pitch)


class MappedHostSlide:
    """A slide layer that stays in PAGE-LOCKED HOST memory and is read by the gather kernels in place, through its device-visible
    address (dh_host_device_pointer): every bulk row copy of the gather then travels over PCIe, and only the rows of the patches
    actually drawn do. Same attributes as DeviceSlide where the gather needs them (ptr, H, W, pitch); `device` is the GPU that reads."""

    def __init__(self, host: torch.Tensor, H: int, W: int, pitch: int, device="cuda"):
        if host.is_cuda or host.dtype != torch.uint8 or not host.is_pinned() or host.numel() < H * pitch:
            raise ValueError("MappedHostSlide needs a pinned host uint8 tensor of at least H * pitch bytes")
        lib = _lib.require_device()
        out = C.c_uint64(0)
        with torch.cuda.device(device):
            check(lib.dh_host_device_pointer(host.data_ptr(), C.byref(out)), "dh_host_device_pointer")
        self.host, self.H, self.W, self.pitch = host, int(H), int(W), int(pitch)
        self._ptr, self._device = int(out.value), torch.device(device)

    @property
    def ptr(self) -> int:
        return self._ptr

    @property
    def device(self):
        return self._device

    def view_rows(self, y0: int, y1: int) -> "MappedHostSlide":
        return MappedHostSlide(self.host[y0 * self.pitch : y1 * self.pitch], y1 - y0, self.W, self.pitch, self._device)

    def rows2d(self) -> torch.Tensor:
        return self.host[: self.H * self.pitch].view(self.H, self.pitch)

    def to_numpy(self) -> np.ndarray:
        return self.rows2d()[:, : 3 * self.W].numpy().reshape(self.H, self.W, 3)


def dense_count(H: int, W: int, ps: int, stride: int, batch_size: int) -> tuple[int, int]:
    """(N, N padded to a multiple of batch_size) of full_samplers.py:374-404."""
    lib = _lib.load()
    npad = C.c_int64(0)
    n = lib.dh_dense_count(H, W, ps, stride, batch_size, C.byref(npad))
    if n < 0:
        raise ValueError(f"dh_dense_count: {_lib.last_error()}")
    return int(n), int(npad.value)


def dense_coords(H: int, W: int, ps: int, stride: int, batch_size: int, first: int = 0, count: Optional[int] = None,
                 device="cuda") -> torch.Tensor:
    """int32 [count, 2] (y, x) of the padded dense enumeration, generated on the device."""
    lib = _lib.require_device()
    n, npad = dense_count(H, W, ps, stride, batch_size)
    if count is None:
        count = npad - first
    out = torch.empty((count, 2), dtype=torch.int32, device=device)
    with torch.cuda.device(out.device):
        check(lib.dh_dense_coords(H, W, ps, stride, batch_size, first, count, out.data_ptr(), _stream()), "dh_dense_coords")
    return out


def gather_normalize(slide: "DeviceSlide | MappedHostSlide", coords: torch.Tensor, ps: int, *, dtype=torch.float32, layout: str = "NHWC",
                     scale255: bool = True, mean: Optional[Sequence[float]] = None, std: Optional[Sequence[float]] = None,
                     flip: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                     out_index: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Patches at int32 `coords` [B,2] (y,x) -> [B,ps,ps,3] (NHWC) or [B,3,ps,ps] (NCHW), or -- layout "S2D16", bfloat16 -- the
    2x2 space-to-depth image [B, ps/2+3, ps/2+3, 16] with a zero border (DH_S2D16 in include/deephisto_b200.h). The kernel writes
    only the interior: an `out` buffer passed by the caller must have a zero border (it stays zero across calls). Layout "S2D48",
    bfloat16: the 4x4 space-to-depth image [B, ps/4, ps/4, 48] (DH_S2D48; no border, no padding)."""
    lib = _lib.require_device()
    _need_cuda(coords, "coords", torch.int32)
    if coords.ndim != 2 or coords.shape[1] != 2:
        raise ValueError("coords must be [B, 2]")
    B = coords.shape[0]
    lay = {"NHWC": DH_NHWC, "NCHW": DH_NCHW, "S2D16": DH_S2D16, "S2D48": DH_S2D48}[layout]
    shape = {DH_NHWC: (B, ps, ps, 3), DH_NCHW: (B, 3, ps, ps), DH_S2D16: (B, ps // 2 + 3, ps // 2 + 3, 16), DH_S2D48: (B, ps // 4, ps // 4, 48)}[lay]
    if out is None:
        out = (torch.zeros if lay == DH_S2D16 else torch.empty)(shape, dtype=dtype, device=coords.device)
    else:
        _need_cuda(out, "out", dtype)
        if out_index is None and tuple(out.shape) != shape:
            raise ValueError(f"out must have shape {shape}")
    if flip is not None:
        _need_cuda(flip, "flip", torch.uint8)
    if out_index is not None:
        _need_cuda(out_index, "out_index", torch.int32)
    m = s = None
    if mean is not None or std is not None:
        m = (C.c_float * 3)(*(mean if mean is not None else (0.0, 0.0, 0.0)))
        s = (C.c_float * 3)(*(std if std is not None else (1.0, 1.0, 1.0)))
    with torch.cuda.device(coords.device):
        check(
            lib.dh_gather_normalize(slide.ptr, slide.H, slide.W, slide.pitch, coords.data_ptr(), _ptr(out_index), B,
                                    ps, out.data_ptr(), _DTYPES[dtype], lay, int(bool(scale255)), m, s, _ptr(flip), _stream()),
            "dh_gather_normalize",
        )
    return out


class SlideTable:
    """Descriptor table of several resident slides for gather_normalize_multi (one launch over a multi-image dataset)."""

    def __init__(self, slides: Sequence[DeviceSlide], device=None):
        self.slides = list(slides)
        rows = [[s.ptr, s.H, s.W, s.pitch] for s in self.slides]
        self.host = np.ascontiguousarray(np.asarray(rows, dtype=np.int64))
        self.dev = torch.from_numpy(self.host).to(self.slides[0].device if device is None else device)


def gather_normalize_multi(table: SlideTable, images: torch.Tensor, coords: torch.Tensor, ps: int, *, dtype=torch.float32, layout: str = "NHWC",
                           scale255: bool = True, mean: Optional[Sequence[float]] = None, std: Optional[Sequence[float]] = None,
                           flip: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Patch b is read from slide images[b] of the table at coords[b]: ONE launch for a batch that spans several slides."""
    lib = _lib.require_device()
    _need_cuda(coords, "coords", torch.int32)
    _need_cuda(images, "images", torch.int32)
    B = coords.shape[0]
    lay = {"NHWC": DH_NHWC, "NCHW": DH_NCHW}[layout]
    shape = (B, ps, ps, 3) if lay == DH_NHWC else (B, 3, ps, ps)
    if out is None:
        out = torch.empty(shape, dtype=dtype, device=coords.device)
    if flip is not None:
        _need_cuda(flip, "flip", torch.uint8)
    m = s = None
    if mean is not None or std is not None:
        m = (C.c_float * 3)(*(mean if mean is not None else (0.0, 0.0, 0.0)))
        s = (C.c_float * 3)(*(std if std is not None else (1.0, 1.0, 1.0)))
    with torch.cuda.device(coords.device):
        check(lib.dh_gather_normalize_multi(table.host.ctypes.data, table.dev.data_ptr(), len(table.slides), images.data_ptr(), coords.data_ptr(),
                                            None, B, ps, out.data_ptr(), _DTYPES[dtype], lay, int(bool(scale255)), m, s, _ptr(flip), _stream()),
              "dh_gather_normalize_multi")
    return out


def stitch_dense(logits: torch.Tensor, H: int, W: int, ps: int, stride: int, d: int, batch_size: int, *, row_begin: int = 0,
                 row_end: Optional[int] = None, want_sum: bool = True, want_count: bool = False, want_argmax: bool = False):
    """Deterministic stitch of dense-sampler logits [Npad, n] -> (sum [rows,dw,n] f32, count u32, argmax u8)."""
    lib = _lib.require_device()
    _need_cuda(logits, "logits", torch.float32)
    n = logits.shape[1]
    dh, dw = H // d, W // d
    if row_end is None:
        row_end = dh
    rows = row_end - row_begin
    N, npad = dense_count(H, W, ps, stride, batch_size)
    if logits.shape[0] < npad:
        raise ValueError(f"logits has {logits.shape[0]} rows, the padded enumeration needs {npad}")
    dev = logits.device
    sum_map = torch.empty((rows, dw, n), dtype=torch.float32, device=dev) if want_sum else None
    cnt = torch.empty((rows, dw), dtype=torch.int32, device=dev) if want_count else None
    amax = torch.empty((rows, dw), dtype=torch.uint8, device=dev) if want_argmax else None
    with torch.cuda.device(dev):
        check(
            lib.dh_stitch_dense_ex(logits.data_ptr(), H, W, ps, stride, d, n, batch_size, _ptr(sum_map), _ptr(cnt), _ptr(amax), row_begin,
                                   row_end, _stream()),
            "dh_stitch_dense_ex",
        )
    return sum_map, cnt, amax


def stitch_scatter(logits: torch.Tensor, coords: torch.Tensor, ps: int, d: int, sum_map: Optional[torch.Tensor],
                   count_map: Optional[torch.Tensor], row_offset: int = 0) -> None:
    """Accumulate logits [P,n] of patches at int32 coords [P,2] into sum_map [rows,dw,n] / count_map [rows,dw] (atomics)."""
    lib = _lib.require_device()
    _need_cuda(logits, "logits", torch.float32)
    _need_cuda(coords, "coords", torch.int32)
    ref = sum_map if sum_map is not None else count_map
    if sum_map is not None:
        _need_cuda(sum_map, "sum_map", torch.float32)
    if count_map is not None:
        _need_cuda(count_map, "count_map", torch.int32)
    rows, dw = ref.shape[0], ref.shape[1]
    with torch.cuda.device(logits.device):
        check(
            lib.dh_stitch_scatter(logits.data_ptr(), coords.data_ptr(), logits.shape[0], ps, d, logits.shape[1], _ptr(sum_map),
                                  _ptr(count_map), rows, dw, row_offset, _stream()),
            "dh_stitch_scatter",
        )


def stitch_binned(logits: torch.Tensor, coords: torch.Tensor, ps: int, d: int, rows: int, dw: int, *, row_offset: int = 0,
                  want_sum: bool = True, want_count: bool = False, want_argmax: bool = False):
    """Deterministic stitch of an arbitrary coordinate list: logits [P,n] f32 + int32 coords [P,2] -> (sum [rows,dw,n] f32,
    count i32 [rows,dw], argmax u8 [rows,dw]), each None unless requested. Bit-identical to the reference's loop over the
    patches in list order (predict_full_patched.py:47-54); the maps cover rows [row_offset, row_offset+rows)."""
    lib = _lib.require_device()
    _need_cuda(logits, "logits", torch.float32)
    _need_cuda(coords, "coords", torch.int32)
    if logits.dim() != 2 or coords.shape != (logits.shape[0], 2):
        raise ValueError(f"logits [P,n] and coords [P,2] expected, got {tuple(logits.shape)} and {tuple(coords.shape)}")
    P, n = logits.shape
    dev = logits.device
    sum_map = torch.empty((rows, dw, n), dtype=torch.float32, device=dev) if want_sum else None
    cnt = torch.empty((rows, dw), dtype=torch.int32, device=dev) if want_count else None
    amax = torch.empty((rows, dw), dtype=torch.uint8, device=dev) if want_argmax else None
    nbytes = int(lib.dh_stitch_binned_scratch_bytes(P, ps, d, n, rows, dw))
    scratch = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib.dh_stitch_binned(logits.data_ptr(), coords.data_ptr(), P, ps, d, n, _ptr(sum_map), _ptr(cnt), _ptr(amax), rows, dw,
                                   row_offset, scratch.data_ptr(), scratch.numel(), _stream()), "dh_stitch_binned")
    return sum_map, cnt, amax


def stitch_finalize(sum_map: torch.Tensor, count_map: Optional[torch.Tensor] = None, *, want_norm: bool = False,
                    want_argmax: bool = True):
    lib = _lib.require_device()
    _need_cuda(sum_map, "sum_map", torch.float32)
    n = sum_map.shape[-1]
    cells = sum_map.numel() // n
    norm = torch.empty_like(sum_map) if want_norm else None
    amax = torch.empty(sum_map.shape[:-1], dtype=torch.uint8, device=sum_map.device) if want_argmax else None
    with torch.cuda.device(sum_map.device):
        check(lib.dh_stitch_finalize(sum_map.data_ptr(), _ptr(count_map), cells, n, _ptr(norm), _ptr(amax), _stream()),
              "dh_stitch_finalize")
    return norm, amax


def maxpool3x3s2_nhwc(x: torch.Tensor) -> torch.Tensor:
    """max_pool2d(kernel 3, stride 2, padding 1) of a channels_last bf16 tensor [B,C,H,W] (NHWC in memory) -> channels_last [B,C,OH,OW]."""
    lib = _lib.require_device()
    if not x.is_cuda or x.dtype != torch.bfloat16 or x.dim() != 4 or not x.is_contiguous(memory_format=torch.channels_last):
        raise ValueError("maxpool3x3s2_nhwc needs a CUDA bfloat16 [B,C,H,W] tensor in channels_last memory format")
    B, Cc, H, W = x.shape
    out = torch.empty((B, Cc, (H - 1) // 2 + 1, (W - 1) // 2 + 1), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
    with torch.cuda.device(x.device):
        check(lib.dh_maxpool3x3s2_nhwc(x.data_ptr(), B, H, W, Cc, out.data_ptr(), DH_BF16, _stream()), "dh_maxpool3x3s2_nhwc")
    return out


def maxpool3x3s2_d2s(x: torch.Tensor) -> torch.Tensor:
    """max_pool2d(kernel 3, stride 2, padding 1) of a convolution output stored depth-to-space: x channels_last bf16 [B, 4*C, H, W] with
    channel (P*2 + Q)*C + o of block (Y, X) = pixel (2Y + P, 2X + Q), channel o -> channels_last [B, C, H, W]."""
    lib = _lib.require_device()
    if not x.is_cuda or x.dtype != torch.bfloat16 or x.dim() != 4 or not x.is_contiguous(memory_format=torch.channels_last) or x.shape[1] % 4:
        raise ValueError("maxpool3x3s2_d2s needs a CUDA bfloat16 [B,4C,H,W] tensor in channels_last memory format")
    B, C4, H, W = x.shape
    out = torch.empty((B, C4 // 4, H, W), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
    with torch.cuda.device(x.device):
        check(lib.dh_maxpool3x3s2_d2s(x.data_ptr(), B, H, W, C4 // 4, out.data_ptr(), DH_BF16, _stream()), "dh_maxpool3x3s2_d2s")
    return out


def colorize_overlay(argmax: torch.Tensor, lut_rgb: torch.Tensor, slide: Optional[DeviceSlide] = None, d: int = 1, alpha: float = 0.6,
                     want_mask: bool = True, want_thumb: bool = False, want_overlay: bool = False):
    """Class map u8 [dh,dw] -> (mask, thumbnail, overlay), each uint8 [dh,dw,3] or None (predict_full_patched.py:81-113)."""
    lib = _lib.require_device()
    _need_cuda(argmax, "argmax", torch.uint8)
    _need_cuda(lut_rgb, "lut_rgb", torch.uint8)
    if lut_rgb.numel() != 768:
        raise ValueError("lut_rgb must be uint8 [256, 3]")
    if (want_thumb or want_overlay) and slide is None:
        raise ValueError("the thumbnail and the overlay need the slide")
    dh, dw = argmax.shape
    mk = lambda want: torch.empty((dh, dw, 3), dtype=torch.uint8, device=argmax.device) if want else None  # noqa: E731
    mask, thumb, over = mk(want_mask), mk(want_thumb), mk(want_overlay)
    with torch.cuda.device(argmax.device):
        check(lib.dh_colorize_overlay(argmax.data_ptr(), None if slide is None else slide.storage.data_ptr(), 0 if slide is None else slide.H,
                                      0 if slide is None else slide.W, 0 if slide is None else slide.pitch, dh, dw, d, lut_rgb.data_ptr(),
                                      float(alpha), _ptr(mask), _ptr(thumb), _ptr(over), _stream()), "dh_colorize_overlay")
    return mask, thumb, over


class CoverState:
    """Device state of the coverage-driven random sampler (accumulator + scratch)."""

    def __init__(self, H: int, W: int, ps: int, speedup: int, dense_level: int, batch_size: int, seed: int, device="cuda"):
        lib = _lib.require_device()
        self.H, self.W, self.ps, self.speedup, self.dense_level, self.B, self.seed = H, W, ps, speedup, dense_level, batch_size, seed
        self.dh, self.dw = H // speedup, W // speedup
        self.accum = torch.zeros((self.dh, self.dw), dtype=torch.int32, device=device)
        self.scratch = torch.empty(int(lib.dh_cover_scratch_words(self.dh, self.dw)), dtype=torch.int32, device=device)
        self.nonzero = torch.zeros(1, dtype=torch.int32, device=device)
        self.batch_index = 0

    def restore(self, accum: torch.Tensor, batch_index: int) -> None:
        """Resume a run: adopt a saved accumulator and batch counter and rebuild the sampler state (dh_cover_init). The Philox
        stream is keyed by (seed, batch_index), so the continuation is identical to the uninterrupted run."""
        lib = _lib.require_device()
        self.accum.copy_(accum.to(self.accum.dtype))
        self.batch_index = int(batch_index)
        with torch.cuda.device(self.accum.device):
            check(lib.dh_cover_init(self.accum.data_ptr(), self.dh, self.dw, self.dense_level, self.scratch.data_ptr(), _stream()), "dh_cover_init")

    def next_coords(self, *, stop_when_full: bool = False, coords_out: Optional[torch.Tensor] = None,
                    nonzero_out: Optional[torch.Tensor] = None) -> tuple[torch.Tensor, torch.Tensor]:
        """One batch (one launch): int32 coords [B,2] and the device counter of non-zero accumulator cells. With stop_when_full a call
        after coverage is complete leaves the accumulator and `coords_out` untouched (batches can then be enqueued ahead of the
        read-back of the counter)."""
        lib = _lib.require_device()
        coords = torch.empty((self.B, 2), dtype=torch.int32, device=self.accum.device) if coords_out is None else coords_out
        nonzero = self.nonzero if nonzero_out is None else nonzero_out
        with torch.cuda.device(self.accum.device):
            check(
                lib.dh_cover_sample(self.accum.data_ptr(), self.dh, self.dw, self.H, self.W, self.ps, self.speedup, self.dense_level, self.B,
                                    self.seed, self.batch_index, coords.data_ptr(), nonzero.data_ptr(), self.scratch.data_ptr(),
                                    1 if stop_when_full else 0, _stream()),
                "dh_cover_sample",
            )
        self.batch_index += 1
        return coords, nonzero

    def next_group(self, n_batches: int, read_back: bool = False):
        """`n_batches` consecutive batches enqueued without a host synchronisation: coords int32 [n, B, 2] (zeros for batches after
        coverage completed) and their non-zero counters int32 [n] -- one read-back serves the whole group. With read_back the
        second value is (pinned host int32 [n], event): the asynchronous copy of the counters and the event that marks it done."""
        lib = _lib.require_device()
        dev = self.accum.device
        coords = torch.zeros((n_batches, self.B, 2), dtype=torch.int32, device=dev)
        counts = torch.zeros(n_batches, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            check(lib.dh_cover_sample_group(self.accum.data_ptr(), self.dh, self.dw, self.H, self.W, self.ps, self.speedup, self.dense_level,
                                            self.B, self.seed, self.batch_index, n_batches, coords.data_ptr(), counts.data_ptr(),
                                            self.scratch.data_ptr(), _stream()), "dh_cover_sample_group")
        self.batch_index += n_batches
        if not read_back:
            return coords, counts
        # the counters travel right behind the group's own launches: a group enqueued later does not delay their read-back
        host = torch.empty(n_batches, dtype=torch.int32, pin_memory=True)
        host.copy_(counts, non_blocking=True)
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(dev))
        return coords, (host, ready)


def region_accept_dense(edges: torch.Tensor, edge_begin: int, edge_end: int, y0: int, x0: int, ny: int, nx: int, stride: int,
                        ps: int, threshold: float, want_area: bool = False):
    """Acceptance mask u8 [ny*nx] (+ float64 clip areas) for the candidate grid of one region."""
    lib = _lib.require_device()
    _need_cuda(edges, "edges", torch.float64)
    mask = torch.empty(max(ny * nx, 0), dtype=torch.uint8, device=edges.device)
    area = torch.empty(max(ny * nx, 0), dtype=torch.float64, device=edges.device) if want_area else None
    if ny <= 0 or nx <= 0:
        return mask, area
    with torch.cuda.device(edges.device):
        check(lib.dh_region_accept_dense(edges.data_ptr(), edge_begin, edge_end, y0, x0, ny, nx, stride, ps, float(threshold),
                                         mask.data_ptr(), _ptr(area), _stream()), "dh_region_accept_dense")
    return mask, area


def compact_coords(mask: torch.Tensor, y0: int, x0: int, ny: int, nx: int, stride: int) -> torch.Tensor:
    """Accepted candidates in row-major order as int32 [n,2] (one small D2H sync for n)."""
    lib = _lib.require_device()
    _need_cuda(mask, "mask", torch.uint8)
    coords = torch.empty((max(ny * nx, 1), 2), dtype=torch.int32, device=mask.device)
    n_out = torch.zeros(1, dtype=torch.int32, device=mask.device)
    if ny * nx == 0:
        return coords[:0]
    with torch.cuda.device(mask.device):
        check(lib.dh_compact_coords(mask.data_ptr(), y0, x0, ny, nx, stride, coords.data_ptr(), n_out.data_ptr(), _stream()),
              "dh_compact_coords")
    return coords[: int(n_out.item())]


def region_sample(tables: "_lib.RegionTables", n_slots: int, k: int, ps: int, threshold: float, *, miss_limit: int = 500,
                  max_redraw: int = 64, fixed_class: int = -1, slots_per_table_draw: int = 1, seed: int = 0, slot_offset: int = 0,
                  device="cuda", out=None):
    """Random region sampling: (coords int32 [S,2], labels int64 [S], images int32 [S], status u8 [S])."""
    lib = _lib.require_device()
    if out is None:
        coords = torch.empty((n_slots, 2), dtype=torch.int32, device=device)
        labels = torch.empty(n_slots, dtype=torch.int64, device=device)
        images = torch.empty(n_slots, dtype=torch.int32, device=device)
        status = torch.empty(n_slots, dtype=torch.uint8, device=device)
    else:
        coords, labels, images, status = out
    with torch.cuda.device(coords.device):
        check(
            lib.dh_region_sample(C.byref(tables), n_slots, k, ps, float(threshold), miss_limit, max_redraw, fixed_class,
                                 slots_per_table_draw, seed, slot_offset, coords.data_ptr(), labels.data_ptr(), images.data_ptr(),
                                 status.data_ptr(), _stream()),
            "dh_region_sample",
        )
    return coords, labels, images, status


def rasterize_polygons(edges: torch.Tensor, edge_off: torch.Tensor, reg_bbox: torch.Tensor, scale: float, mh: int, mw: int) -> torch.Tensor:
    """int32 [mh, mw] label map: 1 + index of the last polygon containing the pixel centre, 0 = background."""
    lib = _lib.require_device()
    _need_cuda(edges, "edges", torch.float64)
    _need_cuda(edge_off, "edge_off", torch.int32)
    _need_cuda(reg_bbox, "reg_bbox", torch.float64)
    out = torch.empty((mh, mw), dtype=torch.int32, device=edges.device)
    with torch.cuda.device(edges.device):
        check(lib.dh_rasterize_polygons(edges.data_ptr(), edge_off.data_ptr(), reg_bbox.data_ptr(), edge_off.numel() - 1, float(scale),
                                        out.data_ptr(), mh, mw, _stream()), "dh_rasterize_polygons")
    return out


def set_gather_variant(variant: str = "auto") -> None:
    """Profiling switch: "auto" (TMA-staged kernel when the shape allows), "direct" (LDG/STG kernel), "tma" (fail if unsupported)."""
    lib = _lib.load()
    check(lib.dh_gather_set_variant({"auto": 0, "direct": 1, "tma": 2, "tma_noload": 3, "tma_nostore": 4, "tma_nomem": 5, "tma_plainstore": 6, "tma_blocked": 7, "tma_evictfirst": 8, "tma_blocked_evictfirst": 9}[variant]),
          "dh_gather_set_variant")
