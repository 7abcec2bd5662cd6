"""Slide source seam.

The reference talks to `psimage.PSImage` (full_samplers.py:35-38,55,176-180; region_samplers.py:216,229,
501,513-520; predict_full_patched.py:37-38,103-105): context manager, `_assert_layer`, `layer_size`,
`get_region_from_layer(layer, (y0,x0), (y1,x1))`, `get_region`, `height`, `width`, `close`. psimage is not
part of the reference tree, so every sampler here accepts, wherever the reference takes a `.psi` path:
  * a `DeviceSlide` (already resident in HBM), a `SyntheticSlide`, a `PinnedSlide` (pinned host memory), a uint8 numpy array [H,W,3],
  * a path to a `.npy` file holding such an array,
  * any object with the PSImage duck type above (including a real psimage.PSImage if it is installed),
  * a `.psi` path when the `psimage` package is importable.
The chosen layer is uploaded ONCE to HBM (row pitch padded to 16 B) and all patch extraction runs there.
`layer` is a downscale factor (1, 2, 4, ...), as in the reference (region_samplers.py:68,792)."""

from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path
from typing import Any

import numpy as np

from .ops import DeviceSlide


@dataclass
class Patch:
    """Field-compatible stand-in for psimage.core.patches.Patch (fields used by the reference:
    predict_full_patched.py:51-52, full_samplers.py:87-88,286-288)."""

    layer: int
    pos_x: int
    pos_y: int
    patch_size: int
    data: Any = None


class ArraySlide:
    """numpy-backed object with the PSImage duck type (layer n = every n-th pixel of layer 1)."""

    def __init__(self, array: np.ndarray):
        if array.ndim != 3 or array.shape[2] != 3 or array.dtype != np.uint8:
            raise ValueError("ArraySlide needs a uint8 [H, W, 3] array")
        self._a = array
        self.height, self.width = array.shape[:2]

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def close(self):
        pass

    def _assert_layer(self, layer: int):
        if layer < 1 or int(layer) != layer:
            raise ValueError(f"invalid layer {layer}")

    def layer_size(self, layer: int):
        return self.height // layer, self.width // layer

    def get_region_from_layer(self, layer: int, p0, p1) -> np.ndarray:
        (y0, x0), (y1, x1) = p0, p1
        a = self._a if layer == 1 else self._a[::layer, ::layer][: self.height // layer, : self.width // layer]
        return a[y0:y1, x0:x1, :]

    def get_region(self, p0, p1, target_hw=None) -> np.ndarray:
        r = self.get_region_from_layer(1, p0, p1)
        if target_hw is None:
            return r
        th, tw = target_hw
        ys = (np.arange(th) * (r.shape[0] / th)).astype(np.int64)
        xs = (np.arange(tw) * (r.shape[1] / tw)).astype(np.int64)
        return r[ys][:, xs]


class SyntheticSlide:
    """Seeded synthetic slide generated directly in HBM by dh_synth_slide; layer must be 1."""

    def __init__(self, H: int, W: int, seed: int = 0):
        self.height, self.width, self.seed = int(H), int(W), int(seed)
        self._dev: dict[str, DeviceSlide] = {}

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def close(self):
        pass

    def _assert_layer(self, layer: int):
        if layer != 1:
            raise ValueError("SyntheticSlide only has layer 1")

    def layer_size(self, layer: int):
        self._assert_layer(layer)
        return self.height, self.width

    def device_slide(self, device="cuda") -> DeviceSlide:
        key = str(device)
        if key not in self._dev:
            self._dev[key] = DeviceSlide.synthetic(self.height, self.width, self.seed, device)
        return self._dev[key]

    def device_band(self, y0: int, y1: int, device="cuda") -> DeviceSlide:
        """Rows [y0, y1) generated directly in HBM (no full-slide buffer)."""
        return DeviceSlide.synthetic(self.height, self.width, self.seed, device, y0=y0, rows=y1 - y0)

    def get_region_from_layer(self, layer: int, p0, p1) -> np.ndarray:
        self._assert_layer(layer)
        (y0, x0), (y1, x1) = p0, p1
        s = self.device_slide()
        return s.rows2d()[y0:y1, 3 * x0 : 3 * x1].cpu().numpy().reshape(y1 - y0, x1 - x0, 3)


class PinnedSlide:
    """A slide layer held in PINNED host memory with the device row pitch (uint8 [H, pitch], pitch % 16 == 0): the upload to
    HBM is one asynchronous cudaMemcpy at PCIe speed (a pageable numpy array goes through a staging copy first). Layer 1 only."""

    def __init__(self, host, H: int, W: int, pitch: int, y_origin: int = 0, full_height: int = None):
        """`y_origin` / `full_height`: the buffer holds rows [y_origin, y_origin + H) of a layer of `full_height` rows (one rank's
        row band of a sharded slide); row arguments of every method stay in layer coordinates."""
        import torch

        if not (isinstance(host, torch.Tensor) and host.dtype == torch.uint8 and not host.is_cuda and host.numel() >= H * pitch):
            raise ValueError("PinnedSlide needs a host uint8 tensor of at least H * pitch bytes")
        self.host, self.rows, self.width, self.pitch = host, int(H), int(W), int(pitch)
        self.y_origin = int(y_origin)
        self.height = int(full_height) if full_height is not None else self.y_origin + self.rows
        self.pinned = host.is_pinned()      # False only when page-locking failed (_host_buffer): uploads then go through staging

    def _rows(self, y0: int, y1: int):
        if not (self.y_origin <= y0 <= y1 <= self.y_origin + self.rows):
            raise ValueError(f"rows [{y0}, {y1}) are outside the band [{self.y_origin}, {self.y_origin + self.rows}) this buffer holds")
        return y0 - self.y_origin, y1 - self.y_origin

    @staticmethod
    def _host_buffer(nbytes: int):
        import torch

        try:
            return torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)     # page-locked at allocation: no pageable staging copy
        except RuntimeError:               # page-locking refused (ulimit / small host, or no CUDA runtime): a pageable buffer
            return torch.empty(nbytes, dtype=torch.uint8)

    @classmethod
    def from_numpy(cls, arr: np.ndarray) -> "PinnedSlide":
        import torch

        H, W, _ = arr.shape
        pitch = DeviceSlide.pitch_for(W)
        host = cls._host_buffer(H * pitch)
        host.view(H, pitch)[:, : 3 * W].copy_(torch.from_numpy(np.ascontiguousarray(arr)).view(H, 3 * W))
        return cls(host, H, W, pitch)

    @classmethod
    def from_device(cls, dev: DeviceSlide, y_origin: int = 0, full_height: int = None) -> "PinnedSlide":
        import torch

        host = cls._host_buffer(dev.H * dev.pitch)
        host.copy_(dev.storage[: dev.H * dev.pitch])
        return cls(host, dev.H, dev.W, dev.pitch, y_origin, full_height)

    @property
    def nbytes(self) -> int:
        return self.rows * self.pitch

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def close(self):
        pass

    def _assert_layer(self, layer: int):
        if layer != 1:
            raise ValueError("a PinnedSlide holds exactly one layer; pass layer=1")

    def layer_size(self, layer: int):
        self._assert_layer(layer)
        return self.height, self.width

    def get_region_from_layer(self, layer: int, p0, p1) -> np.ndarray:
        self._assert_layer(layer)
        (y0, x0), (y1, x1) = p0, p1
        a, b = self._rows(y0, y1)
        return self.host.view(-1, self.pitch)[a:b, 3 * x0 : 3 * x1].numpy().reshape(y1 - y0, x1 - x0, 3)

    def to_device(self, device="cuda", y0: int = None, y1: int = None) -> DeviceSlide:
        """Rows [y0, y1) (layer coordinates; default: everything this buffer holds) as a DeviceSlide whose row 0 is row y0."""
        import torch

        y0 = self.y_origin if y0 is None else y0
        y1 = self.y_origin + self.rows if y1 is None else y1
        a, b = self._rows(y0, y1)
        storage = torch.empty((b - a) * self.pitch, dtype=torch.uint8, device=device)
        storage.copy_(self.host[a * self.pitch : b * self.pitch], non_blocking=True)
        return DeviceSlide(storage, b - a, self.width, self.pitch)


def open_slide(source):
    """Return a PSImage-duck-typed object for `source` (see module docstring)."""
    if isinstance(source, (ArraySlide, SyntheticSlide, DeviceSlideSource, PinnedSlide)):
        return source
    if isinstance(source, DeviceSlide) or type(source).__name__ == "MappedHostSlide":
        return DeviceSlideSource(source)
    if isinstance(source, np.ndarray):
        return ArraySlide(source)
    if isinstance(source, (str, Path)):
        p = Path(source)
        if p.suffix == ".npy":
            return ArraySlide(np.load(p, mmap_mode="r"))
        try:
            from psimage.core.image import PSImage  # type: ignore
        except ImportError as e:
            raise RuntimeError(
                f"cannot open {p}: the `psimage` package is not installed; pass a .npy file, a numpy array, "
                "a DeviceSlide or any object with the PSImage interface"
            ) from e
        return PSImage(p)
    if all(hasattr(source, a) for a in ("layer_size", "get_region_from_layer")):
        return source
    raise TypeError(f"unsupported slide source {type(source)!r}")


def tile_spans(rects, H: int, W: int, pitch: int, tile: int = 512):
    """Cover pixel rectangles (y0, y1, x0, x1) with `tile` x `tile` tiles and return the covered area as an int64 array [n][4] =
    {y0, y1, byte_x0, byte_x1} (byte columns 16-byte aligned, vertically adjacent identical runs merged) for dh_upload_rects."""
    ty, tx = -(-H // tile), -(-W // tile)
    occ = np.zeros((ty, tx), dtype=bool)
    for y0, y1, x0, x1 in rects:
        y0, y1, x0, x1 = max(0, int(y0)), min(H, int(y1)), max(0, int(x0)), min(W, int(x1))
        if y1 > y0 and x1 > x0:
            occ[y0 // tile : (y1 - 1) // tile + 1, x0 // tile : (x1 - 1) // tile + 1] = True
    out, open_runs = [], {}                                    # open_runs: (b0, b1) -> first tile row of a vertical stack of equal runs
    for i in range(ty + 1):
        runs = set()
        if i < ty:
            row = occ[i]
            j = 0
            while j < tx:
                if row[j]:
                    k = j
                    while k < tx and row[k]:
                        k += 1
                    runs.add(((j * tile * 3) // 16 * 16, min(pitch, -(-(min(W, k * tile) * 3) // 16) * 16)))
                    j = k
                else:
                    j += 1
        for key in [k for k in open_runs if k not in runs]:
            out.append((open_runs.pop(key) * tile, min(H, i * tile), key[0], key[1]))
        for key in runs:
            open_runs.setdefault(key, i)
    return np.asarray(sorted(out), dtype=np.int64).reshape(-1, 4)


def upload_rects(host: "PinnedSlide", rects, device="cuda", tile: int = 512):
    """DeviceSlide holding only the tiles of `host` that intersect `rects` (the rest is zero) and the number of bytes that
    travelled: what an annotated sampler needs of a slide -- the reference reads just the patches it draws from storage
    (region_samplers.py:513-520)."""
    import torch

    from . import _lib

    if host.y_origin != 0 or host.rows != host.height:
        raise ValueError("upload_rects needs a PinnedSlide holding the whole layer")
    lib = _lib.require_device()
    spans = np.ascontiguousarray(tile_spans(rects, host.height, host.width, host.pitch, tile))
    dev = torch.device(device)
    storage = torch.zeros(host.height * host.pitch, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.dh_upload_rects(storage.data_ptr(), host.height, host.pitch, host.host.data_ptr(), len(spans), spans.ctypes.data,
                                       torch.cuda.current_stream(dev).cuda_stream), "dh_upload_rects")
    nbytes = int(((spans[:, 1] - spans[:, 0]) * (spans[:, 3] - spans[:, 2])).sum()) if len(spans) else 0
    return DeviceSlide(storage, host.height, host.width, host.pitch), nbytes


def alloc_bytes_on(alloc_stream, nbytes: int, device):
    """A uint8 device buffer for work on the CURRENT stream, allocated from `alloc_stream`'s pool of torch's caching allocator (pools
    are per stream: a background copy stream never finds the block the application freed on its own stream and pays a cudaMalloc
    -- 2.5 ms for 3.2 GB, once measured at 126 ms next to another rank's allocation, profiles/r02_e2e.md). The current stream waits
    for everything enqueued on `alloc_stream` so far (the block's previous owner), and the allocator is told about the new user."""
    import torch

    cur = torch.cuda.current_stream(device)
    if alloc_stream is None or alloc_stream == cur:
        return torch.empty(nbytes, dtype=torch.uint8, device=device)
    with torch.cuda.stream(alloc_stream):
        storage = torch.empty(nbytes, dtype=torch.uint8, device=device)
        ev = torch.cuda.Event()
        ev.record(alloc_stream)
    cur.wait_event(ev)
    storage.record_stream(cur)
    return storage


def sharded_upload(host: "PinnedSlide", device="cuda", group=None, stats: dict = None, alloc_stream=None):
    """Collective slide ingestion for data-parallel sampling: every rank of `group` holds the same layer in host memory and needs
    it resident. Each rank copies only its 1/world share of the rows over ITS PCIe link and one all-gather over NVLink replicates
    the shares (in place: a rank's share sits at its offset of the full buffer) -- instead of `world` full uploads competing for
    the host's memory and PCIe bandwidth. Returns (DeviceSlide, bytes this rank copied from the host). Works with NCCL (CUDA) and
    gloo (CPU tensors; used by the world_size-2 CPU tests). `stats` (optional dict) receives "alloc_ms" (host time of the device
    allocation) and, on CUDA, the event pairs "upload" and "allgather" for the two phases. `alloc_stream`: see alloc_bytes_on."""
    import time

    import torch
    import torch.distributed as dist

    if host.y_origin != 0 or host.rows != host.height:
        raise ValueError("sharded_upload needs a PinnedSlide holding the whole layer")
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    per = -(-host.rows // world)                                   # rows per rank; the last shares are padded
    pitch = host.pitch
    t0 = time.perf_counter()
    if torch.device(device).type == "cuda":
        storage = alloc_bytes_on(alloc_stream, per * world * pitch, device)
    else:
        storage = torch.empty(per * world * pitch, dtype=torch.uint8, device=device)
    timed = stats is not None and storage.is_cuda
    if stats is not None:
        stats["alloc_ms"] = 1e3 * (time.perf_counter() - t0)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if timed else None
    a, b = min(host.rows, rank * per), min(host.rows, (rank + 1) * per)
    mine = storage[rank * per * pitch : (rank + 1) * per * pitch]
    if timed:
        ev[0].record()
    if b > a:
        mine[: (b - a) * pitch].copy_(host.host[a * pitch : b * pitch], non_blocking=True)
    if (b - a) < per:
        mine[(b - a) * pitch :].zero_()
    if timed:
        ev[1].record()
    dist.all_gather_into_tensor(storage, mine, group=group)
    if timed:
        ev[2].record()
        stats["upload"], stats["allgather"] = (ev[0], ev[1]), (ev[1], ev[2])
    if storage.is_cuda:
        return DeviceSlide(storage, host.rows, host.width, pitch), (b - a) * pitch
    return storage[: host.rows * pitch], (b - a) * pitch          # CPU (gloo) callers get the assembled bytes


class DeviceSlideSource:
    """PSImage duck type over a slide that already lives in HBM (layer 1 only)."""

    def __init__(self, dev: DeviceSlide):
        self.dev = dev
        self.height, self.width = dev.H, dev.W

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def close(self):
        pass

    def _assert_layer(self, layer: int):
        if layer != 1:
            raise ValueError("a DeviceSlide holds exactly one layer; pass layer=1")

    def layer_size(self, layer: int):
        self._assert_layer(layer)
        return self.dev.H, self.dev.W

    def get_region_from_layer(self, layer: int, p0, p1) -> np.ndarray:
        self._assert_layer(layer)
        (y0, x0), (y1, x1) = p0, p1
        s = self.dev
        return s.rows2d()[y0:y1, 3 * x0 : 3 * x1].cpu().numpy().reshape(y1 - y0, x1 - x0, 3)


def layer_to_device(src, layer: int, device="cuda", resident: bool = True):
    """Upload layer `layer` of an opened slide to HBM once (the analogue of full_samplers.py:53-55). resident=False (PinnedSlide
    sources only): no upload -- a MappedHostSlide that the gather kernels read in place over PCIe (slides larger than HBM)."""
    if not resident:
        from .ops import MappedHostSlide

        if not (isinstance(src, PinnedSlide) and src.pinned and src.y_origin == 0 and src.rows == src.height):
            raise ValueError("resident=False needs a PinnedSlide that holds the whole layer in page-locked memory")
        src._assert_layer(layer)
        return MappedHostSlide(src.host, src.rows, src.width, src.pitch, device)
    if isinstance(src, DeviceSlideSource):
        src._assert_layer(layer)
        return src.dev
    if isinstance(src, SyntheticSlide):
        src._assert_layer(layer)
        return src.device_slide(device)
    if isinstance(src, PinnedSlide):
        src._assert_layer(layer)
        return src.to_device(device)
    h, w = src.layer_size(layer)
    arr = np.asarray(src.get_region_from_layer(layer, (0, 0), (h, w)))
    return DeviceSlide.from_numpy(arr, device)


def band_to_device(src, layer: int, y0: int, y1: int, device="cuda") -> DeviceSlide:
    """Upload rows [y0, y1) of layer `layer` (a row band with its halo); row 0 of the result is slide row y0."""
    if isinstance(src, DeviceSlideSource):
        src._assert_layer(layer)
        s = src.dev
        return DeviceSlide(s.storage[y0 * s.pitch : y1 * s.pitch], y1 - y0, s.W, s.pitch)
    if isinstance(src, SyntheticSlide):
        src._assert_layer(layer)
        return src.device_band(y0, y1, device)
    if isinstance(src, PinnedSlide):
        src._assert_layer(layer)
        return src.to_device(device, y0, y1)
    h, w = src.layer_size(layer)
    arr = np.asarray(src.get_region_from_layer(layer, (y0, 0), (y1, w)))
    return DeviceSlide.from_numpy(arr, device)
