"""Whole-slide samplers -- drop-in for the reference's `patch_samplers/full_samplers.py`.

Same class names, constructor arguments, iterator protocol and yielded shapes as the reference
(FullImageRndSampler full_samplers.py:21-299, FullImageDenseSampler :302-452); the work is done by the
sm_100a kernels of libdeephisto_b200.so on a slide that is uploaded to HBM once:
  coordinates   dh_dense_coords (:374-404) / dh_cover_sample (:81-162)
  pixels        dh_gather_normalize (:187-202, :353-369, :437-452)
`generator_torch()` yields CUDA tensors (the reference yields CPU tensors that the caller moves to the
device, predict_full_patched.py:70, train.py:166). `generator()` / `__iter__` still yield `list[Patch]`
with numpy uint8 data, materialised lazily with one device->host copy per batch.
Both `SamplerExecutionMode`s run the same HBM-resident path; the reference's ProcessPool / shared-memory
machinery (:57-60,229-261,406-423) has no equivalent here and no worker processes are created."""

from __future__ import annotations

from enum import Enum
from pathlib import Path
from typing import Iterable, Iterator

import numpy as np
import torch

from .. import ops
from ..slide import Patch, band_to_device, layer_to_device, open_slide


class SamplerExecutionMode(Enum):
    INMEMORY_SINGLEPROC = 1
    ONDISK_MULTIPROC = 2


def _patches_from_device(slide, coords_dev: torch.Tensor, layer: int, ps: int) -> list[Patch]:
    """list[Patch] with uint8 [ps,ps,3] numpy data for int32 device coords (one D2H copy)."""
    raw = ops.gather_normalize(slide, coords_dev, ps, dtype=torch.uint8, layout="NHWC")
    data = raw.cpu().numpy()
    yx = coords_dev.cpu().numpy()
    return [Patch(layer=layer, pos_x=int(x), pos_y=int(y), patch_size=ps, data=data[i]) for i, (y, x) in enumerate(yx.tolist())]


class FullImageRndSampler:
    """Coverage-driven random sampling until every coarse (1/speedup) cell is covered (full_samplers.py:21-299).

    Extra keyword-only arguments (not in the reference): `seed` (Philox key; the reference uses the unseeded
    global numpy RNG), `device`, `lazy_slide` (row-band sharded prediction: the layer is not uploaded by the constructor, every
    rank makes only its band resident through `band_sampler`), `resident=False` (PinnedSlide sources: the layer is never uploaded,
    patches are gathered in place from pinned host memory over PCIe -- slides that do not fit HBM)."""

    def __init__(self, psimage_path: Path, layer: int, patch_size: int, batch_size: int, mode: SamplerExecutionMode,
                 dense_level: int = 2, speedup: int = 16, *, seed: int = 0, device="cuda", lazy_slide: bool = False, quiet: bool = False,
                 resident: bool = True):
        self.mode = mode
        self._psim_path = psimage_path
        self._src = open_slide(psimage_path)
        self._slide_dev = None
        self._device = device
        self._resident = resident
        with self._src as psim:
            self.layer = layer
            psim._assert_layer(layer)
            self.h, self.w = psim.layer_size(self.layer)
            if not lazy_slide:
                self._slide_dev = layer_to_device(psim, layer, device, resident=resident)
        self.dh = self.h // speedup
        self.dw = self.w // speedup
        if not quiet:
            print(f"Image {self.h} x {self.w} at {speedup}x -> {self.dh} x {self.dw}")
        self.patch_size = patch_size
        self.batch_size = batch_size
        self._downscale = speedup
        self.dense_level = dense_level
        self._filled_ratio: list[float] = []
        self._seed = seed
        self._device = device
        self._state: ops.CoverState | None = None
        self._resume = None
        self._consumed = 0             # batches handed to the consumer by the running / last iteration

    @property
    def _slide(self):
        if self._slide_dev is None:
            with self._src as psim:
                self._slide_dev = layer_to_device(psim, self.layer, self._device, resident=self._resident)
        return self._slide_dev

    def band_sampler(self, y0: int, y1: int, stream_index: int) -> "FullImageRndSampler":
        """A coverage sampler over slide rows [y0, y1) only (one rank's row band of a sharded prediction): same parameters, its own
        accumulator, Philox key = (seed, stream_index) so that the bands draw from disjoint substreams. Coordinates it yields are
        relative to row y0. A resident slide is viewed in place, a lazy one uploads / generates just these rows."""
        if not (0 <= y0 < y1 <= self.h) or y1 - y0 < self.patch_size:
            raise ValueError(f"band [{y0}, {y1}) must lie inside the {self.h}-row layer and hold at least one patch")
        if self._slide_dev is not None:
            band = self._slide_dev.view_rows(y0, y1)
        elif not self._resident:
            band = self._slide.view_rows(y0, y1)
        else:
            with self._src as psim:
                band = band_to_device(psim, self.layer, y0, y1, self._device)
        seed = (int(self._seed) & 0xFFFFFFFF) | ((int(stream_index) + 1) << 32)
        return FullImageRndSampler(band, 1, self.patch_size, self.batch_size, self.mode, self.dense_level, self._downscale, seed=seed,
                                   device=self._device, quiet=True)

    # -- resume: the Philox draws are keyed by (seed, batch index) and the accumulator is a pure function of the batches drawn, so
    #    (seed, number of batches handed to the consumer) is the whole state of a run
    def state_dict(self) -> dict:
        """State as of the LAST BATCH HANDED TO THE CONSUMER. The device runs up to two groups ahead of the consumer; those batches are
        not part of the state -- a restored sampler draws them again (bit-identically) instead of counting their footprints as covered."""
        if self._resume is not None:                                # restored but not iterated yet: still the restored state
            return dict(self._resume)
        k = int(self._consumed)
        return {"seed": int(self._seed), "batch_index": k, "accum": None, "filled_ratio": list(self._filled_ratio[:k])}

    def load_state_dict(self, state: dict) -> None:
        """The next `generator*()` call continues the saved run: the first `batch_index` batches are replayed on the device (their
        coordinates discarded; ~40 us each) to rebuild the accumulator, then iteration continues with batch `batch_index`. States that
        carry an accumulator (saved by an earlier version after a complete iteration) are adopted directly (dh_cover_init)."""
        if int(state["seed"]) != int(self._seed):
            raise ValueError(f"state was saved with seed {state['seed']}, this sampler uses seed {self._seed}")
        self._resume = state if (state.get("accum") is not None or int(state.get("batch_index", 0)) > 0) else None

    # -- device-side iteration ------------------------------------------------------------------------
    def _group_generator(self, group: int = 16) -> Iterator[tuple[torch.Tensor, list[float]]]:
        """Groups of up to `group` batches: (int32 device coords [g, B, 2], their filled ratios), ending with the batch that reaches
        filled_ratio >= 1 (:263-274). The batches of a group are enqueued back to back (one launch each) and ONE read-back of their
        non-zero counters serves the group; launches that find the slide already covered are no-ops (dh_cover_sample,
        stop_when_full), so after a complete run the accumulator is exactly the footprint histogram of the yielded batches. The
        device runs up to two groups ahead of the batch being consumed (the reference's worker pool runs ahead as well)."""
        self._state = ops.CoverState(self.h, self.w, self.patch_size, self._downscale, self.dense_level, self.batch_size,
                                     self._seed, self._device)
        cells = self.dh * self.dw
        self._consumed = 0
        self._filled_ratio = []
        if self._resume is not None:
            res, self._resume = self._resume, None
            k = int(res["batch_index"])
            if res.get("accum") is not None:
                self._state.restore(res["accum"].to(self._device), k)
            else:
                for a in range(0, k, 256):                                       # replay: same launches, coordinates discarded
                    self._state.next_group(min(256, k - a))
            self._filled_ratio = list(res["filled_ratio"])
            self._consumed = k
            if self._filled_ratio and self._filled_ratio[-1] >= 1:
                return
        done = False
        ahead = self._state.next_group(group, read_back=True)
        while not done:
            (coords, (counts, ready)), ahead = ahead, self._state.next_group(group, read_back=True)   # next group enqueued before this one is read
            ready.synchronize()
            ratios = [c / cells for c in counts.tolist()]
            keep = len(ratios)
            for i, r in enumerate(ratios):
                if r >= 1:
                    keep, done = i + 1, True
                    break
            self._filled_ratio.extend(ratios[:keep])
            yield coords[:keep], ratios[:keep]

    def coords_generator(self) -> Iterator[tuple[torch.Tensor, float]]:
        """(int32 device coords [B,2], filled_ratio) per batch until filled_ratio >= 1 (:263-274)."""
        for coords, ratios in self._group_generator():
            for g, r in enumerate(ratios):
                self._consumed += 1                                              # state_dict(): resume behind the batch being handed out
                yield coords[g], r

    def generator(self) -> Iterator[tuple[list[Patch], float]]:
        for coords, filled_ratio in self.coords_generator():
            yield _patches_from_device(self._slide, coords, self.layer, self.patch_size), filled_ratio

    def __iter__(self) -> Iterator[tuple[list[Patch], float]]:
        return self.generator()

    def generator_torch(self, normalize: bool = False, dtype=torch.float32,
                        layout: str = "NHWC") -> Iterator[tuple[torch.Tensor, torch.Tensor, float]]:
        """features [B,ps,ps,3] float32 with values 0..255 -- the reference does NOT divide by 255 here
        (:286); pass normalize=True for [0,1]. coords float32 [B,2] (y, x)."""
        for coords, ratios in self._group_generator():                       # one gather launch per group of batches
            g = len(ratios)
            features = ops.gather_normalize(self._slide, coords.reshape(g * self.batch_size, 2), self.patch_size, dtype=dtype, layout=layout,
                                            scale255=normalize)
            features = features.view((g, self.batch_size) + tuple(features.shape[1:]))
            coords_f = coords.to(torch.float32)
            for i, r in enumerate(ratios):
                self._consumed += 1
                yield features[i], coords_f[i], r

    # -- reporting helpers of the reference -------------------------------------------------------------
    @property
    def _accum(self):
        return None if self._state is None else self._state.accum.cpu().numpy().astype(np.float32)

    def plot_empty_area_history(self, filename: str):
        """Coverage per iteration as a JPEG (reference :292-298: same title, axis labels, 300 dpi)."""
        try:
            from matplotlib.figure import Figure
        except ImportError as e:
            raise RuntimeError("plot_empty_area_history needs matplotlib") from e
        fig = Figure()
        ax = fig.add_subplot()
        ax.plot(range(len(self._filled_ratio)), self._filled_ratio)
        ax.set(title="Empty area", xlabel="iteration", ylabel="empty area percentage")
        fig.savefig(filename, format="jpg", dpi=300)

    def visualize_heatmap(self, name: str):
        """Two images like the reference's (:292-299): the coverage counts scaled to 0..255 under `name`, the covered / uncovered
        mask under "_" + name."""
        from PIL import Image

        counts = self._accum
        if counts is None:
            return
        peak = float(counts.max())
        scaled = (counts / peak * 255).astype(np.uint8)                  # an untouched accumulator divides by zero, as in the reference
        Image.fromarray(scaled).save(name)
        Image.fromarray(((scaled > 0) * np.uint8(255)).astype(np.uint8)).save("_" + name, quality=98)


class FullImageDenseSampler:
    """Dense grid + last column / row / corner, batched, last batch padded with the corner (:302-452).

    `stride=None` means stride = patch_size (the reference would raise a TypeError in `range`)."""

    def __init__(self, psimage_path: Path, layer: int, patch_size: int, batch_size: int, mode: SamplerExecutionMode,
                 stride: int = None, *, device="cuda", lazy_slide: bool = False):
        """`lazy_slide=True` (row-band sharding, BASELINE config 4) does not upload the whole layer: each rank uploads only
        the rows of its band through `band_slide`."""
        self._psim_path = psimage_path
        self.mode = mode
        self._src = open_slide(psimage_path)
        self._slide_dev = None
        with self._src as psim:
            self.layer = layer
            psim._assert_layer(layer)
            self.h, self.w = psim.layer_size(self.layer)
            if not lazy_slide:
                self._slide_dev = layer_to_device(psim, layer, device)
        self.patch_size = patch_size
        self.batch_size = batch_size
        self.stride = patch_size if stride is None else stride
        self._device = device
        print(f"Image {self.h} x {self.w}")
        self.n_patches, self.n_padded = ops.dense_count(self.h, self.w, self.patch_size, self.stride, self.batch_size)

    def __len__(self) -> int:
        return self.n_padded // self.batch_size

    @property
    def _slide(self):
        if self._slide_dev is None:
            with self._src as psim:
                self._slide_dev = layer_to_device(psim, self.layer, self._device)
        return self._slide_dev

    def band_slide(self, y0: int, y1: int):
        """(DeviceSlide holding rows [y0, y1) of the layer, y0): what a rank needs for its row band plus the patch-size halo.
        A slide that is already resident is used as it is (offset 0)."""
        if self._slide_dev is not None:
            return self._slide_dev, 0
        with self._src as psim:
            return band_to_device(psim, self.layer, y0, y1, self._device), y0

    def coords_device(self) -> torch.Tensor:
        """int32 [n_padded, 2] device tensor of all (y, x) in reference order, padding included."""
        return ops.dense_coords(self.h, self.w, self.patch_size, self.stride, self.batch_size, device=self._device)

    def _create_batched_coords(self) -> list[list[tuple[int, int]]]:
        """Same return value as the reference's method (:374-404), produced by the device enumeration."""
        c = self.coords_device().cpu().numpy().tolist()
        b = self.batch_size
        return [[(y, x) for y, x in c[i : i + b]] for i in range(0, len(c), b)]

    def coords_generator(self) -> Iterator[tuple[torch.Tensor, float]]:
        coords = self.coords_device()
        n_batches = len(self)
        for i in range(n_batches):
            yield coords[i * self.batch_size : (i + 1) * self.batch_size], i / n_batches

    def generator(self) -> Iterable[tuple[list[Patch], float]]:
        for coords, progress in self.coords_generator():
            yield _patches_from_device(self._slide, coords, self.layer, self.patch_size), progress

    def __iter__(self) -> Iterable[tuple[list[Patch], float]]:
        return self.generator()

    def generator_torch(self, dtype=torch.float32, layout: str = "NHWC", mean=None,
                        std=None) -> Iterator[tuple[torch.Tensor, torch.Tensor, float]]:
        """features [B,ps,ps,3] float32 in [0,1] (bit-identical to `.astype(float32) / 255`, :441-443),
        coords float32 [B,2] (y, x), progress i / n_batches (never reaches 1.0, like the reference).

        Up to 64 batches (<= ~2.5 GB of features) are gathered by ONE launch and yielded as contiguous slices: short launches
        cannot fill HBM (profiles/r01_gather.md)."""
        coords = self.coords_device()
        coords_f = coords.to(torch.float32)
        n_batches, B, ps = len(self), self.batch_size, self.patch_size
        per_batch = B * ps * ps * 3 * torch.empty((), dtype=dtype).element_size()
        ahead = max(1, min(64, (5 << 29) // max(per_batch, 1)))
        for b0 in range(0, n_batches, ahead):
            nb = min(ahead, n_batches - b0)
            features = ops.gather_normalize(self._slide, coords[b0 * B : (b0 + nb) * B], ps, dtype=dtype, layout=layout, scale255=True,
                                            mean=mean, std=std)
            for i, f in enumerate(features.view((nb, B) + tuple(features.shape[1:])).unbind(0)):
                yield f, coords_f[(b0 + i) * B : (b0 + i + 1) * B], (b0 + i) / n_batches
