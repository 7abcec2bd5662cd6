"""Annotation-driven samplers -- drop-in for the reference's `patch_samplers/region_samplers.py`.

Same classes and call signatures as the reference (RegionAnnotation :18-191, AnnoRegionRndSampler :252-796,
AnnoRegionDenseSampler :799-871, extract_and_save_subset :874-909). Differences, all deliberate:
  * polygons are packed once into device tables (deephisto_b200/geometry.py); candidate generation,
    the exact clip-area acceptance test and the patch gather run in sm_100a kernels;
  * randomness is Philox4x32-10 keyed by `seed` (keyword argument, default 0) with documented counters, not the
    unseeded global numpy RNG; `max_workers` / `batches_per_worker` are accepted, and no worker process is ever
    spawned (the reference's mp.set_start_method("spawn", force=True), :314-323, is dropped -- SURVEY Q12);
  * `torch_generator` yields CUDA tensors;
  * random origins are clamped so that patches never overhang the slide; `cls_idx=0` means class 0,
    not "any class" (Q5); `torch_iterable_dataset` yields (y, x), not (y, y) (Q6);
  * invalid (self-intersecting) polygons: shapely's buffer(0) repair (:69-71) needs GEOS; by default such a ring is kept with
    winding-number area semantics and named in a warning (`invalid_polygons="skip"` drops it, also with a warning)."""

from __future__ import annotations

import json
import warnings
from collections import defaultdict
from pathlib import Path
from typing import Iterator

import numpy as np
import torch
from torch.utils.data import IterableDataset

from .. import geometry, ops
from ..slide import Patch, PinnedSlide, alloc_bytes_on, layer_to_device, open_slide, sharded_upload, tile_spans, upload_rects


_STREAMS: dict = {}


def _shared_stream(device, role: str) -> "torch.cuda.Stream":
    """One side stream per (device, role) for ALL samplers of the process. torch's caching allocator keeps a separate pool of freed
    blocks per stream: a sampler that created its own producer stream would pay a cudaMalloc of its multi-GB prefetch buffers (~2 ms
    per GB, measured in the bench's e2e leg) even when an earlier sampler has just released buffers of the same size."""
    dev = torch.device(device)
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device(), role)
    if key not in _STREAMS:
        _STREAMS[key] = torch.cuda.Stream(dev)
    return _STREAMS[key]


class RegionAnnotation:
    """One annotated polygon (reference :18-191). `polygon` is replaced by `vertices_scaled` + an edge table."""

    def __init__(self, img_path: Path, region_idx: int, class_: str, vertices: np.ndarray, layer: int,
                 layer_size: tuple[int, int], *, image_index: int = 0, seed: int = 0, device="cuda", invalid: str = "approximate"):
        self.file_path = img_path
        self.region_idx = region_idx
        self.class_ = class_
        self.vertices = vertices
        self._layer = layer
        self._layer_size = layer_size
        self._region = geometry.make_region(image_index, region_idx, class_, vertices, layer, invalid=invalid)  # raises like :64-67
        self.area = self._region.area
        self.bounds = self._region.bounds
        self._seed, self._device, self._calls = seed, device, 0
        self._edges_dev = None
        self._tables = None

    def __str__(self) -> str:
        stem = Path(self.file_path).stem if isinstance(self.file_path, (str, Path)) else type(self.file_path).__name__
        return f"Region [{stem}, {self.region_idx}, {self.class_}, {self.vertices.shape}, {round(self.area, 0)}]"

    def _edges(self) -> torch.Tensor:
        if self._edges_dev is None:
            self._edges_dev = torch.from_numpy(self._region.edges.reshape(-1)).to(self._device)
        return self._edges_dev

    def _extract_patch_coords_rnd(self, patch_size: int, n_patches: int, region_intersection: float = 0.75,
                                  miss_limit: int = 500) -> list[tuple[int, int]]:
        """Reference :82-143. Raises RuntimeError("Region is too small.") / ("Miss limit reached...") alike."""
        ps = patch_size
        thr = ps * ps * region_intersection
        if self.area < thr:
            raise RuntimeError("Region is too small.")
        if self._tables is None:
            reg = self._region
            one = geometry.Region(0, reg.region_idx, reg.class_, reg.vertices, reg.area, reg.bounds, reg.edges)
            self._tables = geometry.RegionTables([one], [reg.class_], [tuple(self._layer_size)], [{reg.class_: [0]}], np.ones(1), 0.0,
                                                 device=self._device)
        offset = self._calls * (1 << 20)
        self._calls += 1
        coords, _, _, status = ops.region_sample(self._tables.struct, n_patches, 1, ps, thr, miss_limit=miss_limit, max_redraw=1,
                                                 seed=self._seed ^ (self.region_idx << 32), slot_offset=offset, device=self._device)
        st = status.cpu().numpy()
        if (st != 0).any():
            raise RuntimeError("Miss limit reached. Probably region is too small.")
        return [(int(y), int(x)) for y, x in coords.cpu().numpy().tolist()]

    def _dense_grid(self, patch_size: int, stride: int):
        """Candidate grid of reference :171-179 (Python round = banker's rounding; upper clamps only)."""
        h, w = self._layer_size
        x0, y0, x1, y1 = self.bounds
        x0, y0, x1, y1 = round(x0), round(y0), round(x1), round(y1)
        x1 = min(x1, w - patch_size)
        y1 = min(y1, h - patch_size)
        return y0, x0, len(range(y0, y1, stride)), len(range(x0, x1, stride))

    def _extract_patch_coords_dense_device(self, patch_size: int, stride: int, region_intersection: float = 0.75) -> torch.Tensor:
        y0, x0, ny, nx = self._dense_grid(patch_size, stride)
        if ny == 0 or nx == 0:
            return torch.zeros((0, 2), dtype=torch.int32, device=self._device)
        e = self._edges()
        mask, _ = ops.region_accept_dense(e, 0, len(self._region.edges), y0, x0, ny, nx, stride, patch_size,
                                          patch_size * patch_size * region_intersection)
        return ops.compact_coords(mask, y0, x0, ny, nx, stride)

    def _extract_patch_coords_dense(self, patch_size: int, stride: int, region_intersection: float = 0.75) -> list[tuple[int, int]]:
        """Reference :145-191: accepted (y, x) in row-major order."""
        c = self._extract_patch_coords_dense_device(patch_size, stride, region_intersection)
        return [(int(y), int(x)) for y, x in c.cpu().numpy().tolist()]


def _load_annotation(anno) -> list[dict]:
    if isinstance(anno, (str, Path)):
        with open(anno) as f:
            return json.load(f)
    return list(anno)


def _parse_annotations(img_anno_paths, layer: int, classes: list[str] = None, *, seed: int = 0, device="cuda", verbose: bool = True,
                       invalid: str = "approximate"):
    """Reference :194-249. Returns (regions_all, regions_per_image, layer sizes, opened slide sources)."""
    regions_all = defaultdict(list)
    regions_per_image = [defaultdict(list) for _ in img_anno_paths]
    sizes, sources = [], []
    regions_failed = 0
    for j, (psim_path, anno_path) in enumerate(img_anno_paths):
        src = open_slide(psim_path)
        sources.append(src)
        with src as psim:
            size = tuple(psim.layer_size(layer))
            sizes.append(size)
            for i, a in enumerate(_load_annotation(anno_path)):
                cls = a["class"]
                if classes is not None and cls not in classes:
                    continue
                try:
                    reg = RegionAnnotation(img_path=psim_path, region_idx=i, class_=cls, vertices=np.array(a["vertices"], dtype=np.float64),
                                           layer=layer, layer_size=size, image_index=j, seed=seed, device=device, invalid=invalid)
                    regions_per_image[j][cls].append(reg)
                    regions_all[cls].append(reg)
                except Exception as e:
                    regions_failed += 1
                    # the reference only counts failures (:235-236); a dropped region changes class areas and sampling weights, so name it
                    warnings.warn(f"annotation dropped: image {j}, region {i} (class {cls!r}): {e}", RuntimeWarning, stacklevel=2)
    if verbose:
        if regions_failed > 0:
            print(f"Failed to parse {regions_failed} regions.")
        print(f"regions all: { {cls: len(r) for cls, r in regions_all.items()} }")
        print("regions per image:")
        for i, rpi in enumerate(regions_per_image):
            print(f"\timage {i}: { {cls: len(r) for cls, r in rpi.items()} }")
    return regions_all, regions_per_image, sizes, sources


def build_tables(images, layer: int, area_influence: float, classes, one_image_for_batch: bool, device="cuda", invalid: str = "approximate"):
    """(RegionTables, flat region list, sorted class names) from [(layer_hw, [ {"class","vertices"} ...]) ...].
    Table layout: struct dh_region_tables; weights: reference _calc_weights :395-482."""
    regions: list[geometry.Region] = []
    per_image: list[dict[str, list[int]]] = [dict() for _ in images]
    everything: dict[str, list[int]] = {}
    for j, (hw, annos) in enumerate(images):
        for i, a in enumerate(annos):
            cls = a["class"]
            if classes is not None and cls not in classes:
                continue
            try:
                with warnings.catch_warnings():                       # callers that parsed the annotations first have warned already
                    warnings.simplefilter("ignore", RuntimeWarning)
                    reg = geometry.make_region(j, i, cls, np.array(a["vertices"], dtype=np.float64), layer, invalid=invalid)
            except Exception:
                continue
            per_image[j].setdefault(cls, []).append(len(regions))
            everything.setdefault(cls, []).append(len(regions))
            regions.append(reg)
    names = sorted(everything.keys())
    if one_image_for_batch:
        img_areas = [sum(sum(regions[r].area for r in regs) for regs in t.values()) for t in per_image]
        weights = geometry.area_weights(img_areas, area_influence)                                  # :469-475
        tables = geometry.RegionTables(regions, names, [tuple(hw) for hw, _ in images], per_image, weights, area_influence, device)
    else:
        tables = geometry.RegionTables(regions, names, [tuple(hw) for hw, _ in images], [everything], np.ones(1), area_influence, device,
                                       sorted_classes_per_table=True)
    return tables, regions, names


class AnnoRegionRndSampler:
    """Random patches inside annotated regions (reference :252-796).

    Keyword-only extras: seed, device, out_dtype, out_layout ("NHWC" like the reference, or "NCHW"), flips (fused
    random H/V flip per BATCH, train.py:71-81), mean / std."""

    def __init__(self, img_anno_paths, layer: int, patch_size: int, region_intersection: float = 0.75,
                 patches_from_one_region: int = 4, region_area_influence: float = 0.5, classes: list[str] = None,
                 one_image_for_batch: bool = False, *, seed: int = 0, device="cuda", out_dtype=torch.float32, out_layout: str = "NHWC",
                 flips: bool = False, mean=None, std=None, verbose: bool = True, sparse_upload: bool = None, shard_upload=None,
                 prefetch_bytes: int = 5 << 30, prefetch_batches: int = 32, zero_copy: bool = None, zero_copy_fraction: float = 0.5,
                 resident: bool = True, invalid_polygons: str = "approximate"):
        self.img_anno_paths = img_anno_paths
        self.layer = layer
        self.patch_size = patch_size
        self.region_intersection = region_intersection
        self.patches_from_one_region = patches_from_one_region
        self.region_area_influence = region_area_influence
        self.one_image_for_batch = one_image_for_batch
        self._seed, self._device = seed, device
        self._out_dtype, self._out_layout, self._flips, self._mean, self._std = out_dtype, out_layout, flips, mean, std
        if invalid_polygons not in ("approximate", "skip"):
            raise ValueError("invalid_polygons must be 'approximate' or 'skip'")
        self.regions, self.regions_per_image, self._sizes, self._sources = _parse_annotations(
            img_anno_paths, layer=layer, classes=classes, seed=seed, device=device, verbose=verbose, invalid=invalid_polygons)
        self.classes = sorted(list(self.regions.keys()))
        annos = [[{"class": r.class_, "vertices": r.vertices} for rs in rpi.values() for r in rs] for rpi in self.regions_per_image]
        # keep the per-image dict order of the reference (class first-appearance order, then file order)
        images = []
        for j, rpi in enumerate(self.regions_per_image):
            ordered = sorted((r for rs in rpi.values() for r in rs), key=lambda r: r.region_idx)
            images.append((self._sizes[j], [{"class": r.class_, "vertices": r.vertices} for r in ordered]))
        del annos
        self._tables, self._flat_regions, names = build_tables(images, layer, region_area_influence, None, one_image_for_batch, device,
                                                               invalid=invalid_polygons)
        assert names == self.classes
        self._slides = [None] * len(img_anno_paths)
        self._prefetch_bytes, self._prefetch_batches = int(prefetch_bytes), int(prefetch_batches)   # torch_generator: features per prefetch group
        self._sparse_upload = sparse_upload    # PinnedSlide sources: upload only the tiles annotated regions can reach (None = when < 1/5 of the layer)
        self._shard_upload = shard_upload      # True / a process group: COLLECTIVE ingestion of pinned slides (slide.sharded_upload); every rank
        #                                        of the group must then iterate the sampler (first use of an image is a collective call)
        self.uploaded_bytes = 0                # bytes copied host -> device for pinned sources so far
        # Zero-copy ingestion of PINNED host slides (torch_generator): a job whose patches add up to less than `zero_copy_fraction`
        # of the slide bytes gathers them straight from host memory (the gather's bulk row copies read the mapped pinned buffer over
        # PCIe) instead of waiting for the whole layer to be uploaded; the layer is then made resident in the background, behind the
        # job's last gather. None = that cost rule, True = always while a slide is not resident, False = never.
        # resident=False: the slides are NEVER uploaded (datasets larger than HBM): every job is gathered in place, at PCIe speed.
        self._resident = bool(resident)
        if not self._resident:
            zero_copy = True
        self._zero_copy, self._zero_copy_fraction = zero_copy, float(zero_copy_fraction)
        self.zero_copy_bytes = 0               # patch bytes (ps * ps * 3 each) read in place from pinned host memory so far
        self._mapped = [None] * len(img_anno_paths)        # ops.MappedHostSlide views of pinned sources
        self._mapped_table = None
        self._upload_done = [None] * len(img_anno_paths)   # event of a background upload still in flight (resident slides wait on it)
        self._copy_stream = None
        self.ingest_events: list[dict] = []    # one record per upload: host ms of the allocation, CUDA events of copy / all-gather
        # Redraws of (class, region) after a failed region (too small / miss limit): the reference's worker loop retries without bound
        # (region_samplers.py:571-572,589-590) -- and hangs when no region can ever succeed. Here a slot gives up after max_redraw
        # redraws (early exit on success, so the bound costs nothing): 4096 makes a spurious abort in a dataset with many
        # below-threshold regions astronomically unlikely (p_fail^4096) while a dataset that cannot be sampled at all still raises.
        self.max_redraw = 4096
        self._slot_cursor = 0
        self._yield_cursor = None      # slot cursor behind the last batch a running torch_generator has handed out
        self._producer = None          # CUDA stream the gathers of torch_generator's prefetch groups run on
        self._drawer = None            # CUDA stream their coordinate draws run on
        self._slide_table = None       # ops.SlideTable over all images (multi-image datasets)
        self._fail = torch.zeros(1, dtype=torch.uint8, device=device) if torch.device(device).type == "cuda" else None
        if verbose:
            self._print_anno_stats(self.regions)

    # -- resume: the Philox draws are keyed by (seed, global slot index), so two integers are the whole sampler state ---------
    def state_dict(self) -> dict:
        """Taken while a torch_generator is running, the state is that of the last batch handed to the consumer: the batches the
        device has prefetched beyond it are drawn again after a restore (the draw is a pure function of the slot index)."""
        cur = self._slot_cursor if self._yield_cursor is None else self._yield_cursor
        return {"seed": int(self._seed), "slot_cursor": int(cur)}

    def load_state_dict(self, state: dict) -> None:
        """Continue a run: the batches drawn after this call are bit-identical to those the saved sampler would have drawn next."""
        if int(state["seed"]) != int(self._seed):
            raise ValueError(f"state was saved with seed {state['seed']}, this sampler uses seed {self._seed}")
        self._slot_cursor = int(state["slot_cursor"])
        self._yield_cursor = None

    # -- bookkeeping identical to the reference -------------------------------------------------------
    def _print_anno_stats(self, regions):
        areas = {cls: sum(i.area for i in regs) for cls, regs in regions.items()}
        print("Total area per class:")
        for cls in areas:
            print(f"\t{cls}: {round(areas[cls] / 1e9, 2)} Gpx ({round(areas[cls] / sum(areas.values()) * 100, 2)}%)")
        print(f"Approximate number of patches in dataset: {len(self)}")

    def _calc_area_weights(self, areas: list[float], area_influence: float) -> np.ndarray:
        return geometry.area_weights(areas, area_influence)

    def __len__(self):
        ps = self.patch_size * self.layer                                                          # :788-796
        return int(sum(sum(r.area for r in lst) for lst in self.regions.values()) / (ps * ps))

    def _split_chunks(self, n, k):
        q = [k] * (n // k)
        if n % k > 0:
            q.append(n % k)
        return q

    # -- device pipeline --------------------------------------------------------------------------------
    def _reachable_rects(self, j: int):
        """Pixel rectangles (y0, y1, x0, x1) of image j that a patch of this sampler can touch: origins are drawn inside a region's
        bounding box (region_samplers.py:123-124), so a patch stays within the box grown to at least patch_size + 1."""
        ps, rects = self.patch_size, []
        for regs in self.regions_per_image[j].values():
            for r in regs:
                v = np.asarray(r.vertices, dtype=np.float64)
                x0, y0 = np.floor(v.min(axis=0)) - 1
                x1, y1 = np.ceil(v.max(axis=0)) + 1
                rects.append((y0, max(y1, y0 + ps + 2), x0, max(x1, x0 + ps + 2)))
        return rects

    def _whole_pinned(self, j: int) -> bool:
        src = self._sources[j]
        return isinstance(src, PinnedSlide) and src.pinned and src.y_origin == 0 and src.rows == src.height

    def _mapped_slide(self, j: int):
        """Image j read in place from its pinned host buffer (zero-copy ingestion)."""
        if self._mapped[j] is None:
            src = self._sources[j]
            src._assert_layer(self.layer)
            self._mapped[j] = ops.MappedHostSlide(src.host, src.rows, src.width, src.pitch, self._device)
        return self._mapped[j]

    def ingest_stats(self) -> dict:
        """Milliseconds spent making slides resident so far: {"alloc_ms" (host time of the device allocations), "upload_ms" (host ->
        device copies, CUDA events), "allgather_ms" (NVLink replication of sharded uploads), "bytes"}. Synchronises on the events."""
        out = {"alloc_ms": 0.0, "upload_ms": 0.0, "allgather_ms": 0.0, "bytes": 0}
        for rec in self.ingest_events:
            out["alloc_ms"] += rec.get("alloc_ms", 0.0)
            out["bytes"] += rec.get("bytes", 0)
            for key in ("upload", "allgather"):
                if key in rec:
                    a, b = rec[key]
                    b.synchronize()
                    out[key + "_ms"] += a.elapsed_time(b)
        return out

    def upload_in_flight_bytes(self) -> int:
        """Bytes of background uploads (behind a zero-copy job) that have not completed yet."""
        n = 0
        for j, ev in enumerate(self._upload_done):
            if ev is not None and not ev.query():
                n += self._sources[j].nbytes
        return n

    def _slide(self, j: int, alloc_stream=None):
        if not self._resident:
            if not self._whole_pinned(j):
                raise ValueError("resident=False needs every slide as a PinnedSlide in page-locked host memory")
            return self._mapped_slide(j)
        if self._slides[j] is not None and self._upload_done[j] is not None:
            if self._upload_done[j].query():
                self._upload_done[j] = None
            else:
                cur = torch.cuda.current_stream(self._device)
                cur.wait_event(self._upload_done[j])                            # background upload still in flight
                self._slides[j].storage.record_stream(cur)                      # allocated on the copy stream, read on this one
        if self._slides[j] is None:
            with self._sources[j] as psim:
                whole_pinned = isinstance(psim, PinnedSlide) and psim.y_origin == 0 and psim.rows == psim.height
                if self._shard_upload is not None and whole_pinned:
                    # data-parallel ranks with the same host slide: 1/world of the rows over each rank's PCIe link + one NVLink all-gather
                    psim._assert_layer(self.layer)
                    group = None if self._shard_upload is True else self._shard_upload
                    rec = {}
                    self._slides[j], n = sharded_upload(psim, self._device, group, stats=rec, alloc_stream=alloc_stream)
                    rec["bytes"] = n
                    self.ingest_events.append(rec)
                    self.uploaded_bytes += n
                    return self._slides[j]
                sparse = self._sparse_upload
                if sparse is not False and isinstance(psim, PinnedSlide) and psim.y_origin == 0 and psim.rows == psim.height:
                    rects = self._reachable_rects(j)
                    if sparse is None:
                        # measured on a B200 box: the strided (2-D) copies of dh_upload_rects run at ~13 GB/s, one contiguous copy of
                        # the layer at ~53 GB/s -> the sparse path pays off when less than ~1/5 of the layer is reachable
                        spans = tile_spans(rects, psim.height, psim.width, psim.pitch)
                        frac = float(((spans[:, 1] - spans[:, 0]) * (spans[:, 3] - spans[:, 2])).sum()) / max(psim.nbytes, 1) if len(spans) else 0.0
                        sparse = frac < 0.2
                if sparse is True and isinstance(psim, PinnedSlide) and psim.y_origin == 0 and psim.rows == psim.height:
                    # host-resident slide: only the tiles a region can reach travel over PCIe
                    psim._assert_layer(self.layer)
                    self._slides[j], n = upload_rects(psim, rects, self._device)
                    self.uploaded_bytes += n
                elif isinstance(psim, PinnedSlide):
                    import time

                    psim._assert_layer(self.layer)
                    rec = {}
                    t0 = time.perf_counter()
                    storage = alloc_bytes_on(alloc_stream, psim.rows * psim.pitch, self._device)
                    rec["alloc_ms"] = 1e3 * (time.perf_counter() - t0)
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    storage.copy_(psim.host[: psim.rows * psim.pitch], non_blocking=True)
                    b.record()
                    rec["upload"], rec["bytes"] = (a, b), psim.rows * psim.pitch
                    self.ingest_events.append(rec)
                    self._slides[j] = ops.DeviceSlide(storage, psim.rows, psim.width, psim.pitch)
                    self.uploaded_bytes += psim.rows * psim.pitch
                else:
                    self._slides[j] = layer_to_device(psim, self.layer, self._device)
        return self._slides[j]

    def _background_upload(self, after_event=None):
        """Make every pinned slide that is not resident yet resident on the copy stream, behind `after_event` (the last gather that
        reads host memory in place, so the two do not share the PCIe link); later gathers wait on the upload's event (_slide)."""
        if self._copy_stream is None:
            self._copy_stream = _shared_stream(self._device, "copy")
        home = torch.cuda.current_stream(self._device)        # the slide buffer comes from the caller's allocator pool (alloc_bytes_on)
        with torch.cuda.stream(self._copy_stream):
            if after_event is not None:
                self._copy_stream.wait_event(after_event)
            for j in range(len(self._slides)):
                if self._slides[j] is None and self._whole_pinned(j):
                    sl = self._slide(j, alloc_stream=home)
                    sl.storage.record_stream(self._copy_stream)
                    ev = torch.cuda.Event()
                    ev.record(self._copy_stream)
                    self._upload_done[j] = ev

    def _check_failures(self):
        if self._fail is not None and int(self._fail.item()) != 0:
            raise RuntimeError("region sampling failed for some slots after max_redraw redraws "
                               "(regions too small for the patch size / intersection, or miss limit reached)")

    def _sample_raw(self, n_slots: int, slots_per_image_draw: int, cls_idx: int = None, slot_offset: int = None):
        if cls_idx is not None and not (0 <= cls_idx < len(self.classes)):
            raise ValueError(f"cls_idx {cls_idx} out of range")
        if slot_offset is None:
            slot_offset = self._slot_cursor
            self._slot_cursor += n_slots
        ps = self.patch_size
        return ops.region_sample(
            self._tables.struct, n_slots, self.patches_from_one_region, ps, ps * ps * self.region_intersection, miss_limit=500,
            max_redraw=self.max_redraw, fixed_class=-1 if cls_idx is None else cls_idx, slots_per_table_draw=max(slots_per_image_draw, 1),
            seed=self._seed, slot_offset=slot_offset, device=self._device)

    def sample_coords(self, n_slots: int, slots_per_image_draw: int, cls_idx: int = None, slot_offset: int = None):
        """(coords int32 [n,2], labels int64 [n], images int32 [n]) on the device for one worker-sized chunk (:525-591)."""
        coords, labels, images, status = self._sample_raw(n_slots, slots_per_image_draw, cls_idx, slot_offset)
        torch.maximum(self._fail, status.max().reshape(1), out=self._fail)
        return coords, labels, images

    def _gather(self, coords, images, dtype, layout, flip=None, scale255=True, mapped: bool = False):
        """`mapped`: read the patches in place from the pinned host slides (zero-copy ingestion) instead of the resident copies."""
        ps = self.patch_size
        get = self._mapped_slide if mapped else self._slide
        if mapped:
            self.zero_copy_bytes += int(coords.shape[0]) * ps * ps * 3
        if len(self._slides) == 1:
            return ops.gather_normalize(get(0), coords, ps, dtype=dtype, layout=layout, scale255=scale255, mean=self._mean,
                                        std=self._std, flip=flip)
        if dtype != torch.uint8 and ps % (4 if dtype == torch.float32 else 8) == 0:
            # one launch over all source slides (descriptor table in HBM); every image of the dataset is made resident on first use
            if mapped:
                if self._mapped_table is None:
                    self._mapped_table = ops.SlideTable([get(j) for j in range(len(self._slides))], device=self._device)
                table = self._mapped_table
            else:
                if self._slide_table is None or any(e is not None for e in self._upload_done):
                    slides = [self._slide(j) for j in range(len(self._slides))]      # waits on background uploads still in flight
                    if self._slide_table is None:
                        self._slide_table = ops.SlideTable(slides)
                table = self._slide_table
            return ops.gather_normalize_multi(table, images, coords, ps, dtype=dtype, layout=layout, scale255=scale255,
                                              mean=self._mean, std=self._std, flip=flip)
        shape = (len(coords), ps, ps, 3) if layout == "NHWC" else (len(coords), 3, ps, ps)
        out = torch.empty(shape, dtype=dtype, device=coords.device)
        for j in torch.unique(images).tolist():                                                  # one gather per source slide
            idx = torch.nonzero(images == j).reshape(-1).to(torch.int32)
            ops.gather_normalize(get(j), coords[idx.long()].contiguous(), ps, dtype=dtype, layout=layout, scale255=scale255,
                                 mean=self._mean, std=self._std, flip=None if flip is None else flip[idx.long()].contiguous(), out=out,
                                 out_index=idx)
        return out

    def _batch_flip(self, batch_global_index: int, B: int):
        """One H and one V coin per batch (torchvision flips the whole [B,3,H,W] tensor, train.py:71-81)."""
        if not self._flips:
            return None
        g = torch.Generator().manual_seed((self._seed << 20) ^ batch_global_index)
        bits = int(torch.randint(0, 4, (1,), generator=g).item())
        return torch.full((B,), bits, dtype=torch.uint8, device=self._device)

    def torch_generator(self, batch_size: int, n_batches: int, batches_per_worker: int = 2, transforms: callable = None,
                        max_workers: int = None, cls_idx: int = None) -> Iterator[tuple[torch.Tensor, torch.Tensor, torch.Tensor]]:
        """Reference :685-738: yields (features [B,ps,ps,3] float32 in [0,1], labels int64 [B], coords float32 [B,2] (y,x))."""
        chunk = batch_size * batches_per_worker
        if cls_idx is not None and not (0 <= cls_idx < len(self.classes)):
            raise ValueError(f"cls_idx {cls_idx} out of range")
        # Coordinates are counter-based (Philox keyed by the global slot index), so several worker-sized chunks can be drawn by
        # ONE launch with identical results -- as long as the groups of k slots do not straddle a chunk boundary. The features of
        # all prefetched batches are then written by ONE gather launch (short launches cannot fill HBM: profiles/r01_gather.md);
        # every yielded batch is a contiguous slice of that buffer. Prefetch depth: <= 32 batches and <= ~5 GB of features
        # (measured through the bench's consumer loop: 6.7-7.0 M patches/s with 16-batch groups, 7.55 M with 32; profiles/r01_gather.md).
        # The launches run on a PRODUCER STREAM, one prefetch group ahead of the batches being yielded: the consumer's own work on
        # the current stream (its CNN step, its read-back of labels) overlaps with the sampling of the next group instead of
        # queueing behind it -- the role the reference gives to its ProcessPoolExecutor workers (:721-738).
        ps = self.patch_size
        per_batch = batch_size * ps * ps * 3 * torch.empty((), dtype=self._out_dtype).element_size()
        ahead = batches_per_worker
        if chunk % self.patches_from_one_region == 0:
            ahead = max(1, min(self._prefetch_batches, self._prefetch_bytes // max(per_batch, 1)) // batches_per_worker) * batches_per_worker
        groups = self._split_chunks(n_batches, ahead)
        on_gpu = torch.device(self._device).type == "cuda"
        cur = torch.cuda.current_stream(self._device) if on_gpu else None
        # zero-copy ingestion: slides that are still only in pinned host memory are read in place when this job touches less of them
        # than an upload would move (cost rule above); the upload then runs in the background behind the job's last gather
        pending_up = [j for j in range(len(self._slides)) if self._slides[j] is None or not self._resident]
        mapped = bool(on_gpu and pending_up and self._zero_copy is not False and all(self._whole_pinned(j) for j in pending_up)
                      and (len(pending_up) == len(self._slides) or len(self._slides) == 1)
                      and (self._sparse_upload is not True or self._zero_copy is True))
        if mapped and self._zero_copy is None:
            job = n_batches * batch_size * ps * ps * 3
            upload = sum(self._sources[j].nbytes for j in pending_up)
            if self._shard_upload is not None:
                # collective ingestion moves only 1/world of the layer over this rank's PCIe link (the rest arrives over NVLink), while
                # in-place gathers of all ranks share the host's links: measured at N = 4, 20 batches: 604 k patches/s in place vs the
                # sharded upload's 1/4 of 3.2 GB per rank (profiles/r02_e2e.md)
                import torch.distributed as dist

                upload //= max(1, dist.get_world_size(None if self._shard_upload is True else self._shard_upload))
            mapped = job <= self._zero_copy_fraction * upload
        if on_gpu and self._producer is None:
            self._producer = _shared_stream(self._device, "producer")
            self._drawer = _shared_stream(self._device, "drawer")
            self._producer.wait_stream(cur)                 # tables (and anything else set up on the caller's stream) are complete
            self._drawer.wait_stream(cur)

        def launch(nb):
            """Enqueue one prefetch group; returns its tensors and the event that marks them ready. Coordinates are drawn on their
            own stream: the draw of group g+1 does not depend on the gather of group g, so it runs next to it instead of
            extending the producer stream's critical path (measured +4.6 % patches/s, profiles/r01_gather.md)."""
            with torch.cuda.stream(self._drawer):
                first_slot = self._slot_cursor
                coords, labels, images, status = self._sample_raw(batch_size * nb, chunk, cls_idx)
                group = {"labels": labels, "coords": coords.to(torch.float32), "fail": status.max()}
                drawn = torch.cuda.Event()
                drawn.record(self._drawer)
            with torch.cuda.stream(self._producer):
                self._producer.wait_event(drawn)
                coords.record_stream(self._producer)
                images.record_stream(self._producer)
                flip = None
                if self._flips:
                    flip = torch.cat([self._batch_flip(first_slot // max(batch_size, 1) + i, batch_size) for i in range(nb)])
                group["features"] = self._gather(coords, images, self._out_dtype, self._out_layout, flip, mapped=mapped)
                group["first_slot"] = first_slot
                ready = torch.cuda.Event()
                ready.record(self._producer)
            return group, ready

        def launch_next(gi):
            if gi >= len(groups):
                return None
            out = launch(groups[gi])
            if mapped and self._resident and gi == len(groups) - 1:
                self._background_upload(after_event=out[1])
            return out

        pending = launch_next(0)
        for gi, nb in enumerate(groups):
            group, ready = pending
            pending = launch_next(gi + 1)
            cur = torch.cuda.current_stream(self._device)    # the stream the caller consumes this group on
            cur.wait_event(ready)
            first_slot = group.pop("first_slot")
            for t in group.values():
                t.record_stream(cur)                         # allocated on the producer stream, used on the caller's
            if int(group["fail"].item()) != 0:
                raise RuntimeError("region sampling failed for some slots after max_redraw redraws "
                                   "(regions too small for the patch size / intersection, or miss limit reached)")
            feats = group["features"]
            fv = feats.view((nb, batch_size) + tuple(feats.shape[1:])).unbind(0)
            lv = group["labels"].view(nb, batch_size).unbind(0)
            cv = group["coords"].view(nb, batch_size, 2).unbind(0)
            for i, (f, l, c) in enumerate(zip(fv, lv, cv)):
                if transforms is not None:
                    f = transforms(f)
                self._yield_cursor = first_slot + (i + 1) * batch_size     # state_dict(): resume behind the batch being handed out
                yield f, l, c
        self._yield_cursor = None

    def structs_generator(self, batch_size: int, n_batches: int, batches_per_worker: int = 2, max_workers: int = None,
                          cls_idx: int = None) -> Iterator[list[tuple[Patch, int]]]:
        """Reference :641-683: yields list[(Patch, class_index)] per batch, Patch.data uint8 numpy [ps,ps,3]."""
        for nb in self._split_chunks(n_batches, batches_per_worker):
            coords, labels, images = self.sample_coords(batch_size * nb, batch_size * batches_per_worker, cls_idx)
            self._check_failures()
            raw = self._gather(coords, images, torch.uint8, "NHWC").cpu().numpy()
            yx, lab = coords.cpu().numpy().tolist(), labels.cpu().numpy().tolist()
            lst = [(Patch(self.layer, pos_x=x, pos_y=y, patch_size=self.patch_size, data=raw[i]), int(lab[i])) for i, (y, x) in enumerate(yx)]
            for i in range(0, len(lst), batch_size):
                yield lst[i : i + batch_size]

    def torch_iterable_dataset(self) -> IterableDataset:
        """Reference :740-786: infinite per-sample stream (features [ps,ps,3], label, coords (y, x))."""
        outer = self

        class CustomIterableDataset(IterableDataset):
            def __iter__(self):
                while True:
                    for f, l, c in outer.torch_generator(batch_size=64, n_batches=2, batches_per_worker=2):
                        for i in range(f.shape[0]):
                            yield f[i], l[i], c[i]

        return CustomIterableDataset()


class AnnoRegionDenseSampler:
    """Dense grid inside every annotated region (reference :799-871)."""

    def __init__(self, img_anno_paths, layer: int, patch_size: int, stride: int, region_intersection: float = 0.75,
                 classes: list[str] = None, *, device="cuda", verbose: bool = True, invalid_polygons: str = "approximate"):
        self.img_anno_paths = img_anno_paths
        self.layer = layer
        self.patch_size = patch_size
        self.stride = stride
        self.region_intersection = region_intersection
        self._device = device
        self.regions, _, self._sizes, self._sources = _parse_annotations(img_anno_paths, layer=layer, classes=classes, device=device,
                                                                       verbose=verbose, invalid=invalid_polygons)
        self.classes = sorted(list(self.regions.keys()))
        self._slides = [None] * len(img_anno_paths)

    def _slide(self, j: int):
        if self._slides[j] is None:
            with self._sources[j] as psim:
                self._slides[j] = layer_to_device(psim, self.layer, self._device)
        return self._slides[j]

    def region_batches(self, dtype=torch.float32, layout: str = "NHWC"):
        """Device-side iteration: (features of ALL accepted patches of one region, int32 coords, class index)."""
        for cls_idx, cls in enumerate(self.classes):
            for region in self.regions[cls]:
                coords = region._extract_patch_coords_dense_device(self.patch_size, self.stride, self.region_intersection)
                if len(coords) == 0:
                    continue
                feats = ops.gather_normalize(self._slide(region._region.image), coords, self.patch_size, dtype=dtype, layout=layout)
                yield feats, coords, cls_idx

    def _patches_one_region(self, region: RegionAnnotation) -> list[Patch]:
        coords = region._extract_patch_coords_dense_device(self.patch_size, self.stride, self.region_intersection)
        if len(coords) == 0:
            return []
        raw = ops.gather_normalize(self._slide(region._region.image), coords, self.patch_size, dtype=torch.uint8).cpu().numpy()
        return [Patch(self.layer, pos_x=x, pos_y=y, patch_size=self.patch_size, data=raw[i]) for i, (y, x) in enumerate(coords.cpu().numpy().tolist())]

    def structs_generator(self) -> Iterator[tuple[Patch, int]]:
        for cls_idx, cls in enumerate(self.classes):
            for region in self.regions[cls]:
                for p in self._patches_one_region(region):
                    yield p, cls_idx


def extract_and_save_subset(img_anno_paths, out_folder: Path, patch_size: int, layer: int, patches_per_class: int, intersection=0.95):
    """Reference :874-909 (JPEG dump of test patches; IO-bound tooling around the path)."""
    from PIL import Image

    sampler = AnnoRegionRndSampler(img_anno_paths=img_anno_paths, layer=layer, patch_size=patch_size, region_intersection=intersection,
                                   region_area_influence=0, patches_from_one_region=1)
    batch_size = 4
    for cls_idx, cls in enumerate(sampler.classes):
        (out_folder / str(cls_idx)).mkdir(parents=True, exist_ok=True)
        g = sampler.structs_generator(batch_size=batch_size, n_batches=patches_per_class // batch_size, cls_idx=cls_idx)
        count = 0
        for batch in g:
            for patch, _ in batch:
                Image.fromarray(patch.data).save(out_folder / str(cls_idx) / f"{count}.jpg")
                count += 1
