"""The reference's `anno/utils.py` over the device path: `AnnoClass` (:19-42) and `AnnoDescription` (:45-140), used by
`examples.predict_full_patched` for `len(anno.anno_classes)` and the class colours (predict_full_patched.py:43,94-95,142-151);
`AnnoVisualizerParams`, `PatchVisAccent` and `AnnoVisualizer` (:193-408) with the polygon fill done by
`dh_rasterize_polygons` (pixel-centre even-odd rule, painter's order) on a thumbnail averaged on the device -- the PIL fill
of the reference is edge-inclusive, so this is a visual equivalent, not a parity target. Palette generation
(distinctipy, :143-190) is replaced by a deterministic hue walk."""

from __future__ import annotations

import json
from dataclasses import dataclass
from pathlib import Path
from typing import Iterable


@dataclass
class AnnoClass:
    id: int
    label: str
    alternate_labels: tuple = ()
    description: str = None
    color: tuple = None

    @property
    def label_full(self) -> str:
        """Main label followed by its alternatives in parentheses (same text as the reference's property, anno/utils.py:37-42)."""
        alts = ", ".join(self.alternate_labels or ())
        return f"{self.label} ({alts})" if alts else self.label

    def __str__(self) -> str:
        """Same text as the reference's AnnoClass.__str__ (anno/utils.py:30-35)."""
        fields = [str(self.id), self.label_full, str(self.color)]
        if self.description:
            fields.append(self.description)
        return "AnnoClass [" + ", ".join(fields) + "]"


def _spread_colors(n: int) -> list[tuple[int, int, int]]:
    """Deterministic, well separated colours (golden-ratio hue walk); stands in for distinctipy (anno/utils.py:143-190)."""
    import colorsys

    out = []
    for i in range(n):
        r, g, b = colorsys.hsv_to_rgb((0.11 + 0.61803398875 * i) % 1.0, 0.75, 0.95)
        out.append((int(r * 255), int(g * 255), int(b * 255)))
    return out


class AnnoDescription:
    """Set of annotation classes (reference :45-140)."""

    def __init__(self, _anno_classes) -> None:
        self.anno_classes = _anno_classes
        self.anno_classes_dict = {c.label: c for c in _anno_classes}
        for c in _anno_classes:
            for alt in c.alternate_labels or ():
                self.anno_classes_dict[alt] = c

    @classmethod
    def with_known_colors(cls, labels_with_color: dict) -> "AnnoDescription":
        return AnnoDescription([AnnoClass(id=i, label=lbl, color=color) for i, (lbl, color) in enumerate(labels_with_color.items())])

    @classmethod
    def with_auto_colors(cls, labels: Iterable[str]) -> "AnnoDescription":
        labels = list(labels)
        colors = _spread_colors(len(labels))
        return AnnoDescription([AnnoClass(id=i, label=lbl, color=colors[i]) for i, lbl in enumerate(labels)])

    @classmethod
    def auto_from_files(cls, path: Path) -> "AnnoDescription":
        path = Path(path)
        files = [f for f in path.iterdir() if f.suffix == ".json"] if path.is_dir() else ([path] if path.suffix == ".json" else [])
        if not files:
            raise RuntimeError("No annotation files found")
        labels = set()
        for f in files:
            with f.open("r") as fh:
                for anno in json.load(fh):
                    if isinstance(anno, dict):
                        labels.add(anno["class"])
        return cls.with_auto_colors(sorted(labels))

    def color_by_label(self, label: str):
        return self.anno_classes_dict[label].color


@dataclass
class AnnoVisualizerParams:
    """Parameters of the annotation visualisation (reference :193-227)."""

    fill: bool
    fill_transparency: float
    line_width: int
    show_legend: bool
    legend_placement: str
    legend_size: int

    @classmethod
    def default(cls) -> "AnnoVisualizerParams":
        return AnnoVisualizerParams(fill=True, fill_transparency=0.3, line_width=2, show_legend=True, legend_placement="TR", legend_size=20)

    @classmethod
    def no_legend(cls) -> "AnnoVisualizerParams":
        return AnnoVisualizerParams(fill=True, fill_transparency=0.3, line_width=2, show_legend=False, legend_placement=None, legend_size=None)


@dataclass
class PatchVisAccent:
    """A patch to highlight on the preview (reference :230-246)."""

    layer: int
    size: int
    x: int
    y: int
    label: str = None

    @classmethod
    def parse(cls, code_str: str, layer: int, patch_s: int) -> "PatchVisAccent":
        s = code_str.split("_")                       # "r28_LP_7_x17311_y14066"
        return PatchVisAccent(layer=layer, size=patch_s, x=int(s[3][1:]), y=int(s[4][1:]), label=s[1])


class AnnoVisualizer:
    """Preview of a slide with its polygon annotations (reference :249-408): area-averaged thumbnail, polygons filled with the class
    colour at `fill_transparency`, opaque outline of `line_width` pixels, later polygons over earlier ones, optional patch accents
    and legend. Thumbnail, fill and blend run on the device; the result is a PIL image like the reference's."""

    def __init__(self, anno_description: AnnoDescription, vis_params: AnnoVisualizerParams = None, device="cuda") -> None:
        self.anno_description = anno_description
        self.vis_params = vis_params if vis_params is not None else AnnoVisualizerParams.default()
        self._device = device

    @staticmethod
    def _downscale(h: int, w: int, scale, max_side, auto_downscale: bool, limit: int = 16384) -> int:
        if scale is not None:
            if not (0 < scale <= 1):
                raise ValueError("scale must be in (0, 1]")
            d = max(1, round(1 / scale))
        elif max_side is not None:
            d = max(1, -(-max(h, w) // int(max_side)))
        else:
            d = 1
        if max(h, w) // d > limit:
            if not auto_downscale:
                raise RuntimeError(f"preview of {h // d} x {w // d} pixels is too big; pass scale / max_side or auto_downscale=True")
            d = -(-max(h, w) // limit)
        return d

    def visualize_device(self, psimage, polygon_annotations, scale: float = None, max_side: int = None, auto_downscale: bool = False,
                         patch_accents=None):
        """uint8 [h, w, 3] device tensor of the preview (no legend)."""
        import numpy as np
        import torch

        from .. import geometry, ops
        from ..slide import layer_to_device, open_slide

        vp = self.vis_params
        with open_slide(psimage) as src:
            slide = layer_to_device(src, 1, self._device)
        d = self._downscale(slide.H, slide.W, scale, max_side, auto_downscale)
        mh, mw = slide.H // d, slide.W // d
        dummy = torch.zeros((mh, mw), dtype=torch.uint8, device=slide.device)
        lut = torch.zeros((256, 3), dtype=torch.uint8, device=slide.device)
        _, thumb, _ = ops.colorize_overlay(dummy, lut, slide, d, want_mask=False, want_thumb=True)   # integer area average, one pass
        # polygons (and accents, as squares drawn after them) -> one label map, last polygon on top
        polys, colors, alphas = [], [], []
        fill_a = int(255 * vp.fill_transparency) if vp.fill else 0
        for lbl, poly in polygon_annotations:
            polys.append(np.asarray(poly, dtype=np.float64).reshape(-1, 2))
            colors.append(self.anno_description.color_by_label(lbl))
            alphas.append(fill_a)
        n_outlined_wide = len(polys)
        for pa in patch_accents or ():
            c = self.anno_description.color_by_label(pa.label)
            x, y, sz = pa.layer * pa.x, pa.layer * pa.y, pa.layer * pa.size
            polys.append(np.asarray([[x, y], [x + sz, y], [x + sz, y + sz], [x, y + sz]], dtype=np.float64))
            colors.append((min(255, c[0] + 20), max(0, c[1] - 10), min(255, c[2] + 10)))
            alphas.append(min(255, fill_a + 80))
        if not polys:
            return thumb
        edges = [geometry.build_edges(p) for p in polys]
        off = np.zeros(len(edges) + 1, np.int32)
        off[1:] = np.cumsum([len(e) for e in edges])
        bbox = np.asarray([geometry.polygon_bounds(p) for p in polys], dtype=np.float64).reshape(-1)
        cat = np.concatenate(edges).reshape(-1) if off[-1] else np.zeros(8, np.float64)
        dev = slide.device
        label = ops.rasterize_polygons(torch.from_numpy(cat).to(dev), torch.from_numpy(off).to(dev), torch.from_numpy(bbox).to(dev),
                                       float(d), mh, mw).long()
        col = torch.tensor([(0, 0, 0)] + list(colors), dtype=torch.float32, device=dev)
        alpha = torch.tensor([0] + alphas, dtype=torch.float32, device=dev)
        a = alpha[label]
        # outline: a labelled pixel with a different label within `width` pixels (4-neighbourhood, inside the polygon)
        def outline(width: int, first: int, last: int):
            sel = (label > first) & (label <= last)
            edge = torch.zeros_like(sel)
            for k in range(1, max(int(width), 0) + 1):
                for dy, dx in ((k, 0), (-k, 0), (0, k), (0, -k)):
                    sh = torch.roll(label, shifts=(dy, dx), dims=(0, 1))
                    diff = sh != label
                    if dy > 0: diff[:dy] = True                     # noqa: E701  (the slide border closes the outline)
                    if dy < 0: diff[dy:] = True                     # noqa: E701
                    if dx > 0: diff[:, :dx] = True                  # noqa: E701
                    if dx < 0: diff[:, dx:] = True                  # noqa: E701
                    edge |= diff
            return sel & edge

        a = torch.where(outline(vp.line_width, 0, n_outlined_wide), torch.full_like(a, 255.0), a)
        if len(polys) > n_outlined_wide:
            a = torch.where(outline(1, n_outlined_wide, len(polys)), torch.full_like(a, 255.0), a)
        a = (a / 255.0).unsqueeze(-1)
        out = thumb.float() * (1 - a) + col[label] * a                 # Image.alpha_composite of the overlay over the thumbnail
        return out.round().clamp_(0, 255).to(torch.uint8)

    def visualize(self, psimage, polygon_annotations, scale: float = None, max_side: int = None, auto_downscale=False, patch_accents=None):
        """PIL image with the annotations drawn (same signature as the reference, :260-332)."""
        from PIL import Image

        img = Image.fromarray(self.visualize_device(psimage, polygon_annotations, scale, max_side, auto_downscale, patch_accents).cpu().numpy())
        if self.vis_params.show_legend:
            img = self._add_legend(img)
        return img.convert("RGB")

    def _add_legend(self, img):
        """Colour swatches + full labels in a corner (the reference renders this through matplotlib, :371-408)."""
        from PIL import ImageDraw, ImageFont

        vp = self.vis_params
        size = int(vp.legend_size or 20)
        try:
            font = ImageFont.load_default(size=size)
        except TypeError:
            font = ImageFont.load_default()
        draw = ImageDraw.Draw(img)
        rows = [(c.color, c.label_full) for c in self.anno_description.anno_classes]
        pad, sw = size // 2, size
        tw = max((draw.textlength(lbl, font=font) for _, lbl in rows), default=0)
        bw, bh = int(3 * pad + sw + tw), int(pad + len(rows) * (size + pad))
        place = (vp.legend_placement or "TR").upper()
        x0 = pad if place.endswith("L") else max(pad, img.width - bw - pad)
        y0 = pad if place.startswith("T") else max(pad, img.height - bh - pad)
        draw.rectangle([x0, y0, x0 + bw, y0 + bh], fill=(255, 255, 255), outline=(128, 128, 128))
        for i, (color, lbl) in enumerate(rows):
            y = y0 + pad + i * (size + pad)
            draw.rectangle([x0 + pad, y, x0 + pad + sw, y + size], fill=tuple(color))
            draw.text((x0 + 2 * pad + sw, y), lbl, fill=(0, 0, 0), font=font)
        return img
