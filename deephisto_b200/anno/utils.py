"""The slice of the reference's `anno/utils.py` that the hot path touches: `AnnoClass` (:19-42) and
`AnnoDescription` (:45-140), used by `examples.predict_full_patched` for `len(anno.anno_classes)` and the class
colours (predict_full_patched.py:43,94-95,142-151). Palette generation (distinctipy) and the PIL / matplotlib
visualiser (:143-408) are presentation code outside the path."""

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

    def __str__(self) -> str:
        label = self.label
        if self.alternate_labels:
            label += " (" + ", ".join(self.alternate_labels) + ")"
        description = ", " + self.description if self.description else ""
        return f"AnnoClass [{self.id}, {label}, {self.color}{description}]"

    @property
    def label_full(self) -> str:
        if not self.alternate_labels:
            return self.label
        return self.label + " (" + ", ".join(self.alternate_labels) + ")"


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
