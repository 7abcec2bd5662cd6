from dataclasses import dataclass
from typing import Any


@dataclass
class Patch:
    layer: int
    pos_x: int
    pos_y: int
    patch_size: int
    data: Any = None
