"""Import the UNMODIFIED reference (/root/reference) with stub third-party modules -- only possible in the
build container (the GPU box has no /root/reference); used by oracle/make_golden.py and by the CPU tests
that pin the oracle when the reference tree is present."""

import importlib
import sys
from pathlib import Path

REFERENCE = Path("/root/reference")
STUBS = Path(__file__).resolve().parent / "refstubs"


def available() -> bool:
    return (REFERENCE / "patch_samplers" / "full_samplers.py").exists()


class reference_modules:
    """Context manager: puts the stubs and the reference first on sys.path, imports the requested reference
    modules, and restores sys.path / sys.modules afterwards (the product has same-named sub-packages)."""

    NAMES = ("patch_samplers", "examples", "anno", "models", "utils", "psimage", "shapely", "matplotlib", "distinctipy")

    def __init__(self, *modules):
        self.modules = modules

    def __enter__(self):
        if not available():
            raise RuntimeError("reference tree not present")
        self._path = list(sys.path)
        self._saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in self.NAMES}
        for k in self._saved:
            del sys.modules[k]
        sys.path[:0] = [str(STUBS), str(REFERENCE)]
        mods = [importlib.import_module(m) for m in self.modules]
        return mods[0] if len(mods) == 1 else mods

    def __exit__(self, *exc):
        for k in [k for k in sys.modules if k.split(".")[0] in self.NAMES]:
            del sys.modules[k]
        sys.modules.update(self._saved)
        sys.path[:] = self._path
        return False


def register_slide(name: str, array):
    """Make `array` openable as PSImage(name) inside reference_modules."""
    sys.path.insert(0, str(STUBS))
    try:
        from psimage.core import image as im
        im.register(name, array)
        reg = im.REGISTRY
    finally:
        sys.path.pop(0)
    return reg
