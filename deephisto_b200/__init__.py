"""deephisto_b200: B200-native (sm_100a) implementation of the DeepHisto patch-sampling and patched-prediction
hot path, behind the reference's `patch_samplers` iterator API and `examples.predict_full_patched`."""

__version__ = "0.1.0"


def install_dropin() -> None:
    """Register this package's modules under the reference's top-level names so that unmodified reference-style
    scripts (`from patch_samplers.full_samplers import ...`, `from examples.predict_full_patched import ...`) resolve here."""
    import importlib
    import sys

    for name in ("patch_samplers", "patch_samplers.full_samplers", "patch_samplers.region_samplers", "examples",
                 "examples.predict_full_patched", "anno", "anno.utils"):
        try:
            sys.modules[name] = importlib.import_module(f"deephisto_b200.{name}")
        except ModuleNotFoundError:
            pass
