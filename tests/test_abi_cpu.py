"""The C-ABI library loads without a GPU and exports every symbol include/deephisto_b200.h declares."""

import re
import subprocess
from pathlib import Path

import pytest

from deephisto_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "deephisto_b200.h"


def declared_symbols():
    return sorted(set(re.findall(r"^DH_API [\w\s\*]+?\b(dh_\w+)\(", HEADER.read_text(), flags=re.M)))


def test_header_symbols_all_bound_and_exported():
    syms = declared_symbols()
    assert len(syms) >= 18
    assert sorted(_lib.SIGNATURES) == syms
    lib = _lib.load()  # no GPU needed
    for s in syms:
        assert hasattr(lib, s), s
    exported = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    names = {line.split()[-1] for line in exported.splitlines() if " T " in line}
    assert set(syms) <= names
    assert {n for n in names if n.startswith("dh_")} == set(syms), "exported dh_* symbol missing from the header"


def test_host_only_entry_points():
    lib = _lib.load()
    assert lib.dh_version() == 100
    from deephisto_b200 import ops

    assert ops.dense_count(8192, 8192, 224, 224, 16) == (1369, 1376)
    assert ops.dense_count(40000, 40000, 224, 112, 64) == (127449, 127488)
    assert ops.dense_count(100000, 100000, 224, 112, 64) == (795664, 795712)
    with pytest.raises(ValueError):
        ops.dense_count(100, 100, 224, 112, 64)
    assert "smaller than patch" in _lib.last_error()
    assert lib.dh_cover_scratch_words(2500, 2500) > 2500 * 2500 // 32


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.DeepHistoError):
        _lib.require_device()
    from deephisto_b200 import ops

    with pytest.raises(_lib.DeepHistoError):
        ops.gather_normalize(None, torch.zeros((1, 2), dtype=torch.int32), 8)
