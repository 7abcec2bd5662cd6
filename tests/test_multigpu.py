"""Row-band sharded prediction over NCCL on >= 2 real GPUs (skipped on a single-GPU box; the host logic and the exchange
step are covered on CPU by tests/test_bands_cpu.py)."""

import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_banded_predict_equals_single_gpu_over_nccl():
    n = min(torch.cuda.device_count(), 4)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1", "--master-port", str(port),
           str(ROOT / "tests" / "helpers" / "multigpu_predict_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, OMP_NUM_THREADS="4"))
    assert out.returncode == 0, (out.stdout + out.stderr)[-4000:]
    for r in range(n):
        assert f"MULTIGPU OK rank {r}/{n}" in out.stdout
