"""Multi-GPU (NCCL) checks of the time-sharded operators: needs at least two CUDA devices on the box.

Runs tests/multi_gpu/sharded_checks.py under torchrun with one process per GPU (2, and 4 when available) and expects
every rank to find its sharded results identical to the unsharded ones.
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_devices():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_equals_unsharded_nccl(world):
    if _n_devices() < world:
        pytest.skip(f"needs {world} CUDA devices")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29610 + world),
           os.path.join(ROOT, "tests", "multi_gpu", "sharded_checks.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "sharded checks passed" in r.stdout
