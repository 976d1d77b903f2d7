"""Multi-GPU parity, launched the way the driver launches bench.py: one process per GPU under
torch.distributed.run over NCCL.  Skipped on boxes with fewer than 2 GPUs (the world-size-2 host logic is
covered on CPU / gloo by tests/test_sharding_gloo.py)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_fused_gather_equals_nccl_and_single_gpu(world):
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29650 + world),
           os.path.join(ROOT, "tests", "_multigpu_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "MULTIGPU OK" in res.stdout
