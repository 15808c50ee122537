"""Multi-GPU checks that need real NCCL between two B200s: skipped on one-GPU boxes (the column-shard logic itself is
covered on one GPU by test_gpu_parity.py::test_column_sharded_estimation_matches_one_device)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _gpus():
    import torch
    return torch.cuda.device_count()


def test_column_sharded_estimation_over_nccl():
    if _gpus() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "tools", "check_sharded_estimation.py"), "300", "2000", "5"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "replicated results identical on all ranks: True" in r.stdout
