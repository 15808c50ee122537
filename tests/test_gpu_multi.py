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


def test_draw_and_tile_partitions_over_nccl():
    """Partition A (draws + one NCCL sum on the device, rank 0 alone downloads) and partition B (tiles, compact per-rank
    shares, no collective) against the one-device path and the oracle (tools/check_partitions.py)."""
    n_gpus = _gpus()
    if n_gpus < 2:
        pytest.skip("needs two GPUs")
    for world, shape in ((2, ("200", "3000", "5", "37")), (min(n_gpus, 8), ("120", "2100", "3", "19"))):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
               "127.0.0.1", "--master-port", "29613", os.path.join(ROOT, "tools", "check_partitions.py"), *shape]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
        assert "partitions ok: True" in r.stdout
