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


@pytest.mark.parametrize("ndev,MC", [(2, 37), (2, 1), (0, 50)])
def test_multi_device_sampler_in_one_process(ndev, MC):
    """fable_multi_sampler: CPPFABLESampler with the draws sharded over the devices of one process and the accumulators
    summed over NVLink peer memory -- samples bit-identical to one device, sums to rounding (ndev = 0: every visible
    device; MC = 1: a device without work)."""
    import numpy as np
    import fable_b200 as fb
    from fable_b200 import synth
    from oracle import fable_oracle as orc
    if _gpus() < 2:
        pytest.skip("needs two GPUs")
    ndev = ndev or _gpus()
    n, p, k = 150, 1234, 4
    Y = synth.factor_data(n, p, k, rep=3)
    u, d, v = orc.svd_thin(Y)
    one_ctx = fb.Context(0, 777)
    multi = fb.MultiContext(ndev, 777)
    try:
        for inj in (False, True):
            gam = z = None
            if inj:
                gam, z = synth.injected_draws(MC, p, k, 0.5 * (1.0 + n), seed=MC)
            a = fb.CPPFABLESampler(Y, 1.0, 1.0, MC, u, v, d, k, 1.4, ctx=one_ctx, gam=gam, z=z, m0=5, psi_moments=True)
            b = fb.CPPFABLESampler(Y, 1.0, 1.0, MC, u, v, d, k, 1.4, ctx=multi, gam=gam, z=z, m0=5, psi_moments=True)
            for key in ("LambdaSamples", "SigmaSqSamples", "SigmaSqEstimatePostMean", "SigmaSqEstimatePlugIn", "G"):
                np.testing.assert_array_equal(a[key], b[key], err_msg=key)
            assert a["tauSqEstimate"] == b["tauSqEstimate"]
            for key in ("PsiSum", "PsiSumSq"):
                assert np.max(np.abs(a[key] - b[key])) <= 1e-12 * np.max(np.abs(a[key])), key
                assert np.array_equal(b[key], b[key].T)
            if inj:
                o = orc.sampler(Y, 1.0, 1.0, MC, u, v, d, k, 1.4, gam, z)
                s1, s2 = orc.psi_moments(o["LambdaSamples"], o["SigmaSqSamples"], k)
                assert np.max(np.abs(b["PsiSum"] - s1)) <= 1e-10 * np.max(np.abs(s1))
                assert np.max(np.abs(b["LambdaSamples"] - o["LambdaSamples"])) <= 1e-10 * np.max(np.abs(o["LambdaSamples"]))
        assert multi.launches > 0
    finally:
        multi.close(); one_ctx.close()
