"""Small driver for ncu: the sampler + covariance kernels only (no SVD / rank search launches).
usage: python tools/prof_psi.py [p] [B] [steps] [mode]   mode: mat_acc | acc | mat"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fable_b200 as fb
from fable_b200 import synth

p = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
mode = sys.argv[4] if len(sys.argv) > 4 else "mat_acc"
n, k = 1000, 10
Y = synth.factor_data(n, p, k, rep=0)
lam, Q = np.linalg.eigh(Y @ Y.T)
U = np.asfortranarray(Q[:, ::-1][:, :k])
T = Y.T @ U
d = np.linalg.norm(T, axis=0)
V = np.asfortranarray(T / d)
ctx = fb.Context(0, 12345)
post = fb.Posterior(Y, 1.0, 1.0, U, V, d, k, 4.9, ctx=ctx)
slots = B
ring = torch.empty((slots, p, p), dtype=torch.float64, device="cuda") if mode != "acc" else None
post.accum_reset()
ctx.kernel_timing(True)
m0 = 0
for s in range(steps):
    post.draw_batch(m0, B, dev_psi=ring.data_ptr() if ring is not None else 0, psi_slots=slots if ring is not None else 0,
                    accumulate=(mode != "mat"))
    m0 += B
ms, cnt = ctx.kernel_time()
print(f"mode={mode} p={p} B={B}: {ms / cnt:.3f} ms per launch, {B * cnt / ms * 1e3:.1f} draws/s, "
      f"{8.0 * p * p * B * cnt / ms / 1e6:.1f} GB/s (8p^2 per draw)")
