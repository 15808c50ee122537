"""Small driver for ncu: one rank's share of the tile partition at BASELINE configs[3] (p = 50000, k = 20) on ONE GPU.
usage: python tools/prof_psi_tiles.py [world] [rank] [B] [steps]   (default: rank 3 of 8, 32 draws per launch, 3 launches)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fable_b200 as fb
from fable_b200 import synth, parallel

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 3
B = int(sys.argv[3]) if len(sys.argv) > 3 else 32
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
n, p, k = 500, 50000, 20
Y = synth.factor_data(n, p, k, rep=0)
lam, Q = np.linalg.eigh(Y @ Y.T)
U = np.asfortranarray(Q[:, ::-1][:, :k]); T = Y.T @ U
d = np.linalg.norm(T, axis=0); V = np.asfortranarray(T / d)
ctx = fb.Context(0, 12345)
post = fb.Posterior(Y, 1.0, 1.0, U, V, d, k, 4.9, ctx=ctx)
lo, hi = parallel.tile_range(p, rank, world)
post.set_tile_range(lo, hi)
nt = fb.api.tile_count(p, lo, hi)
ring = torch.empty(B * 2 * nt * 4096, dtype=torch.float64, device="cuda")
post.accum_reset()
ctx.kernel_timing(True)
m0 = 0
for s in range(steps):
    post.draw_batch_tiles(m0, B, dev_psi_tiles=ring.data_ptr(), psi_slots=B, accumulate=True); m0 += B
ms, cnt = ctx.kernel_time()
gb = 8.0 * 2 * nt * 4096 * B / 1e9
print(f"tiles rank {rank}/{world} p={p} k={k} B={B}: {nt} tiles, {ms / cnt:.3f} ms per launch, {gb / (ms / cnt) * 1e3:.1f} GB/s stored ({gb:.1f} GB per launch)")
