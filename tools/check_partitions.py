"""Multi-GPU partitions of the sampling path over NCCL, one process per GPU, checked on rank 0 against the one-device
path and the CPU oracle (test infrastructure: this script may import oracle/).
launch: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/check_partitions.py [n p k MC]

  A (draws, BASELINE configs[2]): parallel.distributed_sampler -- every rank draws its share of the MC axis through the
    C-ABI, the accumulators are summed on the device, rank 0 downloads them.  Checked: the union of the ranks' sample rows
    equals the one-device CPPFABLESampler bit for bit; the reduced PsiSum / PsiSumSq equal it to 1e-12 (summation order).
  B (tiles, BASELINE configs[3]): every rank processes all draws for its slice of the tile enumeration and stores only
    its share.  Checked: the shares (expanded to dense with zeros elsewhere and summed over ranks) equal the one-device
    draws and accumulators bit for bit, and sampled entries equal the oracle's psi_draw.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import fable_b200 as fb
from fable_b200 import parallel, synth
from oracle import fable_oracle as orc


def main():
    n, p, k, MC = (int(a) for a in sys.argv[1:5]) if len(sys.argv) > 4 else (200, 3000, 5, 37)
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = fb.Context(local, 4242)
    Y = synth.factor_data(n, p, k, rep=1)                  # same seed -> same matrix on every rank
    u, d, v = orc.svd_thin(Y)
    vi = 1.37
    ok = True

    # ---- A: draws ---------------------------------------------------------------------------------------------
    res = parallel.distributed_sampler(Y, 1.0, 1.0, MC, u, v, d, k, vi, ctx, dev, m0=11, dst=0)
    lo, hi = res["draw_range"]
    assert (lo, hi) == parallel.draw_range(MC, rank, world)
    parts_L = [None] * world; parts_S = [None] * world
    dist.all_gather_object(parts_L, np.ascontiguousarray(res["LambdaSamples"]))
    dist.all_gather_object(parts_S, np.ascontiguousarray(res["SigmaSqSamples"]))
    if rank == 0:
        one = fb.CPPFABLESampler(Y, 1.0, 1.0, MC, u, v, d, k, vi, ctx=ctx, m0=11, want_G=False, psi_moments=True)
        same_rows = np.array_equal(np.concatenate(parts_L, axis=0), one["LambdaSamples"]) and \
            np.array_equal(np.concatenate(parts_S, axis=0), one["SigmaSqSamples"])
        e1 = float(np.max(np.abs(res["PsiSum"] - one["PsiSum"])) / np.max(np.abs(one["PsiSum"])))
        e2 = float(np.max(np.abs(res["PsiSumSq"] - one["PsiSumSq"])) / np.max(np.abs(one["PsiSumSq"])))
        sym = np.array_equal(res["PsiSum"], res["PsiSum"].T)
        r1, r2 = orc.psi_moments(one["LambdaSamples"], one["SigmaSqSamples"], k)
        eo = float(np.max(np.abs(res["PsiSum"] - r1)) / np.max(np.abs(r1)))
        print(f"A draws: union of sample rows identical to one device: {same_rows}; reduced sums vs one device "
              f"{e1:.2e} / {e2:.2e}; vs oracle {eo:.2e}; symmetric: {sym}")
        ok &= same_rows and e1 < 1e-12 and e2 < 1e-12 and eo < 1e-10 and sym
    else:
        assert "PsiSum" not in res                         # only rank 0 downloads the p x p sums

    # ---- B: tiles ---------------------------------------------------------------------------------------------
    B = 5
    post = fb.Posterior(Y, 1.0, 1.0, u, v, d, k, vi, ctx=ctx)
    lo_t, hi_t = parallel.tile_range(p, rank, world)
    post.set_tile_range(lo_t, hi_t)
    nt = fb.api.tile_count(p, lo_t, hi_t)
    ring = torch.full((B, 2 * nt, 64, 64), float("nan"), dtype=torch.float64, device=dev)
    post.draw_batch_tiles(3, B, dev_psi_tiles=ring.data_ptr(), psi_slots=B, accumulate=True)
    a_sum, a_sq, cnt = post.accum_dev()
    assert cnt == nt * 4096
    dense = torch.empty((B + 2, p, p), dtype=torch.float64, device=dev)
    for b in range(B):
        post.tiles_to_dense(ring[b].data_ptr(), True, 0, p, dense[b].data_ptr())
    post.tiles_to_dense(a_sum, False, 0, p, dense[B].data_ptr())
    post.tiles_to_dense(a_sq, False, 0, p, dense[B + 1].data_ptr())
    ctx.synchronize()
    nz = (dense != 0).to(torch.float64)
    dist.all_reduce(dense, op=dist.ReduceOp.SUM)           # disjoint supports: the sum places every share
    dist.all_reduce(nz, op=dist.ReduceOp.SUM)
    post.set_tile_range(0, 0)
    if rank == 0:
        full = torch.empty((B, p, p), dtype=torch.float64, device=dev)
        post.draw_batch(3, B, dev_psi=full.data_ptr(), psi_slots=B, accumulate=True)
        f1, f2 = post.accum_fetch()
        same_draws = bool(torch.equal(dense[:B], full))
        same_acc = np.array_equal(dense[B].cpu().numpy().T, f1) and np.array_equal(dense[B + 1].cpu().numpy().T, f2)
        disjoint = float(nz.max().item()) <= 1.0
        L, S = post.samples(3, B)
        eo = 0.0
        for b in range(B):
            want = orc.psi_draw(L[b], S[b], k)
            eo = max(eo, float(np.max(np.abs(full[b].cpu().numpy().T - want)) / np.max(np.abs(want))))
        print(f"B tiles: {world} shares of {[fb.api.tile_count(p, *parallel.tile_range(p, g, world)) for g in range(world)]} tiles; "
              f"union identical to one device: draws {same_draws}, accumulators {same_acc}; disjoint: {disjoint}; vs oracle {eo:.2e}")
        ok &= same_draws and same_acc and disjoint and eo < 1e-10
    post.close()
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.broadcast(flag, 0)
    if rank == 0:
        print("partitions ok:", bool(flag.item() == 1.0))
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
