"""Column-sharded estimation over NCCL, one process per GPU, checked against the one-device path on rank 0.
launch: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_sharded_estimation.py [n p k]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import fable_b200 as fb
from fable_b200 import parallel, synth


def main():
    n, p, k = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (1000, 20000, 10)
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = fb.Context(local, 12345)
    Y = synth.factor_data(n, p, k, rep=0)                  # same seed -> same matrix on every rank
    j0, j1 = parallel.column_range(p, rank, world)
    ar = parallel.host_allreduce(dev)
    Yl = np.asfortranarray(Y[:, j0:j1])
    ard = parallel.device_allreduce(dev)                               # the n x n Gram is summed in HBM (NCCL on the library's stream)
    parallel.sharded_estimation(Yl, p, j0, rank, world, ctx, ar, allreduce_dev=ard)      # warm-up (module load, NCCL channels)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = parallel.sharded_estimation(Yl, p, j0, rank, world, ctx, ar, allreduce_dev=ard)
    torch.cuda.synchronize(); dist.barrier()
    t_sh = time.perf_counter() - t0
    post = fb.Posterior.from_rows(n, 1.0, 1.0, res["rows"], res["varInflation"], ctx=ctx)
    L, S = post.samples(3, 4)
    post.close()
    ok = True
    if rank == 0:
        fb.FABLEPosteriorSampler(Y, MC=1, ctx=ctx, want_G=False)
        t0 = time.perf_counter()
        one = fb.FABLEPosteriorSampler(Y, MC=1, ctx=ctx, want_G=False)
        t_one = time.perf_counter() - t0
        s1 = one["svdY"]
        kk = one["estRank"]
        rel = lambda a, b: float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
        checks = {
            "estRank equal": res["estRank"] == kk,
            "d": rel(res["svd"]["d"], s1["d"]) < 1e-10,
            "u_k": rel(res["svd"]["u"][:, :kk], s1["u"][:, :kk]) < 1e-8,
            "v_k (local rows)": rel(res["svd"]["v"][:, :kk], s1["v"][j0:j1, :kk]) < 1e-8,
            "varInflation": abs(res["varInflation"] / one["varInflation"] - 1) < 1e-10,
            "tauSq": abs(res["tauSqEstimate"] / one["FABLEHyperParameters"]["tauSqEstimate"] - 1) < 1e-10,
            "SigmaSqEstimate": rel(res["SigmaSqEstimate"], one["FABLEHyperParameters"]["SigmaSqEstimate"]) < 1e-10,
        }
        post1 = fb.Posterior(Y, 1.0, 1.0, s1["u"], s1["v"], s1["d"], kk, one["varInflation"], ctx=ctx)
        L1, S1 = post1.samples(3, 4)
        post1.close()
        checks["samples from gathered rows"] = rel(L, L1) < 1e-9 and rel(S, S1) < 1e-9
        ok = all(checks.values())
        print(f"sharded estimation n={n} p={p} k={k} on {world} GPUs: {t_sh * 1e3:.1f} ms (one device, incl. 1 draw: {t_one * 1e3:.1f} ms); "
              f"kEst={res['estRank']} kMax={res['kMax']}; checks: {checks}", flush=True)
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.broadcast(flag, 0)
    # every rank must hold the same replicated results
    h = torch.tensor([float(np.sum(res["svd"]["u"])), float(np.sum(res["rows"]["G0"])), float(res["estRank"])], dtype=torch.float64, device=dev)
    hs = [torch.zeros_like(h) for _ in range(world)]
    dist.all_gather(hs, h)
    same = all(torch.equal(hs[0], x) for x in hs)
    if rank == 0:
        print("replicated results identical on all ranks:", same, flush=True)
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if (flag.item() == 1.0 and same) else 1)


if __name__ == "__main__":
    main()
