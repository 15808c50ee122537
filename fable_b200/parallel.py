"""Multi-GPU sharding of the sampling path: one process per GPU (torch.distributed), NCCL only
where the path has a real exchange step.

Partition A (draws, BASELINE config 3): rank g owns a contiguous range of the MC axis.  The
counter-based stream makes draw m a pure function of (seed, m, row), so the union over ranks is
bit-identical to a single-GPU run; LambdaSamples / SigmaSqSamples row-ranges are disjoint (no
communication) and the only collective is one sum all-reduce of the p x p accumulators.

Partition B (tiles, BASELINE config 4): rank g owns a set of upper-triangle 64 x 64 output tiles
and processes every draw for them; no collective.  `tile_range` gives the contiguous slice of the
column-major upper-triangle tile enumeration used by the covariance kernel.

Partition C (columns, estimation: SURVEY.md §8e "Gram/SVD" and "Rank / row-posterior"): rank g holds a block of
the variables (columns of Y).  `column_range` gives the block, `host_allreduce` the callback the library completes
its sums over variables with (`Context.set_column_shard`), `sharded_estimation` runs svd -> kMax -> rank search ->
row posterior -> varInflation on the shards and returns the gathered O(p k) row posterior every rank builds its
sampler from (`Posterior.from_rows`).  Exchanged: the n x n Gram once, a handful of scalars per criterion
evaluation, the row posterior once.

The reference has no multi-process code at all: there is nothing to mirror here,
only its MC loop (src/updated-FABLE-functions.cpp:599-618) and pair loop (:650-688) to shard.
"""
from __future__ import annotations

import os

import numpy as np


def share_host_threads(local_world, cpu_count=None):
    """One process per GPU on one host: split the host cores between the ranks' copy pipelines.  The library moves large
    results into pageable caller memory with up to 16 host threads per context (`FABLE_B200_COPY_THREADS`); N ranks
    doing so at once would run 16 N threads.  Sets the variable only if the user has not; returns the value in force."""
    n = cpu_count if cpu_count is not None else (os.cpu_count() or 1)
    per_rank = max(1, min(16, n // max(1, int(local_world))))
    return int(os.environ.setdefault("FABLE_B200_COPY_THREADS", str(per_rank)))


def draw_range(MC, rank, world):
    """Contiguous [m_begin, m_end) of the MC axis owned by `rank`; sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world) or MC < 0:
        raise ValueError("draw_range: need 0 <= rank < world and MC >= 0")
    base, rem = divmod(int(MC), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def tile_slots(p, tile=64, sb=16):
    """Length of the covariance kernel's tile enumeration: upper-triangle 64 x 64 tiles in 16 x 16
    super-block order, padding slots included (same count as fable_tile_slots)."""
    nT = -(-int(p) // tile)
    nSB = -(-nT // sb)
    return nSB * (nSB + 1) // 2 * sb * sb


def tile_range(p, rank, world):
    """Contiguous [t_begin, t_end) of the tile enumeration owned by `rank`: equal shares of the VALID tiles
    (fable_tile_partition; a contiguous slice keeps the operand locality of the kernel's order and owns a
    contiguous range of the compact tile storage)."""
    from . import api
    return api.tile_partition(p, rank, world)


class _DevArray:
    """Zero-copy view of a raw device allocation for torch (via __cuda_array_interface__)."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def as_torch(ptr, count, device):
    import torch
    return torch.as_tensor(_DevArray(ptr, count), device=device)


def allreduce_moments(tensors, group=None):
    """Sum all-reduce of the accumulator tensors (NCCL on GPUs; gloo in the CPU tests)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    for t in tensors:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)


def reduce_posterior_accumulators(post, device, dst=0, group=None):
    """The one exchange step of partition A: sum the device-resident sum / sumsq accumulators of a
    fable_posterior over ranks.  They are stored as compact upper tiles (the lower triangle is never kept), so
    the buffers are reduced in place as they are, to rank `dst` (dst=None: all-reduce, every rank gets the
    sums).  The library's stream is drained first and torch's current stream after, so the two stream domains
    never overlap on the buffers."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    a, b, n = post.accum_dev()
    post.ctx.synchronize()
    ts = [as_torch(a, n, device), as_torch(b, n, device)]
    for t in ts:
        if dst is None:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        else:
            dist.reduce(t, dst=dst, op=dist.ReduceOp.SUM, group=group)
    torch.cuda.synchronize(device)


def allreduce_posterior_accumulators(post, device, group=None):
    reduce_posterior_accumulators(post, device, dst=None, group=group)


def distributed_sampler(Y, gamma0, delta0sq, MC, U_Y, V_Y, svalsY, kEst, varInflation, ctx, device, m0=0, out=None,
                        dst=0, group=None, psi_moments=True):
    """CPPFABLESampler (cpp:532-629) on the draw partition, through the C-ABI with host buffers: every rank calls this
    with the same inputs.  Rank g uploads them, draws its contiguous share [lo, hi) of the MC axis (its rows of
    LambdaSamples / SigmaSqSamples land in its own host arrays; the union over ranks is what one GPU would draw),
    the sum / sumsq accumulators are summed ON THE DEVICE (one NCCL reduce of the compact tiles) and only rank `dst`
    (None: every rank) downloads the dense p x p PsiSum / PsiSumSq.  `out` may supply column-major host arrays
    ("LambdaSamples" (hi-lo) x k p, "SigmaSqSamples" (hi-lo) x p, "PsiSum" / "PsiSumSq" p x p)."""
    import torch.distributed as dist
    from . import api
    on = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank(group) if on else 0
    world = dist.get_world_size(group) if on else 1
    lo, hi = draw_range(MC, rank, world)
    out = out or {}
    post = api.Posterior(Y, gamma0, delta0sq, U_Y, V_Y, svalsY, kEst, varInflation, ctx=ctx)
    try:
        p, k = post.p, post.k
        L = out.get("LambdaSamples"); S = out.get("SigmaSqSamples")
        if L is None:
            L = np.empty((hi - lo, k * p), order="F")
        if S is None:
            S = np.empty((hi - lo, p), order="F")
        if psi_moments:
            post.accum_reset()
        if hi > lo:
            post.sample_range(m0 + lo, hi - lo, L, S, psi_pass=psi_moments)
        res = {"LambdaSamples": L, "SigmaSqSamples": S, "draw_range": (lo, hi), "EstimatedRank": k,
               "VarianceInflation": float(varInflation)}
        if psi_moments:
            reduce_posterior_accumulators(post, device, dst=dst, group=group)
            if dst is None or rank == dst:
                s1 = out.get("PsiSum"); s2 = out.get("PsiSumSq")
                if s1 is None:
                    s1 = np.empty((p, p), order="F")
                if s2 is None:
                    s2 = np.empty((p, p), order="F")
                res["PsiSum"], res["PsiSumSq"] = post.accum_fetch(out=(s1, s2))
        return res
    finally:
        post.close()


def gather_sample_rows(local, MC, rank, world, group=None):
    """Assemble the MC x q sample matrix on every rank from disjoint contiguous row ranges
    (host tensors; used by tests and small runs -- large runs keep the ranges where they are)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return local
    parts = [None] * world
    dist.all_gather_object(parts, np.ascontiguousarray(local), group=group)
    out = np.concatenate(parts, axis=0)
    assert out.shape[0] == MC
    return np.asfortranarray(out)


# ---- partition C: column-sharded estimation ---------------------------------------------------------------
def column_range(p, rank, world):
    """Contiguous [j_begin, j_end) of the variables (columns of Y) owned by `rank`."""
    return draw_range(p, rank, world)


def host_allreduce(device=None, group=None):
    """In-place sum of a float64 numpy array over the ranks of `group`: the callback `Context.set_column_shard`
    needs.  NCCL groups go through a device staging tensor (the library hands over host buffers: the n x n Gram
    once, a few scalars per criterion evaluation); gloo groups reduce the host buffer directly."""
    import torch
    import torch.distributed as dist

    def fn(a):
        if dist.get_backend(group) == "nccl":
            t = torch.from_numpy(a).to(device)
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            a[...] = t.cpu().numpy()
        else:
            dist.all_reduce(torch.from_numpy(a), op=dist.ReduceOp.SUM, group=group)
    return fn


def device_allreduce(device, group=None):
    """In-place NCCL sum of `count` float64 at a device address, ordered on the library's stream: the device-pointer
    callback of `Context.set_column_shard` (the n x n Gram of the sharded SVD stays in HBM).  NCCL groups only."""
    import torch
    import torch.distributed as dist

    def fn(ptr, count, stream):
        t = as_torch(ptr, count, device)
        with torch.cuda.stream(torch.cuda.ExternalStream(stream, device=device)):
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return fn


def gather_rows(local, p_total, col_offset, allreduce):
    """Every rank's rows of a per-variable array (p_local, ...) -> the full (p_total, ...) array on every rank:
    each rank writes its block into zeros and the blocks are summed (disjoint, so the sum is exact)."""
    local = np.asarray(local, dtype=np.float64)
    full = np.zeros((int(p_total),) + local.shape[1:], dtype=np.float64)
    full[col_offset:col_offset + local.shape[0]] = local
    flat = full.reshape(-1)
    allreduce(flat)
    return full


def sharded_estimation(Y_local, p_total, col_offset, rank, world, ctx, allreduce, gamma0=1.0, delta0sq=1.0, maxProp=0.95,
                       variant="wrapper", allreduce_dev=None):
    """The estimation half of FABLEPosteriorSampler (R/FABLE_wrapper.R:20-44) on column shards: every rank passes
    its own columns of Y.  Returns, identical on every rank: svd d / u (and the local rows of v), kMax, kEst, the
    gathered row posterior (G0, gnd, sig, sigmahat, tausq: what `Posterior.from_rows` takes), the plug-in
    hyper-parameters' sigma^2 and unshrunk G0, and varInflation."""
    from . import api
    n = Y_local.shape[0]
    ctx.set_column_shard(p_total, col_offset, rank, world, allreduce, allreduce_dev=allreduce_dev)
    try:
        s = api.svd(Y_local, ctx=ctx)                                              # wrapper:23-26
        kMax = api.kmax_rule(s["d"], maxProp)                                      # wrapper:27
        kEst = api.CPPRankEstimator(Y_local, s["u"], s["v"], s["d"], kMax, ctx=ctx)   # wrapper:31
        hyp = api.row_posterior(Y_local, 1.0, 1.0, s["u"], s["v"], s["d"], kEst, shrunk=False, ctx=ctx)   # wrapper:36 (cpp:371-391)
        rows = api.row_posterior(Y_local, gamma0, delta0sq, s["u"], s["v"], s["d"], kEst, shrunk=True, ctx=ctx)   # cpp:564-592
    finally:
        ctx.set_column_shard(0)
    g = lambda a: gather_rows(a, p_total, col_offset, allreduce)
    sig_plug = g(hyp["sig"])
    G0_plug = np.asfortranarray(g(hyp["G0"]))
    full = {"G0": np.asfortranarray(g(rows["G0"])), "gnd": g(rows["gnd"]), "sig": g(rows["sig"]),
            "sigmahat": g(rows["sigmahat"]), "tausq": rows["tausq"]}
    vi = api.var_inflation(sig_plug, G0_plug, variant, ctx=ctx)                    # wrapper:40-44 (replicated: O(p^2 k) flops, no p x p)
    return {"svd": s, "kMax": kMax, "estRank": kEst, "rows": full, "SigmaSqEstimate": sig_plug, "G0_unshrunk": G0_plug,
            "tauSqEstimate": hyp["tausq"], "varInflation": vi, "n": n}
