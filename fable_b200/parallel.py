"""Multi-GPU sharding of the sampling path: one process per GPU (torch.distributed), NCCL only
where the path has a real exchange step.

Partition A (draws, BASELINE config 3): rank g owns a contiguous range of the MC axis.  The
counter-based stream makes draw m a pure function of (seed, m, row), so the union over ranks is
bit-identical to a single-GPU run; LambdaSamples / SigmaSqSamples row-ranges are disjoint (no
communication) and the only collective is one sum all-reduce of the p x p accumulators.

Partition B (tiles, BASELINE config 4): rank g owns a set of upper-triangle 64 x 64 output tiles
and processes every draw for them; no collective.  `tile_range` gives the contiguous slice of the
column-major upper-triangle tile enumeration used by the covariance kernel.

The reference has no multi-process code at all: there is nothing to mirror here,
only its MC loop (src/updated-FABLE-functions.cpp:599-618) and pair loop (:650-688) to shard.
"""
from __future__ import annotations

import numpy as np


def draw_range(MC, rank, world):
    """Contiguous [m_begin, m_end) of the MC axis owned by `rank`; sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world) or MC < 0:
        raise ValueError("draw_range: need 0 <= rank < world and MC >= 0")
    base, rem = divmod(int(MC), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def tile_slots(p, tile=64, sb=16):
    """Length of the covariance kernel's tile enumeration: upper-triangle 64 x 64 tiles in 16 x 16
    super-block order, padding slots included (same count as fable_posterior_tile_slots)."""
    nT = -(-int(p) // tile)
    nSB = -(-nT // sb)
    return nSB * (nSB + 1) // 2 * sb * sb


def tile_range(p, rank, world, tile=64, sb=16):
    """Contiguous [t_begin, t_end) of the tile enumeration owned by `rank` (multiples of a super-block
    so every rank keeps the operand locality of the kernel's order)."""
    total = tile_slots(p, tile, sb)
    blocks = total // (sb * sb)
    lo, hi = draw_range(blocks, rank, world)
    return lo * sb * sb, hi * sb * sb


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
    fable_posterior over ranks.  Only their upper triangles carry information, so those are packed
    (p(p+1)/2 doubles each) and reduced to rank `dst` (dst=None: all-reduce, every rank gets the sums);
    the result is unpacked in place on the receiving rank(s).  The library's stream is drained first and
    torch's current stream after, so the two stream domains never overlap on the buffers."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    a, b, n = post.accum_pack()
    post.ctx.synchronize()
    ts = [as_torch(a, n, device), as_torch(b, n, device)]
    for t in ts:
        if dst is None:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        else:
            dist.reduce(t, dst=dst, op=dist.ReduceOp.SUM, group=group)
    torch.cuda.synchronize(device)
    if dst is None or dist.get_rank(group) == dst:
        post.accum_unpack()
        post.ctx.synchronize()


def allreduce_posterior_accumulators(post, device, group=None):
    reduce_posterior_accumulators(post, device, dst=None, group=group)


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
