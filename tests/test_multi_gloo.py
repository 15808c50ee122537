"""CPU tests of the multi-GPU host logic (fable_b200/parallel.py) with world_size-2 gloo:
draw sharding + sum all-reduce of the accumulators reproduces the single-process result, sample row
ranges are disjoint and complete, tile ranges partition the upper triangle.  The per-rank arithmetic
is done by the oracle's C twin here (no GPU in this container); on the GPU box the same helpers drive
libfable_b200 (bench.py under torchrun)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from fable_b200 import parallel  # noqa: E402


def test_draw_range_partitions_the_mc_axis():
    for MC in (0, 1, 7, 10, 1000, 10001):
        for world in (1, 2, 3, 8):
            r = [parallel.draw_range(MC, g, world) for g in range(world)]
            assert r[0][0] == 0 and r[-1][1] == MC
            assert all(r[g][1] == r[g + 1][0] for g in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.draw_range(10, 2, 2)


def _enumerate_tiles(p):
    """Brute-force statement of the covariance pass's tile order (fable_b200/csrc/tiles.cuh): 16 x 16 super-blocks over
    their own upper triangle, column-major; tiles column-major inside; None for padding slots."""
    nT = -(-p // 64); nSB = -(-nT // 16)
    out = []
    for sJ in range(nSB):
        for sI in range(sJ + 1):
            for r in range(256):
                I, J = sI * 16 + r % 16, sJ * 16 + r // 16
                out.append((I, J) if I <= J < nT else None)
    return out


def test_tile_range_partitions_the_upper_triangle():
    from fable_b200 import api
    for p in (1, 64, 65, 1000, 1025, 5000, 20000, 50000):
        enum = _enumerate_tiles(p)
        total = parallel.tile_slots(p)
        nT = -(-p // 64)
        assert total == len(enum) == api.tile_slots(p) and total % 256 == 0
        valid = [e for e in enum if e is not None]
        assert len(valid) == nT * (nT + 1) // 2 == api.tile_count(p) and len(set(valid)) == len(valid)
        prefix = np.concatenate([[0], np.cumsum([e is not None for e in enum])])
        for world in (1, 2, 3, 8):
            r = [parallel.tile_range(p, g, world) for g in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[g][1] == r[g + 1][0] for g in range(world - 1))
            counts = [api.tile_count(p, a, b) if b > a else 0 for a, b in r]
            assert counts == [int(prefix[b] - prefix[a]) for a, b in r]          # the library counts what the enumeration holds
            assert sum(counts) == len(valid) and max(counts) - min(counts) <= 1  # equal shares of the stored tiles
        rng = np.random.default_rng(p)
        for _ in range(50):                                   # slot and compact index of a tile, slices at arbitrary slots
            J = int(rng.integers(nT)); I = int(rng.integers(J + 1))
            slot, c = api.tile_locate(p, I, J)
            assert enum[slot] == (I, J) and c == prefix[slot]
            a, b = sorted(int(x) for x in rng.integers(0, total + 1, 2))
            if b > a:
                assert api.tile_count(p, a, b) == prefix[b] - prefix[a]
    with pytest.raises(fb_api_error()):
        api.tile_locate(100, 1, 0)


def fb_api_error():
    from fable_b200 import api
    return api.FableError


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    return port


def _worker(rank, world, port, MC, p, k, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import c_oracle
    rng = np.random.default_rng(0)                       # same posterior on every rank
    G0 = rng.standard_normal((p, k)); gnd = rng.uniform(20.0, 60.0, p)
    a_n, shape, seed = 0.07, 15.5, 99
    lo, hi = parallel.draw_range(MC, rank, world)
    gam, z = c_oracle.rng_fill(seed, lo, hi - lo, p, k, shape)        # counter-based: draw m is the same on any rank
    L, S = c_oracle.sampler(hi - lo, p, k, G0, gnd, a_n, gam, z)
    s1, s2 = c_oracle.psi_moments(L, S, k)
    t1, t2 = torch.from_numpy(np.ascontiguousarray(s1)), torch.from_numpy(np.ascontiguousarray(s2))
    parallel.allreduce_moments([t1, t2])
    Lall = parallel.gather_sample_rows(L, MC, rank, world)
    if rank == 0:
        np.savez(os.path.join(out_dir, "multi.npz"), s1=t1.numpy(), s2=t2.numpy(), L=Lall)
    dist.destroy_process_group()


def test_sharded_draws_and_allreduce_match_single_process(tmp_path):
    from oracle import c_oracle
    MC, p, k, world = 11, 37, 3, 2
    port = _free_port()
    mp.start_processes(_worker, args=(world, port, MC, p, k, str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    got = np.load(tmp_path / "multi.npz")
    rng = np.random.default_rng(0)
    G0 = rng.standard_normal((p, k)); gnd = rng.uniform(20.0, 60.0, p)
    gam, z = c_oracle.rng_fill(99, 0, MC, p, k, 15.5)
    L, S = c_oracle.sampler(MC, p, k, G0, gnd, 0.07, gam, z)
    s1, s2 = c_oracle.psi_moments(L, S, k)
    np.testing.assert_array_equal(got["L"], L)                     # samples: bit-identical, disjoint row ranges
    assert np.max(np.abs(got["s1"] - s1)) <= 1e-12 * np.max(np.abs(s1))   # sums: order of addition differs
    assert np.max(np.abs(got["s2"] - s2)) <= 1e-12 * np.max(np.abs(s2))


# ---- partition C (columns): the all-reduce callback and the row gather --------------------------------------
def _worker_columns(rank, world, port, n, p, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    Y = rng.standard_normal((n, p))                       # same matrix on every rank; each uses its own columns
    j0, j1 = parallel.column_range(p, rank, world)
    ar = parallel.host_allreduce()
    gram = np.ascontiguousarray(Y[:, j0:j1] @ Y[:, j0:j1].T).reshape(-1)     # what the library hands to the callback
    ar(gram)
    s = np.array([np.sum(np.log(np.sum(Y[:, j0:j1] ** 2, axis=0)))])          # a per-variable sum (criterion pattern)
    ar(s)
    rows = parallel.gather_rows(Y[:, j0:j1].T.copy(), p, j0, ar)              # per-variable rows (p_local, n) -> (p, n)
    vec = parallel.gather_rows(np.arange(j0, j1, dtype=np.float64), p, j0, ar)
    np.savez(os.path.join(out_dir, f"cols{rank}.npz"), gram=gram, s=s, rows=rows, vec=vec)
    dist.destroy_process_group()


def test_column_shards_allreduce_and_gather(tmp_path):
    n, p, world = 7, 23, 2
    for P in (0, 1, 5, 23):
        r = [parallel.column_range(P, g, 3) for g in range(3)]
        assert r[0][0] == 0 and r[-1][1] == P and all(r[g][1] == r[g + 1][0] for g in range(2))
    port = _free_port()
    mp.start_processes(_worker_columns, args=(world, port, n, p, str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    rng = np.random.default_rng(5)
    Y = rng.standard_normal((n, p))
    got = [np.load(tmp_path / f"cols{g}.npz") for g in range(world)]
    for g in range(world):
        np.testing.assert_allclose(got[g]["gram"].reshape(n, n), Y @ Y.T, rtol=1e-13, atol=1e-13)
        np.testing.assert_allclose(got[g]["s"][0], np.sum(np.log(np.sum(Y ** 2, axis=0))), rtol=1e-13)
        np.testing.assert_array_equal(got[g]["rows"], Y.T)                    # disjoint blocks: the gather is exact
        np.testing.assert_array_equal(got[g]["vec"], np.arange(p, dtype=np.float64))
        for key in ("gram", "s"):
            np.testing.assert_array_equal(got[g][key], got[0][key])           # same bits on every rank


def test_share_host_threads(monkeypatch):
    from fable_b200 import parallel
    monkeypatch.delenv("FABLE_B200_COPY_THREADS", raising=False)
    assert parallel.share_host_threads(8, cpu_count=16) == 2
    assert os.environ["FABLE_B200_COPY_THREADS"] == "2"
    assert parallel.share_host_threads(1, cpu_count=64) == 2          # an explicit / earlier setting is kept
    monkeypatch.delenv("FABLE_B200_COPY_THREADS")
    assert parallel.share_host_threads(1, cpu_count=64) == 16
    monkeypatch.delenv("FABLE_B200_COPY_THREADS")
    assert parallel.share_host_threads(8, cpu_count=4) == 1
