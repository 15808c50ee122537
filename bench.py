#!/usr/bin/env python
"""bench.py -- posterior samples/sec of the FABLE sampling hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            (N=1: plain python; N>1: torchrun)
    python bench.py --impl reference ...                     (CPU restatement on the host cores)
    python bench.py --workload c4 [--partition tiles]        (BASELINE configs[3]: covariance tiles over the GPUs)

One STEP = one batch of B posterior draws per GPU through the hot path:
    Philox sampler kernel  (sigma2_m, Lambda_m; reference cpp:599-618)
 -> covariance kernel      (Psi_m = Lambda_m Lambda_m' + diag(sigma2_m) formed in 64x64 blocks,
                            every draw MATERIALISED as a dense p x p FP64 matrix into an HBM ring,
                            and folded into entry-wise sum / sum-of-squares accumulators;
                            reference entry formula cpp:672-684, mean cpp:688).
`value`  = draws/s with the row posterior already resident in HBM (CUDA events, max over ranks).
`e2e`    = the same metric through the reference-facing C-ABI (`fable_sampler` = CPPFABLESampler, cpp:532-629) with HOST
           buffers: Y, U, V, d uploaded from pinned memory and LambdaSamples, SigmaSqSamples, Psi sum / sumsq read
           back, every step.  N > 1: the distributed job -- every rank draws its share of the MC axis through the C-ABI,
           the accumulators are summed on the device (NCCL) and rank 0 alone downloads the dense p x p sums.
`e2e_pageable` = the same call with plain (pageable) numpy inputs and freshly allocated outputs, as R hands them over.

Partitions:
  draws (default; c1-c3): draws are sharded over ranks (weak scaling: B draws per GPU per step); the only collective is
         one NCCL sum of the compact accumulator tiles to rank 0 at the end of the timed region.
  tiles (c4): every rank processes ALL draws for its own slice of the 64 x 64 covariance tiles and stores only that
         share (compact per-rank layout); no collective at all (strong scaling: the job's draws/s).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs; the metric is quoted on c3, which fits one B200 (3.2 GB per p x p)
    "c1": dict(n=500, p=1000, k=10, batch=64, e2e_mc=1000, ref_mc=256),
    "c2": dict(n=1000, p=5000, k=10, batch=32, e2e_mc=2000, ref_mc=128),
    "c3": dict(n=1000, p=20000, k=10, batch=40, e2e_mc=8192, ref_mc=16),    # 40 slots x 3.2 GB = 128 GB of the 183 GB
    "c4": dict(n=500, p=50000, k=20, batch=0, e2e_mc=1024, ref_mc=4),    # 20 GB per p x p: tiles over the GPUs
}
RING_BUDGET_GB = 120.0     # tile partition: HBM given to the ring of materialised draws (per GPU)


def metric_name(w):
    return f"posterior samples/sec (n={w['n']},p={w['p']},k={w['k']})"


UNIT = "samples/s"


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            pk = json.load(f)
        return float(pk["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clocks and throttle reasons during the timed region (NVML, 50 ms period)."""

    def __init__(self, index):
        self.index = index; self.samples = []; self.reasons = set(); self.max_mhz = None
        self._stop = threading.Event(); self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {}
        for nm in ("HwSlowdown", "HwThermalSlowdown", "SwThermalSlowdown", "SwPowerCap", "HwPowerBrakeSlowdown",
                   "SyncBoost", "ApplicationsClocksSetting"):
            bit = getattr(nv, "nvmlClocksThrottleReason" + nm, None)
            if bit is not None:
                names[bit] = nm
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.05)

    def start(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._run, daemon=True); self._t.start()

    def stop(self):
        self._stop.set()
        if self._t is not None:
            self._t.join()
        snake = {"HwSlowdown": "hw_slowdown", "HwThermalSlowdown": "hw_thermal_slowdown", "SwThermalSlowdown": "sw_thermal_slowdown",
                 "SwPowerCap": "sw_power_cap", "HwPowerBrakeSlowdown": "hw_power_brake", "SyncBoost": "sync_boost",
                 "ApplicationsClocksSetting": "applications_clocks_setting"}
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(snake[r] for r in self.reasons), "samples": len(self.samples)}


def _leading_triplets(Y, k):
    """Leading k singular triplets by host LAPACK (untimed set-up input of the timed regions)."""
    n, p = Y.shape
    lam, Q = np.linalg.eigh(Y @ Y.T if n <= p else Y.T @ Y)
    if n <= p:
        U = np.asfortranarray(Q[:, ::-1][:, :k]); T = Y.T @ U
        d = np.linalg.norm(T, axis=0); V = np.asfortranarray(T / d)
    else:
        V = np.asfortranarray(Q[:, ::-1][:, :k]); T = Y @ V
        d = np.linalg.norm(T, axis=0); U = np.asfortranarray(T / d)
    return U, d, V


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle's C twin (OpenMP, all host threads) on a bounded sample of the workload
# ------------------------------------------------------------------------------------------------
def _cpu_inputs(w):
    """Row posterior for the CPU arm from the same synthetic Y (estimation done by the numpy oracle
    with the consistent-SVD shortcut-free formulas; untimed)."""
    from fable_b200 import synth
    from oracle import fable_oracle as orc
    n, p, k = w["n"], w["p"], w["k"]
    Y = synth.factor_data(n, p, k, rep=0)
    U, d, V = _leading_triplets(Y, k)       # Gram + eigh is exact enough for a timing input
    st = orc._shrunk_stats(Y, U, V, d, k, 1.0, 1.0)
    return st, n, p, k


def _cpu_threads():
    """The CPU arm uses every host core this process may run on -- set explicitly: torch.distributed.run exports
    OMP_NUM_THREADS=1 to its workers, which would make the baseline single-threaded."""
    from oracle import c_oracle
    return c_oracle.set_num_threads()


def cpu_sample_rate(w, mc, repeats=1):
    """draws/s of the CPU restatement: sampler loop (cpp:599-618) + Psi entry formation and
    entry-wise summaries over the draws (cpp:650-688), `mc` draws at full p."""
    from oracle import c_oracle
    cores = _cpu_threads()
    st, n, p, k = _cpu_inputs(w)
    a_n = np.sqrt(1.26) / np.sqrt(st["w"])
    shape = 0.5 * st["gamma_n"]
    gam, z = c_oracle.rng_fill(12345, 0, mc, p, k, shape)      # draw generation is not timed on the CPU arm
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        L, S = c_oracle.sampler(mc, p, k, st["G0"], st["gnd"], a_n, gam, z)
        s1, s2 = c_oracle.psi_moments(L, S, k)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return mc / best, best, cores


def run_reference(args, w, wname):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    mc = w["ref_mc"]
    from oracle import c_oracle
    cores = _cpu_threads()
    st, n, p, k = _cpu_inputs(w)
    a_n = np.sqrt(1.26) / np.sqrt(st["w"]); shape = 0.5 * st["gamma_n"]
    gam, z = c_oracle.rng_fill(12345, 0, mc, p, k, shape)
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        L, S = c_oracle.sampler(mc, p, k, st["G0"], st["gnd"], a_n, gam, z)
        c_oracle.psi_moments(L, S, k)
        if it >= args.warmup:
            times.append(time.perf_counter() - t0)
    total = float(np.sum(times))
    val = mc * args.steps / total
    sample = f"{mc} draws per step at full p={p}: sampler loop + Psi entries + entry-wise sum/sumsq (C restatement, OpenMP)"
    print(json.dumps({
        "impl": "reference", "metric": metric_name(w), "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
        "scaling": "strong" if _partition(args, wname) == "tiles" else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": _workload_label(wname, n, p, k, _partition(args, wname), int(os.environ.get("WORLD_SIZE", "1"))),
                   "draws_per_step": mc},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def _partition(args, wname):
    return args.partition or ("tiles" if wname == "c4" else "draws")


def _workload_label(wname, n, p, k, partition, world):
    s = f"{wname}: n={n}, p={p}, k={k}"
    if partition == "tiles":
        return s + "; covariance tiles partitioned over ranks"
    return s + ("; draws sharded over ranks" if world > 1 else "")


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
class Job:
    """Process-group and device plumbing shared by the two partitions."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        from fable_b200 import parallel
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            parallel.share_host_threads(int(os.environ.get("LOCAL_WORLD_SIZE", self.world)))
            torch.cuda.set_device(self.local)
            # NCCL prints its version banner to stdout when the communicator is created: keep stdout to the one
            # JSON line by pointing fd 1 at stderr while the group and its first collective come up
            sys.stdout.flush()
            saved = os.dup(1); os.dup2(2, 1)
            try:
                dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
                dist.barrier()
                torch.cuda.synchronize(self.local)
            finally:
                sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)
        else:
            torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def _pin(torch, *shape):
    return torch.empty(shape, dtype=torch.float64).pin_memory()


def _estimation_rooflines(job, ctx, Y, U, V, d, kEst, stream):
    """Gram step and row pass timed inside the run (CUDA events on the library's stream, resident inputs):
    roofline_gram = m(m+1)K flops / fable_dev_syrk time / measured DMMA peak; roofline_rows = 8 n p bytes /
    fable_dev_rowstats time / measured HBM bandwidth."""
    import fable_b200 as fb
    torch = job.torch
    n, p = Y.shape
    dY = torch.from_numpy(np.ascontiguousarray(Y.T)).to(job.dev)           # (p, n) row-major == n x p column-major
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)

    def timed(fn, reps=5, warm=2):
        for _ in range(warm):
            fn()
        ctx.synchronize()
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream); ctx.synchronize()
        return e0.elapsed_time(e1) / reps

    m, K = (n, p) if n <= p else (p, n)
    dC = torch.empty((m, m), dtype=torch.float64, device=job.dev)
    gram_ms = timed(lambda: fb.api.dev_syrk(ctx, 0 if n <= p else 1, m, K, 1.0, dY.data_ptr(), n, dC.data_ptr()))
    peak64 = ctx.measure_fp64_peak()
    flops = float(m) * (m + 1) * K
    gram = {"bound": "tensor", "kernel": "k_syrk_tma (fable_dev_syrk: Gram Y Y' incl. split-K sum and mirror)",
            "achieved": flops / (gram_ms * 1e-3) / 1e12, "peak": peak64, "unit": "TFLOP/s",
            "frac": flops / (gram_ms * 1e-3) / 1e12 / peak64, "traffic": None,
            "peak_source": "fable_ctx_measure_fp64_peak: DMMA.8x8x4 issue peak measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
            "algorithmic_gflop_per_launch": flops / 1e9, "avg_launch_ms": gram_ms}
    del dC
    dU = torch.from_numpy(np.ascontiguousarray(U[:, :kEst].T)).to(job.dev)
    c = np.asfortranarray(V[:, :kEst] * d[:kEst])                           # YtU = V_k o d_k (cpp:374-376)
    dc = torch.from_numpy(np.ascontiguousarray(c.T)).to(job.dev)
    ds = torch.empty(p, dtype=torch.float64, device=job.dev); dr = torch.empty(p, dtype=torch.float64, device=job.dev)
    rows_ms = timed(lambda: fb.api.dev_rowstats(ctx, dY.data_ptr(), n, p, dU.data_ptr(), dc.data_ptr(), p, kEst, ds.data_ptr(), dr.data_ptr()),
                    reps=20, warm=3)
    hbm, src = _peaks()
    nbytes = 8.0 * n * p + 8.0 * p * kEst
    rows = {"bound": "hbm", "kernel": "k_rowstats_mma (fable_dev_rowstats: residual + column sums of squares, one pass over Y)",
            "achieved": nbytes / (rows_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s", "frac": nbytes / (rows_ms * 1e-3) / 1e9 / hbm,
            "traffic": None, "peak_source": src, "algorithmic_mb_per_launch": nbytes / 1e6, "avg_launch_ms": rows_ms,
            "l2": "Y (%.0f MB) exceeds the 126 MB L2" % (8e-6 * n * p) if 8.0 * n * p > 126e6 else "Y fits L2: not an HBM figure"}
    return gram, rows


def _column_shard_check(job, ctx):
    """Partition C (column-sharded Gram/SVD, rank search, row posterior) over NCCL on the job's ranks, on a
    small problem, against the one-device path on rank 0: worst relative error of the sign-invariant outputs, or -1 if
    kEst / kMax differ.  Untimed; part of the multi-rank line so that the driver's scaling run exercises it."""
    import fable_b200 as fb
    from fable_b200 import synth, parallel
    n, p, k = 300, 4000, 5
    Y = synth.factor_data(n, p, k, rep=7)
    j0, j1 = parallel.column_range(p, job.rank, job.world)
    res = parallel.sharded_estimation(np.asfortranarray(Y[:, j0:j1]), p, j0, job.rank, job.world, ctx,
                                      parallel.host_allreduce(job.dev), allreduce_dev=parallel.device_allreduce(job.dev))
    worst = 0.0
    if job.rank == 0:
        one = fb.FABLEPosteriorSampler(Y, MC=1, ctx=ctx, want_G=False)
        rel = lambda a, b: float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / np.max(np.abs(b)))
        if res["estRank"] != one["estRank"]:
            worst = -1.0
        else:
            worst = max(rel(res["svd"]["d"], one["svdY"]["d"]), abs(res["varInflation"] / one["varInflation"] - 1),
                        rel(res["SigmaSqEstimate"], one["FABLEHyperParameters"]["SigmaSqEstimate"]),
                        abs(res["tauSqEstimate"] / one["FABLEHyperParameters"]["tauSqEstimate"] - 1))
    return job.max_over_ranks(abs(worst)) * (-1.0 if job.sum_over_ranks(1.0 if worst < 0 else 0.0) > 0 else 1.0)


def run_draws(args, w, wname, job):
    import fable_b200 as fb
    from fable_b200 import synth, parallel
    torch, dist = job.torch, job.dist
    world, rank, local, dev = job.world, job.rank, job.local, job.dev
    n, p, k = w["n"], w["p"], w["k"]
    B = args.batch or w["batch"]
    slots = B          # one dense p x p slot per draw of a batch (tile-major kernel order)

    # ---- untimed setup: data and the leading-k singular triplets (host LAPACK; the GPU estimation
    # pipeline -- SVD, rank search -- is timed as separate stages AFTER the timed regions and checked
    # against these inputs) ------------------------------------------------------------------------
    ctx = fb.Context(local, 12345)
    stages = {}
    Y = synth.factor_data(n, p, k, rep=0)
    kEst = k
    U, d, V = _leading_triplets(Y, kEst)

    def timed(name, fn):
        # host-timed stage through the C-ABI with pageable buffers; the first call of a process also pays one-off costs
        # (work-buffer allocation, graph instantiation, kernel attribute set-up), so it is reported beside the repeat
        ctx.synchronize(); t0 = time.perf_counter(); out = fn(); ctx.synchronize()
        stages[name.replace("_ms", "_first_call_ms")] = round(1e3 * (time.perf_counter() - t0), 3)
        ctx.synchronize(); t0 = time.perf_counter(); out = fn(); ctx.synchronize()
        stages[name] = round(1e3 * (time.perf_counter() - t0), 3)
        return out
    hyp = timed("hyper_ms", lambda: fb.FABLEHyperParameters(Y, U, V, d, kEst, ctx=ctx, want_G=False))
    vi = timed("var_inflation_ms", lambda: fb.var_inflation(hyp["SigmaSqEstimate"], hyp["G0"], "wrapper", ctx=ctx))
    stages.update(kEst=kEst, varInflation=vi)
    post = fb.Posterior(Y, 1.0, 1.0, U, V, d, kEst, vi, ctx=ctx)
    ring = torch.empty((slots, p, p), dtype=torch.float64, device=dev)
    post.accum_reset()
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)

    # ---- timed region 1: resident-in-HBM draws/s ---------------------------------------------------
    base = rank * 10_000_000           # disjoint draw ranges per rank (counter-based stream)
    m0 = base
    for _ in range(args.warmup):
        post.draw_batch(m0, B, dev_psi=ring.data_ptr(), psi_slots=slots, accumulate=True); m0 += B
    if world > 1:      # warm the reduction too: NCCL sets its rings / buffers up on the first call of a collective
        parallel.reduce_posterior_accumulators(post, dev, dst=0)
    post.accum_reset()
    ctx.synchronize()
    clocks = ClockSampler(local); clocks.start()
    ctx.kernel_timing(True)
    l0 = ctx.launches
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    job.barrier()
    m_timed = m0
    e0.record(stream)
    for _ in range(args.steps):
        post.draw_batch(m0, B, dev_psi=ring.data_ptr(), psi_slots=slots, accumulate=True); m0 += B
    a_sum, a_sq, a_cnt = post.accum_dev()            # folds the batches' partial sums that are still pending: inside the timed region
    e1.record(stream)
    ctx.synchronize()
    local_checksum = float(parallel.as_torch(a_sum, a_cnt, dev).sum().item()) if world > 1 else 0.0   # untimed (between the events)
    if world > 1:       # the one exchange step of the path: sum the accumulators over ranks (NCCL)
        ea = torch.cuda.Event(enable_timing=True); eb = torch.cuda.Event(enable_timing=True)
        ea.record()
        parallel.reduce_posterior_accumulators(post, dev, dst=0)      # compact upper tiles, in place -> rank 0
        eb.record(); torch.cuda.synchronize(dev)
        allreduce_ms = ea.elapsed_time(eb)
    else:
        allreduce_ms = 0.0
    job.barrier()
    dev_ms = e0.elapsed_time(e1) + allreduce_ms
    kern_ms, kern_launches = ctx.kernel_time()
    ctx.kernel_timing(False)
    launches = ctx.launches - l0
    clk = clocks.stop()
    dev_ms_max = job.max_over_ranks(dev_ms)
    value = world * args.steps * B / (dev_ms_max * 1e-3)

    # ---- correctness of the multi-GPU sum (untimed) ----------------------------------------------------
    # (i) checksum of the reduced accumulators on rank 0 against the sum over ranks of the per-rank checksums taken
    # before the reduce; (ii) a few tiles recomputed on rank 0 over EVERY rank's draw range with a one-tile slice
    reduce_check = None
    if world > 1:
        want = job.sum_over_ranks(local_checksum)
        worst = 0.0
        if rank == 0:
            red = parallel.as_torch(a_sum, a_cnt, dev)
            got = float(red.sum().item())
            worst = abs(got - want) / max(abs(want), 1e-300)
            rows = post.get()
            rows["sigmahat"] = rows["gnd"] / (rows["gamma_n"] - 2.0)
            chk = fb.Posterior.from_rows(n, 1.0, 1.0, rows, vi, ctx=ctx)
            nT = -(-p // 64)
            for (I, J) in ((0, 0), (1, nT - 1), (nT // 2, nT // 2 + 1)):
                slot, cidx = fb.api.tile_locate(p, I, J)
                chk.set_tile_range(slot, slot + 1)
                chk.accum_reset()
                for r in range(world):
                    mm = r * 10_000_000 + (m_timed - base)
                    for _ in range(args.steps):
                        chk.draw_batch(mm, B, accumulate=True); mm += B
                c_sum, _, c_cnt = chk.accum_dev()
                ctx.synchronize()
                a = parallel.as_torch(c_sum, c_cnt, dev)
                b = red[cidx * 4096:(cidx + 1) * 4096]
                worst = max(worst, float(((a - b).abs().max() / b.abs().max()).item()))
            chk.close()
        reduce_check = job.max_over_ranks(worst)

    shard_check = _column_shard_check(job, ctx) if world > 1 else None

    # ---- timed region 2: end to end through the C-ABI with host buffers --------------------------------
    mc = args.e2e_mc or w["e2e_mc"]                   # per rank (weak scaling, like `value`); pinned host memory scales with the ranks
    try:                                               # ... so it is halved where the host could not hold two copies of the buffers
        import psutil
        need = 8.0 * world * (mc * (kEst + 1) * p) + 16.0 * p * p
        while mc > 512 and 2.5 * need > psutil.virtual_memory().available:
            mc //= 2; need = 8.0 * world * (mc * (kEst + 1) * p) + 16.0 * p * p
    except Exception:
        pass
    mc = int(job.max_over_ranks(-float(mc)) * -1)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    ctx.set_psi_ring(slots)
    del ring
    torch.cuda.empty_cache()
    h2d = 8 * (n * p + n * kEst + p * kEst + kEst)
    d2h_samples = 8 * (mc * kEst * p + mc * p + 2 * p)
    d2h = d2h_samples + 8 * 2 * p * p                 # rank 0 (N > 1: the other ranks download their samples only)

    def e2e_leg(pinned, steps, m_start):
        alloc = (lambda *s: _pin(torch, *s).numpy()) if pinned else (lambda *s: np.empty(s))
        Yh = alloc(p, n); Yh[...] = Y.T; Yh = Yh.T                          # (p, n) row-major == n x p column-major
        Uh = alloc(kEst, n); Uh[...] = U[:, :kEst].T; Uh = Uh.T
        Vh = alloc(kEst, p); Vh[...] = V[:, :kEst].T; Vh = Vh.T
        dh = alloc(kEst); dh[...] = d[:kEst]

        def outputs():
            o = {"LambdaSamples": alloc(kEst * p, mc).T, "SigmaSqSamples": alloc(p, mc).T}
            if rank == 0:
                o.update(PsiSum=alloc(p, p).T, PsiSumSq=alloc(p, p).T)
            if world == 1:
                o.update(SigmaSqEstimatePostMean=alloc(p), SigmaSqEstimatePlugIn=alloc(p))
            return o

        def call(m_begin, o):
            if world == 1:
                return fb.CPPFABLESampler(Yh, 1.0, 1.0, mc, Uh, Vh, dh, kEst, vi, ctx=ctx, m0=m_begin, want_G=False,
                                          psi_moments=True, out=o)
            return parallel.distributed_sampler(Yh, 1.0, 1.0, mc * world, Uh, Vh, dh, kEst, vi, ctx, dev, m0=m_begin, out=o, dst=0)
        mm = m_start
        o = outputs()
        if pinned:
            call(mm, o); mm += mc * world                              # warm-up call
        job.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            if not pinned:
                o = outputs()                                          # R allocates its result vectors per call
            res = call(mm, o); mm += mc * world
        loss = float(res["PsiSum"][0, 0]) if rank == 0 else float(res["SigmaSqSamples"][0, 0])   # result read on the host
        torch.cuda.synchronize(dev)
        secs = job.max_over_ranks(time.perf_counter() - t0)
        return world * steps * mc / secs, mm, loss

    e2e_val, m0, _ = e2e_leg(True, e2e_steps, m0)
    try:
        torch._C._host_emptyCache()           # give the pinned buffers back before allocating pageable ones of the same size
    except Exception:
        pass
    e2e_pageable, m0, _ = (None, m0, None) if args.no_pageable else e2e_leg(False, 1, m0)
    ctx.set_psi_ring(0)

    # ---- accumulate-only mode (no materialisation), for information -------------------------------------
    post.accum_reset()
    for _ in range(2):
        post.draw_batch(m0, B, accumulate=True); m0 += B
    ctx.synchronize()
    e0.record(stream)
    for _ in range(args.steps):
        post.draw_batch(m0, B, accumulate=True); m0 += B
    e1.record(stream); ctx.synchronize()
    acc_only = args.steps * B / (e0.elapsed_time(e1) * 1e-3)
    post.close()

    # ---- estimation stages: rooflines of the Gram step and the row pass (timed in this run), wall times of the
    # SVD and the rank search ----------------------------------------------------------------------------
    gram = rows = None
    if rank == 0 and not args.no_stages:
        gram, rows = _estimation_rooflines(job, ctx, Y, U, V, d, kEst, stream)
        s = timed("svd_ms", lambda: fb.svd(Y, ctx=ctx))
        kMax = fb.kmax_rule(s["d"], 0.95)
        k_gpu = timed("rank_ms", lambda: fb.CPPRankEstimator(Y, s["u"], s["v"], s["d"], kMax, ctx=ctx))
        stages.update(kMax=kMax, kEst_gpu_pipeline=k_gpu,
                      d_relerr_vs_host_lapack=float(np.max(np.abs(s["d"][:kEst] - d) / d)))
        del s
        # the single-upload wrapper pipeline (FABLEPosteriorMean, R/FABLE_wrapper.R:80-128) with the library's own stage
        # marks: the rank search and the row posterior on resident data, apart from the host copies around them
        ctx.stage_timing(True)
        best = None
        for rep in range(4):                  # first call: warm-up; then the fastest of three (the p x p return into fresh pageable
            pm = fb.FABLEPosteriorMean(Y, ctx=ctx)          # memory depends on what the host is doing: 126-229 ms seen for the same code)
            rec = dict(total=round(1e3 * pm["runTime"], 3), kEst=int(pm["estRank"]), **ctx.stage_times())
            del pm
            if rep and (best is None or rec["total"] < best["total"]):
                best = rec
        stages["FABLEPosteriorMean_ms"] = best
        ctx.stage_timing(False)

    if rank == 0:
        peak, peak_src = _peaks()
        alg_bytes_per_launch = 8.0 * p * p * B + 8.0 * p * (kEst + 1) * B       # SURVEY §8(d)(i) x draws per launch
        avg_launch_s = (kern_ms / max(kern_launches, 1)) * 1e-3
        achieved = alg_bytes_per_launch / avg_launch_s / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get(wname)      # GB per launch, from the committed ncu capture
            except Exception:
                traffic = None
        config = {"workload": _workload_label(wname, n, p, kEst, "draws", world), "draws_per_step_per_gpu": B, "psi_slots": slots,
                  "mode": "materialise every Psi draw (dense p x p FP64) + sum/sumsq accumulators (compact upper tiles)",
                  "l2": "each step writes %.1f GB >> 126 MB L2; no flush needed" % (8e-9 * p * p * B),
                  "nccl_reduce_ms_in_timed_region": round(allreduce_ms, 3)}
        if reduce_check is not None:
            config["reduce_check_relerr"] = reduce_check
            config["column_shard_check_relerr"] = shard_check      # partition C over NCCL vs one device (-x: kEst differs)
        e2e = {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "draws_per_step_per_gpu": mc, "steps": e2e_steps,
               "call": "fable_sampler (CPPFABLESampler) with pinned host buffers" if world == 1 else
                       "per rank: fable_posterior_create + fable_posterior_sample_range (its share of the MC axis), "
                       "NCCL sum of the accumulators on the device, fable_posterior_accum_fetch on rank 0 only; pinned host buffers"}
        if world > 1:
            e2e["d2h_bytes_per_step_other_ranks"] = d2h_samples
        line = {
            "metric": metric_name(w), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config, "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": {"sm_mhz": clk["sm_mhz"], "sm_max_mhz": clk["sm_max_mhz"], "reasons": clk["reasons"]},
            "roofline": {"bound": "hbm", "kernel": "k_psi<MAT,ACC>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_unit": "GB per launch",
                         "traffic_source": "static: ncu --set full capture committed under profiles/ (not measured in this run)",
                         "peak_source": peak_src,
                         "algorithmic_gb_per_launch": alg_bytes_per_launch / 1e9, "avg_launch_ms": avg_launch_s * 1e3,
                         "launches_timed": int(kern_launches), "kernel_share_of_step": kern_ms / max(dev_ms - allreduce_ms, 1e-9)},
            "accumulate_only_samples_per_s_per_gpu": acc_only,
            "stages": stages,
        }
        if e2e_pageable is not None:
            line["e2e_pageable"] = {"value": e2e_pageable, "unit": UNIT, "steps": 1, "draws_per_step_per_gpu": mc,
                                    "call": "same call, pageable numpy inputs and freshly allocated outputs (what R hands over)"}
        if gram is not None:
            line["roofline_gram"] = gram; line["roofline_rows"] = rows
        if world == 1 and not args.no_cpu:
            v, secs, cores = cpu_sample_rate(w, w["ref_mc"])
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{w['ref_mc']} draws at full p={p} in {secs:.1f} s: sampler loop + Psi entries + "
                                              "entry-wise sum/sumsq (C restatement of cpp:599-618, 650-688; OpenMP)"}
        print(json.dumps(line))
    ctx.close()


def _slice_entries(p, lo, hi):
    """Entries of the dense symmetric p x p matrix that the slice [lo, hi) of the tile enumeration owns (both
    triangles), and its stored (valid) tiles -- the numpy statement of fable_b200/csrc/tiles.cuh."""
    nT = -(-p // 64)
    t = np.arange(lo, hi, dtype=np.int64)
    s, r = t // 256, t % 256
    sJ = ((np.sqrt(8.0 * s + 1.0) - 1.0) * 0.5).astype(np.int64)
    sJ = np.where(sJ * (sJ + 1) // 2 > s, sJ - 1, sJ); sJ = np.where((sJ + 1) * (sJ + 2) // 2 <= s, sJ + 1, sJ)
    sI = s - sJ * (sJ + 1) // 2
    I, J = sI * 16 + r % 16, sJ * 16 + r // 16
    ok = (I <= J) & (J < nT)
    I, J = I[ok], J[ok]
    hI = np.minimum(64, p - 64 * I); wJ = np.minimum(64, p - 64 * J)
    return int(np.sum(hI * wJ * np.where(I == J, 1, 2))), int(ok.sum())


def run_tiles(args, w, wname, job):
    """BASELINE configs[3]: every rank processes all draws for its own slice of the covariance tiles and keeps only
    that share, compactly (2 x tiles x 32 KB per draw: the block and its mirror image); no collective."""
    import fable_b200 as fb
    from fable_b200 import synth, parallel
    torch = job.torch
    world, rank, local, dev = job.world, job.rank, job.local, job.dev
    n, p, k = w["n"], w["p"], w["k"]
    ctx = fb.Context(local, 12345)
    Y = synth.factor_data(n, p, k, rep=0)
    kEst = k
    U, d, V = _leading_triplets(Y, kEst)
    hyp = fb.FABLEHyperParameters(Y, U, V, d, kEst, ctx=ctx, want_G=False)
    vi = fb.var_inflation(hyp["SigmaSqEstimate"], hyp["G0"], "wrapper", ctx=ctx)
    post = fb.Posterior(Y, 1.0, 1.0, U, V, d, kEst, vi, ctx=ctx)
    lo, hi = parallel.tile_range(p, rank, world)
    post.set_tile_range(lo, hi)
    entries, nt = _slice_entries(p, lo, hi)
    assert nt == fb.api.tile_count(p, lo, hi)
    slot_doubles = 2 * nt * 4096
    B = args.batch or max(1, min(48, int(RING_BUDGET_GB * 1e9 // (8 * slot_doubles))))
    B = int(job.max_over_ranks(-float(B)) * -1)                     # the same batch everywhere (min over ranks)
    ring = torch.empty(B * slot_doubles, dtype=torch.float64, device=dev)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)

    # ---- untimed check of this rank's share: a small batch against numpy on the returned samples (two owned tiles:
    # the materialised block, its stored mirror image, and the accumulated sums) ---------------------------
    Bc = min(B, 3)
    post.accum_reset()
    post.draw_batch_tiles(0, Bc, dev_psi_tiles=ring.data_ptr(), psi_slots=B, accumulate=True)
    ctx.synchronize()
    L, S = post.samples(0, Bc)
    Lr = L.reshape(Bc, p, kEst)
    s1, s2 = post.accum_fetch_tiles()
    tiles = ring[:Bc * slot_doubles].view(Bc, 2 * nt, 64, 64)
    worst = 0.0
    # first and last stored tile of the slice
    nT = -(-p // 64)
    owned = []
    for t in (lo, hi - 1):
        step = 1 if t == lo else -1
        while True:
            ss, r = divmod(t, 256)
            sJ = int((np.sqrt(8.0 * ss + 1.0) - 1.0) * 0.5)
            while sJ * (sJ + 1) // 2 > ss: sJ -= 1
            while (sJ + 1) * (sJ + 2) // 2 <= ss: sJ += 1
            I, J = (ss - sJ * (sJ + 1) // 2) * 16 + r % 16, sJ * 16 + r // 16
            if I <= J < nT:
                owned.append((I, J)); break
            t += step
    for (I, J) in owned:
        _, cg = fb.api.tile_locate(p, I, J)
        c = cg - fb.api.tile_locate(p, *owned[0])[1]
        ru, rv = slice(64 * I, min(p, 64 * I + 64)), slice(64 * J, min(p, 64 * J + 64))
        acc = np.zeros((ru.stop - ru.start, rv.stop - rv.start))
        for b in range(Bc):
            blk = Lr[b, ru, :] @ Lr[b, rv, :].T
            if I == J:
                blk = blk + np.diag(S[b, ru])
            acc += blk
            got = tiles[b, 2 * c].cpu().numpy().T[:blk.shape[0], :blk.shape[1]]          # [v][u] stored -> (u, v)
            worst = max(worst, float(np.max(np.abs(got - blk)) / np.max(np.abs(blk))))
            if I != J:
                gotm = tiles[b, 2 * c + 1].cpu().numpy()[:blk.shape[0], :blk.shape[1]]   # mirror image: block (J, I) = blk'
                worst = max(worst, float(np.max(np.abs(gotm - blk)) / np.max(np.abs(blk))))
        gots = s1[c].T[:acc.shape[0], :acc.shape[1]]
        worst = max(worst, float(np.max(np.abs(gots - acc)) / np.max(np.abs(acc))))
    tile_check = job.max_over_ranks(worst)
    del L, S, Lr, s1, s2

    # ---- timed region 1: resident-in-HBM draws/s of the whole job (all ranks process the same draws) -------
    m0 = 1000
    post.accum_reset()
    for _ in range(args.warmup):
        post.draw_batch_tiles(m0, B, dev_psi_tiles=ring.data_ptr(), psi_slots=B, accumulate=True); m0 += B
    post.accum_reset()
    ctx.synchronize()
    clocks = ClockSampler(local); clocks.start()
    ctx.kernel_timing(True)
    l0 = ctx.launches
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    job.barrier()
    e0.record(stream)
    for _ in range(args.steps):
        post.draw_batch_tiles(m0, B, dev_psi_tiles=ring.data_ptr(), psi_slots=B, accumulate=True); m0 += B
    post.accum_dev()                                  # pending partial sums are folded inside the timed region
    e1.record(stream)
    ctx.synchronize()
    job.barrier()
    dev_ms = e0.elapsed_time(e1)
    kern_ms, kern_launches = ctx.kernel_time()
    ctx.kernel_timing(False)
    launches = ctx.launches - l0
    clk = clocks.stop()
    dev_ms_max = job.max_over_ranks(dev_ms)
    value = args.steps * B / (dev_ms_max * 1e-3)                   # a draw is done when every rank has done its tiles
    alg_bytes = 8.0 * entries * B + 8.0 * p * (kEst + 1) * B        # this GPU's share of 8 p^2 per draw + the operands
    achieved = alg_bytes / ((kern_ms / max(kern_launches, 1)) * 1e-3) / 1e9
    frac_min = -job.max_over_ranks(-achieved)                       # the slowest GPU's figure goes on the line
    per_rank = [achieved]
    if world > 1:
        t = torch.tensor([achieved], dtype=torch.float64, device=dev)
        g = [torch.zeros_like(t) for _ in range(world)]
        job.dist.all_gather(g, t)
        per_rank = [float(x.item()) for x in g]
    post.close()

    # ---- timed region 2: end to end through the C-ABI with host buffers: inputs up, all draws over the slice,
    # the rank's share of sum / sumsq (compact tiles) back ------------------------------------------------
    mc = args.e2e_mc or w["e2e_mc"]
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    hY = _pin(torch, p, n); hY.numpy()[...] = Y.T
    hU = _pin(torch, kEst, n); hU.numpy()[...] = U[:, :kEst].T
    hV = _pin(torch, kEst, p); hV.numpy()[...] = V[:, :kEst].T
    hd = _pin(torch, kEst); hd.numpy()[...] = d[:kEst]
    o1 = _pin(torch, nt * 4096).numpy(); o2 = _pin(torch, nt * 4096).numpy()
    Yh, Uh, Vh, dh = hY.numpy().T, hU.numpy().T, hV.numpy().T, hd.numpy()

    def e2e_call(m_begin):
        pe = fb.Posterior(Yh, 1.0, 1.0, Uh, Vh, dh, kEst, vi, ctx=ctx)          # fable_posterior_create: H2D of Y, U_k, V_k, d
        try:
            pe.set_tile_range(lo, hi)
            pe.accum_reset()
            for b0 in range(0, mc, B):
                nb = min(B, mc - b0)
                pe.draw_batch_tiles(m_begin + b0, nb, dev_psi_tiles=ring.data_ptr(), psi_slots=B, accumulate=True)
            pe.accum_fetch_tiles(out=(o1, o2))                                   # D2H of the share
        finally:
            pe.close()
        return float(o1[0])
    e2e_call(m0); m0 += mc
    job.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        loss = e2e_call(m0); m0 += mc
    torch.cuda.synchronize(dev)
    e2e_val = e2e_steps * mc / job.max_over_ranks(time.perf_counter() - t0)

    if rank == 0:
        peak, peak_src = _peaks()
        line = {
            "metric": metric_name(w), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": _workload_label(wname, n, p, kEst, "tiles", world), "draws_per_step": B, "psi_slots": B,
                       "partition": "tiles: rank g owns slice g of the upper 64x64 tile enumeration (equal tile counts) and "
                                    "processes every draw for it; no collective",
                       "tiles_per_rank": nt, "ring_gb_per_rank": 8e-9 * B * slot_doubles,
                       "mode": "materialise the rank's share of every Psi draw (block + mirror image, compact) + sum/sumsq accumulators",
                       "l2": "each step writes %.1f GB per GPU >> 126 MB L2; no flush needed" % (8e-9 * entries * B),
                       "tile_check_relerr": tile_check},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": 8 * (n * p + n * kEst + p * kEst + kEst),
                    "d2h_bytes_per_step": 8 * 2 * nt * 4096, "draws_per_step": mc, "steps": e2e_steps,
                    "call": "per rank: fable_posterior_create + set_tile_range + draw_batch_tiles over all draws + "
                            "fable_posterior_accum_fetch_tiles (its share of sum / sumsq); pinned host buffers"},
            "gpu_launches": int(launches),
            "clocks": {"sm_mhz": clk["sm_mhz"], "sm_max_mhz": clk["sm_max_mhz"], "reasons": clk["reasons"]},
            "roofline": {"bound": "hbm", "kernel": "k_psi<MAT,ACC> (compact per-rank output)", "achieved": frac_min, "peak": peak,
                         "unit": "GB/s", "frac": frac_min / peak, "traffic": None, "peak_source": peak_src,
                         "per_gpu": "slowest rank's figure: its share of 8 p^2 bytes per draw / its kernel time",
                         "per_rank_frac": [round(x / peak, 4) for x in per_rank],
                         "algorithmic_gb_per_launch": alg_bytes / 1e9, "avg_launch_ms": kern_ms / max(kern_launches, 1),
                         "launches_timed": int(kern_launches), "kernel_share_of_step": kern_ms / max(dev_ms, 1e-9)},
        }
        if world == 1 and not args.no_cpu:
            v, secs, cores = cpu_sample_rate(w, w["ref_mc"])
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{w['ref_mc']} draws at full p={p} in {secs:.1f} s: sampler loop + Psi entries + "
                                              "entry-wise sum/sumsq (C restatement of cpp:599-618, 650-688; OpenMP)"}
        print(json.dumps(line))
    ctx.close()


def run_b200(args, w, wname):
    job = Job()
    try:
        if _partition(args, wname) == "tiles":
            run_tiles(args, w, wname, job)
        else:
            run_draws(args, w, wname, job)
    finally:
        job.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--partition", default="", choices=["", "draws", "tiles"],
                    help="multi-GPU partition (default: draws; tiles for c4)")
    ap.add_argument("--batch", type=int, default=0, help="draws per step per GPU (default: per workload)")
    ap.add_argument("--e2e-mc", type=int, default=0, help="draws per end-to-end call (per rank)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-stages", action="store_true", help="skip the Gram / row-pass rooflines and the SVD / rank-search stage timings")
    ap.add_argument("--no-pageable", action="store_true", help="skip the pageable end-to-end leg")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w, args.workload)
    else:
        run_b200(args, w, args.workload)


if __name__ == "__main__":
    main()
