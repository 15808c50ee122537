#!/usr/bin/env python
"""bench.py -- posterior samples/sec of the FABLE sampling hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            (N=1: plain python; N>1: torchrun)
    python bench.py --impl reference ...                     (CPU restatement on the host cores)

One STEP = one batch of B posterior draws per GPU through the hot path:
    Philox sampler kernel  (sigma2_m, Lambda_m; reference cpp:599-618)
 -> covariance kernel      (Psi_m = Lambda_m Lambda_m' + diag(sigma2_m) formed in 64x64 blocks,
                            every draw MATERIALISED as a dense p x p FP64 matrix into an HBM ring,
                            and folded into entry-wise sum / sum-of-squares accumulators;
                            reference entry formula cpp:672-684, mean cpp:688).
`value`  = draws/s with the row posterior already resident in HBM (CUDA events, max over ranks).
`e2e`    = the same metric through the reference-facing C-ABI call `fable_sampler`
           (= CPPFABLESampler, cpp:532-629) with HOST buffers: Y, U, V, d uploaded from pinned memory
           and LambdaSamples, SigmaSqSamples, Psi sum / sumsq read back, every step.
Multi-GPU: draws are sharded over ranks (weak scaling: B draws per GPU per step); the only
collective is one NCCL sum (reduce to rank 0 of the packed upper triangles of the accumulators) at the
end of the timed region.
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
    "c4": dict(n=500, p=50000, k=20, batch=4, e2e_mc=256, ref_mc=4),     # 20 GB per p x p: 4 slots + 2 accumulators
}
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
    # leading k triplets only (the sampler needs nothing else): Gram + eigh is exact enough for a timing input
    G = Y @ Y.T
    lam, Q = np.linalg.eigh(G)
    U = Q[:, ::-1][:, :k].copy()
    T = Y.T @ U
    d = np.linalg.norm(T, axis=0)
    V = T / d
    st = orc._shrunk_stats(Y, U, V, d, k, 1.0, 1.0)
    return st, n, p, k


def cpu_sample_rate(w, mc, repeats=1):
    """draws/s of the CPU restatement: sampler loop (cpp:599-618) + Psi entry formation and
    entry-wise summaries over the draws (cpp:650-688), `mc` draws at full p."""
    from oracle import c_oracle
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
    return mc / best, best, c_oracle.num_threads()


def run_reference(args, w, wname):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    mc = w["ref_mc"]
    from oracle import c_oracle
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
    cores = c_oracle.num_threads()
    sample = f"{mc} draws per step at full p={p}: sampler loop + Psi entries + entry-wise sum/sumsq (C restatement, OpenMP)"
    print(json.dumps({
        "impl": "reference", "metric": metric_name(w), "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{wname}: n={n}, p={p}, k={k}", "draws_per_step": mc},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def run_b200(args, w, wname):
    import torch
    import torch.distributed as dist
    import fable_b200 as fb
    from fable_b200 import synth, parallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        parallel.share_host_threads(int(os.environ.get("LOCAL_WORLD_SIZE", world)))
        torch.cuda.set_device(local)
        # NCCL prints its version banner to stdout when the communicator is created: keep stdout to the one
        # JSON line by pointing fd 1 at stderr while the group and its first collective come up
        sys.stdout.flush()
        saved = os.dup(1); os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize(local)
        finally:
            sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)
    else:
        torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    n, p, k = w["n"], w["p"], w["k"]
    B = args.batch or w["batch"]
    slots = B          # one dense p x p slot per draw of a batch (tile-major kernel order)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- untimed setup: data and the leading-k singular triplets (host LAPACK; the GPU estimation
    # pipeline -- SVD, rank search -- is timed as separate stages AFTER the timed regions and checked
    # against these inputs) ------------------------------------------------------------------------
    ctx = fb.Context(local, 12345)
    stages = {}
    Y = synth.factor_data(n, p, k, rep=0)
    kEst = k
    lam, Q = np.linalg.eigh(Y @ Y.T if n <= p else Y.T @ Y)
    if n <= p:
        U = np.asfortranarray(Q[:, ::-1][:, :kEst]); T = Y.T @ U
        d = np.linalg.norm(T, axis=0); V = np.asfortranarray(T / d)
    else:
        V = np.asfortranarray(Q[:, ::-1][:, :kEst]); T = Y @ V
        d = np.linalg.norm(T, axis=0); U = np.asfortranarray(T / d)

    def timed(name, fn):
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
    m0 = rank * 10_000_000           # disjoint draw ranges per rank (counter-based stream)
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
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        post.draw_batch(m0, B, dev_psi=ring.data_ptr(), psi_slots=slots, accumulate=True); m0 += B
    e1.record(stream)
    ctx.synchronize()
    if world > 1:       # the one exchange step of the path: sum the accumulators over ranks (NCCL)
        ea = torch.cuda.Event(enable_timing=True); eb = torch.cuda.Event(enable_timing=True)
        ea.record()
        parallel.reduce_posterior_accumulators(post, dev, dst=0)      # packed upper triangles -> rank 0
        eb.record(); torch.cuda.synchronize(dev)
        allreduce_ms = ea.elapsed_time(eb)
    else:
        allreduce_ms = 0.0
    barrier()
    dev_ms = e0.elapsed_time(e1) + allreduce_ms
    kern_ms, kern_launches = ctx.kernel_time()
    ctx.kernel_timing(False)
    launches = ctx.launches - l0
    clk = clocks.stop()
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max = float(t.item())
    value = world * args.steps * B / (dev_ms_max * 1e-3)

    # ---- timed region 2: end to end through fable_sampler with host buffers ---------------------------
    mc = args.e2e_mc or (w["e2e_mc"] if world == 1 else max(256, w["e2e_mc"] // 2))   # pinned host memory scales with ranks
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    pin = lambda *shape: torch.empty(shape, dtype=torch.float64).pin_memory()
    hY = pin(p, n); hY.numpy()[...] = Y.T                      # (p, n) row-major == n x p column-major
    hU = pin(kEst, n); hU.numpy()[...] = U[:, :kEst].T
    hV = pin(kEst, p); hV.numpy()[...] = V[:, :kEst].T
    hd = pin(kEst); hd.numpy()[...] = d[:kEst]
    out_t = {"LambdaSamples": pin(kEst * p, mc), "SigmaSqSamples": pin(p, mc), "SigmaSqEstimatePostMean": pin(p),
             "SigmaSqEstimatePlugIn": pin(p), "PsiSum": pin(p, p), "PsiSumSq": pin(p, p)}
    out = {kk: (vv.numpy().T if vv.dim() == 2 else vv.numpy()) for kk, vv in out_t.items()}
    Yh, Uh, Vh, dh = hY.numpy().T, hU.numpy().T, hV.numpy().T, hd.numpy()
    h2d = 8 * (n * p + n * kEst + p * kEst + kEst)
    d2h = 8 * (mc * kEst * p + mc * p + 2 * p + 2 * p * p)
    ctx.set_psi_ring(slots)
    del ring
    torch.cuda.empty_cache()

    def e2e_call(m_begin):
        return fb.CPPFABLESampler(Yh, 1.0, 1.0, mc, Uh, Vh, dh, kEst, vi, ctx=ctx, m0=m_begin, want_G=False,
                                  psi_moments=True, out=out)
    e2e_call(m0); m0 += mc                                      # warm-up call
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        res = e2e_call(m0); m0 += mc
    loss = float(res["PsiSum"][0, 0])                           # result read on the host
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = world * e2e_steps * mc / float(t.item())
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

    # ---- estimation stages (untimed for the metric; wall times for information) ------------------------
    if rank == 0 and not args.no_stages:
        post.close()
        s = timed("svd_ms", lambda: fb.svd(Y, ctx=ctx))
        kMax = fb.kmax_rule(s["d"], 0.95)
        k_gpu = timed("rank_ms", lambda: fb.CPPRankEstimator(Y, s["u"], s["v"], s["d"], kMax, ctx=ctx))
        stages.update(kMax=kMax, kEst_gpu_pipeline=k_gpu,
                      d_relerr_vs_host_lapack=float(np.max(np.abs(s["d"][:kEst] - d) / d)))
        del s

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
        line = {
            "metric": metric_name(w), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{wname}: n={n}, p={p}, k={kEst}; draws sharded over ranks" if world > 1 else
                       f"{wname}: n={n}, p={p}, k={kEst}", "draws_per_step_per_gpu": B, "psi_slots": slots,
                       "mode": "materialise every Psi draw (dense p x p FP64) + sum/sumsq accumulators",
                       "l2": "each step writes %.1f GB >> 126 MB L2; no flush needed" % (8e-9 * p * p * B),
                       "nccl_reduce_ms_in_timed_region": round(allreduce_ms, 3)},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "draws_per_step_per_gpu": mc, "steps": e2e_steps, "call": "fable_sampler (CPPFABLESampler) with pinned host buffers"},
            "gpu_launches": int(launches),
            "clocks": {"sm_mhz": clk["sm_mhz"], "sm_max_mhz": clk["sm_max_mhz"], "reasons": clk["reasons"]},
            "roofline": {"bound": "hbm", "kernel": "k_psi<MAT,ACC>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_unit": "GB per launch (ncu)", "peak_source": peak_src,
                         "algorithmic_gb_per_launch": alg_bytes_per_launch / 1e9, "avg_launch_ms": avg_launch_s * 1e3,
                         "launches_timed": int(kern_launches), "kernel_share_of_step": kern_ms / max(dev_ms - allreduce_ms, 1e-9)},
            "accumulate_only_samples_per_s_per_gpu": acc_only,
            "stages": stages,
        }
        if world == 1 and not args.no_cpu:
            v, secs, cores = cpu_sample_rate(w, w["ref_mc"])
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{w['ref_mc']} draws at full p={p} in {secs:.1f} s: sampler loop + Psi entries + "
                                              "entry-wise sum/sumsq (C restatement of cpp:599-618, 650-688; OpenMP)"}
        print(json.dumps(line))
    post.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="draws per step per GPU (default: per workload)")
    ap.add_argument("--e2e-mc", type=int, default=0, help="draws per end-to-end fable_sampler call")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-stages", action="store_true", help="skip the SVD / rank-search stage timings")
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
