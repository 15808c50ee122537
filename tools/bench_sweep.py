"""BASELINE.json configs[4]: rank estimation + posterior mean over p in {2000, 10000, 50000} x n in {200, 1000, 5000}
through the single-call wrapper (fable_wrapper_posterior_mean, host buffers), with the stage split and the
Gram step's FP64 tensor-core fraction.  Prints one markdown table."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import fable_b200 as fb
from fable_b200 import synth

PEAK = 37.0
ctx = fb.Context(0, 1)
stream = torch.cuda.ExternalStream(ctx.stream)
shapes = [(int(a), int(b)) for a, b in (s.split("x") for s in sys.argv[1:])] or \
    [(n, p) for n in (200, 1000, 5000) for p in (2000, 10000, 50000)]
print("| n | p | kMax (0.5) | kEst | Gram ms | Gram TFLOP/s (m(m+1)K) | of DMMA peak | SVD total s | rank + CovPostMean s | wrapper total s | sum d^2 / ||Y||_F^2 - 1 |")
print("|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
for n, p in shapes:
    Y = synth.factor_data(n, p, 10, rep=0)
    m, K = min(n, p), max(n, p)
    dY = torch.from_numpy(np.ascontiguousarray(Y.T)).cuda()          # (p, n) row-major == n x p column-major
    G = torch.empty((m, m), dtype=torch.float64, device="cuda")
    run = (lambda: fb.api.dev_syrk(ctx, 0, n, p, 1.0, dY.data_ptr(), n, G.data_ptr())) if n <= p else \
          (lambda: fb.api.dev_syrk(ctx, 1, p, n, 1.0, dY.data_ptr(), n, G.data_ptr()))
    for _ in range(2): run()
    ctx.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(5): run()
    e1.record(stream); ctx.synchronize()
    gram_ms = e0.elapsed_time(e1) / 5
    tf = m * (m + 1) * K / (gram_ms * 1e-3) / 1e12
    del dY, G
    torch.cuda.empty_cache()
    t0 = time.perf_counter(); s = fb.svd(Y, ctx=ctx); ctx.synchronize(); t_svd = time.perf_counter() - t0
    frob = float(np.sum(s["d"] ** 2) / np.sum(Y * Y) - 1.0)
    kMax = fb.kmax_rule(s["d"], 0.5)
    t0 = time.perf_counter(); pm = fb.CPPFABLEPostMean(Y, 1.0, 1.0, s["u"], s["v"], s["d"], kMax, ctx=ctx, want_cov=(p <= 20000)); ctx.synchronize()
    t_pm = time.perf_counter() - t0
    del s
    if p <= 20000:
        t0 = time.perf_counter(); w = fb.FABLEPosteriorMean(Y, ctx=ctx); ctx.synchronize(); t_w = time.perf_counter() - t0
        assert w["estRank"] == pm["kEst"]
        del w
    else:
        t_w = float("nan")
    print(f"| {n} | {p} | {kMax} | {pm['kEst']} | {gram_ms:.3f} | {tf:.2f} | {tf / PEAK:.2f} | {t_svd:.3f} | {t_pm:.3f} | {t_w:.3f} | {frob:.1e} |", flush=True)
