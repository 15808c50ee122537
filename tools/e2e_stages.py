"""Wall clock of the estimation entry points at one configuration with PAGEABLE host buffers (what R hands over),
GPU side only (tools/stage_table.py puts the CPU restatement beside them).  best of 3 after one warm call.
usage: python tools/e2e_stages.py [c3|c2|c1]        FABLE_B200_COPY_THREADS=0: driver copies (A/B)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fable_b200 as fb
from fable_b200 import synth

CFG = {"c1": (500, 1000, 10), "c2": (1000, 5000, 10), "c3": (1000, 20000, 10)}
name = sys.argv[1] if len(sys.argv) > 1 else "c3"
n, p, k = CFG[name]
ctx = fb.Context(0, 12345)
Y = synth.factor_data(n, p, k, rep=0)

def tm(fn, reps=3):
    fn(); best = 1e9
    for _ in range(reps):
        ctx.synchronize(); t0 = time.perf_counter(); out = fn(); ctx.synchronize()
        best = min(best, time.perf_counter() - t0)
    return out, best

s, t = tm(lambda: fb.svd(Y, ctx=ctx)); print(f"{name} svd                      {t * 1e3:8.1f} ms")
u, d, v = s["u"], s["d"], s["v"]
kMax = fb.kmax_rule(d, 0.95); kMax50 = fb.kmax_rule(d, 0.5)
kk, t = tm(lambda: fb.CPPRankEstimator(Y, u, v, d, kMax, ctx=ctx)); print(f"{name} rank (kMax={kMax})          {t * 1e3:8.1f} ms  -> {kk}")
_, t = tm(lambda: fb.CPPCriterionFunction(kk, Y, u, v, d, ctx=ctx)); print(f"{name} criterion(k={kk})           {t * 1e3:8.1f} ms")
_, t = tm(lambda: fb.FABLEHyperParameters(Y, u, v, d, kk, ctx=ctx, want_G=False)); print(f"{name} hyper (no G)              {t * 1e3:8.1f} ms")
_, t = tm(lambda: fb.CPPFABLEPostMean(Y, 1.0, 1.0, u, v, d, kMax50, ctx=ctx)); print(f"{name} CPPFABLEPostMean          {t * 1e3:8.1f} ms")
_, t = tm(lambda: fb.FABLEPosteriorMean(Y, ctx=ctx)); print(f"{name} FABLEPosteriorMean        {t * 1e3:8.1f} ms")
_, t = tm(lambda: fb.FABLEPosteriorSampler(Y, MC=1000, ctx=ctx)); print(f"{name} FABLEPosteriorSampler 1000 {t * 1e3:8.1f} ms")
