import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
import fable_b200 as fb
from fable_b200 import synth
from oracle import fable_oracle as orc
ctx = fb.Context(0, 1)
for (n, p) in [(1000, 5000), (1000, 20000)]:
    Y = synth.factor_data(n, p, 10, rep=0)
    u, d, v = orc.svd_thin(Y)
    kMax = orc.kmax_rule(d, 0.95)
    for rep in range(3):
        ctx.synchronize(); t0 = time.perf_counter(); l0 = ctx.launches
        k = fb.CPPRankEstimator(Y, u, v, d, kMax, ctx=ctx)
        ctx.synchronize(); print(n, p, "rank", k, "kMax", kMax, round(time.perf_counter() - t0, 4), "s launches", ctx.launches - l0, flush=True)
    for kk in (1, 10, 33, 457, 900):
        ctx.synchronize(); t0 = time.perf_counter()
        fb.CPPCriterionFunction(kk, Y, u, v, d, ctx=ctx)
        ctx.synchronize(); print("  criterion k=", kk, round(time.perf_counter() - t0, 4), flush=True)
