"""A small tour of the C-ABI for compute-sanitizer (memcheck): every kernel family once, tiny shapes, no torch.
usage: compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fable_b200 as fb
from fable_b200 import synth

ctx = fb.Context(0, 7)
n, p, k = 48, 150, 3
Y = synth.factor_data(n, p, k, rep=1)
Yc = np.asfortranarray(Y - Y.mean(axis=0, keepdims=True))                  # rank-deficient: deflation + completion
s = fb.svd(Yc, ctx=ctx)
assert s["d"][-1] == 0.0
pm = fb.FABLEPosteriorMean(Y, ctx=ctx)                                      # Gram, Jacobi, rank search (steered), row pass, k_psi
s = pm["svdY"]
kMax = fb.kmax_rule(s["d"], 0.95)
kk = fb.CPPRankEstimator(Y, s["u"], s["v"], s["d"], kMax, ctx=ctx)
fb.RankEstimator(Y, s, kMax, ctx=ctx)
hyp = fb.FABLEHyperParameters(Y, s["u"], s["v"], s["d"], kk, ctx=ctx)
fb.CPPcov_correct_matrix(hyp["SigmaSqEstimate"], hyp["G"], ctx=ctx)
vi = fb.var_inflation(hyp["SigmaSqEstimate"], hyp["G0"], "wrapper", ctx=ctx)
for ring in (0, 4):                                                         # accumulate-only and materialising + partial sums
    ctx.set_psi_ring(ring)
    out = fb.CPPFABLESampler(Y, 1.0, 1.0, 21, s["u"], s["v"], s["d"], kk, vi, ctx=ctx, psi_moments=True)
    assert np.all(np.isfinite(out["PsiSum"]))
ctx.set_psi_ring(0)
post = fb.Posterior(Y, 1.0, 1.0, s["u"], s["v"], s["d"], kk, vi, ctx=ctx)
pp = post.post_processing(40, 0.1)                                           # tile-resident quantiles
pp2 = post.post_processing(40, 0.1, SelectedIndices=[3, 9, 3, 150])
L, S = post.samples(0, 40)
fb.CPPCCFABLEPostProcessing({"LambdaSamples": L, "SigmaSqSamples": S, "EstimatedRank": kk}, 0.1, ctx=ctx)
os.environ["FABLE_B200_POST_TILES"] = "0"
fb.CCFABLEPostProcessingSubmatrix({"LambdaSamples": L, "SigmaSqSamples": S, "EstimatedRank": kk}, 0.1, [3, 9, 3, 150], ctx=ctx)   # per-pair kernels
lo, hi = fb.api.tile_partition(p, 1, 2)
post.set_tile_range(lo, hi)
post.draw_batch(0, 3, accumulate=True)
t1, t2 = post.accum_fetch_tiles()
post.close()
fb.CCFABLE_DirectSampler(Y, 1.0, 1.0, 3, ctx=ctx)
fb.CPPsvd(Y[:, :20], 1, ctx=ctx)
fb.CPPLogLikelihoodEval(Y, np.sqrt(n) * s["u"][:, :kk], hyp["G0"], hyp["SigmaSqEstimate"], ctx=ctx)
print("sanitize tour ok, launches:", ctx.launches)
ctx.close()
