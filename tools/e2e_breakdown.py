"""Where the end-to-end fable_sampler call spends its time (C3, pinned host buffers): wall clock of the whole
call, of the same call without the covariance pass, and the summed device time of the covariance kernels.
usage: python tools/e2e_breakdown.py [mc] [slots]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fable_b200 as fb
from fable_b200 import synth

mc = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
slots = int(sys.argv[2]) if len(sys.argv) > 2 else 40
n, p, k = 1000, 20000, 10
Y = synth.factor_data(n, p, k, rep=0)
lam, Q = np.linalg.eigh(Y @ Y.T)
U = np.asfortranarray(Q[:, ::-1][:, :k]); T = Y.T @ U; d = np.linalg.norm(T, axis=0); V = np.asfortranarray(T / d)
ctx = fb.Context(0, 12345)
pin = lambda *shape: torch.empty(shape, dtype=torch.float64).pin_memory()
hY = pin(p, n); hY.numpy()[...] = Y.T
hU = pin(k, n); hU.numpy()[...] = U.T
hV = pin(k, p); hV.numpy()[...] = V.T
hd = pin(k); hd.numpy()[...] = d
out_t = {"LambdaSamples": pin(k * p, mc), "SigmaSqSamples": pin(p, mc), "SigmaSqEstimatePostMean": pin(p),
         "SigmaSqEstimatePlugIn": pin(p), "PsiSum": pin(p, p), "PsiSumSq": pin(p, p)}
out = {kk: (vv.numpy().T if vv.dim() == 2 else vv.numpy()) for kk, vv in out_t.items()}
Yh, Uh, Vh, dh = hY.numpy().T, hU.numpy().T, hV.numpy().T, hd.numpy()
ctx.set_psi_ring(slots)

def call(moments):
    return fb.CPPFABLESampler(Yh, 1.0, 1.0, mc, Uh, Vh, dh, k, 4.9, ctx=ctx, m0=0, want_G=False, psi_moments=moments, out=out)

call(True)
for rep in range(2):
    ctx.kernel_timing(True)
    ctx.synchronize(); t0 = time.perf_counter(); call(True); ctx.synchronize(); t_full = time.perf_counter() - t0
    ms, cnt = ctx.kernel_time()
    ctx.kernel_timing(False)
    ctx.synchronize(); t0 = time.perf_counter(); call(False); ctx.synchronize(); t_samp = time.perf_counter() - t0
    print(f"mc={mc} slots={slots}: full call {t_full:.3f} s ({mc / t_full:.0f} draws/s); covariance kernels {ms / 1e3:.3f} s in {cnt} launches "
          f"({mc / ms * 1e3:.0f} draws/s); samples-only call {t_samp:.3f} s; full - kernels = {t_full - ms / 1e3:.3f} s")
