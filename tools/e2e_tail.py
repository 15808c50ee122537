"""Where the end-to-end time of fable_sampler goes at c3 (pinned host buffers): combinations of outputs, tail on / off.
usage: python tools/e2e_tail.py [MC]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fable_b200 as fb
from fable_b200 import synth


def main():
    MC = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    n, p, k, B = 1000, 20000, 10, 40
    Y = synth.factor_data(n, p, k, rep=0)
    lam, Q = np.linalg.eigh(Y @ Y.T)
    U = np.asfortranarray(Q[:, ::-1][:, :k]); T = Y.T @ U
    d = np.linalg.norm(T, axis=0); V = np.asfortranarray(T / d)
    ctx = fb.Context(0, 12345)
    pin = lambda *s: torch.empty(s, dtype=torch.float64).pin_memory().numpy()
    out = {"LambdaSamples": pin(k * p, MC).T, "SigmaSqSamples": pin(p, MC).T, "PsiSum": pin(p, p).T, "PsiSumSq": pin(p, p).T,
           "SigmaSqEstimatePostMean": pin(p), "SigmaSqEstimatePlugIn": pin(p)}
    ctx.set_psi_ring(B)
    rows = []
    for tail in ("256", "0"):
        os.environ["FABLE_B200_TAIL_DRAWS"] = tail
        for samples, moments in ((True, True), (False, True), (True, False)):
            best = None
            for rep in range(3):
                ctx.synchronize(); t0 = time.perf_counter()
                fb.CPPFABLESampler(Y, 1.0, 1.0, MC, U, V, d, k, 1.26, ctx=ctx, m0=rep * MC, want_G=False, want_samples=samples,
                                   psi_moments=moments, out=out)
                dt = time.perf_counter() - t0
                best = dt if best is None or dt < best else best
            rows.append((tail, samples, moments, best))
            print(f"tail={tail:>3} samples={samples!s:5} moments={moments!s:5}: {best:.3f} s  ({MC / best:.0f} draws/s)", flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
