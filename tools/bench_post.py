"""Time the tile-resident posterior summaries (f1: CPPCCFABLEPostProcessing, cpp:631-713) on resident draws.
usage: python tools/bench_post.py [n p k MC]   (default BASELINE configs[1]: 1000 5000 10 5000)
Prints wall time of fable_posterior_post_processing (draw generation + selection + 3 p x p results to the host) and the
achieved rate in entry-values formed per second; checks a few entries against numpy."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fable_b200 as fb
from fable_b200 import synth


def main():
    n, p, k, MC = (int(a) for a in sys.argv[1:5]) if len(sys.argv) > 4 else (1000, 5000, 10, 5000)
    alpha = 0.05
    Y = synth.factor_data(n, p, k, rep=0)
    lam, Q = np.linalg.eigh(Y @ Y.T)
    U = np.asfortranarray(Q[:, ::-1][:, :k]); T = Y.T @ U
    d = np.linalg.norm(T, axis=0); V = np.asfortranarray(T / d)
    ctx = fb.Context(0, 12345)
    post = fb.Posterior(Y, 1.0, 1.0, U, V, d, k, 1.26, ctx=ctx)
    post.post_processing(min(MC, 64), alpha)                # warm-up (module load, buffers)
    for rep in range(2):
        ctx.synchronize(); t0 = time.perf_counter()
        got = post.post_processing(MC, alpha)
        dt = time.perf_counter() - t0
        vals = p * (p + 1) / 2 * MC
        print(f"post_processing n={n} p={p} k={k} MC={MC}: {dt:.3f} s  ({vals / dt / 1e9:.1f} G entry-values/s per pass-equivalent; "
              f"{3 * 8 * p * p / 1e9:.2f} GB of results to the host)")
    L, S = post.samples(0, MC)
    Lr = L.reshape(MC, p, k)
    rng = np.random.default_rng(0)
    worst = 0.0
    for _ in range(50):
        a, b = rng.integers(0, p, 2)
        v = np.einsum("ml,ml->m", Lr[:, a, :], Lr[:, b, :]) + (S[:, a] if a == b else 0.0)
        for key, pr in (("LowerQuantileMatrix", alpha / 2), ("UpperQuantileMatrix", 1 - alpha / 2)):
            want = np.quantile(v, pr, method="hazen")
            worst = max(worst, abs(got[key][a, b] - want) / max(abs(want), 1e-300))
    print(f"max rel err of 100 sampled quantiles vs numpy (hazen = Hyndman-Fan 5): {worst:.2e}")
    post.close(); ctx.close()


if __name__ == "__main__":
    main()
