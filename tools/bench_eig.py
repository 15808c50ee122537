"""Eigensolver timing: fable_dev_syevj on the n x n Gram of synthetic factor data (the a1 step after the SYRK).
Prints sweeps, ms, eigenvalue error vs numpy eigh and the orthogonality of Q.
usage: python tools/bench_eig.py [n x p ...]        FABLE_B200_BJ_TOL / FABLE_B200_BJ_INNER: experiments"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fable_b200 as fb
from fable_b200 import synth

ctx = fb.Context(0, 1)
stream = torch.cuda.ExternalStream(ctx.stream)
shapes = [tuple(int(x) for x in s.split("x")) for s in sys.argv[1:]] or [(1000, 20000), (1000, 5000), (500, 1000), (2000, 10000)]
for n, p in shapes:
    Y = synth.factor_data(n, p, 10, rep=0)
    G = Y @ Y.T
    lam = np.linalg.eigvalsh(G)[::-1]
    dG0 = torch.from_numpy(G).cuda()
    dW = torch.empty(n, dtype=torch.float64, device="cuda"); dQ = torch.empty((n, n), dtype=torch.float64, device="cuda")
    best = 1e9
    for rep in range(3):
        dG = dG0.clone(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        sweeps = fb.api.dev_syevj(ctx, dG.data_ptr(), n, dW.data_ptr(), dQ.data_ptr())
        e1.record(stream); ctx.synchronize()
        best = min(best, e0.elapsed_time(e1))
    W = dW.cpu().numpy(); Q = dQ.cpu().numpy().T          # columns of the column-major Q
    err = float(np.max(np.abs(W - lam) / lam)); orth = float(np.max(np.abs(Q.T @ Q - np.eye(n))))
    res = float(np.max(np.abs(G @ Q - Q * W)) / lam[0])
    print(f"n={n} p={p}: sweeps {sweeps}  {best:.1f} ms  eig relerr {err:.2e}  orth {orth:.2e}  resid {res:.2e}", flush=True)
