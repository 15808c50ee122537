"""Gram step timing: fable_dev_syrk on Y Y' (n <= p) / Y'Y shapes of BASELINE config 5.
Reports algorithmic SYRK TFLOP/s = m (m+1) K / t   against the measured DMMA peak."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fable_b200 as fb

PEAK = 37.0
ctx = fb.Context(0, 1)
stream = torch.cuda.ExternalStream(ctx.stream)
shapes = [(int(a), int(b)) for a, b in (s.split("x") for s in sys.argv[1:])] or \
    [(200, 2000), (200, 10000), (200, 50000), (1000, 2000), (1000, 10000), (1000, 20000), (1000, 50000), (5000, 10000), (5000, 50000)]
out = []
for n, p in shapes:
    Y = torch.randn((p, n), dtype=torch.float64, device="cuda")      # (p, n) row-major == n x p column-major
    m, K = min(n, p), max(n, p)
    G = torch.empty((m, m), dtype=torch.float64, device="cuda")
    def run():
        if n <= p:
            fb.api.dev_syrk(ctx, 0, n, p, 1.0, Y.data_ptr(), n, G.data_ptr())
        else:
            fb.api.dev_syrk(ctx, 1, p, n, 1.0, Y.data_ptr(), n, G.data_ptr())
    torch.cuda.synchronize()
    for _ in range(3): run()
    ctx.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record(stream)
    for _ in range(reps): run()
    e1.record(stream); ctx.synchronize()
    ms = e0.elapsed_time(e1) / reps
    ref = (Y.T @ Y) if n <= p else (Y @ Y.T)          # torch (cuBLAS) as a yardstick for correctness only
    err = float((G - ref).abs().max() / ref.abs().max())
    tf = m * (m + 1) * K / (ms * 1e-3) / 1e12
    out.append(dict(n=n, p=p, ms=round(ms, 4), syrk_tflops=round(tf, 2), frac_of_dmma_peak=round(tf / PEAK, 3), relerr=err))
    print(out[-1], flush=True)
print(json.dumps(out))
