"""Row-pass timing: fable_dev_rowstats (per-row conjugate posterior / small-rank criterion residual) on resident inputs.
HBM-read bound: algorithmic bytes = 8 n p (SURVEY 8d "Row-posterior"); peak = MEASURED_PEAKS.json hbm_gbs.
Y (>= 160 MB at C3) is larger than the 126 MB L2, and three different Y buffers are cycled, so every pass reads HBM.

usage: python tools/bench_rows.py [n x p x k ...]      FABLE_B200_ROWS_MMA=0 times the DFMA kernel"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fable_b200 as fb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK = 6547.2
ctx = fb.Context(0, 1)
stream = torch.cuda.ExternalStream(ctx.stream)
shapes = [tuple(int(x) for x in s.split("x")) for s in sys.argv[1:]] or \
    [(1000, 20000, 10), (1000, 5000, 10), (500, 50000, 20), (1000, 50000, 10), (200, 50000, 10), (1000, 20000, 4), (1000, 20000, 24)]
for n, p, k in shapes:
    Ys = [torch.randn((p, n), dtype=torch.float64, device="cuda") for _ in range(3)]
    U = torch.linalg.qr(torch.randn((n, k), dtype=torch.float64, device="cuda"))[0].T.contiguous()
    c = torch.randn((k, p), dtype=torch.float64, device="cuda")
    s = torch.empty(p, dtype=torch.float64, device="cuda"); r = torch.empty(p, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    res = {}
    for mode in ("1", "0"):
        os.environ["FABLE_B200_ROWS_MMA"] = mode
        def run(i):
            fb.api.dev_rowstats(ctx, Ys[i % 3].data_ptr(), n, p, U.data_ptr(), c.data_ptr(), p, k, s.data_ptr(), r.data_ptr())
        for i in range(3000): run(i)          # ~0.2 s: a 50 us kernel timed from an idle GPU is timed at idle clocks
        ctx.synchronize()
        reps = 300
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(reps): run(i)
        e1.record(stream); ctx.synchronize()
        us = e0.elapsed_time(e1) / reps * 1e3
        ref = ((Ys[(reps - 1) % 3].T - U.T @ c) ** 2).sum(0)
        err = float(((r - ref).abs() / ref).max())
        res["dmma" if mode == "1" else "dfma"] = dict(us=round(us, 1), GBps=round(8 * n * p / us / 1e3, 1), frac=round(8 * n * p / us / 1e3 / PEAK, 3), relerr=err)
    print(json.dumps(dict(n=n, p=p, k=k, **res)), flush=True)
