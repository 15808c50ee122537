"""ncu driver for the row pass: a few launches of both kernels at one shape (default C3).
ncu --set full --clock-control none --import-source on -k regex:k_rowstats -o gpurun_out/rows python tools/prof_rows.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fable_b200 as fb

n, p, k = (int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1000x20000x10").split("x"))
ctx = fb.Context(0, 1)
Ys = [torch.randn((p, n), dtype=torch.float64, device="cuda") for _ in range(2)]
U = torch.linalg.qr(torch.randn((n, k), dtype=torch.float64, device="cuda"))[0].T.contiguous()
c = torch.randn((k, p), dtype=torch.float64, device="cuda")
s = torch.empty(p, dtype=torch.float64, device="cuda"); r = torch.empty(p, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
for mode in ("1", "0"):
    os.environ["FABLE_B200_ROWS_MMA"] = mode
    for i in range(2):
        fb.api.dev_rowstats(ctx, Ys[i % 2].data_ptr(), n, p, U.data_ptr(), c.data_ptr(), p, k, s.data_ptr(), r.data_ptr())
    ctx.synchronize()
