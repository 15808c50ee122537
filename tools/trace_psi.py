"""Phase clocks of the covariance kernel (debug build tools/bin/libfable_trace.so, -DFB_PSI_TRACE)."""
import os, sys, subprocess
import numpy as np
mode = sys.argv[1] if len(sys.argv) > 1 else "mat_acc"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
out = "/tmp/psi_trace.bin"
env = dict(os.environ, FABLE_B200_LIB=os.path.join(os.path.dirname(os.path.abspath(__file__)), "bin", "libfable_trace.so"), FABLE_B200_TRACE_OUT=out)
print(subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "prof_psi.py"), "20000", str(B), "1", mode], env=env, capture_output=True, text=True).stdout.strip())
t = np.fromfile(out, dtype=np.int64).reshape(-1, 2, 8)
live = t[:, 0, 7] != 0
t = t[live]
names = ["prologue", "math+acc", "wait drain (barrier A)", "staging + barrier B", "TMA issue + zero", "tile switch (acc out/in)"]
for w, lab in ((0, "warp 0 (issuing thread)"), (1, "warp 7")):
    x = t[:, w, :6].astype(float)
    tot = x.sum(axis=1)
    print(f"{lab}: {len(x)} CTAs, clocks per CTA mean {tot.mean():.0f} min {tot.min():.0f} max {tot.max():.0f}")
    for i, nm in enumerate(names):
        print(f"   {nm:28s} {x[:, i].mean():12.0f} clk per CTA   {100 * x[:, i].sum() / tot.sum():5.1f}%")
