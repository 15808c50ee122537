import os, sys, time
sys.argv = [sys.argv[0], "2048", "40"]
exec(open("tools/e2e_breakdown.py").read().split("call(True)\nfor rep")[0])
for i in range(5):
    ctx.synchronize(); t0 = time.perf_counter(); call(True); ctx.synchronize(); print(f"call {i}: {time.perf_counter() - t0:.3f} s", flush=True)
for i in range(2):
    ctx.synchronize(); t0 = time.perf_counter(); call(False); ctx.synchronize(); print(f"samples-only call {i}: {time.perf_counter() - t0:.3f} s", flush=True)
for i in range(2):
    ctx.synchronize(); t0 = time.perf_counter(); call(True); ctx.synchronize(); print(f"call after samples-only {i}: {time.perf_counter() - t0:.3f} s", flush=True)
