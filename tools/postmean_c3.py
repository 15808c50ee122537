"""FABLEPosteriorMean at C3 through pageable host buffers, stage split by the library's own marks; best of 4 calls.
usage: FABLE_B200_PREFAULT=0|1 python tools/postmean_c3.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fable_b200 as fb
from fable_b200 import synth
n, p, k = 1000, 20000, 10
Y = synth.factor_data(n, p, k, rep=0)
ctx = fb.Context(0, 12345)
ctx.stage_timing(True)
best = None
for rep in range(5):
    r = fb.FABLEPosteriorMean(Y, ctx=ctx)
    st = ctx.stage_times()
    if rep and (best is None or r["runTime"] < best[0]):
        best = (r["runTime"], st, float(r["FABLEPostMean"][::997, ::991].sum()))
    del r
print(f"PREFAULT={os.environ.get('FABLE_B200_PREFAULT', '1')}: {1e3 * best[0]:.1f} ms  {best[1]}  checksum {best[2]:.10e}")
