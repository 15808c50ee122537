"""FABLEPosteriorSampler(MC = 1000) at C3 through pageable host buffers (every list element returned); best of 3 calls.
usage: FABLE_B200_PREFAULT=0|1 python tools/sampler_c3.py [MC]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fable_b200 as fb
from fable_b200 import synth
MC = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
n, p, k = 1000, 20000, 10
Y = synth.factor_data(n, p, k, rep=0)
ctx = fb.Context(0, 12345)
best = None
for rep in range(4):
    r = fb.FABLEPosteriorSampler(Y, MC=MC, ctx=ctx)
    t = r["runTime"]
    chk = float(r["CCFABLESamples"]["LambdaSamples"][::97, ::991].sum()) if "CCFABLESamples" in r else 0.0
    if rep and (best is None or t < best[0]):
        best = (t, chk)
    del r
print(f"PREFAULT={os.environ.get('FABLE_B200_PREFAULT', '1')}: FABLEPosteriorSampler(MC={MC}) {1e3 * best[0]:.1f} ms  checksum {best[1]:.10e}")
