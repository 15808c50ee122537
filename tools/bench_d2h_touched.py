"""Return path of a dense p x p result (fable_post_mean at C3: 3.2 GB) into pageable memory that is ALREADY mapped (second
call into the same numpy array) against fresh memory and pinned memory: what the staged copy pipeline costs without page faults.
usage: python tools/bench_d2h_touched.py   (FABLE_B200_COPY_THREADS=n)"""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fable_b200 as fb
from fable_b200 import synth
from fable_b200.api import _load, _ptr, _i64
import torch

n, p, k = 1000, 20000, 10
ctx = fb.Context(0, 1)
Y = synth.factor_data(n, p, k, rep=0)
s = fb.svd(Y, ctx=ctx)
U, V, d = s["u"], s["v"], s["d"]; r = len(d); kMax = fb.kmax_rule(d, 0.5)

def call(cov_ptr):
    kk = C.c_int()
    t0 = time.perf_counter()
    ctx.check(_load().fable_post_mean(ctx._h, _ptr(Y), _i64(n), _i64(p), C.c_double(1.0), C.c_double(1.0), _ptr(U), _ptr(V), _ptr(d),
                                      _i64(r), C.c_int(kMax), cov_ptr, C.byref(kk), None, None))
    return 1e3 * (time.perf_counter() - t0)

call(_ptr(np.empty((p, p), order="F")))
base = min(call(None) for _ in range(3))
pin = torch.empty(p * p, dtype=torch.float64).pin_memory()
t_pin = min(call(C.c_void_p(pin.data_ptr())) for _ in range(3)); del pin
fresh, touched = [], []
for rep in range(3):
    a = np.empty((p, p), order="F")
    fresh.append(call(_ptr(a))); touched.append(call(_ptr(a))); touched.append(call(_ptr(a)))
    del a
print(f"threads {os.environ.get('FABLE_B200_COPY_THREADS', 'default')}: without the p x p result {base:.0f} ms; + result into pinned {t_pin - base:.0f} ms, "
      f"into fresh pageable {min(fresh) - base:.0f} ms, into mapped pageable {min(touched) - base:.0f} ms  (3.2 GB: 51 GB/s = 63 ms)")
