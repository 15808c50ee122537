"""Return path of a dense p x p result into pageable host memory (fable_post_mean at C3: 3.2 GB): destination from
numpy (asks for transparent huge pages itself) vs a plain anonymous mmap (what R's allocVector gets from malloc),
with and without the library's own huge-page hint.   usage: python tools/bench_d2h.py   (FABLE_B200_THP=0: no hint)"""
import ctypes as C, mmap, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fable_b200 as fb
from fable_b200 import synth
from fable_b200.api import _load, _ptr, _i64

n, p, k = 1000, 20000, 10
ctx = fb.Context(0, 1)
Y = synth.factor_data(n, p, k, rep=0)
s = fb.svd(Y, ctx=ctx)
U, V, d = s["u"], s["v"], s["d"]; r = len(d); kMax = fb.kmax_rule(d, 0.5)

def call(cov_ptr):
    kk = C.c_int()
    ctx.check(_load().fable_post_mean(ctx._h, _ptr(Y), _i64(n), _i64(p), C.c_double(1.0), C.c_double(1.0), _ptr(U), _ptr(V), _ptr(d),
                                      _i64(r), C.c_int(kMax), cov_ptr, C.byref(kk), None, None))

call(_ptr(np.empty((p, p), order="F")))              # warm the context, the block cache, the staging buffers
import torch
pin = torch.empty(p * p, dtype=torch.float64).pin_memory()
for rep in range(2):
    t0 = time.perf_counter(); call(C.c_void_p(pin.data_ptr())); t_pin = time.perf_counter() - t0
    t0 = time.perf_counter(); ctx.check(_load().fable_post_mean(ctx._h, _ptr(Y), _i64(n), _i64(p), C.c_double(1.0), C.c_double(1.0), _ptr(U), _ptr(V), _ptr(d),
                                      _i64(r), C.c_int(kMax), None, C.byref(C.c_int()), None, None)); t_none = time.perf_counter() - t0
    print(f"into pinned memory {t_pin * 1e3:.0f} ms; without the p x p result {t_none * 1e3:.0f} ms", flush=True)
del pin
for rep in range(3):
    a = np.empty((p, p), order="F")
    t0 = time.perf_counter(); call(_ptr(a)); t_np = time.perf_counter() - t0
    del a
    libc = C.CDLL(None); libc.mmap.restype = C.c_void_p; libc.munmap.argtypes = [C.c_void_p, C.c_size_t]
    libc.mmap.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_long]
    addr = libc.mmap(None, p * p * 8, mmap.PROT_READ | mmap.PROT_WRITE, mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS, -1, 0)   # fresh, no madvise
    t0 = time.perf_counter(); call(C.c_void_p(addr)); t_mm = time.perf_counter() - t0
    libc.munmap(C.c_void_p(addr), p * p * 8)
    print(f"THP hint {os.environ.get('FABLE_B200_THP', '1')}: fable_post_mean into numpy.empty {t_np * 1e3:.0f} ms, into a fresh mmap {t_mm * 1e3:.0f} ms", flush=True)
