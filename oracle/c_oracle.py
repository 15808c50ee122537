"""ctypes loader for the C twin of the oracle (oracle/fable_oracle.c).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

_dp = C.POINTER(C.c_double)


def build(force=False):
    src = os.path.join(_HERE, "fable_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp)


def _f(a):
    """column-major float64 copy (the C side indexes i + ld*j)."""
    return np.asfortranarray(a, dtype=np.float64)


def num_threads():
    return int(lib().orc_num_threads())


def set_num_threads(n=None):
    """OpenMP threads of the timed CPU arm: all host cores this process may run on unless told otherwise
    (torchrun exports OMP_NUM_THREADS=1 to its workers, which would silently make the baseline single-threaded)."""
    if n is None:
        try:
            n = len(os.sched_getaffinity(0))
        except AttributeError:
            n = os.cpu_count() or 1
    lib().orc_set_num_threads(C.c_int(int(n)))
    return num_threads()


def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return [int(x) for x in o]


def rng_fill(seed, m0, MC, p, k, shape, want_gam=True, want_z=True):
    gam = np.empty((MC, p)) if want_gam else None
    z = np.empty((MC, k, p)) if want_z else None
    lib().orc_rng_fill(C.c_uint64(seed), C.c_int64(m0), C.c_int64(MC), C.c_int64(p), C.c_int(k),
                       C.c_double(shape), _p(gam) if want_gam else None, _p(z) if want_z else None)
    return gam, z


def sampler(MC, p, k, G0, gnd, a_n, gam, z):
    G0 = _f(G0); gnd = np.ascontiguousarray(gnd, dtype=np.float64)
    gam = np.ascontiguousarray(gam, dtype=np.float64); z = np.ascontiguousarray(z, dtype=np.float64)
    L = np.empty((MC, k * p), order="F"); S = np.empty((MC, p), order="F")
    lib().orc_sampler(C.c_int64(MC), C.c_int64(p), C.c_int(k), _p(G0), _p(gnd), C.c_double(a_n),
                      _p(gam), _p(z), _p(L), _p(S))
    return L, S


def psi_moments(L, S, k):
    L = _f(L); S = _f(S)
    MC, p = S.shape
    s1 = np.empty((p, p), order="F"); s2 = np.empty((p, p), order="F")
    lib().orc_psi_moments(C.c_int64(MC), C.c_int64(p), C.c_int(k), _p(L), _p(S), _p(s1), _p(s2))
    return s1, s2


def psi_draw(L, S, k, m):
    L = _f(L); S = _f(S)
    MC, p = S.shape
    out = np.empty((p, p), order="F")
    lib().orc_psi_draw(C.c_int64(MC), C.c_int64(p), C.c_int(k), C.c_int64(m), _p(L), _p(S), _p(out))
    return out


def cov_correct_sums(sig, G):
    sig = np.ascontiguousarray(sig, dtype=np.float64); G = _f(G)
    p = sig.size
    a = C.c_double(); b = C.c_double()
    lib().orc_cov_correct_sums(C.c_int64(p), _p(sig), _p(G), C.byref(a), C.byref(b))
    return a.value, b.value


def resid_sigsq(Y, U, V, d, k):
    Y = _f(Y); U = _f(U); V = _f(V); d = np.ascontiguousarray(d, dtype=np.float64)
    n, p = Y.shape
    sig = np.empty(p)
    lib().orc_resid_sigsq(C.c_int64(n), C.c_int64(p), C.c_int(k), _p(Y), _p(U), C.c_int64(U.shape[0]),
                          _p(V), C.c_int64(V.shape[0]), _p(d), _p(sig))
    return sig
