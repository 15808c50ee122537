"""ctypes loader for oracle/_ref/libfable_ref.so: the REFERENCE'S OWN src/updated-FABLE-functions.cpp compiled
(where it lies) against the minimal Armadillo/Rcpp stand-ins of oracle/shim/.  TEST INFRASTRUCTURE ONLY.

Used to pin the restated oracle (oracle/fable_oracle.py) against the reference's own function bodies.  What
this does and does not establish is stated in DESIGN.md "Oracle": formulas, index conventions, bisection
control flow, list layouts and the random-draw consumption order are the reference's code executing; the
linear-algebra primitives under them are the shim's restatement of Armadillo's documented semantics."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libfable_ref.so")
_lib = None
_dp = C.POINTER(C.c_double)
_i64 = C.c_int64


def available():
    if not os.path.exists(_SO) and os.path.exists("/root/reference/src/updated-FABLE-functions.cpp"):
        subprocess.call(["make", "-C", _HERE, "-s", "ref"])
    return os.path.exists(_SO)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libfable_ref.so is not built (needs /root/reference)")
        _lib = C.CDLL(_SO)
        _lib.ref_last_error.restype = C.c_char_p
    return _lib


def _f(a):
    return np.asfortranarray(a, dtype=np.float64)


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _chk(rc):
    if rc != 0:
        raise RuntimeError("reference (shim build): " + lib().ref_last_error().decode())


def _svd_args(Y, U, V, d):
    Y = _f(Y); d = np.ascontiguousarray(d, dtype=np.float64)
    r = d.size
    U = _f(np.asarray(U)[:, :r]); V = _f(np.asarray(V)[:, :r])
    n, p = Y.shape
    return Y, U, V, d, n, p, r


def criterion(kInd, Y, U, V, d):
    Y, U, V, d, n, p, r = _svd_args(Y, U, V, d)
    o = C.c_double()
    _chk(lib().ref_criterion(C.c_int(kInd), _p(Y), _i64(n), _i64(p), _p(U), _p(V), _p(d), _i64(r), C.byref(o)))
    return o.value


def bisection(lower, upper, Y, U, V, d):
    Y, U, V, d, n, p, r = _svd_args(Y, U, V, d)
    o = np.empty(2)
    _chk(lib().ref_bisection(C.c_int(lower), C.c_int(upper), _p(Y), _i64(n), _i64(p), _p(U), _p(V), _p(d), _i64(r), _p(o)))
    return int(o[0]), int(o[1])


def rank_estimator(Y, U, V, d, kMax):
    Y, U, V, d, n, p, r = _svd_args(Y, U, V, d)
    k = C.c_int()
    _chk(lib().ref_rank(_p(Y), _i64(n), _i64(p), _p(U), _p(V), _p(d), _i64(r), C.c_int(kMax), C.byref(k)))
    return k.value


def hyper_parameters(Y, U, V, d, kEst):
    Y, U, V, d, n, p, r = _svd_args(Y, U, V, d)
    sig = np.empty(p); G = np.empty((p, p), order="F"); G0 = np.empty((p, kEst), order="F"); tau = C.c_double()
    _chk(lib().ref_hyper(_p(Y), _i64(n), _i64(p), _p(U), _p(V), _p(d), _i64(r), C.c_int(kEst), _p(sig), _p(G), _p(G0), C.byref(tau)))
    return {"SigmaSqEstimate": sig, "G": G, "G0": G0, "tauSqEstimate": tau.value, "EstimatedRank": kEst}


def cov_correct_matrix(sig, G):
    sig = np.ascontiguousarray(sig, dtype=np.float64); G = _f(G)
    p = sig.size
    o = np.empty(p * (p + 1) // 2)
    _chk(lib().ref_cov_correct(_p(sig), _p(G), _i64(p), _p(o)))
    return o


def post_mean(Y, gamma0, delta0sq, U, V, d, kMax):
    Y, U, V, d, n, p, r = _svd_args(Y, U, V, d)
    Cov = np.empty((p, p), order="F"); k = C.c_int()
    _chk(lib().ref_post_mean(_p(Y), _i64(n), _i64(p), C.c_double(gamma0), C.c_double(delta0sq), _p(U), _p(V), _p(d), _i64(r),
                             C.c_int(kMax), _p(Cov), C.byref(k)))
    return {"CovPostMean": Cov, "kEst": k.value}


def sampler(Y, gamma0, delta0sq, MC, U, V, d, kEst, varInflation, gam, z):
    """gam (MC, p) and z (MC, k, p) flattened C-order = the order in which the reference's loop consumes its
    gamma and normal variates (draw m: p gammas, then the p x k ZSample in column-major memory order)."""
    Y, U, V, d, n, p, r = _svd_args(Y, U, V, d)
    g = np.ascontiguousarray(gam, dtype=np.float64).reshape(-1); zz = np.ascontiguousarray(z, dtype=np.float64).reshape(-1)
    L = np.empty((MC, kEst * p), order="F"); S = np.empty((MC, p), order="F"); pm = np.empty(p); pi = np.empty(p)
    G = np.empty((p, p), order="F"); tau = C.c_double(); used = (C.c_int64 * 2)()
    _chk(lib().ref_sampler(_p(Y), _i64(n), _i64(p), C.c_double(gamma0), C.c_double(delta0sq), C.c_int(MC), _p(U), _p(V), _p(d),
                           _i64(r), C.c_int(kEst), C.c_double(varInflation), _p(g), _p(zz), _p(L), _p(S), _p(pm), _p(pi),
                           C.byref(tau), _p(G), used))
    return {"LambdaSamples": L, "SigmaSqSamples": S, "SigmaSqEstimatePostMean": pm, "SigmaSqEstimatePlugIn": pi,
            "tauSqEstimate": tau.value, "G": G, "EstimatedRank": kEst, "consumed": (used[0], used[1])}


def post_processing(out, alpha, selected_1based=None, optimized=False):
    L = _f(out["LambdaSamples"]); S = _f(out["SigmaSqSamples"]); k = int(out["EstimatedRank"])
    MC, p = S.shape
    if selected_1based is None:
        which, sel = 0, np.arange(1, p + 1, dtype=np.int64)
        ns = p
    else:
        which, sel = (2 if optimized else 1), np.ascontiguousarray(selected_1based, dtype=np.int64)
        ns = sel.size
    M = np.empty((ns, ns), order="F"); Lo = np.empty((ns, ns), order="F"); Up = np.empty((ns, ns), order="F")
    _chk(lib().ref_post_processing(C.c_int(which), _p(L), _p(S), _i64(MC), _i64(p), C.c_int(k), C.c_double(alpha),
                                   sel.ctypes.data_as(C.POINTER(C.c_int64)), _i64(sel.size), _p(M), _p(Lo), _p(Up)))
    return {"PostMeanMatrix": M, "LowerQuantileMatrix": Lo, "UpperQuantileMatrix": Up}


def symmetrize(A):
    A = _f(A); o = np.empty_like(A, order="F")
    _chk(lib().ref_symmetrize(_p(A), _i64(A.shape[0]), _p(o)))
    return o


def log_likelihood_eval(Y, M, Lam, sig):
    Y = _f(Y); M = _f(M); Lam = _f(Lam); sig = np.ascontiguousarray(sig, dtype=np.float64)
    o = C.c_double()
    _chk(lib().ref_loglik(_p(Y), _i64(Y.shape[0]), _i64(Y.shape[1]), _p(M), _p(Lam), C.c_int(M.shape[1]), _p(sig), C.byref(o)))
    return o.value
