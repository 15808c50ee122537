"""Host-side mirror of the reference's R interface over the C-ABI of libfable_b200.so.

Function names, argument order and returned element names follow the reference:
  R/FABLE_wrapper.R:12-68   FABLEPosteriorSampler(Y, gamma0, delta0sq, maxProp, MC)
  R/FABLE_wrapper.R:80-128  FABLEPosteriorMean(Y, gamma0, delta0sq, maxProp)
  R/RcppExports.R:20-75     CPPCriterionFunction, CPPBisectionRecursion, CPPRankEstimator,
                            FABLEHyperParameters, CPPcov_correct_matrix, CPPFABLEPostMean,
                            CPPFABLESampler, CPPsvd
Every function marshals numpy float64 arrays as column-major `double*` into the C-ABI and raises
`FableError` (R: `stop()`) on a non-zero status.  No arithmetic on matrix-sized data happens here.
"""
from __future__ import annotations

import ctypes as C
import os
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.environ.get("FABLE_B200_LIB") or os.path.join(_HERE, "libfable_b200.so")
_lib = None
_dp = C.POINTER(C.c_double)
_i64 = C.c_int64


class FableError(RuntimeError):
    """Raised for every non-zero fable_status (the Rcpp layer raises Rcpp::stop with the same text)."""

    def __init__(self, code, msg):
        super().__init__(f"libfable_b200 status {code}: {msg}")
        self.code = code


def library_path():
    return _SO


def _load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        raise FableError(2, f"{_SO} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU fallback)")
    lib = C.CDLL(_SO)
    lib.fable_last_error.restype = C.c_char_p
    lib.fable_last_error.argtypes = [C.c_void_p]
    lib.fable_ctx_launch_count.restype = C.c_int64
    lib.fable_ctx_launch_count.argtypes = [C.c_void_p]
    lib.fable_ctx_stream.restype = C.c_void_p
    lib.fable_ctx_stream.argtypes = [C.c_void_p]
    lib.fable_ctx_destroy.restype = None
    lib.fable_ctx_destroy.argtypes = [C.c_void_p]
    lib.fable_posterior_destroy.restype = None
    lib.fable_posterior_destroy.argtypes = [C.c_void_p]
    lib.fable_multi_destroy.restype = None
    lib.fable_multi_destroy.argtypes = [C.c_void_p]
    lib.fable_multi_last_error.restype = C.c_char_p
    lib.fable_multi_last_error.argtypes = [C.c_void_p]
    lib.fable_multi_ctx.restype = C.c_void_p
    lib.fable_multi_ctx.argtypes = [C.c_void_p, C.c_int]
    _lib = lib
    return lib


def _F(a, name="matrix"):
    """float64 column-major view/copy (R's memory order)."""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        return np.ascontiguousarray(a)
    if a.ndim != 2:
        raise FableError(1, f"{name} must be a matrix")
    return np.asfortranarray(a)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_dp)


_ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_double), C.c_int64)
_ALLREDUCE_DEV_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p)


class Context:
    """fable_ctx: one device, one stream set.  `seed` keys the Philox stream (fable-philox-v1)."""

    def __init__(self, device=0, seed=12345):
        lib = _load()
        h = C.c_void_p()
        rc = lib.fable_ctx_create(C.c_int(device), C.c_uint64(seed), C.byref(h))
        if rc != 0:
            raise FableError(rc, lib.fable_last_error(None).decode())
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            _load().fable_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc != 0:
            raise FableError(rc, _load().fable_last_error(self._h).decode())

    def set_seed(self, seed):
        self.check(_load().fable_ctx_set_seed(self._h, C.c_uint64(seed)))

    @property
    def launches(self):
        return int(_load().fable_ctx_launch_count(self._h))

    @property
    def stream(self):
        return int(_load().fable_ctx_stream(self._h) or 0)

    def synchronize(self):
        self.check(_load().fable_ctx_synchronize(self._h))

    shard = None

    def set_column_shard(self, p_total, col_offset=0, rank=0, world=1, allreduce=None, allreduce_dev=None):
        """Column-sharded estimation (fable_ctx_set_column_shard): this context holds columns
        [col_offset, col_offset + p_local) of a p_total-column problem; `allreduce(a)` must sum the float64 numpy
        array `a` over all ranks in place, with identical bits on every rank.  Optional `allreduce_dev(ptr, count,
        stream)`: the same for `count` doubles at a DEVICE address, ordered on the given CUDA stream (the n x n Gram
        then never leaves HBM).  p_total = 0 switches it off."""
        lib = _load()
        if not p_total:
            self.check(lib.fable_ctx_set_column_shard(self._h, _i64(0), _i64(0), C.c_int(0), C.c_int(1), None, None))
            self.shard = None
            self._shard_cb = None
            self._shard_dev_cb = None
            return

        def _cb(_user, vals, count):
            try:
                allreduce(np.ctypeslib.as_array(vals, shape=(count,)))
                return 0
            except Exception:            # never let an exception cross the C boundary
                import traceback
                traceback.print_exc()
                return 1
        cb = _ALLREDUCE_FN(_cb)
        self.check(lib.fable_ctx_set_column_shard(self._h, _i64(p_total), _i64(col_offset), C.c_int(rank), C.c_int(world), cb, None))
        self._shard_cb = cb             # keep the trampoline alive as long as the context uses it
        self.shard = (int(p_total), int(col_offset), int(rank), int(world))
        self._shard_dev_cb = None
        if allreduce_dev is not None:
            def _dcb(_user, ptr, count, stream):
                try:
                    allreduce_dev(int(ptr), int(count), int(stream or 0))
                    return 0
                except Exception:
                    import traceback
                    traceback.print_exc()
                    return 1
            dcb = _ALLREDUCE_DEV_FN(_dcb)
            self.check(lib.fable_ctx_set_column_shard_dev(self._h, dcb, None))
            self._shard_dev_cb = dcb

    def trim_cache(self):
        """Return the library's idle cached device blocks to the driver (they are reused across calls otherwise)."""
        self.check(_load().fable_ctx_trim_cache(self._h))

    def kernel_timing(self, enable=True):
        """Start (and clear) / stop per-launch CUDA-event timing of the covariance kernel."""
        self.check(_load().fable_ctx_kernel_timing(self._h, C.c_int(1 if enable else 0)))

    def kernel_time(self):
        """(summed device milliseconds, launches) of the covariance kernel since kernel_timing(True)."""
        ms = C.c_double(); cnt = C.c_int64()
        self.check(_load().fable_ctx_kernel_time(self._h, C.byref(ms), C.byref(cnt)))
        return ms.value, cnt.value

    def stage_timing(self, enable=True):
        """Make the wrapper pipelines record their stage split (the stream is drained at every stage boundary)."""
        self.check(_load().fable_ctx_stage_timing(self._h, C.c_int(1 if enable else 0)))

    def stage_times(self):
        """Stage milliseconds of the last FABLEPosteriorMean / FABLEPosteriorSampler pipeline call on this context."""
        ms = (C.c_double * 6)()
        self.check(_load().fable_ctx_stage_times(self._h, ms, C.c_int(6)))
        names = ("upload_Y", "svd_on_device", "svd_to_host_and_kmax", "rank_search", "row_posterior", "pxp_results_to_host")
        return {nm: round(float(v), 3) for nm, v in zip(names, ms)}

    def measure_fp64_peak(self):
        """DMMA.8x8x4 issue peak of this device in TFLOP/s (diagnostics; the Gram step's roofline denominator)."""
        v = C.c_double()
        self.check(_load().fable_ctx_measure_fp64_peak(self._h, C.byref(v)))
        return v.value

    def set_psi_ring(self, slots):
        """slots > 0: the sampler's Psi pass also materialises every draw into a device ring."""
        self.check(_load().fable_ctx_set_psi_ring(self._h, C.c_int(int(slots))))


class MultiContext:
    """fable_multi: several devices of one box driven from this process (one context per device, peer access all
    round).  Pass it as `ctx` to CPPFABLESampler: the draws are sharded over the devices and the accumulators are summed
    over NVLink peer memory (fable_multi_sampler)."""

    def __init__(self, devices, seed=12345):
        lib = _load()
        devs = list(range(devices)) if isinstance(devices, int) else [int(x) for x in devices]
        arr = (C.c_int * len(devs))(*devs)
        h = C.c_void_p()
        rc = lib.fable_multi_create(C.c_int(len(devs)), arr, C.c_uint64(seed), C.byref(h))
        if rc != 0:
            raise FableError(rc, lib.fable_multi_last_error(None).decode())
        self._h = h
        self.devices = devs

    def close(self):
        if getattr(self, "_h", None):
            _load().fable_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc != 0:
            raise FableError(rc, _load().fable_multi_last_error(self._h).decode())

    @property
    def launches(self):
        lib = _load()
        return sum(int(lib.fable_ctx_launch_count(C.c_void_p(lib.fable_multi_ctx(self._h, g)))) for g in range(len(self.devices)))


_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(int(os.environ.get("FABLE_B200_DEVICE", os.environ.get("LOCAL_RANK", "0"))),
                               int(os.environ.get("FABLE_B200_SEED", "12345")))
    return _default_ctx


def _ctx(ctx):
    return ctx if ctx is not None else default_context()


def _inputs(Y, U_Y, V_Y, svalsY):
    Y = _F(Y, "Y"); U = _F(U_Y, "U_Y"); V = _F(V_Y, "V_Y"); d = _F(svalsY, "svalsY")
    n, p = Y.shape
    r = d.shape[0]
    if U.shape != (n, r) and not (U.shape[0] == n and U.shape[1] >= r):
        raise FableError(1, "U_Y must be n x length(svalsY)")
    if V.shape != (p, r) and not (V.shape[0] == p and V.shape[1] >= r):
        raise FableError(1, "V_Y must be p x length(svalsY)")
    return Y, U, V, d, n, p, r


# ------------------------------------------------------------------------------------------------
# a1 / a2
# ------------------------------------------------------------------------------------------------
def svd(Y, ctx=None):
    """Drop-in for base-R `svd(Y)` as the wrappers use it: list(d, u, v)."""
    ctx = _ctx(ctx)
    Y = _F(Y, "Y")
    n, p = Y.shape
    r = n if ctx.shard else min(n, p)          # column shard: Y holds the local columns, r = min(n, p_total) = n
    U = np.empty((n, r), order="F"); d = np.empty(r); V = np.empty((p, r), order="F")
    ctx.check(_load().fable_svd(ctx._h, _ptr(Y), _i64(n), _i64(p), _ptr(U), _ptr(d), _ptr(V)))
    return {"d": d, "u": U, "v": V}


def CPPsvd(X, flag=4, ctx=None):
    """CPPsvd (updated-FABLE-functions.cpp:40-74): list(U, d, V).  flag 1 / 2: the full decomposition (arma::svd: U n x n,
    V p x p); any other flag: the economy form (svd_econ).  The LAPACK driver `flag` also selected there ("std" /
    "dc") has no counterpart: one algorithm."""
    if int(flag) in (1, 2):
        ctx = _ctx(ctx)
        X = _F(X, "X")
        n, p = X.shape
        U = np.empty((n, n), order="F"); d = np.empty(min(n, p)); V = np.empty((p, p), order="F")
        ctx.check(_load().fable_svd_full(ctx._h, _ptr(X), _i64(n), _i64(p), _ptr(U), _ptr(d), _ptr(V)))
        return {"U": U, "d": d, "V": V}
    s = svd(X, ctx)
    return {"U": s["u"], "d": s["d"], "V": s["v"]}


def kmax_rule(svalsY, maxProp):
    d = _F(svalsY)
    out = C.c_int()
    rc = _load().fable_kmax(_ptr(d), _i64(d.shape[0]), C.c_double(maxProp), C.byref(out))
    if rc != 0:
        raise FableError(rc, "kMax rule: no index reaches maxProp")
    return out.value


# ------------------------------------------------------------------------------------------------
# a3 / a4 / a5
# ------------------------------------------------------------------------------------------------
def CPPCriterionFunction(kInd, Y, U_Y, V_Y, svalsY, ctx=None):
    ctx = _ctx(ctx)
    Y, U, V, d, n, p, r = _inputs(Y, U_Y, V_Y, svalsY)
    out = C.c_double()
    ctx.check(_load().fable_criterion(ctx._h, C.c_int(int(kInd)), _ptr(Y), _i64(n), _i64(p), _ptr(U), _ptr(V),
                                      _ptr(d), _i64(r), C.byref(out)))
    return out.value


def CPPLogLikelihoodEval(Y, M, Lambda, SigmaSq, ctx=None):
    """cpp:95-125 (exported by the reference, not used by its wrappers)."""
    ctx = _ctx(ctx)
    Y = _F(Y, "Y"); M = _F(M, "M"); L = _F(Lambda, "Lambda"); s = _F(SigmaSq)
    n, p = Y.shape
    k = M.shape[1]
    if M.shape[0] != n or L.shape != (p, k) or s.shape[0] != p:
        raise FableError(1, "CPPLogLikelihoodEval: M must be n x k, Lambda p x k, SigmaSq length p")
    out = C.c_double()
    ctx.check(_load().fable_log_likelihood_eval(ctx._h, _ptr(Y), _i64(n), _i64(p), _ptr(M), _ptr(L), C.c_int(k), _ptr(s),
                                                C.byref(out)))
    return out.value


def CPPBisectionRecursion(lower, upper, Y, U_Y, V_Y, svalsY, ctx=None):
    ctx = _ctx(ctx)
    Y, U, V, d, n, p, r = _inputs(Y, U_Y, V_Y, svalsY)
    lo = C.c_int(); hi = C.c_int()
    ctx.check(_load().fable_bisection(ctx._h, C.c_int(int(lower)), C.c_int(int(upper)), _ptr(Y), _i64(n), _i64(p),
                                      _ptr(U), _ptr(V), _ptr(d), _i64(r), C.byref(lo), C.byref(hi)))
    return np.array([float(lo.value), float(hi.value)])   # arma::vec of doubles (cpp:263-267)


def CPPRankEstimator(Y, U_Y, V_Y, svalsY, kMax, ctx=None):
    ctx = _ctx(ctx)
    Y, U, V, d, n, p, r = _inputs(Y, U_Y, V_Y, svalsY)
    k = C.c_int()
    ctx.check(_load().fable_rank_estimator(ctx._h, _ptr(Y), _i64(n), _i64(p), _ptr(U), _ptr(V), _ptr(d), _i64(r),
                                           C.c_int(int(kMax)), C.byref(k)))
    return k.value


def RankEstimator(Ymat, svdmod, kMax, ctx=None):
    """The pure-R prototype RankEstimator(Ymat, svdmod, kMax) (R/FABLE_code.R:32-186): svdmod = list(d, u, v)."""
    ctx = _ctx(ctx)
    Y, U, V, d, n, p, r = _inputs(Ymat, svdmod["u"], svdmod["v"], svdmod["d"])
    k = C.c_int()
    ctx.check(_load().fable_rank_estimator_proto(ctx._h, _ptr(Y), _i64(n), _i64(p), _ptr(U), _ptr(V), _ptr(d), _i64(r),
                                                 C.c_int(int(kMax)), C.byref(k)))
    return k.value


def RankEstimatorCriterion(kInd, Ymat, svdmod, ctx=None):
    """CriterionFunction(kInd) inside the prototype's RankEstimator (R/FABLE_code.R:57-116)."""
    ctx = _ctx(ctx)
    Y, U, V, d, n, p, r = _inputs(Ymat, svdmod["u"], svdmod["v"], svdmod["d"])
    out = C.c_double()
    ctx.check(_load().fable_criterion_proto(ctx._h, C.c_int(int(kInd)), _ptr(Y), _i64(n), _i64(p), _ptr(U), _ptr(V), _ptr(d),
                                            _i64(r), C.byref(out)))
    return out.value


def CCFABLE_DirectSampler(Y, gamma0=1.0, delta0sq=1.0, MC=1000, ctx=None, gam=None, z=None, m0=0, info=None):
    """R/FABLE_code.R:212-298: the (MC, p, p) array of entry-wise coverage-corrected draws of Psi.  gam / z: injected draws
    (z may be a callable (MC, k, p) -> array: the rank is only known after the prototype's RankEstimator has run).
    `info`, if a dict, receives kMax and k."""
    ctx = _ctx(ctx)
    Y = _F(Y, "Y")
    n, p = Y.shape
    MC = int(MC)
    kM = C.c_int(); k = C.c_int(); ph = C.c_void_p()
    ctx.check(_load().fable_direct_sampler_begin(ctx._h, _ptr(Y), _i64(n), _i64(p), C.c_double(gamma0), C.c_double(delta0sq),
                                                 C.byref(kM), C.byref(k), C.byref(ph)))
    try:
        if info is not None:
            info.update(kMax=kM.value, k=k.value)
        g = zz = None
        if gam is not None or z is not None:
            if gam is None or z is None:
                raise FableError(1, "injected draws need both gam and z")
            zz = z(MC, k.value, p) if callable(z) else z
            g = np.ascontiguousarray(gam, dtype=np.float64).reshape(MC, p)
            zz = np.ascontiguousarray(zz, dtype=np.float64).reshape(MC, k.value, p)
        out = np.empty((MC, p, p), order="F")                         # R array: first index fastest
        ctx.check(_load().fable_direct_sampler_draws(ph, _i64(MC), _i64(m0), _ptr(g), _ptr(zz), out.ctypes.data_as(_dp)))
    finally:
        _load().fable_posterior_destroy(ph)
    return out


# ------------------------------------------------------------------------------------------------
# a6 / a7
# ------------------------------------------------------------------------------------------------
def FABLEHyperParameters(Y, U_Y, V_Y, svalsY, kEst, ctx=None, want_G=True):
    ctx = _ctx(ctx)
    Y, U, V, d, n, p, r = _inputs(Y, U_Y, V_Y, svalsY)
    k = int(kEst)
    if not (1 <= k <= r):
        raise FableError(1, "kEst must be in 1..length(svalsY)")
    sig = np.empty(p); G0 = np.empty((p, k), order="F"); tau = C.c_double()
    G = np.empty((p, p), order="F") if want_G else None
    ctx.check(_load().fable_hyper_parameters(ctx._h, _ptr(Y), _i64(n), _i64(p), _ptr(U), _ptr(V), _ptr(d), _i64(r),
                                             C.c_int(k), _ptr(sig), _ptr(G0), C.byref(tau), _ptr(G)))
    out = {"SigmaSqEstimate": sig, "G": G, "EstimatedRank": k, "G0": G0, "tauSqEstimate": tau.value}
    return out


def CPPcov_correct_matrix(sigsq_hat, llprime_hat, ctx=None):
    """Upper-triangle vector, column-major trimatu order (cpp:401-428)."""
    ctx = _ctx(ctx)
    sig = _F(sigsq_hat); G = _F(llprime_hat)
    p = sig.shape[0]
    if G.shape != (p, p):
        raise FableError(1, "llprime_hat must be p x p")
    out = np.empty(p * (p + 1) // 2)
    ctx.check(_load().fable_cov_correct_matrix(ctx._h, _ptr(sig), _ptr(G), _i64(p), _ptr(out)))
    return out


def cov_correct_matrix(sigsq_hat, llprime_hat, ctx=None):
    """Full p x p B: the pure-R twin the wrapper calls (R/FABLE_code.R:189-206)."""
    ctx = _ctx(ctx)
    sig = _F(sigsq_hat); G = _F(llprime_hat)
    p = sig.shape[0]
    if G.shape != (p, p):
        raise FableError(1, "llprime_hat must be p x p")
    out = np.empty((p, p), order="F")
    ctx.check(_load().fable_cov_correct_full(ctx._h, _ptr(sig), _ptr(G), _i64(p), _ptr(out)))
    return out


def var_inflation(sigsq_hat, G0, variant="wrapper", ctx=None):
    """The scalar the wrappers derive from the correction matrix, as a fused reduction over
    G = G0 G0' that never stores B.  variant: 'wrapper' (R/FABLE_wrapper.R:41-44) or 'script'
    (extras/replicationCodes/runtimeScript.R:75-78)."""
    ctx = _ctx(ctx)
    sig = _F(sigsq_hat); G0 = _F(G0)
    p, k = G0.shape
    out = C.c_double()
    ctx.check(_load().fable_var_inflation(ctx._h, _ptr(sig), _ptr(G0), _i64(p), C.c_int(k),
                                          C.c_int({"wrapper": 0, "script": 1}[variant]), C.byref(out)))
    return out.value


# ------------------------------------------------------------------------------------------------
# a8
# ------------------------------------------------------------------------------------------------
def CPPFABLEPostMean(Y, gamma0, delta0sq, U_Y, V_Y, svalsY, kMax, ctx=None, want_cov=True, internals=False):
    ctx = _ctx(ctx)
    Y, U, V, d, n, p, r = _inputs(Y, U_Y, V_Y, svalsY)
    kMax = int(kMax)
    if not (1 <= kMax <= r):
        raise FableError(1, "kMax must be in 1..length(svalsY)")
    Cov = np.empty((p, p), order="F") if want_cov else None
    k = C.c_int()
    G0s = np.empty((p, kMax), order="F") if internals else None   # only the first kEst columns are written
    Sh = np.empty(p) if internals else None
    ctx.check(_load().fable_post_mean(ctx._h, _ptr(Y), _i64(n), _i64(p), C.c_double(gamma0), C.c_double(delta0sq),
                                      _ptr(U), _ptr(V), _ptr(d), _i64(r), C.c_int(kMax), _ptr(Cov), C.byref(k),
                                      _ptr(G0s), _ptr(Sh)))
    out = {"CovPostMean": Cov, "kEst": k.value}
    if internals:
        out["G0"] = np.asfortranarray(G0s.reshape(-1, order="F")[: p * k.value].reshape((p, k.value), order="F"))
        out["SigmaHat"] = Sh
    return out


# ------------------------------------------------------------------------------------------------
# a9
# ------------------------------------------------------------------------------------------------
def CPPFABLESampler(Y, gamma0, delta0sq, MC, U_Y, V_Y, svalsY, kEst, varInflation, ctx=None, gam=None, z=None,
                    m0=0, want_G=True, want_samples=True, psi_moments=False, out=None):
    """Element names as cpp:620-627.  `gam`/`z`: injected draws (gam[m, j]; z[m, l, j]); default is
    the native Philox stream.  `psi_moments=True` adds PsiSum / PsiSumSq (entry-wise sums over the
    draws of Psi_m = Lambda_m Lambda_m' + diag(sigma2_m))."""
    ctx = _ctx(ctx)
    Y, U, V, d, n, p, r = _inputs(Y, U_Y, V_Y, svalsY)
    k = int(kEst); MC = int(MC)
    if not (1 <= k <= r):
        raise FableError(1, "kEst must be in 1..length(svalsY)")
    if MC < 1:
        raise FableError(1, "MC must be >= 1")
    g = zz = None
    if gam is not None or z is not None:
        if gam is None or z is None:
            raise FableError(1, "injected draws need both gam and z")
        g = np.ascontiguousarray(gam, dtype=np.float64).reshape(MC, p)
        zz = np.ascontiguousarray(z, dtype=np.float64).reshape(MC, k, p)
    out = out or {}

    def _buf(name, shape, want):
        if not want:
            return None
        a = out.get(name)
        if a is None:
            return np.empty(shape, order="F")
        if a.shape != tuple(np.atleast_1d(shape)) or a.dtype != np.float64 or not (a.flags.f_contiguous or a.ndim == 1):
            raise FableError(1, f"out[{name!r}] must be a column-major float64 array of shape {shape}")
        return a
    L = _buf("LambdaSamples", (MC, k * p), want_samples)
    S = _buf("SigmaSqSamples", (MC, p), want_samples)
    pm = _buf("SigmaSqEstimatePostMean", (p,), True); pi = _buf("SigmaSqEstimatePlugIn", (p,), True); tau = C.c_double()
    G = _buf("G", (p, p), want_G)
    s1 = _buf("PsiSum", (p, p), psi_moments)
    s2 = _buf("PsiSumSq", (p, p), psi_moments)
    entry = _load().fable_multi_sampler if isinstance(ctx, MultiContext) else _load().fable_sampler
    ctx.check(entry(ctx._h, _ptr(Y), _i64(n), _i64(p), C.c_double(gamma0), C.c_double(delta0sq),
                                    _i64(MC), _ptr(U), _ptr(V), _ptr(d), _i64(r), C.c_int(k),
                                    C.c_double(varInflation), _ptr(g), _ptr(zz), _i64(m0), _ptr(L), _ptr(S),
                                    _ptr(pm), _ptr(pi), C.byref(tau), _ptr(G), _ptr(s1), _ptr(s2)))
    out = {"LambdaSamples": L, "SigmaSqSamples": S, "SigmaSqEstimatePostMean": pm, "SigmaSqEstimatePlugIn": pi,
           "EstimatedRank": k, "VarianceInflation": float(varInflation), "tauSqEstimate": tau.value, "G": G}
    if psi_moments:
        out["PsiSum"] = s1; out["PsiSumSq"] = s2
    return out


# ------------------------------------------------------------------------------------------------
# f1: post-processing of stored draws (R/RcppExports.R:77-87)
# ------------------------------------------------------------------------------------------------
def _post(FABLEOutput, alpha, sel, ctx):
    ctx = _ctx(ctx)
    L = _F(FABLEOutput["LambdaSamples"]); S = _F(FABLEOutput["SigmaSqSamples"]); k = int(FABLEOutput["EstimatedRank"])
    MC, p = S.shape
    if L.shape != (MC, k * p):
        raise FableError(1, "LambdaSamples must be MC x (k p)")
    if sel is None:
        n = p; sp = None
    else:
        sel = np.ascontiguousarray(sel, dtype=np.int64); n = sel.size
        sp = sel.ctypes.data_as(C.POINTER(C.c_int64))
    M = np.empty((n, n), order="F"); Lo = np.empty((n, n), order="F"); Up = np.empty((n, n), order="F")
    ctx.check(_load().fable_post_processing(ctx._h, _ptr(L), _ptr(S), _i64(MC), _i64(p), C.c_int(k), C.c_double(alpha), sp,
                                            _i64(n), _ptr(M), _ptr(Lo), _ptr(Up)))
    return {"PostMeanMatrix": M, "LowerQuantileMatrix": Lo, "UpperQuantileMatrix": Up}


def CPPCCFABLEPostProcessing(FABLEOutput, alpha, ctx=None):
    """cpp:631-713: mean and (alpha/2, 1-alpha/2) quantiles of every entry of Psi over the draws."""
    return _post(FABLEOutput, alpha, None, ctx)


def CCFABLEPostProcessingSubmatrix(FABLEOutput, alpha, SelectedIndices, ctx=None):
    """cpp:716-818: the same on the sub-matrix of the (1-based) SelectedIndices."""
    return _post(FABLEOutput, alpha, SelectedIndices, ctx)


CCFABLEPostProcessingSubmatrix_Optimized = CCFABLEPostProcessingSubmatrix   # cpp:821-893: same results


def row_posterior(Y, gamma0, delta0sq, U_Y, V_Y, svalsY, k, shrunk=True, ctx=None):
    """fable_row_posterior: the per-variable conjugate posterior -- G0 (p x k), gnd, sig (plug-in
    sigma^2), sigmahat (posterior mean of sigma^2), tauSq.  Under a column shard: the rows of the local variables."""
    ctx = _ctx(ctx)
    Y, U, V, d, n, p, r = _inputs(Y, U_Y, V_Y, svalsY)
    k = int(k)
    G0 = np.empty((p, k), order="F"); gnd = np.empty(p); sig = np.empty(p); sh = np.empty(p); tau = C.c_double()
    ctx.check(_load().fable_row_posterior(ctx._h, _ptr(Y), _i64(n), _i64(p), C.c_double(gamma0), C.c_double(delta0sq),
                                          _ptr(U), _ptr(V), _ptr(d), _i64(r), C.c_int(k), C.c_int(1 if shrunk else 0),
                                          _ptr(G0), _ptr(gnd), _ptr(sig), _ptr(sh), C.byref(tau)))
    return {"G0": G0, "gnd": gnd, "sig": sig, "sigmahat": sh, "tausq": tau.value}


class Posterior:
    """fable_posterior: the row posterior resident in HBM + batched draw / covariance kernels."""

    @classmethod
    def from_rows(cls, n, gamma0, delta0sq, rows, varInflation, ctx=None):
        """fable_posterior_create_from_rows: `rows` = row_posterior(...) output for ALL variables (e.g. gathered
        from column shards); Y is not needed."""
        self = cls.__new__(cls)
        self.ctx = _ctx(ctx)
        G0 = _F(rows["G0"]); gnd = _F(rows["gnd"]); sig = _F(rows["sig"]); sh = _F(rows["sigmahat"])
        p, k = G0.shape
        self.n, self.p, self.k = int(n), p, k
        h = C.c_void_p()
        self.ctx.check(_load().fable_posterior_create_from_rows(self.ctx._h, _i64(n), _i64(p), C.c_int(k), C.c_double(gamma0),
                                                                C.c_double(delta0sq), _ptr(G0), _ptr(gnd), _ptr(sig), _ptr(sh),
                                                                C.c_double(rows["tausq"]), C.c_double(varInflation), C.byref(h)))
        self._h = h
        return self

    def __init__(self, Y, gamma0, delta0sq, U_Y, V_Y, svalsY, kEst, varInflation, ctx=None):
        self.ctx = _ctx(ctx)
        Y, U, V, d, n, p, r = _inputs(Y, U_Y, V_Y, svalsY)
        self.n, self.p, self.k = n, p, int(kEst)
        h = C.c_void_p()
        self.ctx.check(_load().fable_posterior_create(self.ctx._h, _ptr(Y), _i64(n), _i64(p), C.c_double(gamma0),
                                                      C.c_double(delta0sq), _ptr(U), _ptr(V), _ptr(d), _i64(r),
                                                      C.c_int(self.k), C.c_double(varInflation), C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            _load().fable_posterior_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def get(self):
        G0 = np.empty((self.p, self.k), order="F"); gnd = np.empty(self.p); sig = np.empty(self.p); sc = np.empty(4)
        self.ctx.check(_load().fable_posterior_get(self._h, _ptr(G0), _ptr(gnd), _ptr(sig), _ptr(sc)))
        return {"G0": G0, "gnd": gnd, "sig": sig, "tausq": sc[0], "a_n": sc[1], "gamma_n": sc[2], "w": sc[3]}

    def draw_batch(self, m0, B, dev_psi=0, psi_slots=0, accumulate=True, dev_gam=0, dev_z=0):
        """dev_* are raw device addresses (ints), e.g. torch tensors' data_ptr()."""
        self.ctx.check(_load().fable_posterior_draw_batch(self._h, _i64(m0), C.c_int(B), C.c_void_p(dev_psi or None),
                                                          C.c_int(psi_slots), C.c_int(1 if accumulate else 0),
                                                          C.c_void_p(dev_gam or None), C.c_void_p(dev_z or None)))

    def set_entrywise_correction(self, on=True):
        """f3 (R/FABLE_code.R:212-298): materialised draws become G + B o (Lambda Lambda' - G) + diag(sigma2)."""
        self.ctx.check(_load().fable_posterior_set_entrywise_correction(self._h, C.c_int(1 if on else 0)))

    def draw_batch_tiles(self, m0, B, dev_psi_tiles=0, psi_slots=0, accumulate=True, dev_gam=0, dev_z=0):
        """draw_batch with the draws materialised compactly for the posterior's tile slice (partition B): per slot
        2 * tile_count(p, *slice) * 4096 doubles, tile c as the pair (block (I, J), mirrored block (J, I))."""
        self.ctx.check(_load().fable_posterior_draw_batch_tiles(self._h, _i64(m0), C.c_int(B), C.c_void_p(dev_psi_tiles or None),
                                                                C.c_int(psi_slots), C.c_int(1 if accumulate else 0),
                                                                C.c_void_p(dev_gam or None), C.c_void_p(dev_z or None)))

    def tile_slots(self):
        n = C.c_int64()
        self.ctx.check(_load().fable_posterior_tile_slots(self._h, C.byref(n)))
        return n.value

    def set_tile_range(self, t_begin, t_end):
        """Partition B: restrict draw_batch to a slice of the tile enumeration (0, 0 = everything); the accumulators
        of the previous slice are dropped."""
        self.ctx.check(_load().fable_posterior_set_tile_range(self._h, _i64(t_begin), _i64(t_end)))
        self._slice = (int(t_begin), int(t_end))

    def tiles_to_dense(self, dev_tiles, paired, col0, ncols, dev_out):
        """compact tiles of the slice (device address) -> columns [col0, col0 + ncols) of the dense p x p matrix
        (device address, p x ncols, ld p); zeros outside the slice."""
        self.ctx.check(_load().fable_posterior_tiles_to_dense(self._h, C.c_void_p(dev_tiles), C.c_int(1 if paired else 0),
                                                              _i64(col0), _i64(ncols), C.c_void_p(dev_out)))

    def samples(self, m0, MC, gam=None, z=None, want_lambda=True, want_sigma=True):
        L = np.empty((MC, self.k * self.p), order="F") if want_lambda else None
        S = np.empty((MC, self.p), order="F") if want_sigma else None
        g = None if gam is None else np.ascontiguousarray(gam, dtype=np.float64)
        zz = None if z is None else np.ascontiguousarray(z, dtype=np.float64)
        self.ctx.check(_load().fable_posterior_samples(self._h, _i64(m0), _i64(MC), _i64(MC), _ptr(L), _ptr(S),
                                                       _ptr(g), _ptr(zz)))
        return L, S

    def sample_range(self, m0, MC, out_lambda=None, out_sigma=None, gam=None, z=None, psi_pass=True):
        """fable_posterior_sample_range: draws m0 .. m0+MC-1 -- reference-layout samples into the given column-major
        host arrays (MC x k p, MC x p; None to skip), the covariance pass into the device-resident accumulators."""
        g = None if gam is None else np.ascontiguousarray(gam, dtype=np.float64)
        zz = None if z is None else np.ascontiguousarray(z, dtype=np.float64)
        for a, cols in ((out_lambda, self.k * self.p), (out_sigma, self.p)):
            if a is not None and (a.shape != (MC, cols) or not a.flags.f_contiguous or a.dtype != np.float64):
                raise FableError(1, "sample_range: outputs must be column-major float64 arrays of MC rows")
        self.ctx.check(_load().fable_posterior_sample_range(self._h, _i64(MC), _i64(m0), _i64(MC), _ptr(g), _ptr(zz),
                                                            _ptr(out_lambda), _ptr(out_sigma), C.c_int(1 if psi_pass else 0)))

    def post_processing(self, MC, alpha, SelectedIndices=None, m0=0, gam=None, z=None):
        """CPPCCFABLEPostProcessing / CCFABLEPostProcessingSubmatrix (cpp:631-818) on RESIDENT draws m0 .. m0+MC-1: entry-wise
        mean and (alpha/2, 1-alpha/2) sample quantiles of Psi with the draws generated and consumed on the device."""
        sel = None if SelectedIndices is None else np.ascontiguousarray(SelectedIndices, dtype=np.int64)
        n = self.p if sel is None else sel.size
        g = None if gam is None else np.ascontiguousarray(gam, dtype=np.float64)
        zz = None if z is None else np.ascontiguousarray(z, dtype=np.float64)
        M = np.empty((n, n), order="F"); Lo = np.empty((n, n), order="F"); Up = np.empty((n, n), order="F")
        self.ctx.check(_load().fable_posterior_post_processing(
            self._h, _i64(MC), _i64(m0), C.c_double(alpha), None if sel is None else sel.ctypes.data_as(C.POINTER(C.c_int64)),
            _i64(n), _ptr(g), _ptr(zz), _ptr(M), _ptr(Lo), _ptr(Up)))
        return {"PostMeanMatrix": M, "LowerQuantileMatrix": Lo, "UpperQuantileMatrix": Up}

    def accum_dev(self):
        """(dev_sum, dev_sumsq, count): the compact upper tiles of the slice, `count` doubles each -- what the NCCL sum
        of the draw partition reduces, as it is."""
        a = C.c_void_p(); b = C.c_void_p(); n = C.c_int64()
        self.ctx.check(_load().fable_posterior_accum_dev(self._h, C.byref(a), C.byref(b), C.byref(n)))
        return int(a.value), int(b.value), n.value

    def accum_reset(self):
        self.ctx.check(_load().fable_posterior_accum_reset(self._h))

    def accum_fetch(self, out=None):
        """Dense symmetric p x p sum / sumsq on the host (zeros outside the posterior's tile slice)."""
        s1, s2 = out if out is not None else (np.empty((self.p, self.p), order="F"), np.empty((self.p, self.p), order="F"))
        self.ctx.check(_load().fable_posterior_accum_fetch(self._h, _ptr(s1), _ptr(s2)))
        return s1, s2

    def accum_fetch_tiles(self, out=None):
        """The compact tiles as stored: two arrays (tiles, 64, 64), element [c, v, u] = entry (u, v) of tile c.
        `out`: two flat float64 host arrays of the right length (e.g. pinned) to fill instead."""
        _, _, n = self.accum_dev()
        s1, s2 = out if out is not None else (np.empty(n), np.empty(n))
        if s1.size != n or s2.size != n:
            raise FableError(1, "accum_fetch_tiles: out arrays must hold tiles * 4096 doubles")
        self.ctx.check(_load().fable_posterior_accum_fetch_tiles(self._h, _ptr(s1), _ptr(s2)))
        return s1.reshape(-1, 64, 64), s2.reshape(-1, 64, 64)


# ------------------------------------------------------------------------------------------------
# tile layout of the covariance pass (host-only helpers of the C-ABI; no GPU needed)
# ------------------------------------------------------------------------------------------------
def tile_slots(p):
    n = C.c_int64()
    if _load().fable_tile_slots(_i64(p), C.byref(n)) != 0:
        raise FableError(1, "tile_slots: p must be >= 1")
    return n.value


def tile_count(p, t_begin=0, t_end=0):
    """Valid (stored) tiles in the slice [t_begin, t_end) of the enumeration (0, 0 = everything)."""
    n = C.c_int64()
    if _load().fable_tile_count(_i64(p), _i64(t_begin), _i64(t_end), C.byref(n)) != 0:
        raise FableError(1, "tile_count: slice out of bounds")
    return n.value


def tile_partition(p, rank, world):
    """[t_begin, t_end) of rank `rank`: slices with equal shares of the valid tiles."""
    a = C.c_int64(); b = C.c_int64()
    if _load().fable_tile_partition(_i64(p), C.c_int(rank), C.c_int(world), C.byref(a), C.byref(b)) != 0:
        raise FableError(1, "tile_partition: need p >= 1 and 0 <= rank < world")
    return a.value, b.value


def tile_locate(p, I, J):
    """(slot, compact index) of tile (I, J), I <= J."""
    a = C.c_int64(); b = C.c_int64()
    if _load().fable_tile_locate(_i64(p), _i64(I), _i64(J), C.byref(a), C.byref(b)) != 0:
        raise FableError(1, "tile_locate: need 0 <= I <= J < ceil(p / 64)")
    return a.value, b.value


# ------------------------------------------------------------------------------------------------
# device-pointer primitives (addresses as ints, e.g. torch tensors' data_ptr())
# ------------------------------------------------------------------------------------------------
def dev_rankk_cov(ctx, dA, lda, ds, p, k, dOut):
    ctx.check(_load().fable_dev_rankk_cov(ctx._h, C.c_void_p(dA), _i64(lda), C.c_void_p(ds or None), _i64(p),
                                          C.c_int(k), C.c_void_p(dOut)))


def dev_gemm(ctx, a_kfast, b_kfast, M, N, K, alpha, dA, lda, dB, ldb, dC, ldc):
    ctx.check(_load().fable_dev_gemm(ctx._h, C.c_int(a_kfast), C.c_int(b_kfast), _i64(M), _i64(N), _i64(K),
                                     C.c_double(alpha), C.c_void_p(dA), _i64(lda), C.c_void_p(dB), _i64(ldb),
                                     C.c_void_p(dC), _i64(ldc)))


def dev_syrk(ctx, a_kfast, M, K, alpha, dA, lda, dC):
    ctx.check(_load().fable_dev_syrk(ctx._h, C.c_int(a_kfast), _i64(M), _i64(K), C.c_double(alpha), C.c_void_p(dA),
                                     _i64(lda), C.c_void_p(dC)))


def dev_rowstats(ctx, dY, n, p, dU, dC, ldc, k, ds, dresid):
    ctx.check(_load().fable_dev_rowstats(ctx._h, C.c_void_p(dY), _i64(n), _i64(p), C.c_void_p(dU), C.c_void_p(dC),
                                         _i64(ldc), C.c_int(k), C.c_void_p(ds or None), C.c_void_p(dresid)))


def dev_syevj(ctx, dG, m, dW, dQ):
    sweeps = C.c_int()
    ctx.check(_load().fable_dev_syevj(ctx._h, C.c_void_p(dG), _i64(m), C.c_void_p(dW), C.c_void_p(dQ),
                                      C.byref(sweeps)))
    return sweeps.value


# ------------------------------------------------------------------------------------------------
# a11: the two public wrappers (R/FABLE_wrapper.R)
# ------------------------------------------------------------------------------------------------
def FABLEPosteriorMean(Y, gamma0=1.0, delta0sq=1.0, maxProp=0.5, ctx=None, composed=False):
    """R/FABLE_wrapper.R:80-128.  Default: one C-ABI call (fable_wrapper_posterior_mean: Y uploaded once,
    everything resident); composed=True replays the wrapper's own sequence of exports."""
    t0 = time.perf_counter()                                             # wrapper:85
    Y = _F(Y, "Y")                                                       # as.matrix, wrapper:87
    if composed:
        svdY = svd(Y, ctx)                                               # wrapper:90
        kMax = kmax_rule(svdY["d"], maxProp)                             # wrapper:94
        mod = CPPFABLEPostMean(Y, gamma0, delta0sq, svdY["u"], svdY["v"], svdY["d"], kMax, ctx)   # wrapper:102
        Cov, kEst = mod["CovPostMean"], mod["kEst"]
    else:
        ctx = _ctx(ctx)
        n, p = Y.shape
        r = min(n, p)
        U = np.empty((n, r), order="F"); d = np.empty(r); V = np.empty((p, r), order="F")
        Cov = np.empty((p, p), order="F"); k = C.c_int(); kM = C.c_int()
        ctx.check(_load().fable_wrapper_posterior_mean(ctx._h, _ptr(Y), _i64(n), _i64(p), C.c_double(gamma0),
                                                       C.c_double(delta0sq), C.c_double(maxProp), _ptr(Cov), C.byref(k),
                                                       C.byref(kM), _ptr(U), _ptr(d), _ptr(V)))
        svdY = {"d": d, "u": U, "v": V}; kEst = k.value
    return {"FABLEPostMean": Cov, "svdY": svdY, "estRank": kEst,
            "runTime": time.perf_counter() - t0}                         # wrapper:121-124


def FABLEPosteriorSampler(Y, gamma0=1.0, delta0sq=1.0, maxProp=0.95, MC=1000, ctx=None, gam=None, z=None,
                          composed=False, want_G=True):
    """R/FABLE_wrapper.R:12-68.  Default: fable_wrapper_sampler_begin + fable_posterior_sampler_outputs (Y
    uploaded once, every intermediate resident).  Injected draws
    (gam, z -- z may be a callable (MC, k, p) -> array because k is only known after the rank search) or
    composed=True replay the wrapper's own sequence of exports."""
    t0 = time.perf_counter()                                             # wrapper:18
    Y = _F(Y, "Y")
    if composed or gam is not None or z is not None:
        svdY = svd(Y, ctx)                                               # wrapper:23
        U, V, d = svdY["u"], svdY["v"], svdY["d"]
        kMax = kmax_rule(d, maxProp)                                     # wrapper:27
        kEst = CPPRankEstimator(Y, U, V, d, kMax, ctx)                   # wrapper:29
        hyp = FABLEHyperParameters(Y, U, V, d, kEst, ctx, want_G=want_G)  # wrapper:35
        # wrapper:41-44 -- cov_correct_matrix + (sum(B)/(p(p+1)/2))^2, fused: B is never stored
        varInfl = var_inflation(hyp["SigmaSqEstimate"], hyp["G0"], "wrapper", ctx)
        zz = z(MC, kEst, Y.shape[1]) if callable(z) else z
        smp = CPPFABLESampler(Y, gamma0, delta0sq, MC, U, V, d, kEst, varInfl, ctx, gam=gam, z=zz, want_G=want_G)   # wrapper:46
    else:
        ctx = _ctx(ctx)
        n, p = Y.shape
        r = min(n, p)
        MC = int(MC)
        if MC < 1:
            raise FableError(1, "MC must be >= 1")
        U = np.empty((n, r), order="F"); d = np.empty(r); V = np.empty((p, r), order="F")
        k = C.c_int(); kM = C.c_int(); htau = C.c_double(); vi = C.c_double(); tau = C.c_double()
        hsig = np.empty(p)
        hG = np.empty((p, p), order="F") if want_G else None
        G = np.empty((p, p), order="F") if want_G else None
        # step 1 (wrapper:20-44): everything up to varInflation, row posterior left resident
        ph = C.c_void_p()
        ctx.check(_load().fable_wrapper_sampler_begin(
            ctx._h, _ptr(Y), _i64(n), _i64(p), C.c_double(gamma0), C.c_double(delta0sq), C.c_double(maxProp),
            _ptr(U), _ptr(d), _ptr(V), C.byref(kM), C.byref(k), _ptr(hsig), None, C.byref(htau), _ptr(hG), C.byref(vi),
            C.byref(ph)))
        kEst = k.value; varInfl = vi.value
        svdY = {"d": d, "u": U, "v": V}
        # step 2 (wrapper:46-54): the caller sizes FABLEHyperParameters$G0 and LambdaSamples now that k is known
        hG0 = np.empty((p, kEst), order="F")
        L = np.empty((MC, kEst * p), order="F"); S = np.empty((MC, p), order="F"); pm = np.empty(p); pi = np.empty(p)
        try:
            ctx.check(_load().fable_posterior_hyper_G0(ph, _ptr(hG0)))
            hyp = {"SigmaSqEstimate": hsig, "G": hG, "EstimatedRank": kEst, "G0": hG0, "tauSqEstimate": htau.value}
            ctx.check(_load().fable_posterior_sampler_outputs(ph, _i64(MC), _i64(0), None, None, _ptr(L), _ptr(S), _ptr(pm),
                                                              _ptr(pi), C.byref(tau), _ptr(G), None, None))
        finally:
            _load().fable_posterior_destroy(ph)
        smp = {"LambdaSamples": L, "SigmaSqSamples": S, "SigmaSqEstimatePostMean": pm, "SigmaSqEstimatePlugIn": pi,
               "EstimatedRank": kEst, "VarianceInflation": varInfl, "tauSqEstimate": tau.value, "G": G}
    return {"CCFABLESamples": smp, "FABLEHyperParameters": hyp, "svdY": svdY, "estRank": kEst,
            "varInflation": varInfl, "runTime": time.perf_counter() - t0}   # wrapper:59-64
