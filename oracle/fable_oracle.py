"""CPU oracle for the FABLE estimation + sampling hot path.  TEST INFRASTRUCTURE ONLY.

This module is a plain numpy FP64 restatement of the reference algorithm
(shounakch/FABLE, `src/updated-FABLE-functions.cpp`, `R/FABLE_wrapper.R`, `R/FABLE_code.R`);
every function cites the reference file:line it follows.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import it.
The product (`fable_b200`) never does.

PINNING.  The reference ships no tests, golden vectors or fixtures and cannot be built
as an R package in this image (no R / Rcpp / Armadillo / system LAPACK).  Its single translation unit,
src/updated-FABLE-functions.cpp, IS however compiled here where it lies, against a minimal stand-in for the
Armadillo/Rcpp subset it uses (oracle/shim/, oracle/ref_bridge.cpp -> oracle/_ref/libfable_ref.so), and
tests/test_reference_source.py checks this restatement and the frozen fixtures against it to 1e-12:
formulas, index conventions, bisection control flow, list layouts, quantile interpolation and the
random-draw consumption order are therefore pinned to the reference's own function bodies.  What remains
UNPINNED is the third-party arithmetic under them (Armadillo expression evaluation order, LAPACK dgesdd,
R's RNG streams), restated from published definitions.  Further anchors: the analytic known answers of
SURVEY.md App. A.6 (tests/test_oracle.py) and the frozen fixtures under tests/golden/.

Third-party arithmetic the reference delegates and that is restated here from its published
definition (none of it is under /root/reference, versions unpinned in DESCRIPTION:10-11):
  * base-R `svd()` -> LAPACK dgesdd: numpy.linalg.svd (also dgesdd, bundled OpenBLAS).
  * Armadillo `quantile()`: Hyndman & Fan definition 5 (as recalled from Armadillo's op_quantile;
    unverifiable here) -- see `arma_quantile`.
  * Armadillo `randg` / `randn`: stream not reproducible; the sampler takes INJECTED draws
    (SURVEY.md App. A.4), which is the parity contract (BASELINE.json north_star).

Layouts: all matrices are numpy float64; "column-major" statements refer to the R/Armadillo memory
order that the C-ABI uses -- here we index (row, col) and let numpy own the strides.
"""
from __future__ import annotations

import math
import numpy as np

__all__ = [
    "svd_thin", "kmax_rule", "criterion", "bisection", "rank_estimator", "row_stats",
    "hyper_parameters", "cov_correct_matrix_R", "cpp_cov_correct_matrix", "var_inflation_wrapper",
    "var_inflation_script", "post_mean", "sampler", "psi_draw", "psi_moments", "arma_quantile",
    "criterion_R", "rank_estimator_R", "ccfable_direct_sampler",
    "post_processing", "post_processing_submatrix", "symmetrize", "FABLEPosteriorMean",
    "FABLEPosteriorSampler", "log_likelihood_eval", "direct_sampler_psi",
]


# --------------------------------------------------------------------------------------------
# a1/a2  SVD + kMax rule  (R/FABLE_wrapper.R:23-27, 90-94)
# --------------------------------------------------------------------------------------------
def svd_thin(Y):
    """base-R `svd(Y)` (R/FABLE_wrapper.R:23, 90): thin SVD, d descending, u n x r, v p x r."""
    u, d, vt = np.linalg.svd(np.asarray(Y, dtype=np.float64), full_matrices=False)
    return u, d, vt.T.copy()


def kmax_rule(d, maxProp):
    """`min(which(cumsum(d)/sum(d) >= maxProp))` (R/FABLE_wrapper.R:27, 94); 1-based."""
    d = np.asarray(d, dtype=np.float64)
    # R accumulates both sum() and cumsum() sequentially in long double and rounds to double
    # (R sources summary.c rsum / cum.c cumsum), so cumsum(d)[r] == sum(d) exactly.
    cs = np.cumsum(d.astype(np.longdouble))
    frac = cs.astype(np.float64) / np.float64(cs[-1])
    idx = np.nonzero(frac >= maxProp)[0]
    if idx.size == 0:
        raise ValueError("kMax rule: no index reaches maxProp (R would return Inf with a warning)")
    return int(idx[0]) + 1


# --------------------------------------------------------------------------------------------
# a3  CPPCriterionFunction  (src/updated-FABLE-functions.cpp:127-213)
# --------------------------------------------------------------------------------------------
def _residual_sigsq(Y, U, V, d, k):
    """sigsq_hat_diag = colsum((Y - U_k D_k V_k')^2)/n  (cpp:161-166, 371-379, 470-478, 564-572)."""
    n = Y.shape[0]
    UDVt = (U[:, :k] * d[:k]) @ V[:, :k].T
    R = Y - UDVt
    return np.sum(R * R, axis=0) / n


def criterion(kInd, Y, U, V, d):
    """JIC(k) (cpp:128-213): -2*LL + k*max(n,p)*log(min(n,p)), LL = -n*p/2 - n*sum(log sig)/2.

    cpp:199 evaluates -n*p in int arithmetic (overflow for n*p > 2^31-1); restated
    in double, which is identical for every BASELINE config.
    """
    n, p = Y.shape
    sig = _residual_sigsq(Y, U, V, d, kInd)
    LL = (-n * p / 2.0) - (n * float(np.sum(np.log(sig))) / 2.0)
    return (-2.0 * LL) + (kInd * float(max(n, p)) * math.log(float(min(n, p))))


# --------------------------------------------------------------------------------------------
# a4  CPPBisectionRecursion  (cpp:215-271)
# --------------------------------------------------------------------------------------------
def bisection(lower, upper, crit_fn, trace=None):
    """Window shrink (cpp:216-271). `crit_fn(k)->float`. Integer midpoints (cpp:225, 243).

    `lower` is re-evaluated at every level (cpp:233).  Returns (lower, upper)."""
    while upper - lower > 5:                       # cpp:223
        mid = (lower + upper) // 2                 # cpp:225
        c_lo = crit_fn(lower)                      # cpp:233
        c_mid = crit_fn(mid)                       # cpp:235
        if trace is not None:
            trace.append(("lo", lower, c_lo)); trace.append(("mid", mid, c_mid))
        if c_lo < c_mid:                           # cpp:237
            upper = mid                            # cpp:239
        else:
            tq = (mid + upper) // 2                # cpp:243
            c_tq = crit_fn(tq)                     # cpp:247
            if trace is not None:
                trace.append(("tq", tq, c_tq))
            if c_mid <= c_tq:                      # cpp:249
                upper = mid                        # cpp:251
            else:
                lower = mid                        # cpp:255
    return lower, upper                            # cpp:263-267


# --------------------------------------------------------------------------------------------
# a5  CPPRankEstimator  (cpp:281-341)
# --------------------------------------------------------------------------------------------
def rank_estimator(Y, U, V, d, kMax, trace=None):
    """Bisection from [1,kMax] (cpp:314), then every k in the final window (cpp:321-330),
    first minimum wins (`index_min`, cpp:337)."""
    Y = np.asarray(Y, dtype=np.float64)
    fn = lambda k: criterion(k, Y, U, V, d)
    lo, hi = bisection(1, int(kMax), fn, trace)
    ks = list(range(lo, hi + 1))                   # regspace(lo,1,hi), cpp:322
    vals = [fn(k) for k in ks]
    if trace is not None:
        trace.extend(("win", k, v) for k, v in zip(ks, vals))
    return ks[int(np.argmin(vals))]                # np.argmin = first minimum = index_min


# --------------------------------------------------------------------------------------------
# A.2 row statistics shared by a6/a8/a9
# --------------------------------------------------------------------------------------------
def row_stats(Y, U, V, d, k):
    """c=YtU (p x k) = V_k .* d_k (cpp:374-376); q_j = sum_l c_jl^2; s_j = sum_i Y_ij^2;
    sigsq (explicit residual); tausq = mean_j(q_j/sigsq_j)/(n k) (cpp:383, 483, 577)."""
    n, p = Y.shape
    c = V[:, :k] * d[:k]
    q = np.sum(c * c, axis=1)
    s = np.sum(Y * Y, axis=0)
    sig = _residual_sigsq(Y, U, V, d, k)
    tausq = float(np.mean(q / sig)) / (n * k)
    return c, q, s, sig, tausq


# --------------------------------------------------------------------------------------------
# a6  FABLEHyperParameters  (cpp:343-399)  -- UNSHRUNK G0
# --------------------------------------------------------------------------------------------
def hyper_parameters(Y, U, V, d, kEst, want_G=True):
    Y = np.asarray(Y, dtype=np.float64)
    n, p = Y.shape
    c, q, s, sig, tausq = row_stats(Y, U, V, d, kEst)
    G0 = c / math.sqrt(n)                          # cpp:390
    out = {"SigmaSqEstimate": sig, "EstimatedRank": int(kEst), "G0": G0, "tauSqEstimate": tausq}
    if want_G:
        out["G"] = G0 @ G0.T                       # cpp:391
    return out


# --------------------------------------------------------------------------------------------
# a7  cov_correct_matrix (R/FABLE_code.R:189-206) and CPPcov_correct_matrix (cpp:401-428)
# --------------------------------------------------------------------------------------------
def cov_correct_matrix_R(sigsq_hat, llprime_hat):
    """Full p x p B (R/FABLE_code.R:196-204)."""
    g = np.diag(llprime_hat).astype(np.float64)
    num = np.outer(g, g) + llprime_hat ** 2        # R:196
    den = np.outer(g, sigsq_hat) + np.outer(sigsq_hat, g)   # R:197
    B = num / den                                  # R:199
    B[np.diag_indices_from(B)] = np.diag(B) / 2.0  # R:200
    return np.sqrt(1.0 + B)                        # R:202


def cpp_cov_correct_matrix(sigsq_hat, llprime_hat):
    """Upper-triangle VECTOR in column-major trimatu_ind order (cpp:413-426)."""
    B = cov_correct_matrix_R(sigsq_hat, llprime_hat)
    p = B.shape[0]
    return np.concatenate([B[: j + 1, j] for j in range(p)])


def var_inflation_wrapper(B):
    """(sum(B) / (p(p+1)/2))^2 with the FULL p x p sum (R/FABLE_wrapper.R:44)."""
    p = B.shape[0]
    return float((np.sum(B) / (p * (p + 1) / 2.0)) ** 2)


def var_inflation_script(Bupper):
    """mean(upper-triangle vector)^2 (extras/replicationCodes/runtimeScript.R:78)."""
    return float(np.mean(Bupper) ** 2)


# --------------------------------------------------------------------------------------------
# a8  CPPFABLEPostMean  (cpp:440-519)  -- SHRUNK G0
# --------------------------------------------------------------------------------------------
def _shrunk_stats(Y, U, V, d, k, gamma0, delta0sq):
    n, p = Y.shape
    c, q, s, sig, tausq = row_stats(Y, U, V, d, k)
    w = n + (1.0 / tausq)
    G0 = (math.sqrt(n) / w) * c                    # cpp:490, 584
    gamma_n = gamma0 + n                           # cpp:493, 586
    impCoef = n / w                                # cpp:497, 590
    gnd = (gamma0 * delta0sq) + s - impCoef * q    # cpp:499, 592
    return dict(c=c, q=q, s=s, sig=sig, tausq=tausq, w=w, G0=G0, gamma_n=gamma_n, gnd=gnd)


def post_mean(Y, gamma0, delta0sq, U, V, d, kMax, want_cov=True):
    Y = np.asarray(Y, dtype=np.float64)
    k = rank_estimator(Y, U, V, d, kMax)           # cpp:457
    st = _shrunk_stats(Y, U, V, d, k, gamma0, delta0sq)
    SigmaHat = st["gnd"] / (st["gamma_n"] - 2.0)   # cpp:509
    out = {"kEst": int(k), "G0": st["G0"], "SigmaHat": SigmaHat, "tauSqEstimate": st["tausq"],
           "SigmaSqEstimatePlugIn": st["sig"]}
    if want_cov:
        G = st["G0"] @ st["G0"].T                  # cpp:491
        out["CovPostMean"] = G + np.diag(SigmaHat)  # cpp:512 (diagonal quirk: SURVEY Q4)
    return out


# --------------------------------------------------------------------------------------------
# a9  CPPFABLESampler  (cpp:532-629) with INJECTED draws
# --------------------------------------------------------------------------------------------
def sampler(Y, gamma0, delta0sq, MC, U, V, d, kEst, varInflation, gam, z, want_G=True):
    """gam: (MC, p) Gamma(shape=gamma_n/2, scale=1) draws, gam[m, j] (flat: m*p + j).
    z: (MC, k, p) standard normals, z[m, l, j] (flat: m*p*k + l*p + j) -- Armadillo fills the
    p x k ZSample in column-major order, gammas before normals within a draw (cpp:604, 610)."""
    Y = np.asarray(Y, dtype=np.float64)
    n, p = Y.shape
    k = int(kEst)
    st = _shrunk_stats(Y, U, V, d, k, gamma0, delta0sq)
    a_n = math.sqrt(varInflation) / math.sqrt(st["w"])       # cpp:597
    gam = np.asarray(gam, dtype=np.float64).reshape(MC, p)
    z = np.asarray(z, dtype=np.float64).reshape(MC, k, p)
    sigsq_s = (0.5 * st["gnd"])[None, :] / gam               # cpp:604, 606   (MC, p)
    Zs = z * np.sqrt(sigsq_s)[:, None, :]                    # cpp:613        (MC, k, p)
    Lam = st["G0"].T[None, :, :] + a_n * Zs                  # cpp:615        (MC, k, p)[m,l,j]
    # LambdaSamples row m = vectorise(Lambda_m.t()) (cpp:616): element (m, j*k + l)
    LambdaSamples = np.transpose(Lam, (0, 2, 1)).reshape(MC, p * k)
    out = {"LambdaSamples": LambdaSamples, "SigmaSqSamples": sigsq_s,
           "SigmaSqEstimatePostMean": st["gnd"] / (st["gamma_n"] - 2.0),     # cpp:622
           "SigmaSqEstimatePlugIn": st["sig"], "EstimatedRank": k,           # cpp:623-624
           "VarianceInflation": float(varInflation), "tauSqEstimate": st["tausq"],
           "G0": st["G0"], "gnd": st["gnd"], "a_n": a_n, "gamma_n": st["gamma_n"]}
    if want_G:
        out["G"] = st["G0"] @ st["G0"].T                     # cpp:627
    return out


# --------------------------------------------------------------------------------------------
# Psi entries and summaries  (cpp:631-713, 716-818)
# --------------------------------------------------------------------------------------------
def psi_draw(LambdaSamples_row, SigmaSq_row, k):
    """Psi_m[u,v] = sum_l Lam_m[u,l] Lam_m[v,l] + delta_uv sigma2_m[u]  (cpp:672-684)."""
    p = SigmaSq_row.shape[0]
    L = np.asarray(LambdaSamples_row).reshape(p, k)
    return L @ L.T + np.diag(SigmaSq_row)


def psi_moments(LambdaSamples, SigmaSqSamples, k):
    """Streaming summaries: (sum_m Psi_m, sum_m Psi_m^2) entry-wise, full symmetric p x p.
    The mean is cpp:688 (`accu/nMC`); the second moment has no reference twin."""
    MC, p = SigmaSqSamples.shape
    s1 = np.zeros((p, p)); s2 = np.zeros((p, p))
    for m in range(MC):
        P = psi_draw(LambdaSamples[m], SigmaSqSamples[m], k)
        s1 += P; s2 += P * P
    return s1, s2


def arma_quantile(x, prob):
    """Armadillo `quantile(vec, P)`: Hyndman-Fan definition 5 (as recalled; see module header)."""
    x = np.sort(np.asarray(x, dtype=np.float64))
    N = float(x.size)
    if prob < (1.0 - 0.5) / N:
        return -math.inf if prob < 0 else float(x[0])
    if prob > (N - 0.5) / N:
        return math.inf if prob > 1 else float(x[-1])
    h = N * prob + 0.5
    kk = int(math.floor(h))
    w = h - kk
    if kk >= x.size:           # prob == P_max exactly: weight 0 on the (non-existent) upper neighbour -> the maximum
        return float(x[-1])
    return float((1.0 - w) * x[kk - 1] + w * x[kk])


def symmetrize(A):
    """SymmetrizeMatrix (cpp:28-37): A + A' with the diagonal halved."""
    A1 = A + A.T
    A1[np.diag_indices_from(A1)] = np.diag(A1) / 2.0
    return A1


def post_processing_submatrix(out, alpha, selected_1based):
    """CCFABLEPostProcessingSubmatrix (cpp:716-818): entry-wise mean and (alpha/2, 1-alpha/2)
    sample quantiles over the MC draws on the selected index set."""
    Ls, Ss, k = out["LambdaSamples"], out["SigmaSqSamples"], int(out["EstimatedRank"])
    sel = np.asarray(selected_1based, dtype=np.int64) - 1        # cpp:721
    S = sel.size
    Mn = np.zeros((S, S)); Lo = np.zeros((S, S)); Up = np.zeros((S, S))
    MC = Ss.shape[0]
    for a in range(S):
        ja = sel[a]
        La = Ls[:, ja * k:(ja + 1) * k]                          # cpp:751
        for b in range(a, S):
            jb = sel[b]
            v = np.sum(La * Ls[:, jb * k:(jb + 1) * k], axis=1)  # cpp:764
            if ja == jb:
                v = v + Ss[:, ja]                                # cpp:770-772
            Mn[a, b] = np.sum(v) / MC                            # cpp:782
            Lo[a, b] = arma_quantile(v, alpha / 2.0)             # cpp:786
            Up[a, b] = arma_quantile(v, 1.0 - alpha / 2.0)       # cpp:787
    return {"PostMeanMatrix": symmetrize(Mn), "LowerQuantileMatrix": symmetrize(Lo),
            "UpperQuantileMatrix": symmetrize(Up)}


def post_processing(out, alpha):
    """CPPCCFABLEPostProcessing (cpp:631-713) = the submatrix form on all p indices."""
    p = out["SigmaSqSamples"].shape[1]
    return post_processing_submatrix(out, alpha, np.arange(1, p + 1))


def direct_sampler_psi(Y, gamma0, delta0sq, U, V, d, k, gam, z):
    """CCFABLE_DirectSampler (R/FABLE_code.R:212-298) with injected draws: the MC x p x p array of
    entry-wise coverage-corrected Psi draws.  (The R prototype picks k with its own RankEstimator,
    R:221; here k is an input.)"""
    Y = np.asarray(Y, dtype=np.float64)
    n, p = Y.shape
    st = _shrunk_stats(Y, U, V, d, k, gamma0, delta0sq)       # R:236-257 (same sig, tau, G0, gnd as cpp:564-592)
    G0, gnd, sig = st["G0"], st["gnd"], st["sig"]
    G = G0 @ G0.T                                             # R:249
    CC = cov_correct_matrix_R(sig, G)                         # R:251
    a_n = 1.0 / math.sqrt(st["w"])                            # R:263
    gam = np.asarray(gam, dtype=np.float64).reshape(-1, p)
    MC = gam.shape[0]
    z = np.asarray(z, dtype=np.float64).reshape(MC, k, p)
    out = np.empty((MC, p, p))
    for m in range(MC):
        s2 = (0.5 * gnd) / gam[m]                             # R:271-272: 1 / rgamma(shape, rate = gnd/2)
        Z = z[m].T                                            # p x k (column-major fill, R:279)
        SZ = Z * np.sqrt(s2)[:, None]                         # R:280
        SZG = SZ @ G0.T                                       # R:281
        part2 = a_n * (SZG + SZG.T) + a_n ** 2 * ((Z @ Z.T) * np.outer(np.sqrt(s2), np.sqrt(s2)))   # R:283-284
        out[m] = G + CC * part2 + np.diag(s2)                 # R:286-290
    return out


def log_likelihood_evaluator_R(Y, M, Lambda, SigmaSq):
    """LogLikelihoodEvaluator (R/FABLE_code.R:12-28): the Gaussian log-likelihood scaled by 1 / (n p)."""
    n, p = Y.shape
    sd = np.sqrt(SigmaSq)                                     # R:17
    D = (Y - M @ Lambda.T) / sd[None, :]                      # R:19-20
    return float((-n * np.sum(np.log(sd))) + (-0.5 * np.sum(D * D))) / (n * p)    # R:21-26


def criterion_R(kInd, Y, U, V, d):
    """CriterionFunction inside the pure-R prototype RankEstimator (R/FABLE_code.R:57-116): sigma^2 = resid / (n - k) (:87),
    Lambda = YtU / sqrt(n) (:96), M = sqrt(n) U_k (:102), -2 LogLik + k max(n,p) log(min(n,p)) / (n p) (:113-114)."""
    n, p = Y.shape
    Uk, Vk, dk = U[:, :kInd], V[:, :kInd], d[:kInd]
    UDVt = (Uk * dk) @ Vk.T                                   # R:86
    sig = np.sum((Y - UDVt) ** 2, axis=0) / (n - kInd)        # R:87
    YtU = Vk * dk                                             # R:94
    ll = log_likelihood_evaluator_R(Y, math.sqrt(n) * Uk, YtU / math.sqrt(n), sig)     # R:96-106
    return (-2.0 * ll) + (kInd * max(n, p) * math.log(min(n, p)) / (n * p))             # R:113-114


def rank_estimator_R(Y, U, V, d, kMax):
    """RankEstimator (R/FABLE_code.R:32-186): BisectionRecursion (:118-164, as.integer midpoints) from [1, kMax], then the
    criterion on the whole final window and which.min = the first minimum (:166-184)."""
    fn = lambda k: criterion_R(k, Y, U, V, d)
    lo, hi = bisection(1, int(kMax), fn)
    ks = list(range(lo, hi + 1))
    return ks[int(np.argmin([fn(k) for k in ks]))]


def ccfable_direct_sampler(Y, gamma0, delta0sq, MC, gam, z, svd=None):
    """CCFABLE_DirectSampler(Y, gamma0, delta0sq, MC) (R/FABLE_code.R:212-298) with injected draws: svd, kMax at 0.95
    (:219), the prototype's RankEstimator (:221), then the draws of direct_sampler_psi.  z: array (MC, k, p) or a callable
    (MC, k, p) -> array.  Returns (k, array (MC, p, p)).  `svd` = (u, d, v) replaces svd(Y): for FIXED injected normals the
    draws depend on the signs of the singular vectors, so a comparison needs both sides on the same signs."""
    Y = np.asarray(Y, dtype=np.float64)
    u, d, v = svd if svd is not None else svd_thin(Y)         # R:216
    kMax = kmax_rule(d, 0.95)                                 # R:219
    k = rank_estimator_R(Y, u, v, d, kMax)                    # R:221
    zz = z(MC, k, Y.shape[1]) if callable(z) else z
    return k, direct_sampler_psi(Y, gamma0, delta0sq, u, v, d, k, gam, zz)


def log_likelihood_eval(Y, M, Lambda, SigmaSq):
    """CPPLogLikelihoodEval (cpp:95-125)."""
    n = Y.shape[0]
    sd = np.sqrt(SigmaSq)
    D = (Y - M @ Lambda.T) / sd[None, :]
    return float(-n * np.sum(np.log(sd)) - 0.5 * np.sum(D * D))


# --------------------------------------------------------------------------------------------
# a11  R wrappers  (R/FABLE_wrapper.R:12-68, 80-128)
# --------------------------------------------------------------------------------------------
def FABLEPosteriorMean(Y, gamma0=1.0, delta0sq=1.0, maxProp=0.5):
    Y = np.asarray(Y, dtype=np.float64)
    u, d, v = svd_thin(Y)                                    # wrapper:90
    kMax = kmax_rule(d, maxProp)                             # wrapper:94
    pm = post_mean(Y, gamma0, delta0sq, u, v, d, kMax)       # wrapper:102
    return {"FABLEPostMean": pm["CovPostMean"], "svdY": {"d": d, "u": u, "v": v},
            "estRank": pm["kEst"], "kMax": kMax, "_internals": pm}


def FABLEPosteriorSampler(Y, gamma0=1.0, delta0sq=1.0, maxProp=0.95, MC=1000, gam=None, z=None):
    """Injected draws are mandatory here (the reference's RNG stream is not reproducible).
    `z` may be a callable (MC, k, p) -> array since k is only known after rank estimation."""
    Y = np.asarray(Y, dtype=np.float64)
    n, p = Y.shape
    u, d, v = svd_thin(Y)                                    # wrapper:23
    kMax = kmax_rule(d, maxProp)                             # wrapper:27
    kEst = rank_estimator(Y, u, v, d, kMax)                  # wrapper:29
    hyp = hyper_parameters(Y, u, v, d, kEst)                 # wrapper:35
    B = cov_correct_matrix_R(hyp["SigmaSqEstimate"], hyp["G"])   # wrapper:41
    varInflation = var_inflation_wrapper(B)                  # wrapper:44
    zz = z(MC, kEst, p) if callable(z) else z
    smp = sampler(Y, gamma0, delta0sq, MC, u, v, d, kEst, varInflation, gam, zz)   # wrapper:46
    return {"CCFABLESamples": smp, "FABLEHyperParameters": hyp, "svdY": {"d": d, "u": u, "v": v},
            "estRank": kEst, "varInflation": varInflation, "kMax": kMax}
