"""GPU parity tests: every entry point of the C-ABI (through the Python mirror of the reference's
R interface) against the CPU oracle on the same inputs, the frozen golden fixtures, and -- at
BASELINE.json's full sizes -- size-independent properties.

Tolerance: BASELINE.json north_star -- 1e-10 relative in FP64 (SURVEY.md §8c metrics: matrices
max|x-ref|/max|ref|, positive vectors element-wise relative, integers exact).
"""
import math
import os

import numpy as np
import pytest

import fable_b200 as fb
from fable_b200 import synth
from oracle import fable_oracle as orc
from oracle import c_oracle
from conftest import relerr_mat, relerr_vec, TOL

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu

CASES = ["a", "b", "c"]


def _svd_inputs(n, p, k, rep):
    Y = synth.factor_data(n, p, k, rep=rep)
    u, d, v = orc.svd_thin(Y)
    return Y, u, d, v


# ---- golden fixtures ------------------------------------------------------------------------------
@pytest.mark.parametrize("name", CASES)
def test_golden_estimation(ctx, golden, name):
    g = lambda key: golden[f"{name}_{key}"]
    Y, u, d, v = g("Y"), g("u"), g("d"), g("v")
    kMax95, kMax50 = int(g("kMax95")), int(g("kMax50"))
    crit = np.array([fb.CPPCriterionFunction(k, Y, u, v, d, ctx=ctx) for k in range(1, kMax95 + 1)])
    assert relerr_vec(crit, g("crit")) < TOL
    assert fb.CPPRankEstimator(Y, u, v, d, kMax50, ctx=ctx) == int(g("kEst50"))
    kEst = fb.CPPRankEstimator(Y, u, v, d, kMax95, ctx=ctx)
    assert kEst == int(g("kEst95"))
    np.testing.assert_array_equal(fb.CPPBisectionRecursion(1, kMax95, Y, u, v, d, ctx=ctx), g("bis").astype(float))
    hyp = fb.FABLEHyperParameters(Y, u, v, d, kEst, ctx=ctx)
    assert relerr_vec(hyp["SigmaSqEstimate"], g("hyp_sig")) < TOL
    assert relerr_mat(hyp["G0"], g("hyp_G0")) < TOL
    assert relerr_mat(hyp["G"], g("hyp_G")) < TOL
    assert abs(hyp["tauSqEstimate"] / float(g("hyp_tau")) - 1) < TOL
    assert relerr_vec(fb.CPPcov_correct_matrix(hyp["SigmaSqEstimate"], hyp["G"], ctx=ctx), g("Bupper")) < TOL
    B = fb.cov_correct_matrix(hyp["SigmaSqEstimate"], hyp["G"], ctx=ctx)
    assert relerr_mat(B, orc.cov_correct_matrix_R(g("hyp_sig"), g("hyp_G"))) < TOL
    assert abs(fb.var_inflation(hyp["SigmaSqEstimate"], hyp["G0"], "wrapper", ctx=ctx) / float(g("vi_wrapper")) - 1) < TOL
    assert abs(fb.var_inflation(hyp["SigmaSqEstimate"], hyp["G0"], "script", ctx=ctx) / float(g("vi_script")) - 1) < TOL
    pm = fb.CPPFABLEPostMean(Y, 1.0, 1.0, u, v, d, kMax50, ctx=ctx, internals=True)
    assert pm["kEst"] == int(g("kEst50"))
    assert relerr_mat(pm["CovPostMean"], g("pm_cov")) < TOL
    assert relerr_mat(pm["G0"], g("pm_G0")) < TOL
    assert relerr_vec(pm["SigmaHat"], g("pm_sigmahat")) < TOL


@pytest.mark.parametrize("name", CASES)
def test_golden_sampler_injected(ctx, golden, name):
    g = lambda key: golden[f"{name}_{key}"]
    Y, u, d, v = g("Y"), g("u"), g("d"), g("v")
    kEst = int(g("kEst95"))
    out = fb.CPPFABLESampler(Y, 1.0, 1.0, 7, u, v, d, kEst, float(g("vi_wrapper")), ctx=ctx, gam=g("gam"), z=g("z"),
                             psi_moments=True)
    assert relerr_mat(out["LambdaSamples"], g("smp_L")) < TOL
    assert relerr_vec(out["SigmaSqSamples"], g("smp_S")) < TOL
    assert relerr_vec(out["SigmaSqEstimatePostMean"], g("smp_postmean")) < TOL
    assert relerr_vec(out["SigmaSqEstimatePlugIn"], g("hyp_sig")) < TOL
    assert relerr_mat(out["G"], g("smp_G")) < TOL
    assert abs(out["tauSqEstimate"] / float(g("smp_tau")) - 1) < TOL
    assert relerr_mat(out["PsiSum"], g("psi_sum")) < TOL
    assert relerr_mat(out["PsiSumSq"], g("psi_sumsq")) < TOL
    assert out["EstimatedRank"] == kEst and out["VarianceInflation"] == float(g("vi_wrapper"))


def test_gpu_against_the_reference_source(ctx):
    """The CUDA path against the reference's OWN translation unit (oracle/_ref, compiled with the shim):
    rank, hyper-parameters, posterior mean and the sampler with injected draws."""
    from oracle import ref_oracle as ref
    if not ref.available():
        pytest.skip("oracle/_ref/libfable_ref.so not present")
    n, p, k, MC = 70, 150, 3, 9
    Y, u, d, v = _svd_inputs(n, p, k, 61)
    kMax = orc.kmax_rule(d, 0.95)
    assert fb.CPPRankEstimator(Y, u, v, d, kMax, ctx=ctx) == ref.rank_estimator(Y, u, v, d, kMax)
    for kk in (1, 7, 40):
        assert abs(fb.CPPCriterionFunction(kk, Y, u, v, d, ctx=ctx) / ref.criterion(kk, Y, u, v, d) - 1) < TOL
    hg, hr = fb.FABLEHyperParameters(Y, u, v, d, k, ctx=ctx), ref.hyper_parameters(Y, u, v, d, k)
    assert relerr_mat(hg["G"], hr["G"]) < TOL and relerr_vec(hg["SigmaSqEstimate"], hr["SigmaSqEstimate"]) < TOL
    assert relerr_vec(fb.CPPcov_correct_matrix(hr["SigmaSqEstimate"], hr["G"], ctx=ctx), ref.cov_correct_matrix(hr["SigmaSqEstimate"], hr["G"])) < TOL
    pg, pr = fb.CPPFABLEPostMean(Y, 1.5, 0.7, u, v, d, kMax, ctx=ctx), ref.post_mean(Y, 1.5, 0.7, u, v, d, kMax)
    assert pg["kEst"] == pr["kEst"] and relerr_mat(pg["CovPostMean"], pr["CovPostMean"]) < TOL
    gam, z = synth.injected_draws(MC, p, k, 0.5 * (1.5 + n), seed=8)
    sg = fb.CPPFABLESampler(Y, 1.5, 0.7, MC, u, v, d, k, 1.31, ctx=ctx, gam=gam, z=z)
    sr = ref.sampler(Y, 1.5, 0.7, MC, u, v, d, k, 1.31, gam, z)
    for key in ("LambdaSamples", "SigmaSqSamples", "SigmaSqEstimatePostMean", "SigmaSqEstimatePlugIn", "G"):
        assert relerr_mat(sg[key], sr[key]) < TOL, key
    ppg = fb.CCFABLEPostProcessingSubmatrix(sg, 0.05, [2, 9, 9, 40], ctx=ctx)
    ppr = ref.post_processing(sr, 0.05, [2, 9, 9, 40])
    for key in ppr:
        assert relerr_mat(ppg[key], ppr[key]) < TOL, key


# ---- SVD replacement (a1): sign-invariant comparison + decomposition properties -------------------
@pytest.mark.parametrize("n,p,k", [(40, 120, 3), (64, 48, 2), (33, 77, 4), (200, 1000, 5), (300, 130, 4), (1, 5, 1), (5, 1, 1)])
def test_svd_against_lapack(ctx, n, p, k):
    Y = synth.factor_data(n, p, k, rep=n + p)
    s = fb.svd(Y, ctx=ctx)
    u, d, v = orc.svd_thin(Y)
    assert relerr_vec(s["d"], d) < TOL
    r = min(n, p)
    assert s["u"].shape == (n, r) and s["v"].shape == (p, r)
    assert relerr_mat((s["u"] * s["d"]) @ s["v"].T, Y) < TOL
    assert relerr_mat(s["u"].T @ s["u"], np.eye(r)) < 1e-9 and relerr_mat(s["v"].T @ s["v"], np.eye(r)) < 1e-9
    # canonical signs (largest |entry| of each v column positive) -> vectors comparable where the
    # spectrum is well separated (the leading k factor directions)
    sgn = np.sign(v[np.argmax(np.abs(v), axis=0), np.arange(r)])
    kk = min(k, r)
    assert relerr_mat(s["v"][:, :kk], (v * sgn)[:, :kk]) < 1e-8
    assert relerr_mat(s["u"][:, :kk], (u * sgn)[:, :kk]) < 1e-8


def _rank_deficient(kind, n, p, k, seed):
    Y = synth.factor_data(n, p, k, rep=seed)
    if kind == "centred":          # what the reference asks of its caller (cpp:296): one singular value is exactly 0 when n <= p
        Y = Y - Y.mean(axis=0, keepdims=True)
    elif kind == "dup_rows":       # exactly repeated observations: three zero singular values (n <= p)
        Y[-3:] = Y[:3]
    elif kind == "dup_cols":       # exactly repeated variables (p < n)
        Y[:, -2:] = Y[:, :2]
    elif kind == "low_rank":       # rank k plus nothing
        rng = np.random.default_rng(seed)
        Y = rng.standard_normal((n, k)) @ rng.standard_normal((k, p))
    return np.asfortranarray(Y)


@pytest.mark.parametrize("kind,n,p,k,nz", [("centred", 40, 120, 3, 1), ("centred", 200, 1000, 5, 1), ("centred", 500, 1000, 10, 1),
                                           ("centred", 16, 64, 2, 1), ("centred", 300, 130, 4, 0), ("dup_rows", 60, 200, 3, 3),
                                           ("dup_cols", 300, 50, 3, 2), ("low_rank", 64, 200, 5, 59), ("low_rank", 130, 40, 4, 36)])
def test_svd_rank_deficient(ctx, kind, n, p, k, nz):
    """SURVEY H4 / the reference's own precondition (column-centred Y, cpp:296): numerically zero singular values.  The Gram
    route cannot resolve them; they must come back as d = 0 with singular vectors that complete orthonormal bases on both
    sides (what LAPACK returns), never as normalised rounding noise -- and the eigensolver must still converge."""
    Y = _rank_deficient(kind, n, p, k, n + p)
    s = fb.svd(Y, ctx=ctx)
    u, d, v = orc.svd_thin(Y)
    r = min(n, p)
    good = r - nz
    assert relerr_vec(s["d"][:good], d[:good]) < TOL
    assert np.all(s["d"][good:] == 0.0) and np.all(d[good:] < 1e-9 * d[0])                # deflated exactly where LAPACK has noise
    assert relerr_mat(s["u"].T @ s["u"], np.eye(r)) < 1e-9 and relerr_mat(s["v"].T @ s["v"], np.eye(r)) < 1e-9
    assert relerr_mat((s["u"] * s["d"]) @ s["v"].T, Y) < TOL
    assert np.all(np.diff(s["d"]) <= 0)
    for maxProp in (0.5, 0.95):
        kMax = orc.kmax_rule(d, maxProp)
        assert fb.kmax_rule(s["d"], maxProp) == kMax
        if kind != "low_rank":            # (noise-free low-rank data: log of a zero residual variance, SURVEY A.6 -- no oracle)
            assert fb.CPPRankEstimator(Y, s["u"], s["v"], s["d"], kMax, ctx=ctx) == orc.rank_estimator(Y, u, v, d, kMax)
    if kind == "centred" and n <= p:      # the null left vector of column-centred data is the constant vector
        assert abs(abs(s["u"][:, -1] @ np.ones(n)) / math.sqrt(n) - 1) < 1e-9


def test_wrappers_on_centred_data(ctx):
    """FABLEPosteriorMean / FABLEPosteriorSampler on column-centred Y (n <= p), sign-invariant outputs vs the oracle."""
    n, p, k = 120, 400, 4
    Y = _rank_deficient("centred", n, p, k, 9)
    got = fb.FABLEPosteriorMean(Y, ctx=ctx); ref = orc.FABLEPosteriorMean(Y)
    assert got["estRank"] == ref["estRank"] and relerr_mat(got["FABLEPostMean"], ref["FABLEPostMean"]) < TOL
    assert got["svdY"]["d"][-1] == 0.0 and relerr_vec(got["svdY"]["d"][:-1], ref["svdY"]["d"][:-1]) < TOL
    smp = fb.FABLEPosteriorSampler(Y, MC=3, ctx=ctx)
    kMax = orc.kmax_rule(ref["svdY"]["d"], 0.95)
    assert smp["estRank"] == orc.rank_estimator(Y, ref["svdY"]["u"], ref["svdY"]["v"], ref["svdY"]["d"], kMax)
    hyp = orc.hyper_parameters(Y, ref["svdY"]["u"], ref["svdY"]["v"], ref["svdY"]["d"], smp["estRank"])
    assert relerr_vec(smp["FABLEHyperParameters"]["SigmaSqEstimate"], hyp["SigmaSqEstimate"]) < TOL
    assert relerr_mat(smp["FABLEHyperParameters"]["G"], hyp["G"]) < TOL


@pytest.mark.parametrize("n,p,k,seed", [(60, 200, 3, 1), (120, 90, 4, 2), (200, 1000, 7, 3), (80, 300, 1, 4), (300, 400, 12, 5)])
def test_rank_search_steering_takes_the_reference_branches(ctx, n, p, k, seed, monkeypatch):
    """The rank search is steered by the all-rank criterion (one pass over V when (U, V, d) is a consistent SVD of Y) and
    decided by the explicit residual where two values are close: window and rank must equal the oracle's replay of
    cpp:215-341 and the all-explicit path (FABLE_B200_RANK_FAST=0), for maxProp 0.5 and 0.95, with the caller's SVD."""
    Y, u, d, v = _svd_inputs(n, p, k, 100 + seed)
    for maxProp in (0.5, 0.95):
        kMax = orc.kmax_rule(d, maxProp)
        want_k = orc.rank_estimator(Y, u, v, d, kMax)
        want_w = orc.bisection(1, kMax, lambda kk: orc.criterion(kk, Y, u, v, d))
        for fast in ("1", "0"):
            monkeypatch.setenv("FABLE_B200_RANK_FAST", fast)
            assert fb.CPPRankEstimator(Y, u, v, d, kMax, ctx=ctx) == want_k, (maxProp, fast)
            w = fb.CPPBisectionRecursion(1, kMax, Y, u, v, d, ctx=ctx)
            np.testing.assert_array_equal(w, np.asarray(want_w, dtype=float))
    monkeypatch.delenv("FABLE_B200_RANK_FAST")


def test_rank_search_with_inconsistent_svd_inputs(ctx):
    """At the CPP* boundary U, V, d are the caller's: if they are NOT a singular value decomposition of Y the shortcut does
    not apply (the library checks Y'U = V diag(d) and U'U = I) and every criterion is the explicit residual of cpp:161-166,
    as in the reference -- whatever the inputs mean."""
    n, p, k = 70, 180, 3
    Y, u, d, v = _svd_inputs(n, p, k, 77)
    rng = np.random.default_rng(3)
    cases = {"rotated U": (np.asfortranarray(u @ np.linalg.qr(rng.standard_normal((n, n)))[0]), v, d),
             "noisy V": (u, np.asfortranarray(v + 0.05 * rng.standard_normal(v.shape)), d),
             "scaled d": (u, v, d * np.linspace(1.3, 0.7, d.size)),
             "SVD of another matrix": orc.svd_thin(synth.factor_data(n, p, k, rep=78))[0:1] + (v, d)}
    for name, (uu, vv, dd) in cases.items():
        kMax = 40
        assert fb.CPPRankEstimator(Y, uu, vv, dd, kMax, ctx=ctx) == orc.rank_estimator(Y, uu, vv, dd, kMax), name
        for kk in (1, 5, 40):
            assert abs(fb.CPPCriterionFunction(kk, Y, uu, vv, dd, ctx=ctx) / orc.criterion(kk, Y, uu, vv, dd) - 1) < TOL, name


# ---- C1: n=500, p=1000, k=10 (BASELINE.json configs[0]) end to end vs the oracle ------------------
@pytest.fixture(scope="module")
def c1():
    Y, u, d, v = _svd_inputs(500, 1000, 10, 0)
    return dict(Y=Y, u=u, d=d, v=v)


def test_c1_rank_and_postmean(ctx, c1):
    Y, u, d, v = c1["Y"], c1["u"], c1["d"], c1["v"]
    for maxProp in (0.5, 0.95):
        kMax = orc.kmax_rule(d, maxProp)
        assert fb.kmax_rule(d, maxProp) == kMax
        assert fb.CPPRankEstimator(Y, u, v, d, kMax, ctx=ctx) == orc.rank_estimator(Y, u, v, d, kMax)
    kMax = orc.kmax_rule(d, 0.5)
    ref = orc.post_mean(Y, 1.0, 1.0, u, v, d, kMax)
    got = fb.CPPFABLEPostMean(Y, 1.0, 1.0, u, v, d, kMax, ctx=ctx, internals=True)
    assert got["kEst"] == ref["kEst"]
    assert relerr_mat(got["CovPostMean"], ref["CovPostMean"]) < TOL
    assert relerr_mat(got["G0"], ref["G0"]) < TOL and relerr_vec(got["SigmaHat"], ref["SigmaHat"]) < TOL
    assert np.array_equal(got["CovPostMean"], got["CovPostMean"].T)


def test_c1_criterion_large_ranks_use_the_dmma_path(ctx, c1):
    """Ranks above 32 go through the DMMA residual GEMM: same values as the explicit residual."""
    Y, u, d, v = c1["Y"], c1["u"], c1["d"], c1["v"]
    for k in (1, 32, 33, 100, 257, 440):
        assert abs(fb.CPPCriterionFunction(k, Y, u, v, d, ctx=ctx) / orc.criterion(k, Y, u, v, d) - 1) < TOL


def test_wrapper_stage_split_adds_up(ctx, c1):
    """fable_ctx_stage_timing / fable_ctx_stage_times: the stage marks of the single-upload pipeline cover the call (their sum is
    within the call's wall time, every stage of FABLEPosteriorMean ran) and switching them on does not change the result."""
    Y = c1["Y"]
    plain = fb.FABLEPosteriorMean(Y, ctx=ctx)
    ctx.stage_timing(True)
    try:
        timed = fb.FABLEPosteriorMean(Y, ctx=ctx)
        st = ctx.stage_times()
    finally:
        ctx.stage_timing(False)
    assert timed["estRank"] == plain["estRank"]
    np.testing.assert_array_equal(timed["FABLEPostMean"], plain["FABLEPostMean"])
    assert all(v > 0.0 for v in st.values()), st
    assert sum(st.values()) <= 1e3 * timed["runTime"] * 1.05, (st, timed["runTime"])
    fb.FABLEPosteriorMean(Y, ctx=ctx)
    assert sum(ctx.stage_times().values()) == 0.0                        # switched off: cleared, nothing recorded


def test_c1_wrappers_end_to_end(ctx, c1):
    Y = c1["Y"]
    ref = orc.FABLEPosteriorMean(Y)
    got = fb.FABLEPosteriorMean(Y, ctx=ctx)
    assert got["estRank"] == ref["estRank"]
    assert relerr_vec(got["svdY"]["d"], ref["svdY"]["d"]) < TOL
    assert relerr_mat(got["FABLEPostMean"], ref["FABLEPostMean"]) < TOL       # sign-invariant
    MC = 1000
    p = Y.shape[1]
    gam, zfull = synth.injected_draws(MC, p, 10, 0.5 * (1.0 + Y.shape[0]), seed=777)
    refs = orc.FABLEPosteriorSampler(Y, MC=MC, gam=gam, z=zfull)
    gots = fb.FABLEPosteriorSampler(Y, MC=MC, ctx=ctx, gam=gam, z=zfull)
    # the single-upload pipeline (native draws) returns the same estimation outputs as the composed wrapper
    pipe = fb.FABLEPosteriorSampler(Y, MC=50, ctx=ctx)
    assert pipe["estRank"] == gots["estRank"] and abs(pipe["varInflation"] / gots["varInflation"] - 1) < 1e-13
    assert relerr_vec(pipe["svdY"]["d"], gots["svdY"]["d"]) < 1e-14
    for key in ("SigmaSqEstimate", "G0", "G"):
        assert relerr_mat(pipe["FABLEHyperParameters"][key], gots["FABLEHyperParameters"][key]) < 1e-13, key
    assert relerr_mat(pipe["CCFABLESamples"]["G"], a_G := gots["CCFABLESamples"]["G"]) < 1e-13
    ctx.set_seed(12345)
    nat = fb.CPPFABLESampler(Y, 1.0, 1.0, 50, pipe["svdY"]["u"], pipe["svdY"]["v"], pipe["svdY"]["d"], 10, pipe["varInflation"], ctx=ctx, want_G=False)
    np.testing.assert_array_equal(pipe["CCFABLESamples"]["SigmaSqSamples"], nat["SigmaSqSamples"])
    assert relerr_mat(pipe["CCFABLESamples"]["LambdaSamples"], nat["LambdaSamples"]) < 1e-12
    pm1 = fb.FABLEPosteriorMean(Y, ctx=ctx); pm2 = fb.FABLEPosteriorMean(Y, ctx=ctx, composed=True)
    assert pm1["estRank"] == pm2["estRank"] and relerr_mat(pm1["FABLEPostMean"], pm2["FABLEPostMean"]) < 1e-13
    assert gots["estRank"] == refs["estRank"] == 10
    assert abs(gots["varInflation"] / refs["varInflation"] - 1) < TOL
    a, b = gots["CCFABLESamples"], refs["CCFABLESamples"]
    assert relerr_vec(a["SigmaSqSamples"], b["SigmaSqSamples"]) < TOL          # sign-invariant
    assert relerr_mat(a["G"], b["G"]) < TOL
    assert relerr_vec(a["SigmaSqEstimatePostMean"], b["SigmaSqEstimatePostMean"]) < TOL
    # Lambda draws depend on the singular-vector signs: both SVDs canonicalised the same way
    sgn = np.sign(refs["svdY"]["v"][np.argmax(np.abs(refs["svdY"]["v"]), axis=0), np.arange(refs["svdY"]["v"].shape[1])])[:10]
    Lref = b["LambdaSamples"].reshape(MC, p, 10)
    # Lambda_m = G0 + a_n sd z: flipping column l of V flips G0[:, l] only
    G0 = b["G0"]
    Lflip = (Lref - G0[None]) + (G0 * sgn)[None]
    assert relerr_mat(a["LambdaSamples"], Lflip.reshape(MC, p * 10)) < 1e-8


def test_c1_sampler_cpp_boundary_injected(ctx, c1):
    Y, u, d, v = c1["Y"], c1["u"], c1["d"], c1["v"]
    MC, k = 64, 10
    gam, z = synth.injected_draws(MC, Y.shape[1], k, 0.5 * (1.0 + Y.shape[0]), seed=778)
    ref = orc.sampler(Y, 1.0, 1.0, MC, u, v, d, k, 1.26, gam, z)
    got = fb.CPPFABLESampler(Y, 1.0, 1.0, MC, u, v, d, k, 1.26, ctx=ctx, gam=gam, z=z, psi_moments=True)
    assert relerr_mat(got["LambdaSamples"], ref["LambdaSamples"]) < TOL
    assert relerr_vec(got["SigmaSqSamples"], ref["SigmaSqSamples"]) < TOL
    s1, s2 = c_oracle.psi_moments(ref["LambdaSamples"], ref["SigmaSqSamples"], k)
    assert relerr_mat(got["PsiSum"], s1) < TOL and relerr_mat(got["PsiSumSq"], s2) < TOL
    assert np.array_equal(got["PsiSum"], got["PsiSum"].T)


# ---- native RNG: GPU stream == CPU statement of fable-philox-v1; moments; invariances --------------
def test_native_stream_matches_cpu_specification(ctx, c1):
    Y, u, d, v = c1["Y"], c1["u"], c1["d"], c1["v"]
    MC, k, p, n = 40, 10, Y.shape[1], Y.shape[0]
    ctx.set_seed(2024)
    got = fb.CPPFABLESampler(Y, 1.0, 1.0, MC, u, v, d, k, 1.26, ctx=ctx, want_G=False)
    gam, z = c_oracle.rng_fill(2024, 0, MC, p, k, 0.5 * (1.0 + n))
    ref = orc.sampler(Y, 1.0, 1.0, MC, u, v, d, k, 1.26, gam, z, want_G=False)
    # device log/sincos/sqrt differ from glibc by <= 2 ulp: far inside 1e-10
    assert relerr_vec(got["SigmaSqSamples"], ref["SigmaSqSamples"]) < TOL
    assert relerr_mat(got["LambdaSamples"], ref["LambdaSamples"]) < TOL
    # a draw is a pure function of (seed, m, j): an offset call reproduces the tail
    got2 = fb.CPPFABLESampler(Y, 1.0, 1.0, 10, u, v, d, k, 1.26, ctx=ctx, m0=30, want_G=False)
    np.testing.assert_array_equal(got2["SigmaSqSamples"], got["SigmaSqSamples"][30:])
    np.testing.assert_array_equal(got2["LambdaSamples"], got["LambdaSamples"][30:])
    ctx.set_seed(12345)


def test_native_moments_within_monte_carlo_error(ctx):
    n, p, k, MC = 100, 64, 3, 20000
    Y, u, d, v = _svd_inputs(n, p, k, 21)
    out = fb.CPPFABLESampler(Y, 1.0, 1.0, MC, u, v, d, k, 1.5, ctx=ctx, psi_moments=True)
    st = orc._shrunk_stats(Y, u, v, d, k, 1.0, 1.0)
    a_n = math.sqrt(1.5) / math.sqrt(st["w"])
    a = 0.5 * st["gamma_n"]
    Es = st["gnd"] / (st["gamma_n"] - 2)
    sd_s = np.sqrt((0.5 * st["gnd"]) ** 2 / ((a - 1) ** 2 * (a - 2)) / MC)
    assert np.all(np.abs(out["SigmaSqSamples"].mean(axis=0) - Es) < 5 * sd_s)
    L = out["LambdaSamples"].reshape(MC, p, k)
    assert np.all(np.abs(L.mean(axis=0) - st["G0"]) < 5 * np.sqrt(a_n ** 2 * Es / MC)[:, None])
    # E[Psi_uv] = G_uv (u != v); E[Psi_uu] = G_uu + (1 + k a_n^2) E[sigma2_u]   
    G = st["G0"] @ st["G0"].T
    want = G + np.diag((1 + k * a_n ** 2) * Es)
    mean = out["PsiSum"] / MC
    var = out["PsiSumSq"] / MC - mean ** 2
    assert np.all(np.abs(mean - want) < 6 * np.sqrt(var / MC) + 1e-12)


# ---- resident posterior object: batches, slots, accumulators ---------------------------------------
def test_posterior_batches_accumulate_like_one_pass(ctx, c1):
    import torch
    Y, u, d, v = c1["Y"], c1["u"], c1["d"], c1["v"]
    k, p = 10, Y.shape[1]
    post = fb.Posterior(Y, 1.0, 1.0, u, v, d, k, 1.26, ctx=ctx)
    st = post.get()
    ref = orc._shrunk_stats(Y, u, v, d, k, 1.0, 1.0)
    assert relerr_mat(st["G0"], ref["G0"]) < TOL and relerr_vec(st["gnd"], ref["gnd"]) < TOL
    assert abs(st["tausq"] / ref["tausq"] - 1) < TOL
    MC = 24
    slots = 16
    psi = torch.empty((slots, p, p), dtype=torch.float64, device="cuda")
    with pytest.raises(fb.FableError):                 # one slot per draw of the batch
        post.draw_batch(0, 17, dev_psi=psi.data_ptr(), psi_slots=slots, accumulate=False)
    for m0, B in ((0, 5), (5, 16), (21, 3)):          # ragged batches
        post.draw_batch(m0, B, dev_psi=psi.data_ptr(), psi_slots=slots, accumulate=True)
    ctx.synchronize()
    s1, s2 = post.accum_fetch()
    L, S = post.samples(0, MC)
    r1, r2 = c_oracle.psi_moments(L, S, k)
    assert relerr_mat(s1, r1) < TOL and relerr_mat(s2, r2) < TOL
    # the last batch (m0=21, B=3) left draws 21, 22, 23 in slots 0, 1, 2
    for b in range(3):
        want = c_oracle.psi_draw(L, S, k, 21 + b)
        got = psi[b].cpu().numpy().T       # column-major p x p on the device
        assert relerr_mat(got, want) < TOL
        assert np.array_equal(got, got.T)
    post.accum_reset()
    post.draw_batch(0, MC, accumulate=True)
    t1, t2 = post.accum_fetch()
    assert relerr_mat(t1, s1) < 1e-13 and relerr_mat(t2, s2) < 1e-13
    # the accumulators are stored as compact upper tiles (the payload of the multi-GPU reduction, reduced in place):
    # tile c of the enumeration holds block (I, J) column-major; doubling the device buffer doubles the fetched sums
    a_, b_, cnt = post.accum_dev()
    nT = -(-p // 64)
    assert cnt == nT * (nT + 1) // 2 * 4096 == fb.api.tile_count(p) * 4096
    from fable_b200 import parallel
    ctx.synchronize()
    tiles1, tiles2 = post.accum_fetch_tiles()
    pad = np.zeros((nT * 64, nT * 64)); pad[:p, :p] = t1
    for I, J in ((0, 0), (0, nT - 1), (nT - 1, nT - 1), (3, 7)):
        slot, c = fb.api.tile_locate(p, I, J)
        np.testing.assert_array_equal(tiles1[c].T, pad[64 * I:64 * I + 64, 64 * J:64 * J + 64])
    parallel.as_torch(a_, cnt, "cuda").mul_(2.0); torch.cuda.synchronize()
    u1, u2 = post.accum_fetch()
    np.testing.assert_array_equal(u1, 2.0 * t1); np.testing.assert_array_equal(u2, t2)
    post.close()


@pytest.mark.parametrize("ring", [0, 8])
def test_sampler_tail_runs_chunk_major_with_the_same_bits(ctx, ring, monkeypatch):
    """fable_sampler processes the batches of its last ~256 draws super-block group by group so that finished accumulator
    tiles cross PCIe while the rest computes (dense blocks placed by strided copies).  Every entry still sums its draws
    in the same order: PsiSum / PsiSumSq must be bit-identical with the tail off, partly on and fully on, with and without
    the materialising ring, into pageable and into pinned outputs -- and equal the oracle."""
    import torch
    n, p, k, MC = 50, 2500, 3, 70           # 40 x 40 tiles = 3 super-block columns (the last one narrow), 3 batches of 32
    Y, u, d, v = _svd_inputs(n, p, k, 17)
    ctx.set_psi_ring(ring)
    outs = {}
    try:
        for tail in ("0", "32", "256"):
            monkeypatch.setenv("FABLE_B200_TAIL_DRAWS", tail)
            outs[tail] = fb.CPPFABLESampler(Y, 1.0, 1.0, MC, u, v, d, k, 1.3, ctx=ctx, m0=3, want_G=False, psi_moments=True)
        pinned = {"PsiSum": torch.empty((p, p), dtype=torch.float64).pin_memory().numpy().T,
                  "PsiSumSq": torch.empty((p, p), dtype=torch.float64).pin_memory().numpy().T}
        outs["pinned"] = fb.CPPFABLESampler(Y, 1.0, 1.0, MC, u, v, d, k, 1.3, ctx=ctx, m0=3, want_G=False, psi_moments=True, out=pinned)
    finally:
        ctx.set_psi_ring(0)
        monkeypatch.delenv("FABLE_B200_TAIL_DRAWS")
    for key in ("PsiSum", "PsiSumSq", "LambdaSamples", "SigmaSqSamples"):
        for other in ("32", "256", "pinned"):
            np.testing.assert_array_equal(outs["0"][key], outs[other][key], err_msg=f"{key} tail={other}")
    r1, r2 = c_oracle.psi_moments(outs["0"]["LambdaSamples"], outs["0"]["SigmaSqSamples"], k)
    assert relerr_mat(outs["256"]["PsiSum"], r1) < TOL and relerr_mat(outs["256"]["PsiSumSq"], r2) < TOL
    assert np.array_equal(outs["256"]["PsiSum"], outs["256"]["PsiSum"].T)


@pytest.mark.parametrize("n,p,k,B", [(120, 640, 20, 3), (90, 333, 17, 2), (50, 64, 1, 5), (60, 16, 2, 2), (60, 80, 3, 3),
                                     (60, 96, 5, 2), (70, 144, 12, 2), (70, 1104, 10, 2), (70, 150, 4, 2)])
def test_c4_like_ranks_above_one_chunk(ctx, n, p, k, B):
    """BASELINE configs[3] in miniature (k = 20 > the 16-row rank chunk; p on / off the TMA paths; p a multiple of
    16 but not of 32 / 64, where the half-tile kernel has clipped boxes and empty second halves):
    materialised draws and accumulators against the oracle rebuilt from the returned samples."""
    import torch
    Y, u, d, v = _svd_inputs(n, p, k, 51)
    post = fb.Posterior(Y, 1.0, 1.0, u, v, d, k, 2.0, ctx=ctx)
    psi = torch.zeros((B, p, p), dtype=torch.float64, device="cuda")
    post.draw_batch(7, B, dev_psi=psi.data_ptr(), psi_slots=B, accumulate=True)
    s1, s2 = post.accum_fetch()
    L, S = post.samples(7, B)
    r1, r2 = c_oracle.psi_moments(L, S, k)
    assert relerr_mat(s1, r1) < TOL and relerr_mat(s2, r2) < TOL
    for b in range(B):
        assert relerr_mat(psi[b].cpu().numpy().T, c_oracle.psi_draw(L, S, k, b)) < TOL
    post.close()


@pytest.mark.parametrize("n,p,k,seed", [(60, 150, 3, 5), (80, 64, 2, 6), (120, 400, 6, 7), (50, 257, 1, 8)])
def test_prototype_rank_rule_and_direct_sampler_f3(ctx, n, p, k, seed, monkeypatch):
    """The pure-R prototype path as a callable (R/FABLE_code.R:32-186, 212-298): its own criterion (sigma^2 = resid / (n - k),
    everything scaled by 1 / (n p)) and rank rule at the boundary, and CCFABLE_DirectSampler(Y, ...) end to end -- svd, kMax at
    0.95, that rank rule, entry-wise corrected draws in the (MC, p, p) array -- against the numpy restatement, with
    injected draws.  (The restatement follows the cited R lines; R itself cannot run here: parity unpinned for this row.)"""
    Y, u, d, v = _svd_inputs(n, p, k, seed)
    svdmod = {"u": u, "d": d, "v": v}
    kMax = orc.kmax_rule(d, 0.95)
    for kk in (1, 2, 5, kMax):
        assert abs(fb.RankEstimatorCriterion(kk, Y, svdmod, ctx=ctx) / orc.criterion_R(kk, Y, u, v, d) - 1) < TOL
    want_k = orc.rank_estimator_R(Y, u, v, d, kMax)
    for fast in ("1", "0"):
        monkeypatch.setenv("FABLE_B200_RANK_FAST", fast)
        assert fb.RankEstimator(Y, svdmod, kMax, ctx=ctx) == want_k
    monkeypatch.delenv("FABLE_B200_RANK_FAST")
    MC = 5
    gam = np.random.default_rng(seed).gamma(0.5 * (1.0 + n), 1.0, size=(MC, p))
    zfn = lambda MC_, k_, p_: np.random.default_rng(seed + 1).standard_normal((MC_, k_, p_))
    # for fixed injected normals the draws depend on the signs of the singular vectors: put the oracle's SVD on
    # the library's sign convention (largest |entry| of every right vector positive)
    sgn = np.sign(v[np.argmax(np.abs(v), axis=0), np.arange(v.shape[1])])
    k_ref, want = orc.ccfable_direct_sampler(Y, 1.0, 1.0, MC, gam, zfn, svd=(u * sgn, d, v * sgn))
    info = {}
    got = fb.CCFABLE_DirectSampler(Y, 1.0, 1.0, MC, ctx=ctx, gam=gam, z=zfn, info=info)
    assert info["k"] == k_ref == want_k and info["kMax"] == kMax
    assert got.shape == (MC, p, p) and got.flags.f_contiguous
    assert relerr_mat(got, want) < TOL
    assert np.array_equal(got, got.transpose(0, 2, 1))
    nat = fb.CCFABLE_DirectSampler(Y, 1.0, 1.0, 3, ctx=ctx)                     # native stream: runs, symmetric, finite
    assert nat.shape == (3, p, p) and np.all(np.isfinite(nat)) and np.array_equal(nat, nat.transpose(0, 2, 1))


@pytest.mark.parametrize("n,p,k", [(60, 160, 3), (40, 77, 2), (50, 130, 17)])
def test_entrywise_corrected_draws_f3(ctx, n, p, k):
    """CCFABLE_DirectSampler (R/FABLE_code.R:212-298): G + B o (Lambda Lambda' - G) + diag(sigma2), B on the fly."""
    import torch
    Y, u, d, v = _svd_inputs(n, p, k, 71)
    MC = 4
    gam, z = synth.injected_draws(MC, p, k, 0.5 * (1.0 + n), seed=n + k)
    want = orc.direct_sampler_psi(Y, 1.0, 1.0, u, v, d, k, gam, z)
    post = fb.Posterior(Y, 1.0, 1.0, u, v, d, k, 1.0, ctx=ctx)        # a_n = 1/sqrt(n + 1/tau^2): varInflation 1
    post.set_entrywise_correction(True)
    psi = torch.zeros((MC, p, p), dtype=torch.float64, device="cuda")
    dg = torch.from_numpy(np.ascontiguousarray(gam)).cuda(); dz = torch.from_numpy(np.ascontiguousarray(z)).cuda()
    post.draw_batch(0, MC, dev_psi=psi.data_ptr(), psi_slots=MC, accumulate=True, dev_gam=dg.data_ptr(), dev_z=dz.data_ptr())
    s1, s2 = post.accum_fetch()
    got = psi.cpu().numpy().transpose(0, 2, 1)
    assert relerr_mat(got, want) < TOL
    assert relerr_mat(s1, want.sum(axis=0)) < TOL and relerr_mat(s2, (want ** 2).sum(axis=0)) < TOL
    # switching the correction off restores the plain draws
    post.set_entrywise_correction(False)
    post.draw_batch(0, MC, dev_psi=psi.data_ptr(), psi_slots=MC, accumulate=False, dev_gam=dg.data_ptr(), dev_z=dz.data_ptr())
    ctx.synchronize()
    o = orc.sampler(Y, 1.0, 1.0, MC, u, v, d, k, 1.0, gam, z)
    assert relerr_mat(psi[1].cpu().numpy().T, orc.psi_draw(o["LambdaSamples"][1], o["SigmaSqSamples"][1], k)) < TOL
    post.close()


@pytest.mark.parametrize("p,world", [(1600, 2), (1600, 3), (1000, 2), (333, 2), (2100, 8)])
def test_tile_partition_needs_no_collective(ctx, p, world):
    """Partition B (BASELINE configs[3]): 'ranks' owning disjoint slices of the tile enumeration and processing the same
    draws hold disjoint shares of every Psi_m and of the accumulators, each stored compactly (only the rank's tiles);
    expanded to dense with zeros elsewhere, the shares add up to the one-GPU result bit for bit.  p on and off the
    16-row TMA boxes, odd p, slices that cut super-blocks."""
    import torch
    from fable_b200 import parallel
    n, k, B = 60, 3, 4
    Y, u, d, v = _svd_inputs(n, p, k, 31)
    post = fb.Posterior(Y, 1.0, 1.0, u, v, d, k, 1.4, ctx=ctx)
    assert post.tile_slots() == parallel.tile_slots(p)
    full = torch.zeros((B, p, p), dtype=torch.float64, device="cuda")
    post.draw_batch(0, B, dev_psi=full.data_ptr(), psi_slots=B, accumulate=True)
    f1, f2 = post.accum_fetch()
    g1 = np.zeros_like(f1); g2 = np.zeros_like(f2)
    parts = torch.zeros((B, p, p), dtype=torch.float64, device="cuda")
    dense = torch.empty((p, p), dtype=torch.float64, device="cuda")
    owned = 0
    for rank in range(world):
        lo, hi = parallel.tile_range(p, rank, world)
        post.set_tile_range(lo, hi)
        nt = fb.api.tile_count(p, lo, hi)
        owned += nt
        ring = torch.full((B, 2 * nt, 64, 64), float("nan"), dtype=torch.float64, device="cuda")   # the rank's share only
        post.draw_batch_tiles(0, B, dev_psi_tiles=ring.data_ptr(), psi_slots=B, accumulate=True)
        _, _, cnt = post.accum_dev()
        assert cnt == nt * 4096
        a1, a2 = post.accum_fetch()
        assert not (np.any(g1 * a1) or np.any(g2 * a2))        # disjoint supports
        g1 += a1; g2 += a2
        for b in range(B):
            post.tiles_to_dense(ring[b].data_ptr(), True, 0, p, dense.data_ptr())
            ctx.synchronize()
            assert not torch.any(parts[b] * dense)
            parts[b] += dense
    nT = -(-p // 64)
    assert owned == nT * (nT + 1) // 2
    post.set_tile_range(0, 0)
    assert torch.equal(full, parts)
    np.testing.assert_array_equal(g1, f1); np.testing.assert_array_equal(g2, f2)
    L, S = post.samples(0, B)
    r1, r2 = c_oracle.psi_moments(L, S, k)
    assert relerr_mat(f1, r1) < TOL and relerr_mat(f2, r2) < TOL
    assert relerr_mat(full[2].cpu().numpy().T, c_oracle.psi_draw(L, S, k, 2)) < TOL
    post.close()


@pytest.mark.parametrize("M,K,akf", [(1, 1, 0), (130, 77, 0), (130, 77, 1), (300, 4000, 0), (257, 1000, 1),
                                     # K-slow with an even leading dimension -> the TMA-staged kernel: one box, K shorter than a
                                     # k-tile, fewer k-tiles than pipeline stages, M off the 16-row boxes, split-K, many tiles;
                                     # odd leading dimension -> the register-staged kernel
                                     (2, 5, 0), (16, 3, 0), (128, 16, 0), (128, 48, 0), (1000, 5000, 0), (129, 64, 0),
                                     (2064, 1100, 0)])
def test_syrk_gram(ctx, M, K, akf):
    import torch
    rng = np.random.default_rng(M + K)
    A = rng.standard_normal((M, K))
    dA = torch.from_numpy(A.copy() if akf else A.T.copy()).cuda()
    dC = torch.empty((M, M), dtype=torch.float64, device="cuda")
    fb.api.dev_syrk(ctx, akf, M, K, 2.0, dA.data_ptr(), K if akf else M, dC.data_ptr())
    ctx.synchronize()
    got = dC.cpu().numpy()
    assert relerr_mat(got, 2.0 * A @ A.T) < 1e-13
    assert np.array_equal(got, got.T)


@pytest.mark.parametrize("n,p,k", [
    (1000, 2003, 10),     # C2/C3 geometry: 32 items, all 16 warps on a tile with two items each, ragged last tile
    (500, 1000, 10),      # C1: 16 items, one per warp
    (200, 777, 3),        # 7 items (last one 8 rows): 7 warps per tile, two tiles at a time, two warps without a group
    (64, 100, 1), (2, 5, 1), (66, 9, 4),   # 2, 1 and 3 items: 8, 16 and 5 tiles at a time
    (320, 4100, 7),       # 10 items, two k-steps
    (500, 300, 24),       # six k-steps, one item per warp
    (2000, 300, 4),       # 63 items: up to four per warp
    (3000, 100, 4),       # 94 items: up to six per warp
    (1100, 1200, 8),      # 35 items, two k-steps, three per warp
    (130, 50, 32), (2000, 300, 12), (1000, 1200, 17), (4000, 150, 3),    # fragments beyond the register budget -> DFMA kernel
    (501, 300, 10),       # odd n: rows are not 16-byte aligned -> DFMA kernel
    (5000, 200, 10),      # DFMA kernel, U_k staged in two row chunks
])
def test_dev_rowstats_both_kernels(ctx, n, p, k, monkeypatch):
    """fable_dev_rowstats (cpp:371-379 / :161-166): the tensor-core pass and the DFMA pass against numpy."""
    import torch
    rng = np.random.default_rng(n * 7 + p + k)
    Y = np.asfortranarray(rng.standard_normal((n, p)))
    U = np.asfortranarray(np.linalg.qr(rng.standard_normal((n, k)))[0]) if n >= k else np.asfortranarray(rng.standard_normal((n, k)))
    c = np.asfortranarray(Y.T @ U + 0.1 * rng.standard_normal((p, k)))
    want_r = np.sum((Y - U @ c.T) ** 2, axis=0)
    want_s = np.sum(Y ** 2, axis=0)
    dY = torch.from_numpy(np.ascontiguousarray(Y.T)).cuda(); dU = torch.from_numpy(np.ascontiguousarray(U.T)).cuda()
    dc = torch.from_numpy(np.ascontiguousarray(c.T)).cuda()
    for mma, tma in (("1", "0"), ("1", "1"), ("0", "0")):      # cp.async ring, TMA-streamed tiles, DFMA kernel
        monkeypatch.setenv("FABLE_B200_ROWS_MMA", mma)
        monkeypatch.setenv("FABLE_B200_ROWS_TMA", tma)
        ds = torch.full((p,), -1.0, dtype=torch.float64, device="cuda"); dr = torch.full((p,), -1.0, dtype=torch.float64, device="cuda")
        fb.api.dev_rowstats(ctx, dY.data_ptr(), n, p, dU.data_ptr(), dc.data_ptr(), p, k, ds.data_ptr(), dr.data_ptr())
        ctx.synchronize()
        assert relerr_vec(dr.cpu().numpy(), want_r) < 1e-12, (mma, tma)
        assert relerr_vec(ds.cpu().numpy(), want_s) < 1e-12, (mma, tma)
        dr.fill_(-1.0)
        fb.api.dev_rowstats(ctx, dY.data_ptr(), n, p, dU.data_ptr(), dc.data_ptr(), p, k, 0, dr.data_ptr())   # s not wanted
        ctx.synchronize()
        assert relerr_vec(dr.cpu().numpy(), want_r) < 1e-12, (mma, tma)


@pytest.mark.parametrize("p,k", [(1, 1), (16, 1), (48, 4), (63, 1), (64, 2), (65, 3), (80, 2), (96, 7), (112, 5), (130, 13),
                                 (200, 20), (257, 33), (320, 10), (1040, 3)])
def test_rankk_cov_ragged_shapes(ctx, p, k):
    """dense A A' + diag(s) (G = G0 G0', CovPostMean) at tile-edge sizes and ranks above one chunk."""
    import torch
    rng = np.random.default_rng(p * 100 + k)
    A = np.asfortranarray(rng.standard_normal((p, k))); s = rng.uniform(0.5, 2, p)
    dA = torch.from_numpy(A.T.copy()).cuda()          # (k, p) row-major == p x k column-major
    ds = torch.from_numpy(s).cuda()
    out = torch.empty((p, p), dtype=torch.float64, device="cuda")
    fb.api.dev_rankk_cov(ctx, dA.data_ptr(), p, ds.data_ptr(), p, k, out.data_ptr())
    ctx.synchronize()
    want = A @ A.T + np.diag(s)
    assert relerr_mat(out.cpu().numpy().T, want) < 1e-13


def test_whole_tile_kernel_forced_by_env():
    """FABLE_B200_PSI_TV=64 selects the whole-tile (64 x 64, two CTAs per SM) covariance kernel where the half-tile one
    is the default (p % 16 == 0): both must produce the same bits (same DMMA order per entry)."""
    import subprocess, sys
    code = r"""
import sys, numpy as np, torch
sys.path.insert(0, %r)
import fable_b200 as fb
from fable_b200 import synth
n, p, k, B = 60, 208, 6, 3
Y = synth.factor_data(n, p, k, rep=3)
lam, Q = np.linalg.eigh(Y @ Y.T)
U = np.asfortranarray(Q[:, ::-1][:, :k]); T = Y.T @ U; d = np.linalg.norm(T, axis=0); V = np.asfortranarray(T / d)
ctx = fb.Context(0, 99)
post = fb.Posterior(Y, 1.0, 1.0, U, V, d, k, 1.7, ctx=ctx)
psi = torch.zeros((B, p, p), dtype=torch.float64, device="cuda")
post.draw_batch(0, B, dev_psi=psi.data_ptr(), psi_slots=B, accumulate=True)
s1, s2 = post.accum_fetch()
np.savez(sys.argv[1], psi=psi.cpu().numpy(), s1=s1, s2=s2)
""" % ROOT
    import tempfile
    outs = []
    with tempfile.TemporaryDirectory() as td:
        for tv in ("32", "64"):
            path = os.path.join(td, f"tv{tv}.npz")
            env = dict(os.environ, FABLE_B200_PSI_TV=tv)
            subprocess.run([sys.executable, "-c", code, path], check=True, env=env, timeout=300)
            outs.append(dict(np.load(path)))
    for key in ("psi", "s1", "s2"):
        np.testing.assert_array_equal(outs[0][key], outs[1][key])
    assert np.abs(outs[0]["psi"]).max() > 0


# ---- f1: post-processing (mean + exact type-5 sample quantiles) -------------------------------------
@pytest.mark.parametrize("MC", [1, 2, 7, 100, 1000, 1025, 16384, 16385, 20000, 30011])
def test_post_processing_against_oracle(ctx, MC):
    """<= 16384 draws: shared-memory sort; above: radix selection (keys in shared memory up to 25600 draws, in a
    global scratch beyond)."""
    n, p, k = 40, 9, 2
    Y, u, d, v = _svd_inputs(n, p, k, 41)
    gam, z = synth.injected_draws(MC, p, k, 0.5 * (1.0 + n), seed=MC)
    o = orc.sampler(Y, 1.0, 1.0, MC, u, v, d, k, 1.3, gam, z)
    ref = orc.post_processing(o, 0.05)
    got = fb.CPPCCFABLEPostProcessing(o, 0.05, ctx=ctx)
    for key in ("PostMeanMatrix", "LowerQuantileMatrix", "UpperQuantileMatrix"):
        assert relerr_mat(got[key], ref[key]) < TOL, key
        assert np.array_equal(got[key], got[key].T)
    sel = [3, 1, 9, 3]                     # 1-based, unsorted, with a repeated variable (index equality, cpp:770)
    refs = orc.post_processing_submatrix(o, 0.1, sel)
    gots = fb.CCFABLEPostProcessingSubmatrix(o, 0.1, sel, ctx=ctx)
    for key in refs:
        assert relerr_mat(gots[key], refs[key]) < TOL, key
    with pytest.raises(fb.FableError):
        fb.CCFABLEPostProcessingSubmatrix(o, 0.1, [0, 2], ctx=ctx)
    with pytest.raises(fb.FableError):
        fb.CCFABLEPostProcessingSubmatrix(o, 0.1, [2, p + 1], ctx=ctx)


@pytest.mark.parametrize("MC,alpha", [(100, 0.01), (20, 0.05), (10, 0.1), (16, 0.0625), (1024, 1.0 / 1024), (1000, 0.001),
                                      (32768, 1.0 / 32768)])
def test_post_processing_upper_level_at_p_max(ctx, MC, alpha):
    """alpha * MC == 1: the upper level 1 - alpha/2 equals P_max = (N - 0.5)/N exactly, h = N, weight 0 on a neighbour
    that does not exist -- the result is the maximum (never the sort's +inf padding, never a read past the buffer
    when MC is a power of two); the lower level is the minimum."""
    n, p, k = 40, 6, 2
    Y, u, d, v = _svd_inputs(n, p, k, 43)
    gam, z = synth.injected_draws(MC, p, k, 0.5 * (1.0 + n), seed=MC + 1)
    o = orc.sampler(Y, 1.0, 1.0, MC, u, v, d, k, 1.3, gam, z)
    ref = orc.post_processing(o, alpha)
    got = fb.CPPCCFABLEPostProcessing(o, alpha, ctx=ctx)
    for key in ("PostMeanMatrix", "LowerQuantileMatrix", "UpperQuantileMatrix"):
        assert np.all(np.isfinite(got[key])), key
        assert relerr_mat(got[key], ref[key]) < TOL, key
    L = o["LambdaSamples"].reshape(MC, p, k)
    v01 = np.einsum("ml,ml->m", L[:, 0, :], L[:, 1, :])
    assert got["UpperQuantileMatrix"][0, 1] == pytest.approx(v01.max(), rel=1e-12)
    assert got["LowerQuantileMatrix"][0, 1] == pytest.approx(v01.min(), rel=1e-12)


@pytest.mark.parametrize("MC", [1, 3, 16, 17, 100, 1000, 5000])
def test_post_processing_both_kernels_agree(ctx, MC, monkeypatch):
    """The tile-resident selection (post_tiles.cu) and the per-pair sort / selection kernels (post.cu) are two exact
    algorithms for the same order statistics: identical quantiles, bit for bit (means differ in summation order)."""
    n, p, k = 30, 70, 5                     # two tile columns, the second one ragged
    Y, u, d, v = _svd_inputs(n, p, k, 47)
    gam, z = synth.injected_draws(MC, p, k, 0.5 * (1.0 + n), seed=MC + 9)
    o = orc.sampler(Y, 1.0, 1.0, MC, u, v, d, k, 1.3, gam, z)
    res = {}
    for tiles in ("1", "0"):
        monkeypatch.setenv("FABLE_B200_POST_TILES", tiles)
        res[tiles] = fb.CPPCCFABLEPostProcessing(o, 0.1, ctx=ctx)
    monkeypatch.delenv("FABLE_B200_POST_TILES")
    for key in ("LowerQuantileMatrix", "UpperQuantileMatrix"):
        np.testing.assert_array_equal(res["1"][key], res["0"][key], err_msg=key)
    assert relerr_mat(res["1"]["PostMeanMatrix"], res["0"]["PostMeanMatrix"]) < 1e-13
    if MC <= 1000:
        ref = orc.post_processing(o, 0.1)
        for key in ref:
            assert relerr_mat(res["1"][key], ref[key]) < TOL, key


@pytest.mark.parametrize("n,p,k,MC,alpha", [(40, 150, 3, 400, 0.05), (30, 64, 12, 33, 0.2), (50, 200, 20, 300, 0.05), (40, 96, 1, 1000, 0.01)])
def test_post_processing_on_resident_draws(ctx, n, p, k, MC, alpha):
    """fable_posterior_post_processing: the draws are generated and consumed on the device (no MC x k p matrix crosses PCIe).
    Injected draws: against the oracle's sampler + post-processing.  Native stream: against the stored-sample entry point fed
    with the same draws (bit for bit), on the full matrix and on a selection with a repeated variable."""
    Y, u, d, v = _svd_inputs(n, p, k, 53)
    post = fb.Posterior(Y, 1.0, 1.0, u, v, d, k, 1.2, ctx=ctx)
    gam, z = synth.injected_draws(MC, p, k, 0.5 * (1.0 + n), seed=MC + p)
    got = post.post_processing(MC, alpha, gam=gam, z=z)
    o = orc.sampler(Y, 1.0, 1.0, MC, u, v, d, k, 1.2, gam, z)
    ref = orc.post_processing(o, alpha)
    for key in ref:
        assert relerr_mat(got[key], ref[key]) < TOL, key
        assert np.array_equal(got[key], got[key].T)
    L, S = post.samples(7, MC)
    stored = {"LambdaSamples": L, "SigmaSqSamples": S, "EstimatedRank": k}
    a = post.post_processing(MC, alpha, m0=7)
    b = fb.CPPCCFABLEPostProcessing(stored, alpha, ctx=ctx)
    sel = [5, 1, p, 5, 17]
    a2 = post.post_processing(MC, alpha, SelectedIndices=sel, m0=7)
    b2 = fb.CCFABLEPostProcessingSubmatrix(stored, alpha, sel, ctx=ctx)
    for key in a:
        np.testing.assert_array_equal(a[key], b[key], err_msg=key)
        np.testing.assert_array_equal(a2[key], b2[key], err_msg=key)
    assert a2["PostMeanMatrix"][0, 3] == a2["PostMeanMatrix"][0, 0]          # positions 0 and 3 are the same variable (cpp:770)
    with pytest.raises(fb.FableError):
        post.post_processing(MC, alpha, SelectedIndices=[0, 3])
    post.close()


def test_post_processing_c2_shape_sampled_entries(ctx):
    """A quarter-size BASELINE configs[1] (p = 2500, MC = 2000, k = 10) through the resident path: sampled entries against
    numpy's order statistics (Hyndman-Fan type 5 = numpy 'hazen') on values rebuilt from the returned samples."""
    n, p, k, MC, alpha = 300, 2500, 10, 2000, 0.05
    Y, u, d, v = _svd_inputs(n, p, k, 59)
    post = fb.Posterior(Y, 1.0, 1.0, u, v, d, k, 1.26, ctx=ctx)
    got = post.post_processing(MC, alpha)
    L, S = post.samples(0, MC)
    Lr = L.reshape(MC, p, k)
    rng = np.random.default_rng(1)
    pairs = [(0, 0), (0, 1), (63, 64), (p - 1, p - 1), (17, p - 1), (1024, 1023)] + [tuple(rng.integers(0, p, 2)) for _ in range(200)]
    for a_, b_ in pairs:
        vals = np.einsum("ml,ml->m", Lr[:, a_, :], Lr[:, b_, :]) + (S[:, a_] if a_ == b_ else 0.0)
        assert abs(got["PostMeanMatrix"][a_, b_] - vals.mean()) <= 1e-12 * np.abs(vals).max()
        for key, pr in (("LowerQuantileMatrix", alpha / 2), ("UpperQuantileMatrix", 1 - alpha / 2)):
            want = orc.arma_quantile(vals, pr)
            assert abs(got[key][a_, b_] - want) <= 1e-11 * max(abs(want), np.abs(vals).max() * 1e-3), (key, a_, b_)
    post.close()


@pytest.mark.parametrize("MC", [1000, 17000, 26000])
def test_post_processing_ties_and_extreme_levels(ctx, MC):
    """Draws that take few distinct values (every order statistic sits inside a run of duplicates, some entries are
    constant, signed zeros occur) and levels outside [0.5/N, (N-0.5)/N] (Armadillo returns min / max there):
    the sort path and both selection paths against the oracle."""
    rng = np.random.default_rng(MC)
    p, k = 5, 3
    o = {"LambdaSamples": np.asfortranarray(rng.integers(-2, 3, size=(MC, p * k)).astype(np.float64)),
         "SigmaSqSamples": np.asfortranarray(rng.integers(1, 3, size=(MC, p)).astype(np.float64)), "EstimatedRank": k}
    o["LambdaSamples"][:, :k] = 0.0                       # variable 1: every cross product is +-0, its variance is sigma^2 only
    o["LambdaSamples"][::2, 0] = -0.0
    for alpha in (0.05, 0.5, 1e-7, 1.0):
        ref = orc.post_processing(o, alpha)
        got = fb.CPPCCFABLEPostProcessing(o, alpha, ctx=ctx)
        for key in ("PostMeanMatrix", "LowerQuantileMatrix", "UpperQuantileMatrix"):
            np.testing.assert_allclose(got[key], ref[key], rtol=1e-12, atol=1e-12, err_msg=f"{key} alpha={alpha}")


def test_coverage_pattern_of_the_replication_script(ctx):
    """extras/replicationCodes/coverage_FABLE.R:95-130 in miniature: entry-wise 95% intervals from the GPU
    sampler + post-processing cover the true covariance on a random sub-matrix at roughly the nominal rate."""
    n, p, k, MC = 400, 300, 3, 600
    Lambda0, Sigma0 = synth.truth(p, k, seed=1)
    Psi0 = Lambda0 @ Lambda0.T + np.diag(Sigma0)
    Y = synth.factor_data(n, p, k, rep=0)
    out = fb.FABLEPosteriorSampler(Y, MC=MC, ctx=ctx)
    assert out["estRank"] == k
    sel = np.sort(np.random.default_rng(3).choice(p, 40, replace=False)) + 1
    pp_ = fb.CCFABLEPostProcessingSubmatrix(out["CCFABLESamples"], 0.05, sel, ctx=ctx)
    truth = Psi0[np.ix_(sel - 1, sel - 1)]
    cover = np.mean((pp_["LowerQuantileMatrix"] <= truth) & (truth <= pp_["UpperQuantileMatrix"]))
    assert 0.85 <= cover <= 1.0, cover


@pytest.mark.parametrize("n,p,k", [(9, 5, 2), (50, 300, 7), (64, 200, 40)])
def test_log_likelihood_eval(ctx, n, p, k):
    rng = np.random.default_rng(n * p + k)
    Y = rng.standard_normal((n, p)); M = rng.standard_normal((n, k)); Lam = rng.standard_normal((p, k)); sig = rng.uniform(0.5, 2, p)
    got = fb.CPPLogLikelihoodEval(Y, M, Lam, sig, ctx=ctx)
    assert abs(got / orc.log_likelihood_eval(Y, M, Lam, sig) - 1) < TOL
    from oracle import ref_oracle as ref
    if ref.available():
        assert abs(got / ref.log_likelihood_eval(Y, M, Lam, sig) - 1) < TOL


# ---- DMMA GEMM + eigensolver primitives -------------------------------------------------------------
@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (127, 129, 17), (128, 128, 16), (300, 200, 1000), (64, 1000, 33)])
@pytest.mark.parametrize("akf,bkf", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_dmma_gemm(ctx, M, N, K, akf, bkf):
    import torch
    rng = np.random.default_rng(M + N + K)
    A = rng.standard_normal((M, K)); B = rng.standard_normal((N, K))
    # K-slow = column-major M x K (torch (K, M) row-major); K-fast = torch (M, K) row-major
    dA = torch.from_numpy(A.copy() if akf else A.T.copy()).cuda()
    dB = torch.from_numpy(B.copy() if bkf else B.T.copy()).cuda()
    dC = torch.empty((N, M), dtype=torch.float64, device="cuda")
    fb.api.dev_gemm(ctx, akf, bkf, M, N, K, 0.5, dA.data_ptr(), K if akf else M, dB.data_ptr(), K if bkf else N,
                    dC.data_ptr(), M)
    ctx.synchronize()
    assert relerr_mat(dC.cpu().numpy().T, 0.5 * A @ B.T) < 1e-13


@pytest.mark.parametrize("m", [1, 2, 7, 16, 17, 33, 47, 64, 100, 256, 257, 301, 640, 960, 961, 1000])
def test_jacobi_eigensolver(ctx, m):
    import torch
    rng = np.random.default_rng(m)
    X = rng.standard_normal((m, m + 5))
    G = X @ X.T
    dG = torch.from_numpy(G.copy()).cuda()
    dW = torch.empty(m, dtype=torch.float64, device="cuda"); dQ = torch.empty((m, m), dtype=torch.float64, device="cuda")
    sweeps = fb.api.dev_syevj(ctx, dG.data_ptr(), m, dW.data_ptr(), dQ.data_ptr())
    w = dW.cpu().numpy(); Q = dQ.cpu().numpy().T
    assert sweeps <= 30
    assert relerr_vec(w, np.linalg.eigvalsh(G)[::-1]) < 1e-11
    assert relerr_mat(Q @ np.diag(w) @ Q.T, G) < 1e-11
    assert relerr_mat(Q.T @ Q, np.eye(m)) < 1e-11


@pytest.mark.parametrize("m", [96, 301, 1100])
@pytest.mark.parametrize("env", [{"FABLE_B200_BJ_FUSED": "0"}, {"FABLE_B200_BJ_FUSED": "2"}, {"FABLE_B200_BJ_REPLAY": "0"},
                                 {"FABLE_B200_BJ_FUSED": "0", "FABLE_B200_BJ_ISO": "2"}, {"FABLE_B200_BJ_GRAPH": "0"}])
def test_jacobi_eigensolver_every_round_form(ctx, m, env, monkeypatch):
    """The round of the block Jacobi exists in several forms (one cluster kernel per round where all pairs' clusters are
    resident -- also forced into a second wave here --, four kernels with recorded rotations, the legacy in-kernel R, the
    full-matrix inner steps, plain launches instead of the graph): all of them give the decomposition."""
    import torch
    for key, val in env.items():
        monkeypatch.setenv(key, val)
    rng = np.random.default_rng(m)
    X = rng.standard_normal((m, m // 2 + 3))              # rank-deficient: the deflation runs behind every form
    G = X @ X.T
    dG = torch.from_numpy(G.copy()).cuda()
    dW = torch.empty(m, dtype=torch.float64, device="cuda"); dQ = torch.empty((m, m), dtype=torch.float64, device="cuda")
    sweeps = fb.api.dev_syevj(ctx, dG.data_ptr(), m, dW.data_ptr(), dQ.data_ptr())
    w = dW.cpu().numpy(); Q = dQ.cpu().numpy().T
    lam = np.linalg.eigvalsh(G)[::-1]
    assert sweeps <= 30
    assert np.max(np.abs(w - lam)) < 1e-11 * lam[0]
    assert relerr_mat(Q @ np.diag(w) @ Q.T, G) < 1e-11
    assert relerr_mat(Q.T @ Q, np.eye(m)) < 1e-11


@pytest.mark.parametrize("m,spectrum", [(96, "repeated"), (300, "repeated"), (300, "clustered"), (257, "singular"), (64, "singular"),
                                        (12, "singular"), (500, "rank1")])
def test_jacobi_eigensolver_hard_spectra(ctx, m, spectrum):
    """Repeated and clustered eigenvalues (any orthonormal basis of an eigenspace is right: check the decomposition, not
    the vectors) and exactly singular matrices (zero eigenvalues deflated to 0, vectors completed to an orthonormal
    basis, convergence in the usual number of sweeps)."""
    import torch
    rng = np.random.default_rng(m)
    Q0 = np.linalg.qr(rng.standard_normal((m, m)))[0]
    if spectrum == "repeated":
        w0 = np.repeat(np.array([9.0, 4.0, 4.0, 1.0]), m // 4 + 1)[:m]
    elif spectrum == "clustered":
        w0 = 5.0 + 1e-9 * rng.standard_normal(m); w0[:3] = (100.0, 50.0, 50.0 * (1 + 1e-12))
    elif spectrum == "singular":
        w0 = rng.uniform(1.0, 10.0, m); w0[-(m // 4):] = 0.0
    else:
        w0 = np.zeros(m); w0[0] = 3.0
    w0 = np.sort(w0)[::-1]
    G = (Q0 * w0) @ Q0.T
    G = 0.5 * (G + G.T)
    dG = torch.from_numpy(G.copy()).cuda()
    dW = torch.empty(m, dtype=torch.float64, device="cuda"); dQ = torch.empty((m, m), dtype=torch.float64, device="cuda")
    sweeps = fb.api.dev_syevj(ctx, dG.data_ptr(), m, dW.data_ptr(), dQ.data_ptr())
    w = dW.cpu().numpy(); Q = dQ.cpu().numpy().T
    assert sweeps <= 30
    assert np.max(np.abs(w - w0)) < 1e-12 * w0[0]
    assert np.all(w[w0 == 0.0] == 0.0)
    assert relerr_mat(Q @ np.diag(w) @ Q.T, G) < 1e-11
    assert relerr_mat(Q.T @ Q, np.eye(m)) < 1e-11


# ---- degenerate shapes ---------------------------------------------------------------------------------
@pytest.mark.parametrize("n,p,k", [(3, 1, 1), (2, 2, 1), (5, 3, 2), (4, 70, 4), (70, 4, 3), (8, 8, 8)])
def test_tiny_and_degenerate_shapes(ctx, n, p, k):
    """p = 1, n < 4, tall and wide, k equal to the full rank: every entry point against the oracle."""
    rng = np.random.default_rng(n * 100 + p)
    Y = np.asfortranarray(rng.standard_normal((n, p)) + 3.0 * rng.standard_normal((n, 1)) @ rng.standard_normal((1, p)))
    u, d, v = orc.svd_thin(Y)
    r = d.size
    s = fb.svd(Y, ctx=ctx)
    assert relerr_vec(s["d"], d) < 1e-9
    kk = min(k, r)
    if kk < r:                                # at full rank the residual is rounding noise (log of ~1e-30)
        assert abs(fb.CPPCriterionFunction(kk, Y, u, v, d, ctx=ctx) / orc.criterion(kk, Y, u, v, d) - 1) < TOL
        kMax = r - 1
        assert fb.CPPRankEstimator(Y, u, v, d, kMax, ctx=ctx) == orc.rank_estimator(Y, u, v, d, kMax)
        hg, ho = fb.FABLEHyperParameters(Y, u, v, d, kk, ctx=ctx), orc.hyper_parameters(Y, u, v, d, kk)
        assert relerr_mat(hg["G"], ho["G"]) < TOL and relerr_vec(hg["SigmaSqEstimate"], ho["SigmaSqEstimate"]) < TOL
        MC = 3
        gam, z = synth.injected_draws(MC, p, kk, 0.5 * (3.0 + n), seed=n + p)
        sg = fb.CPPFABLESampler(Y, 3.0, 1.0, MC, u, v, d, kk, 1.1, ctx=ctx, gam=gam, z=z, psi_moments=True)
        so = orc.sampler(Y, 3.0, 1.0, MC, u, v, d, kk, 1.1, gam, z)
        assert relerr_mat(sg["LambdaSamples"], so["LambdaSamples"]) < TOL and relerr_vec(sg["SigmaSqSamples"], so["SigmaSqSamples"]) < TOL
        s1, s2 = orc.psi_moments(so["LambdaSamples"], so["SigmaSqSamples"], kk)
        assert relerr_mat(sg["PsiSum"], s1) < TOL and relerr_mat(sg["PsiSumSq"], s2) < TOL


# ---- error behaviour (reference: Armadillo bounds errors surface as R errors, §8b) ------------------
def test_errors_are_status_codes_not_crashes(ctx, golden):
    Y, u, d, v = golden["a_Y"], golden["a_u"], golden["a_d"], golden["a_v"]
    with pytest.raises(fb.FableError) as ei:
        fb.CPPCriterionFunction(0, Y, u, v, d, ctx=ctx)                       # cols(0, -1)
    assert ei.value.code == 1
    with pytest.raises(fb.FableError):
        fb.CPPCriterionFunction(d.size + 1, Y, u, v, d, ctx=ctx)
    with pytest.raises(fb.FableError):
        fb.CPPRankEstimator(Y, u, v, d, d.size + 1, ctx=ctx)
    with pytest.raises(fb.FableError):
        fb.CPPFABLESampler(Y, 1.0, 1.0, 0, u, v, d, 2, 1.0, ctx=ctx)
    with pytest.raises(fb.FableError):
        fb.CPPFABLESampler(Y, 1.0, 1.0, 3, u, v, d, 2, 1.0, ctx=ctx, gam=np.ones((3, Y.shape[1])))
    # the context stays usable after an error
    assert fb.CPPRankEstimator(Y, u, v, d, int(golden["a_kMax50"]), ctx=ctx) == int(golden["a_kEst50"])


# ---- full-size properties (BASELINE.json configs[1]: n=1000, p=5000, k=10) ---------------------------
def test_c2_full_size_properties(ctx):
    import torch
    n, p, k = 1000, 5000, 10
    Y = synth.factor_data(n, p, k, rep=0)
    s = fb.svd(Y, ctx=ctx)
    # Frobenius identity: sum d^2 == ||Y||_F^2; orthogonality of the leading block
    assert abs(np.sum(s["d"] ** 2) / np.sum(Y * Y) - 1) < 1e-12
    assert relerr_mat(s["u"][:, :50].T @ s["u"][:, :50], np.eye(50)) < 1e-10
    kMax = fb.kmax_rule(s["d"], 0.5)
    assert fb.CPPRankEstimator(Y, s["u"], s["v"], s["d"], kMax, ctx=ctx) == 10
    pm = fb.CPPFABLEPostMean(Y, 1.0, 1.0, s["u"], s["v"], s["d"], kMax, ctx=ctx, internals=True)
    C = pm["CovPostMean"]
    assert np.array_equal(C, C.T)
    # rank-k structure: CovPostMean - diag(SigmaHat) == G0 G0' on a random sample of entries
    rng = np.random.default_rng(0)
    iu = rng.integers(0, p, 2000); iv = rng.integers(0, p, 2000)
    want = np.einsum("il,il->i", pm["G0"][iu], pm["G0"][iv]) + np.where(iu == iv, pm["SigmaHat"][iu], 0.0)
    assert np.max(np.abs(C[iu, iv] - want)) < 1e-12 * np.max(np.abs(want))
    # linearity of the accumulators: sum over two disjoint draw ranges == sum over the union
    post = fb.Posterior(Y, 1.0, 1.0, s["u"], s["v"], s["d"], 10, 1.26, ctx=ctx)
    post.draw_batch(0, 8, accumulate=True); a1, a2 = post.accum_fetch()
    post.accum_reset(); post.draw_batch(0, 3, accumulate=True); post.draw_batch(3, 5, accumulate=True)
    b1, b2 = post.accum_fetch()
    assert relerr_mat(b1, a1) < 1e-13 and relerr_mat(b2, a2) < 1e-13
    # a materialised draw equals Lambda_m Lambda_m' + diag(sigma2_m) rebuilt from the returned samples
    psi = torch.empty((1, p, p), dtype=torch.float64, device="cuda")
    post.draw_batch(5, 1, dev_psi=psi.data_ptr(), psi_slots=1, accumulate=False)
    ctx.synchronize()
    L, S = post.samples(5, 1)
    Lm = L[0].reshape(p, k)
    want = Lm @ Lm.T + np.diag(S[0])
    assert relerr_mat(psi[0].cpu().numpy().T, want) < 1e-12
    post.close()


# ---- full size against the oracle (BASELINE.json configs[2] and configs[3] on one GPU) -----------------
@pytest.mark.parametrize("n,p,k,maxProp", [(1000, 20000, 10, 0.5), (500, 50000, 20, 0.5)], ids=["c3", "c4"])
def test_full_size_against_oracle(ctx, n, p, k, maxProp):
    """The whole path at the size the metric is quoted on (c3) and at the high-dimensional size (c4), against the CPU
    oracle on the same inputs: singular values and kMax (LAPACK dgesdd), kEst (bisection replayed by the oracle), the
    plug-in variances / G0 / tauSq, and -- with injected draws -- sampled entries of every materialised draw (dense
    p x p) and of the sum / sumsq accumulators."""
    import torch
    Y = synth.factor_data(n, p, k, rep=0)
    u, d, v = orc.svd_thin(Y)
    s = fb.svd(Y, ctx=ctx)
    assert relerr_vec(s["d"], d) < TOL
    kMax = orc.kmax_rule(d, maxProp)
    assert fb.kmax_rule(s["d"], maxProp) == kMax
    kref = orc.rank_estimator(Y, u, v, d, kMax)
    assert kref == k
    assert fb.CPPRankEstimator(Y, u, v, d, kMax, ctx=ctx) == kref                       # the oracle's SVD as shared input
    assert fb.CPPRankEstimator(Y, s["u"], s["v"], s["d"], kMax, ctx=ctx) == kref        # the library's own SVD
    del s
    hg = fb.FABLEHyperParameters(Y, u, v, d, kref, ctx=ctx, want_G=False)
    hr = orc.hyper_parameters(Y, u, v, d, kref, want_G=False)
    assert relerr_vec(hg["SigmaSqEstimate"], hr["SigmaSqEstimate"]) < TOL and relerr_mat(hg["G0"], hr["G0"]) < TOL
    assert abs(hg["tauSqEstimate"] / hr["tauSqEstimate"] - 1) < TOL
    vi = fb.var_inflation(hg["SigmaSqEstimate"], hg["G0"], "script", ctx=ctx)
    post = fb.Posterior(Y, 1.0, 1.0, u, v, d, kref, vi, ctx=ctx)
    st = orc._shrunk_stats(Y, u, v, d, kref, 1.0, 1.0)
    got = post.get()
    assert relerr_mat(got["G0"], st["G0"]) < TOL and relerr_vec(got["gnd"], st["gnd"]) < TOL
    B = 3 if p <= 20000 else 2
    gam, z = synth.injected_draws(B, p, kref, 0.5 * (1.0 + n), seed=p)
    o = orc.sampler(Y, 1.0, 1.0, B, u, v, d, kref, vi, gam, z)
    dg = torch.from_numpy(np.ascontiguousarray(gam)).cuda(); dz = torch.from_numpy(np.ascontiguousarray(z)).cuda()
    psi = torch.empty((B, p, p), dtype=torch.float64, device="cuda")
    post.draw_batch(0, B, dev_psi=psi.data_ptr(), psi_slots=B, accumulate=True, dev_gam=dg.data_ptr(), dev_z=dz.data_ptr())
    ctx.synchronize()
    rng = np.random.default_rng(p)
    idx = np.unique(np.concatenate([[0, 1, 15, 16, 63, 64, 65, 1023, 1024, p - 65, p - 64, p - 17, p - 16, p - 1], rng.integers(0, p, 500)]))
    ti = torch.from_numpy(idx).cuda()
    want1 = np.zeros((idx.size, idx.size)); want2 = np.zeros_like(want1)
    for m in range(B):
        Lm = o["LambdaSamples"][m].reshape(p, kref)[idx]
        want = orc.psi_draw(Lm.reshape(-1), o["SigmaSqSamples"][m][idx], kref)          # the entries' own rows suffice
        want1 += want; want2 += want * want
        sub = psi[m].index_select(0, ti).index_select(1, ti).cpu().numpy()
        assert relerr_mat(sub.T, want) < TOL and np.array_equal(sub, sub.T)
    del psi
    torch.cuda.empty_cache()
    a_sum, a_sq, cnt = post.accum_dev()
    dense = torch.empty((p, p), dtype=torch.float64, device="cuda")
    for ptr, want in ((a_sum, want1), (a_sq, want2)):
        post.tiles_to_dense(ptr, False, 0, p, dense.data_ptr())
        ctx.synchronize()
        sub = dense.index_select(0, ti).index_select(1, ti).cpu().numpy()
        assert relerr_mat(sub.T, want) < TOL and np.array_equal(sub, sub.T)
    del dense
    post.close()
    ctx.trim_cache()
    torch.cuda.empty_cache()


@pytest.mark.parametrize("threads", [None, "3", "0"])
def test_pipelined_host_copies_round_trip(ctx, threads, monkeypatch):
    """Pageable inputs / outputs above 64 MB cross PCIe in staged chunks gathered and scattered by host threads
    (api.cu upload_bytes / download); 0 threads = plain driver copies.  Y (144 MB, three chunks, the last one ragged) goes
    up, V (144 MB) comes back: the factors must reproduce Y wherever it is sampled, whatever the thread count."""
    if threads is None:
        monkeypatch.delenv("FABLE_B200_COPY_THREADS", raising=False)
    else:
        monkeypatch.setenv("FABLE_B200_COPY_THREADS", threads)
    n, p = 200, 90001
    rng = np.random.default_rng(5)
    Y = np.asfortranarray(rng.standard_normal((n, 12)) @ rng.standard_normal((12, p)) + 0.1 * rng.standard_normal((n, p)))
    s = fb.svd(Y, ctx=ctx)
    assert abs(np.sum(s["d"] ** 2) / np.sum(Y * Y) - 1) < 1e-12
    cols = np.concatenate([np.arange(0, 40), rng.integers(0, p, 400), np.arange(p - 40, p)])      # first / random / last chunk
    rebuilt = (s["u"] * s["d"]) @ s["v"][cols].T
    assert relerr_mat(rebuilt, Y[:, cols]) < 1e-11


# ---- partition C: column-sharded estimation (SURVEY 8(e) "Gram/SVD", "Rank / row-posterior") -----------------
class _ThreadAllreduce:
    """In-process stand-in for an NCCL all-reduce: `world` threads, one context each, sum in rank order (so every
    rank sees the same bits)."""

    def __init__(self, world):
        import threading
        self.world = world
        self.bar = threading.Barrier(world, timeout=120)
        self.bufs = [None] * world
        self.tot = None

    def make(self, rank):
        def fn(a):
            self.bufs[rank] = a
            self.bar.wait()
            if rank == 0:
                tot = self.bufs[0].copy()
                for g in range(1, self.world):
                    tot += self.bufs[g]
                self.tot = tot
            self.bar.wait()
            a[...] = self.tot
            self.bar.wait()
        return fn


@pytest.mark.parametrize("n,p,k,cuts", [(60, 200, 3, (0, 90, 200)), (40, 333, 4, (0, 100, 101, 333)), (50, 64, 2, (0, 64))])
def test_column_sharded_estimation_matches_one_device(ctx, n, p, k, cuts):
    """Each shard (a thread with its own context, its own columns of Y) runs svd -> kMax -> rank -> row posterior
    with the sums over variables completed by the all-reduce callback; the result must match the unsharded path
    and the oracle, and a sampler built from the gathered rows must draw the same samples."""
    import threading
    from fable_b200 import parallel
    Y = synth.factor_data(n, p, k, rep=17)
    world = len(cuts) - 1
    ar = _ThreadAllreduce(world)
    res = [None] * world
    errs = []

    def work(g):
        c = fb.Context(0, 4242)
        try:
            res[g] = parallel.sharded_estimation(np.asfortranarray(Y[:, cuts[g]:cuts[g + 1]]), p, cuts[g], g, world, c, ar.make(g))
            post = fb.Posterior.from_rows(n, 1.0, 1.0, res[g]["rows"], res[g]["varInflation"], ctx=c)
            res[g]["samples"] = post.samples(5, 7)
            post.close()
        except Exception as e:                       # noqa: BLE001
            errs.append((g, e))
            ar.bar.abort()
        finally:
            c.close()

    th = [threading.Thread(target=work, args=(g,)) for g in range(world)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs

    one = fb.FABLEPosteriorSampler(Y, MC=2, ctx=ctx)          # the unsharded wrapper on one context
    ref = orc.FABLEPosteriorMean(Y, maxProp=0.95)
    s1 = fb.svd(Y, ctx=ctx)
    for g in range(world):
        r = res[g]
        assert r["estRank"] == one["estRank"] == ref["estRank"]
        assert relerr_vec(r["svd"]["d"], s1["d"]) < TOL
        kk = r["estRank"]
        # leading triplets (the ones the model uses) agree entry-wise, signs included; the whole factorisation reconstructs Y
        assert relerr_mat(r["svd"]["u"][:, :kk], s1["u"][:, :kk]) < 1e-8
        assert relerr_mat(r["svd"]["v"][:, :kk], s1["v"][cuts[g]:cuts[g + 1], :kk]) < 1e-8
        assert relerr_mat((r["svd"]["u"] * r["svd"]["d"]) @ r["svd"]["v"].T, Y[:, cuts[g]:cuts[g + 1]]) < 1e-11
        np.testing.assert_array_equal(r["svd"]["u"], res[0]["svd"]["u"])          # replicated pieces: identical bits
        np.testing.assert_array_equal(r["rows"]["G0"], res[0]["rows"]["G0"])
        assert abs(r["varInflation"] / one["varInflation"] - 1) < TOL
        st = orc._shrunk_stats(Y, ref["svdY"]["u"], ref["svdY"]["v"], ref["svdY"]["d"], kk, 1.0, 1.0)
        sgn = np.sign(np.sum(st["G0"] * r["rows"]["G0"], axis=0))                  # oracle SVD signs are LAPACK's
        assert relerr_mat(r["rows"]["G0"] * sgn, st["G0"]) < TOL
        assert relerr_vec(r["rows"]["gnd"], st["gnd"]) < TOL and relerr_vec(r["rows"]["sig"], st["sig"]) < TOL
        assert abs(r["rows"]["tausq"] / st["tausq"] - 1) < TOL
    # same Philox stream, same row posterior -> same draws as the one-device posterior
    post1 = fb.Posterior(Y, 1.0, 1.0, s1["u"], s1["v"], s1["d"], one["estRank"], one["varInflation"], ctx=ctx)
    ctx.set_seed(4242)
    L1, S1 = post1.samples(5, 7)
    ctx.set_seed(12345)
    post1.close()
    Lg, Sg = res[0]["samples"]
    assert relerr_mat(Lg, L1) < 1e-9 and relerr_vec(Sg, S1) < 1e-9


def test_column_shard_rejects_whole_matrix_entry_points(ctx):
    Y = synth.factor_data(30, 40, 2, rep=1)
    s = fb.svd(Y, ctx=ctx)
    ctx.set_column_shard(80, 0, 0, 1, lambda a: None)
    try:
        with pytest.raises(fb.FableError):
            fb.CPPFABLEPostMean(Y, 1.0, 1.0, s["u"], s["v"], s["d"], 5, ctx=ctx)
        with pytest.raises(fb.FableError):
            fb.CPPFABLESampler(Y, 1.0, 1.0, 3, s["u"], s["v"], s["d"], 2, 1.0, ctx=ctx)
        with pytest.raises(fb.FableError):
            fb.FABLEHyperParameters(Y, s["u"], s["v"], s["d"], 2, ctx=ctx)            # asks for the dense G
    finally:
        ctx.set_column_shard(0)
    assert fb.CPPRankEstimator(Y, s["u"], s["v"], s["d"], 5, ctx=ctx) >= 1               # back to normal
