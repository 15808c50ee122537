"""Pins the restated oracle (and the frozen golden fixtures) against the REFERENCE'S OWN translation unit,
src/updated-FABLE-functions.cpp, compiled where it lies against the minimal Armadillo/Rcpp stand-ins of
oracle/shim/ (oracle/_ref/libfable_ref.so, built by oracle/Makefile in the container that has
/root/reference; the prebuilt file travels to the GPU box).  Skipped only if that library is absent."""
import numpy as np
import pytest

from oracle import fable_oracle as orc
from oracle import ref_oracle as ref
from fable_b200 import synth
from conftest import relerr_mat, relerr_vec

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libfable_ref.so not built (needs /root/reference)")

CASES = ["a", "b", "c"]


@pytest.mark.parametrize("name", CASES)
def test_golden_fixtures_match_the_reference_source(golden, name):
    g = lambda key: golden[f"{name}_{key}"]
    Y, u, d, v = g("Y"), g("u"), g("d"), g("v")
    kMax95, kMax50 = int(g("kMax95")), int(g("kMax50"))
    crit = np.array([ref.criterion(k, Y, u, v, d) for k in range(1, kMax95 + 1)])
    assert relerr_vec(crit, g("crit")) < 1e-12
    assert ref.bisection(1, kMax95, Y, u, v, d) == tuple(int(x) for x in g("bis"))
    assert ref.rank_estimator(Y, u, v, d, kMax95) == int(g("kEst95"))
    assert ref.rank_estimator(Y, u, v, d, kMax50) == int(g("kEst50"))
    kEst = int(g("kEst95"))
    hyp = ref.hyper_parameters(Y, u, v, d, kEst)
    assert relerr_vec(hyp["SigmaSqEstimate"], g("hyp_sig")) < 1e-12
    assert relerr_mat(hyp["G0"], g("hyp_G0")) < 1e-12 and relerr_mat(hyp["G"], g("hyp_G")) < 1e-12
    assert abs(hyp["tauSqEstimate"] / float(g("hyp_tau")) - 1) < 1e-12
    assert relerr_vec(ref.cov_correct_matrix(hyp["SigmaSqEstimate"], hyp["G"]), g("Bupper")) < 1e-12
    pm = ref.post_mean(Y, 1.0, 1.0, u, v, d, kMax50)
    assert pm["kEst"] == int(g("kEst50")) and relerr_mat(pm["CovPostMean"], g("pm_cov")) < 1e-12
    smp = ref.sampler(Y, 1.0, 1.0, 7, u, v, d, kEst, float(g("vi_wrapper")), g("gam"), g("z"))
    # every injected variate was consumed, in the documented order (gammas of a draw before its normals)
    assert smp["consumed"] == (g("gam").size, g("z").size)
    assert relerr_mat(smp["LambdaSamples"], g("smp_L")) < 1e-12
    assert relerr_vec(smp["SigmaSqSamples"], g("smp_S")) < 1e-12
    assert relerr_vec(smp["SigmaSqEstimatePostMean"], g("smp_postmean")) < 1e-12
    assert relerr_mat(smp["G"], g("smp_G")) < 1e-12 and abs(smp["tauSqEstimate"] / float(g("smp_tau")) - 1) < 1e-12


@pytest.mark.parametrize("seed", range(6))
def test_rank_search_matches_the_reference_source_on_random_problems(seed):
    rng = np.random.default_rng(seed)
    n, p, k = int(rng.integers(20, 70)), int(rng.integers(20, 90)), int(rng.integers(1, 5))
    Y = synth.factor_data(n, p, k, rep=100 + seed)
    u, d, v = orc.svd_thin(Y)
    for maxProp in (0.5, 0.95):
        kMax = orc.kmax_rule(d, maxProp)
        assert ref.rank_estimator(Y, u, v, d, kMax) == orc.rank_estimator(Y, u, v, d, kMax)
        lo_hi = ref.bisection(1, kMax, Y, u, v, d)
        assert lo_hi == orc.bisection(1, kMax, lambda kk: orc.criterion(kk, Y, u, v, d))
    for kk in (1, min(n, p) // 2, min(n, p) - 3):      # (at full rank the residual is rounding noise)
        assert abs(ref.criterion(kk, Y, u, v, d) / orc.criterion(kk, Y, u, v, d) - 1) < 1e-12


def test_sampler_layout_and_draw_order_are_the_references():
    """The injected-draw contract checked against the reference's own loop (cpp:599-618)."""
    n, p, k, MC = 25, 11, 3, 5
    Y = synth.factor_data(n, p, k, rep=55)
    u, d, v = orc.svd_thin(Y)
    gam, z = synth.injected_draws(MC, p, k, 0.5 * (1.0 + n), seed=4)
    r = ref.sampler(Y, 2.0, 0.5, MC, u, v, d, k, 1.9, gam, z)
    o = orc.sampler(Y, 2.0, 0.5, MC, u, v, d, k, 1.9, gam, z)
    assert r["consumed"] == (MC * p, MC * p * k)
    for key in ("LambdaSamples", "SigmaSqSamples", "SigmaSqEstimatePostMean", "SigmaSqEstimatePlugIn", "G"):
        assert relerr_mat(r[key], o[key]) < 1e-12, key
    # a permuted normal stream must NOT reproduce the oracle: the order matters and is checked
    z_bad = np.ascontiguousarray(np.transpose(z, (0, 2, 1)))
    r_bad = ref.sampler(Y, 2.0, 0.5, MC, u, v, d, k, 1.9, gam, z_bad.reshape(MC, k, p))
    assert relerr_mat(r_bad["LambdaSamples"], o["LambdaSamples"]) > 1e-3


@pytest.mark.parametrize("MC", [1, 2, 9, 100])
def test_post_processing_matches_the_reference_source(MC):
    n, p, k = 30, 8, 2
    Y = synth.factor_data(n, p, k, rep=56)
    u, d, v = orc.svd_thin(Y)
    gam, z = synth.injected_draws(MC, p, k, 0.5 * (1.0 + n), seed=MC)
    o = orc.sampler(Y, 1.0, 1.0, MC, u, v, d, k, 1.3, gam, z)
    a = ref.post_processing(o, 0.05); b = orc.post_processing(o, 0.05)
    sel = [3, 1, 8, 3]
    a1 = ref.post_processing(o, 0.1, sel); a2 = ref.post_processing(o, 0.1, sel, optimized=True)
    b1 = orc.post_processing_submatrix(o, 0.1, sel)
    for key in b:
        assert relerr_mat(a[key], b[key]) < 1e-12, key
        assert relerr_mat(a1[key], b1[key]) < 1e-12 and relerr_mat(a2[key], b1[key]) < 1e-12, key


def test_symmetrize_and_loglik_match_the_reference_source():
    rng = np.random.default_rng(1)
    A = np.triu(rng.standard_normal((6, 6)))
    assert relerr_mat(ref.symmetrize(A), orc.symmetrize(A)) < 1e-15
    Y = rng.standard_normal((9, 5)); M = rng.standard_normal((9, 2)); Lam = rng.standard_normal((5, 2)); sig = rng.uniform(0.5, 2, 5)
    assert abs(ref.log_likelihood_eval(Y, M, Lam, sig) / orc.log_likelihood_eval(Y, M, Lam, sig) - 1) < 1e-12


def test_reference_bounds_errors_surface_as_errors():
    Y = synth.factor_data(10, 8, 2, rep=57)
    u, d, v = orc.svd_thin(Y)
    with pytest.raises(RuntimeError):
        ref.criterion(0, Y, u, v, d)            # cols(0, -1): Armadillo bounds error (cpp:144)
    with pytest.raises(RuntimeError):
        ref.criterion(d.size + 1, Y, u, v, d)
