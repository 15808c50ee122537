"""CPU tests of the oracle (oracle/): frozen golden fixtures, analytic known answers
(SURVEY.md App. A.6), control-flow replay of the bisection (reference
src/updated-FABLE-functions.cpp:215-341), and the C twin against the numpy restatement.

The reference ships no tests or golden vectors and cannot be built in this image, so these pin
the oracle against drift and against implementation-independent identities ("parity unpinned"
with respect to the R package itself -- see oracle/fable_oracle.py header).
"""
import math

import numpy as np
import pytest

from oracle import fable_oracle as orc
from oracle import c_oracle
from fable_b200 import synth
from conftest import relerr_mat, relerr_vec, TOL

CASES = ["a", "b", "c"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_golden(golden, name):
    g = lambda key: golden[f"{name}_{key}"]
    Y, u, d, v = g("Y"), g("u"), g("d"), g("v")
    assert orc.kmax_rule(d, 0.5) == int(g("kMax50"))
    assert orc.kmax_rule(d, 0.95) == int(g("kMax95"))
    kMax95 = int(g("kMax95"))
    crit = np.array([orc.criterion(k, Y, u, v, d) for k in range(1, kMax95 + 1)])
    assert relerr_vec(crit, g("crit")) < 1e-13
    assert orc.rank_estimator(Y, u, v, d, int(g("kMax50"))) == int(g("kEst50"))
    kEst = orc.rank_estimator(Y, u, v, d, kMax95)
    assert kEst == int(g("kEst95"))
    hyp = orc.hyper_parameters(Y, u, v, d, kEst)
    assert relerr_vec(hyp["SigmaSqEstimate"], g("hyp_sig")) < 1e-13
    assert relerr_mat(hyp["G"], g("hyp_G")) < 1e-13
    Bu = orc.cpp_cov_correct_matrix(hyp["SigmaSqEstimate"], hyp["G"])
    assert relerr_vec(Bu, g("Bupper")) < 1e-13
    pm = orc.post_mean(Y, 1.0, 1.0, u, v, d, int(g("kMax50")))
    assert relerr_mat(pm["CovPostMean"], g("pm_cov")) < 1e-13
    smp = orc.sampler(Y, 1.0, 1.0, 7, u, v, d, kEst, float(g("vi_wrapper")), g("gam"), g("z"))
    assert relerr_mat(smp["LambdaSamples"], g("smp_L")) < 1e-13
    assert relerr_vec(smp["SigmaSqSamples"], g("smp_S")) < 1e-13


def test_golden_svd_is_a_decomposition_of_Y(golden):
    for name in CASES:
        Y, u, d, v = (golden[f"{name}_{key}"] for key in ("Y", "u", "d", "v"))
        assert relerr_mat((u * d) @ v.T, Y) < 1e-12
        assert np.all(np.diff(d) <= 0)
        r = d.size
        assert relerr_mat(u.T @ u, np.eye(r)) < 1e-12 and relerr_mat(v.T @ v, np.eye(r)) < 1e-12


# ---- analytic known answers (independent of any implementation) ---------------------------------
def _orthogonal_design(n=8, p=5, k=2):
    """Y = U0 diag(d0) V0' + noise orthogonal to U0, with hand-computable c, q, s."""
    rng = np.random.Generator(np.random.Philox(11))
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    W, _ = np.linalg.qr(rng.standard_normal((p, p)))
    dfull = np.array([9.0, 6.0, 1.0, 0.7, 0.4])
    Y = (Q[:, :p] * dfull) @ W.T
    return Y, Q[:, :p], W, dfull


def test_row_stats_known_answer():
    Y, U, V, d = _orthogonal_design()
    n, p = Y.shape
    k = 2
    c, q, s, sig, tausq = orc.row_stats(Y, U, V, d, k)
    assert relerr_mat(c, V[:, :k] * d[:k]) == 0.0
    # consistent SVD: the explicit residual equals the tail of the spectrum row by row
    tail = ((V[:, k:] * d[k:]) ** 2).sum(axis=1)
    assert relerr_vec(sig, tail / n) < 1e-12
    assert relerr_vec(s, q + tail) < 1e-12
    assert abs(tausq - np.mean(q / (tail / n)) / (n * k)) < 1e-15 * tausq * 10


def test_criterion_formula_known_answer():
    Y, U, V, d = _orthogonal_design()
    n, p = Y.shape
    for k in (1, 2, 3):
        tail = ((V[:, k:] * d[k:]) ** 2).sum(axis=1) / n
        want = n * p + n * np.sum(np.log(tail)) + k * max(n, p) * math.log(min(n, p))
        assert abs(orc.criterion(k, Y, U, V, d) - want) < 1e-10 * abs(want)


def test_post_mean_diagonal_quirk_Q4():
    Y = synth.factor_data(50, 30, 2, rep=3)
    u, d, v = orc.svd_thin(Y)
    pm = orc.post_mean(Y, 1.0, 1.0, u, v, d, 5)
    G = pm["G0"] @ pm["G0"].T
    assert relerr_mat(pm["CovPostMean"] - G, np.diag(pm["SigmaHat"])) < 1e-12   # cpp:509-512


def test_hyper_is_unshrunk_postmean_is_shrunk_Q3():
    Y = synth.factor_data(50, 30, 2, rep=4)
    u, d, v = orc.svd_thin(Y)
    n = Y.shape[0]
    hyp = orc.hyper_parameters(Y, u, v, d, 2)
    assert relerr_mat(hyp["G0"], (v[:, :2] * d[:2]) / math.sqrt(n)) < 1e-15       # cpp:390
    st = orc._shrunk_stats(Y, u, v, d, 2, 1.0, 1.0)
    assert relerr_mat(st["G0"], math.sqrt(n) / (n + 1 / st["tausq"]) * (v[:, :2] * d[:2])) < 1e-15   # cpp:490


def test_sampler_moments_known_answer():
    """E[sigma2_j] = gnd_j/(gamma_n-2); Var[Lambda_jl] = a_n^2 E[sigma2_j]  (SURVEY A.6)."""
    n, p, k, MC = 30, 12, 2, 40000
    Y = synth.factor_data(n, p, k, rep=6)
    u, d, v = orc.svd_thin(Y)
    gam, z = synth.injected_draws(MC, p, k, 0.5 * (1.0 + n), seed=5)
    o = orc.sampler(Y, 1.0, 1.0, MC, u, v, d, k, 1.7, gam, z, want_G=False)
    S = o["SigmaSqSamples"]
    a = 0.5 * o["gamma_n"]
    want_mean = o["gnd"] / (o["gamma_n"] - 2.0)
    assert relerr_vec(o["SigmaSqEstimatePostMean"], want_mean) < 1e-15
    sd = np.sqrt((0.5 * o["gnd"]) ** 2 / ((a - 1) ** 2 * (a - 2)) / MC)
    assert np.all(np.abs(S.mean(axis=0) - want_mean) < 5 * sd)
    L = o["LambdaSamples"].reshape(MC, p, k)
    assert np.all(np.abs(L.mean(axis=0) - o["G0"]) < 5 * np.sqrt(o["a_n"] ** 2 * want_mean / MC)[:, None])
    var = L.var(axis=0)
    assert np.all(np.abs(var / (o["a_n"] ** 2 * want_mean)[:, None] - 1.0) < 0.08)


def test_sampler_layout_matches_reference_indexing():
    """LambdaSamples(m, j*k + l) = Lambda_m[j, l]  (cpp:616: vectorise(LambdaSample.t()).t())."""
    n, p, k, MC = 20, 7, 3, 4
    Y = synth.factor_data(n, p, k, rep=7)
    u, d, v = orc.svd_thin(Y)
    gam, z = synth.injected_draws(MC, p, k, 0.5 * (1.0 + n), seed=9)
    o = orc.sampler(Y, 1.0, 1.0, MC, u, v, d, k, 1.0, gam, z, want_G=False)
    for m in range(MC):
        s2 = 0.5 * o["gnd"] / gam[m]
        Lam = o["G0"] + o["a_n"] * (z[m].T * np.sqrt(s2)[:, None])        # p x k, cpp:613-615
        np.testing.assert_allclose(o["LambdaSamples"][m], Lam.reshape(-1), rtol=0, atol=1e-15)
        np.testing.assert_allclose(o["SigmaSqSamples"][m], s2, rtol=1e-15)


# ---- bisection control flow (cpp:215-271, 314-337) ------------------------------------------------
def _bisect_ref_recursive(lower, upper, f):
    if upper - lower > 5:
        mid = (lower + upper) // 2
        if f(lower) < f(mid):
            return _bisect_ref_recursive(lower, mid, f)
        tq = (mid + upper) // 2
        if f(mid) <= f(tq):
            return _bisect_ref_recursive(lower, mid, f)
        return _bisect_ref_recursive(mid, upper, f)
    return lower, upper


@pytest.mark.parametrize("seed", range(8))
def test_bisection_iterative_equals_recursive(seed):
    rng = np.random.default_rng(seed)
    kMax = int(rng.integers(6, 400))
    kstar = int(rng.integers(1, kMax + 1))
    vals = {k: (k - kstar) ** 2 + 3.0 * rng.standard_normal() for k in range(1, kMax + 1)}
    f = lambda k: vals[k]
    assert orc.bisection(1, kMax, f) == _bisect_ref_recursive(1, kMax, f)


def test_bisection_window_edges_and_ties():
    assert orc.bisection(1, 6, lambda k: 0.0) == (1, 6)          # width 5: no shrink (cpp:223)
    assert orc.bisection(1, 7, lambda k: 0.0) == (1, 4)          # tie: mid <= tq -> [lower, mid] (cpp:249)
    assert orc.bisection(1, 100, lambda k: float(k)) == (1, 4)   # increasing: always left
    lo, hi = orc.bisection(1, 100, lambda k: -float(k))          # decreasing: always right
    assert hi == 100 and hi - lo <= 5


def test_rank_first_minimum_wins():
    Y = synth.factor_data(40, 60, 3, rep=8)
    u, d, v = orc.svd_thin(Y)
    trace = []
    k = orc.rank_estimator(Y, u, v, d, 20, trace)
    win = [(kk, val) for tag, kk, val in trace if tag == "win"]
    assert k == min(win, key=lambda t: (t[1], t[0]))[0]


# ---- cov-correct --------------------------------------------------------------------------------
def test_cov_correct_forms_agree():
    rng = np.random.default_rng(0)
    p, k = 9, 2
    G0 = rng.standard_normal((p, k)); sig = rng.uniform(0.5, 2.0, p)
    G = G0 @ G0.T
    B = orc.cov_correct_matrix_R(sig, G)
    Bu = orc.cpp_cov_correct_matrix(sig, G)
    iu = np.triu_indices(p)
    # trimatu_ind is column-major: sort the (row, col) pairs by column then row
    order = np.lexsort((iu[0], iu[1]))
    np.testing.assert_array_equal(Bu, B[iu[0][order], iu[1][order]])
    assert abs(B[0, 0] - math.sqrt(1 + G[0, 0] / (2 * sig[0]))) < 1e-15           # SURVEY a7 diagonal
    assert B[1, 2] == pytest.approx(math.sqrt(1 + (G[1, 1] * G[2, 2] + G[1, 2] ** 2) / (G[1, 1] * sig[2] + sig[1] * G[2, 2])), rel=1e-15)
    # D5: the two varInflation formulas differ (wrapper sums the FULL matrix)
    assert orc.var_inflation_wrapper(B) > 3.0 * orc.var_inflation_script(Bu)


# ---- quantile / post-processing -------------------------------------------------------------------
def test_arma_quantile_definition5():
    x = np.array([3.0, 1.0, 2.0, 4.0])
    assert orc.arma_quantile(x, 0.5) == 2.5
    assert orc.arma_quantile(x, 0.0) == 1.0 and orc.arma_quantile(x, 1.0) == 4.0
    assert orc.arma_quantile(x, 0.125) == 1.0 and orc.arma_quantile(x, 0.25) == 1.5
    for pr in (0.03, 0.2, 0.5, 0.77, 0.975):
        y = np.random.default_rng(1).standard_normal(101)
        assert orc.arma_quantile(y, pr) == pytest.approx(np.quantile(y, pr, method="hazen"), rel=1e-13)


def test_arma_quantile_at_p_max_is_the_maximum():
    """alpha N == 1: the level 1 - alpha/2 is exactly (N - 0.5)/N, h = N, the weight of the missing upper neighbour is
    0: the value is the maximum (and alpha/2 = 0.5/N gives the minimum)."""
    rng = np.random.default_rng(5)
    for N, alpha in ((100, 0.01), (20, 0.05), (10, 0.1), (16, 0.0625), (1000, 0.001)):
        x = rng.standard_normal(N)
        assert orc.arma_quantile(x, 1.0 - alpha / 2.0) == x.max()
        assert orc.arma_quantile(x, alpha / 2.0) == x.min()


def test_prototype_criterion_known_structure():
    """R prototype criterion (R/FABLE_code.R:57-116): with sigma^2 = resid / (n - k) the quadratic term of the log-likelihood is
    exactly -p (n - k) / 2, so criterion_R(k) = [n sum_j log sigma_j^2 + p (n - k) + k max(n,p) log min(n,p)] / (n p)."""
    from fable_b200 import synth
    n, p = 40, 90
    Y = synth.factor_data(n, p, 3, rep=2)
    u, d, v = orc.svd_thin(Y)
    for k in (1, 3, 7):
        resid = np.sum((Y - (u[:, :k] * d[:k]) @ v[:, :k].T) ** 2, axis=0)
        want = (n * np.sum(np.log(resid / (n - k))) + p * (n - k) + k * max(n, p) * math.log(min(n, p))) / (n * p)
        assert orc.criterion_R(k, Y, u, v, d) == pytest.approx(want, rel=1e-12)
    kMax = orc.kmax_rule(d, 0.95)
    k_r = orc.rank_estimator_R(Y, u, v, d, kMax)
    vals = [orc.criterion_R(k, Y, u, v, d) for k in range(1, kMax + 1)]
    assert 1 <= k_r <= kMax and vals[k_r - 1] <= min(vals) + 1e-3          # bisection lands in a near-minimal window (Q1)


def test_post_processing_mean_matches_moments():
    n, p, k, MC = 25, 6, 2, 11
    Y = synth.factor_data(n, p, k, rep=9)
    u, d, v = orc.svd_thin(Y)
    gam, z = synth.injected_draws(MC, p, k, 0.5 * (1.0 + n), seed=3)
    o = orc.sampler(Y, 1.0, 1.0, MC, u, v, d, k, 1.2, gam, z)
    pp = orc.post_processing(o, 0.05)
    s1, s2 = orc.psi_moments(o["LambdaSamples"], o["SigmaSqSamples"], k)
    assert relerr_mat(pp["PostMeanMatrix"], s1 / MC) < 1e-13
    assert np.all(pp["LowerQuantileMatrix"] <= pp["UpperQuantileMatrix"])
    sub = orc.post_processing_submatrix(o, 0.05, [2, 5])
    assert relerr_mat(sub["PostMeanMatrix"], pp["PostMeanMatrix"][np.ix_([1, 4], [1, 4])]) < 1e-14


# ---- C twin ---------------------------------------------------------------------------------------
def test_c_twin_matches_numpy_oracle(golden):
    g = lambda key: golden[f"a_{key}"]
    Y, u, d, v = g("Y"), g("u"), g("d"), g("v")
    k = int(g("kEst95"))
    assert relerr_vec(c_oracle.resid_sigsq(Y, u, v, d, k), orc._residual_sigsq(Y, u, v, d, k)) < 1e-12
    st = orc._shrunk_stats(Y, u, v, d, k, 1.0, 1.0)
    a_n = math.sqrt(float(g("vi_wrapper"))) / math.sqrt(st["w"])
    L, S = c_oracle.sampler(7, Y.shape[1], k, st["G0"], st["gnd"], a_n, g("gam"), g("z"))
    assert relerr_mat(L, g("smp_L")) < 1e-14 and relerr_vec(S, g("smp_S")) < 1e-14
    s1, s2 = c_oracle.psi_moments(L, S, k)
    assert relerr_mat(s1, g("psi_sum")) < 1e-13 and relerr_mat(s2, g("psi_sumsq")) < 1e-13
    assert relerr_mat(c_oracle.psi_draw(L, S, k, 3), orc.psi_draw(g("smp_L")[3], g("smp_S")[3], k)) < 1e-14
    sf, su = c_oracle.cov_correct_sums(g("hyp_sig"), g("hyp_G"))
    p = Y.shape[1]
    assert ((sf / (p * (p + 1) / 2)) ** 2) == pytest.approx(float(g("vi_wrapper")), rel=1e-12)
    assert ((su / (p * (p + 1) / 2)) ** 2) == pytest.approx(float(g("vi_script")), rel=1e-12)


# ---- fable-philox-v1 (the library's own RNG specification; CPU statement in the C twin) -----------
def test_philox4x32_10_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    assert c_oracle.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert c_oracle.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert c_oracle.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_native_stream_moments_and_batch_invariance():
    MC, p, k, shape = 4000, 16, 3, 15.5
    gam, z = c_oracle.rng_fill(42, 0, MC, p, k, shape)
    assert abs(gam.mean() - shape) < 5 * math.sqrt(shape / gam.size)
    assert abs(gam.var() / shape - 1.0) < 0.05
    assert abs(z.mean()) < 5 / math.sqrt(z.size) and abs(z.var() - 1.0) < 0.02
    # a draw is a pure function of (seed, m, j, slot): any sub-range reproduces the same values
    g2, z2 = c_oracle.rng_fill(42, 1000, 50, p, k, shape)
    np.testing.assert_array_equal(g2, gam[1000:1050]); np.testing.assert_array_equal(z2, z[1000:1050])
    g3, _ = c_oracle.rng_fill(43, 0, 10, p, k, shape, want_z=False)
    assert not np.array_equal(g3, gam[:10])


# ---- schedule of the block-Jacobi eigensolver (fable_b200/csrc/eig.cu): a restatement of its index maps ----------
def _bj_pair(i, rnd, NB):
    """bj_pair of eig.cu: block pair number i of round rnd (round-robin over an even number NB of blocks)."""
    if i == 0:
        ca, cb = NB - 1, rnd
    else:
        ca, cb = (rnd + i) % (NB - 1), (rnd - i + (NB - 1)) % (NB - 1)
    return (ca, cb) if ca < cb else (cb, ca)


def _inner_pairs(step, cross_only, BW=32):
    """column pairs (within the 64-column panel of a block pair) rotated in parallel at one step of k_bj_eig"""
    PW = 2 * BW
    out = []
    for t in range(BW):
        if cross_only:
            a, b = t, BW + ((t + step) % BW)
        elif t == 0:
            a, b = PW - 1, step
        else:
            a, b = (step + t) % (PW - 1), (step - t + (PW - 1)) % (PW - 1)
        out.append((min(a, b), max(a, b)))
    return out


@pytest.mark.parametrize("NB", [2, 4, 8, 32])
def test_block_jacobi_sweep_visits_every_column_pair(NB):
    """One outer sweep = rounds 0 .. NB-2; round 0 runs the full 63-step ordering inside every block pair, the other
    rounds only the 32 x 32 cross pairs.  Claim checked here: every pair of the 32 NB columns is
    rotated at least once per sweep, pairs inside a block exactly once, and the rotations of one step are disjoint."""
    BW = 32
    seen = {}
    for rnd in range(NB - 1):
        blocks = [_bj_pair(i, rnd, NB) for i in range(NB // 2)]
        assert sorted(b for pr in blocks for b in pr) == list(range(NB))            # every block in exactly one pair
        cross_only = rnd > 0
        for bi, bj in blocks:
            col = lambda a: (bi if a < BW else bj) * BW + (a % BW)
            for step in range(BW if cross_only else 2 * BW - 1):
                pairs = _inner_pairs(step, cross_only)
                cols = [c for pr in pairs for c in pr]
                assert len(set(cols)) == 2 * BW                                     # disjoint rotations within a step
                for a, b in pairs:
                    key = tuple(sorted((col(a), col(b))))
                    seen[key] = seen.get(key, 0) + 1
    m = NB * BW
    assert len(seen) == m * (m - 1) // 2                                            # every column pair visited
    for (a, b), cnt in seen.items():
        assert cnt == 1, (a, b, cnt)                                                # ... exactly once: a cyclic sweep


def _sym_blocks(BW=32):
    """thread -> (row pair a, column pair b) of the symmetric cross steps in bj_solve (eig.cu): 17 virtual warps x 32 lanes"""
    out = []
    for v in range(17):
        for lane in range(32):
            if v < 16:
                a, b = (v, lane) if lane >= v else (31 - v, 32 - v + lane)
            elif lane < 16:
                a = b = 16 + lane
            else:
                continue
            out.append((a, b))
    return out


def _rotation(hpq, hpp, hqq):
    if hpq == 0.0:
        return 1.0, 0.0
    d, h2 = hqq - hpp, 2.0 * hpq
    t = h2 / (d + np.copysign(np.sqrt(d * d + h2 * h2), d))
    c = 1.0 / np.sqrt(1.0 + t * t)
    return c, c * t


def test_inner_solve_upper_triangle_scheme_equals_the_full_update():
    """bj_solve keeps H as its upper triangle: entry (r, c) lives at [min][max]; a block (row pair a, column pair b) with
    a <= b is rotated from both sides by one thread (528 blocks), a diagonal block through the same four addresses, and
    warp 0 derives the next rotations from upper entries only.  Restated here with NaN in the lower triangle: every upper
    entry is written exactly once per step, nothing below the diagonal is read or written, and the result equals
    J' H J of the full matrix; the recorded rotations replayed with the lane-shift scheme of k_bj_replay give R."""
    BW, PW = 32, 64
    rng = np.random.default_rng(3)
    A = rng.standard_normal((PW, PW + 7))
    Hf = A @ A.T
    iu = np.triu_indices(PW)
    Hc = np.full((PW, PW), np.nan); Hc[iu] = Hf[iu]
    blocks = _sym_blocks()
    assert len(blocks) == 528 and len(set(blocks)) == 528 and all(a <= b for a, b in blocks)
    rot = [_rotation(Hc[i, BW + i], Hc[i, i], Hc[BW + i, BW + i]) for i in range(BW)]
    recorded, R_full = [], np.eye(PW)
    for step in range(BW):
        recorded.append(list(rot))
        o = [(x + step) % BW for x in range(BW)]
        J = np.eye(PW)
        for i in range(BW):
            c, s = rot[i]; pi, qi = i, BW + o[i]
            J[pi, pi] = c; J[qi, qi] = c; J[pi, qi] = s; J[qi, pi] = -s
        Hf_next = J.T @ Hf @ J
        R_full = R_full @ J
        nxt = None
        if step + 1 < BW:                       # the deriving warp: next pivots from upper entries and the current rotations
            nxt = []
            for i in range(BW):
                i1 = (i + 1) % BW; qi = BW + o[i]; qn = BW + (o[i] + 1) % BW
                r1, r2 = rot[i], rot[i1]
                ap = r1[0] * Hc[i, i] - r1[1] * Hc[i, qi]; aq = r1[0] * Hc[i, qi] - r1[1] * Hc[qi, qi]
                hii = r1[0] * ap - r1[1] * aq
                bp1 = r2[1] * Hc[i1, i1] + r2[0] * Hc[i1, qn]; bq1 = r2[1] * Hc[i1, qn] + r2[0] * Hc[qn, qn]
                hqq = r2[1] * bp1 + r2[0] * bq1
                bp = r2[1] * Hc[min(i, i1), max(i, i1)] + r2[0] * Hc[i, qn]
                bq = r2[1] * Hc[i1, qi] + r2[0] * Hc[min(qi, qn), max(qi, qn)]
                hpq = r1[0] * bp - r1[1] * bq
                assert np.isfinite([hii, hqq, hpq]).all()
                np.testing.assert_allclose([hii, hqq, hpq], [Hf_next[i, i], Hf_next[qn, qn], Hf_next[i, qn]], atol=1e-11)
                nxt.append(_rotation(hpq, hii, hqq))
        Hn = np.full((PW, PW), np.nan)
        for a, b in blocks:
            r1, r2 = rot[a], rot[b]
            lo, hi = min(o[a], o[b]), max(o[a], o[b])
            App, Apq, Aqp, Aqq = (a, b), (a, BW + o[b]), (b, BW + o[a]), (BW + lo, BW + hi)
            hpp, hpq, hqp, hqq = Hc[App], Hc[Apq], Hc[Aqp], Hc[Aqq]
            assert np.isfinite([hpp, hpq, hqp, hqq]).all()                         # nothing below the diagonal is read
            ap = r2[0] * hpp - r2[1] * hpq; bp = r2[1] * hpp + r2[0] * hpq
            aq = r2[0] * hqp - r2[1] * hqq; bq = r2[1] * hqp + r2[0] * hqq
            for addr, val in ((App, r1[0] * ap - r1[1] * aq), (Aqp, r1[1] * ap + r1[0] * aq), (Apq, r1[0] * bp - r1[1] * bq),
                              (Aqq, r1[1] * bp + r1[0] * bq)):
                assert addr[0] <= addr[1] and (np.isnan(Hn[addr]) or a == b)       # one writer per upper entry
                Hn[addr] = val
        assert np.isfinite(Hn[iu]).all() and np.isnan(Hn[np.tril_indices(PW, -1)]).all()
        np.testing.assert_allclose(Hn[iu], Hf_next[iu], atol=1e-11)
        Hc, Hf = Hn, Hf_next
        if nxt is not None:
            rot = nxt
    # k_bj_replay: one warp per row of R; lane i holds column i (p) and column 32 + (i + step) % 32 (q); q moves one lane per step
    R = np.zeros((PW, PW))
    for row in range(PW):
        p_ = np.array([1.0 if row == i else 0.0 for i in range(BW)]); q_ = np.array([1.0 if row == BW + i else 0.0 for i in range(BW)])
        for step in range(BW):
            c = np.array([r[0] for r in recorded[step]]); s = np.array([r[1] for r in recorded[step]])
            np_, nq = c * p_ - s * q_, s * p_ + c * q_
            p_, q_ = np_, np.roll(nq, -1)                                          # lane i takes lane i + 1's value
        R[row, :BW], R[row, BW:] = p_, q_
    np.testing.assert_allclose(R, R_full, atol=1e-13)


@pytest.mark.parametrize("warp", [0, 1, 2, 3])
def test_panel_gram_upper_tiles_nine_per_warp(warp):
    """k_bj_gram: warp q owns tile rows q and 7 - q of the 8 x 8 grid of 8 x 8 tiles, columns from the diagonal on: nine
    tiles each, together the 36 upper tiles exactly once."""
    def tiles(w):
        out = []
        for i in range(9):
            first = i < 8 - w
            out.append((w, w + i) if first else (7 - w, 7 - w + i - (8 - w)))
        return out
    mine = tiles(warp)
    assert len(set(mine)) == 9 and all(0 <= ti <= tj < 8 for ti, tj in mine)
    allt = [t for w in range(4) for t in tiles(w)]
    assert sorted(allt) == [(i, j) for i in range(8) for j in range(i, 8)]


# ---- work split of the tensor-core row pass (fable_b200/csrc/rows.cu, k_rowstats_mma): restated index maps -------
@pytest.mark.parametrize("n,p,grid", [(1000, 20000, 148), (500, 1000, 148), (200, 777, 148), (66, 9, 148), (2, 5, 148),
                                      (2000, 300, 148), (3000, 100, 7), (1000, 2003, 3)])
def test_row_pass_work_split_covers_every_item_once(n, p, grid):
    """16 warps per CTA; G = min(16, items) warps share a tile (8 columns), 16 // G tiles at a time; warp r of a group
    owns items r, r + G, ...; CTA b takes tile groups b, b + grid, ...  Every (tile, item) with tile < ceil(p / 8) must be
    worked on exactly once, and the load cursor must produce the same sequence as the work cursor (the cp.async ring
    is consumed in issue order)."""
    MW, ITEM = 16, 32
    nitems = -(-n // ITEM)
    G = min(nitems, MW)
    TC = MW // G
    ntiles = -(-p // 8)
    nsuper = -(-ntiles // TC)
    grid = min(grid, nsuper)
    IPW = -(-nitems // MW)
    seen = {}
    for b in range(grid):
        for w in range(MW):
            ts, wr = divmod(w, G)
            if ts >= TC:
                continue
            work = []
            for st in range(b, nsuper, grid):                       # work cursor: tiles, then the warp's items
                for q in range(IPW):
                    if q == 0 or wr + q * G < nitems:
                        work.append((st * TC + ts, wr + q * G))
            lst, lit, load = b, wr, []                              # load cursor of issue()
            while len(load) < len(work):
                load.append((lst * TC + ts, lit))
                lit += G
                if lit >= nitems:
                    lit, lst = wr, lst + grid
            assert load == work
            for tile, it in work:
                assert it < nitems
                if tile < ntiles:
                    seen[(tile, it)] = seen.get((tile, it), 0) + 1
    assert len(seen) == ntiles * nitems and set(seen.values()) == {1}
