"""Generates tests/golden/fable_golden.npz from the CPU oracle (oracle/fable_oracle.py).

The reference ships no golden vectors and cannot be built here, so these
fixtures are SELF-GENERATED and frozen: they pin the oracle against drift, not against the R
package ("parity unpinned").  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import fable_oracle as orc  # noqa: E402
from fable_b200 import synth  # noqa: E402

CASES = {"a": dict(n=40, p=120, k=3, rep=0), "b": dict(n=64, p=48, k=2, rep=1), "c": dict(n=33, p=77, k=4, rep=2)}


def main():
    out = {}
    for name, c in CASES.items():
        n, p, k = c["n"], c["p"], c["k"]
        Y = synth.factor_data(n, p, k, rep=c["rep"])
        u, d, v = orc.svd_thin(Y)
        # canonical signs (largest-|entry| of each v column positive) so the fixture is sign-stable
        sgn = np.sign(v[np.argmax(np.abs(v), axis=0), np.arange(v.shape[1])])
        u = u * sgn; v = v * sgn
        kMax50 = orc.kmax_rule(d, 0.5); kMax95 = orc.kmax_rule(d, 0.95)
        crit = np.array([orc.criterion(kk, Y, u, v, d) for kk in range(1, kMax95 + 1)])
        kEst50 = orc.rank_estimator(Y, u, v, d, kMax50); kEst95 = orc.rank_estimator(Y, u, v, d, kMax95)
        lo, hi = orc.bisection(1, kMax95, lambda kk: crit[kk - 1])
        hyp = orc.hyper_parameters(Y, u, v, d, kEst95)
        B = orc.cov_correct_matrix_R(hyp["SigmaSqEstimate"], hyp["G"])
        Bu = orc.cpp_cov_correct_matrix(hyp["SigmaSqEstimate"], hyp["G"])
        pm = orc.post_mean(Y, 1.0, 1.0, u, v, d, kMax50)
        MC = 7
        vi = orc.var_inflation_wrapper(B)
        gam, z = synth.injected_draws(MC, p, kEst95, 0.5 * (1.0 + n), seed=777 + c["rep"])
        smp = orc.sampler(Y, 1.0, 1.0, MC, u, v, d, kEst95, vi, gam, z)
        s1, s2 = orc.psi_moments(smp["LambdaSamples"], smp["SigmaSqSamples"], kEst95)
        pre = name + "_"
        out.update({pre + "Y": Y, pre + "u": u, pre + "d": d, pre + "v": v,
                    pre + "kMax50": kMax50, pre + "kMax95": kMax95, pre + "crit": crit,
                    pre + "kEst50": kEst50, pre + "kEst95": kEst95, pre + "bis": np.array([lo, hi]),
                    pre + "hyp_sig": hyp["SigmaSqEstimate"], pre + "hyp_G0": hyp["G0"], pre + "hyp_tau": hyp["tauSqEstimate"],
                    pre + "hyp_G": hyp["G"], pre + "Bupper": Bu, pre + "vi_wrapper": vi,
                    pre + "vi_script": orc.var_inflation_script(Bu),
                    pre + "pm_cov": pm["CovPostMean"], pre + "pm_G0": pm["G0"], pre + "pm_sigmahat": pm["SigmaHat"],
                    pre + "gam": gam, pre + "z": z, pre + "smp_L": smp["LambdaSamples"], pre + "smp_S": smp["SigmaSqSamples"],
                    pre + "smp_postmean": smp["SigmaSqEstimatePostMean"], pre + "smp_G": smp["G"],
                    pre + "smp_tau": smp["tauSqEstimate"], pre + "psi_sum": s1, pre + "psi_sumsq": s2})
    path = os.path.join(ROOT, "tests", "golden", "fable_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
    for name in CASES:
        print(name, "kMax50/95", out[name + "_kMax50"], out[name + "_kMax95"], "kEst", out[name + "_kEst50"], out[name + "_kEst95"],
              "bis", out[name + "_bis"], "vi", out[name + "_vi_wrapper"], out[name + "_vi_script"])


if __name__ == "__main__":
    main()
