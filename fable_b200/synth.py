"""Synthetic factor-model data: the generator of the reference's only timing harness
(`extras/replicationCodes/runtimeScript.R:13-41`), with numpy's Philox in place of R's RNG
(R's stream cannot be reproduced outside R; SURVEY.md §8d fixes these seeds).

    Lambda0 = N(0, 0.5^2)^{p x k} * Bernoulli(0.5)^{p x k}      (truth seed 1,  :15-22)
    Sigma0  ~ U(0.5, 5)^p                                        (:24)
    E = N(0,1)^{n x p} diag(sqrt(Sigma0)),  M = N(0,1)^{n x k}   (data seed 2001 + rep, :36-39)
    Y = M Lambda0' + E   (no centering)               (:41)
"""
from __future__ import annotations

import numpy as np


def truth(p, k=10, seed=1, pi0=0.5, lambdasd=0.5):
    g = np.random.Generator(np.random.Philox(seed))
    Lambda0 = g.normal(0.0, lambdasd, size=(p, k)) * (g.random((p, k)) < (1.0 - pi0))
    Sigma0 = g.uniform(0.5, 5.0, size=p)
    return Lambda0, Sigma0


def factor_data(n, p, k=10, rep=0, truth_seed=1):
    """Returns Y (n x p, Fortran order = the column-major layout the C-ABI takes)."""
    Lambda0, Sigma0 = truth(p, k, truth_seed)
    g = np.random.Generator(np.random.Philox(2001 + rep))
    E = g.standard_normal((n, p)) * np.sqrt(Sigma0)[None, :]
    M = g.standard_normal((n, k))
    return np.asfortranarray(M @ Lambda0.T + E)


def injected_draws(MC, p, k, shape, seed=777):
    """Injected sampler draws in the reference's consumption order:
    gam (MC, p) ~ Gamma(shape, 1); z (MC, k, p) ~ N(0,1)."""
    g = np.random.Generator(np.random.Philox(seed))
    gam = g.gamma(shape, 1.0, size=(MC, p))
    z = g.standard_normal((MC, k, p))
    return gam, z
