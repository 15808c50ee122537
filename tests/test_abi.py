"""CPU tests of the drop-in boundary: libfable_b200.so loads without a GPU, exports every symbol
include/fable_b200.h declares, refuses loudly to compute without a B200 (no CPU fallback), and
its host-only pieces (kMax rule, argument checks in the Python mirror) behave like the reference."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import fable_b200 as fb
from fable_b200 import api
from oracle import fable_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "fable_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fable_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    syms = _declared_symbols()
    for must in ("fable_svd", "fable_kmax", "fable_criterion", "fable_bisection", "fable_rank_estimator",
                 "fable_hyper_parameters", "fable_cov_correct_matrix", "fable_var_inflation", "fable_post_mean",
                 "fable_sampler", "fable_posterior_draw_batch", "fable_dev_rankk_cov", "fable_dev_gemm"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    lib = api._load()
    missing = [s for s in _declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    assert lib.fable_abi_version() == 2


def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(fb.FableError) as ei:
        fb.Context(0, 1)
    assert ei.value.code == 2 and "no CPU fallback" in str(ei.value)
    with pytest.raises(fb.FableError):
        fb.CPPRankEstimator(np.zeros((4, 3)), np.zeros((4, 3)), np.zeros((3, 3)), np.ones(3), 2)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "fable_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("no oracle", ""), os.path.join(dirpath, f)


@pytest.mark.parametrize("maxProp", [0.05, 0.5, 0.95, 1.0])
def test_kmax_rule_host_matches_oracle(maxProp):
    rng = np.random.default_rng(3)
    d = np.sort(rng.uniform(0.1, 10, 57))[::-1].copy()
    assert fb.kmax_rule(d, maxProp) == orc.kmax_rule(d, maxProp)


def test_kmax_rule_unreachable_prop_is_an_error():
    with pytest.raises(fb.FableError):
        fb.kmax_rule(np.array([3.0, 2.0, 1.0]), 1.5)


def test_argument_validation_in_the_mirror():
    Y = np.zeros((5, 4)); U = np.zeros((5, 4)); V = np.zeros((4, 4)); d = np.ones(4)
    with pytest.raises(fb.FableError):
        api._inputs(Y, np.zeros((6, 4)), V, d)
    with pytest.raises(fb.FableError):
        api._inputs(Y, U, np.zeros((3, 4)), d)
    with pytest.raises(fb.FableError):
        api._F(np.zeros((2, 2, 2)))


def test_r_package_native_layer_type_checks():
    """rpkg/src/*.cpp (the Rcpp layer a maintainer ships) cannot be built here -- no R, no Rcpp -- but it is pure marshalling
    into the C-ABI: parse and type-check it against include/fable_b200.h with a declaration-only Rcpp stand-in (argument
    counts and types of every fable_* call, list elements, the registration table)."""
    import shutil
    import subprocess
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    for src in ("fable_b200_rcpp.cpp", "RcppExports.cpp"):
        r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I", os.path.join(ROOT, "tests", "rcpp_stub"),
                            "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "rpkg", "src", src)],
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-3000:]
    # the registration table carries the reference's fourteen routines at their arities (src/RcppExports.cpp:212-228)
    text = open(os.path.join(ROOT, "rpkg", "src", "RcppExports.cpp")).read()
    want = {"rGamma": 3, "SymmetrizeMatrix": 1, "CPPsvd": 2, "CPPLogLikelihoodEval": 4, "CPPCriterionFunction": 5,
            "CPPBisectionRecursion": 6, "CPPRankEstimator": 5, "FABLEHyperParameters": 5, "CPPcov_correct_matrix": 2,
            "CPPFABLEPostMean": 7, "CPPFABLESampler": 9, "CPPCCFABLEPostProcessing": 2, "CCFABLEPostProcessingSubmatrix": 3,
            "CCFABLEPostProcessingSubmatrix_Optimized": 3}
    for name, arity in want.items():
        assert re.search(r"FB_REG\(%s, %d\)" % (name, arity), text), name
        m = re.search(r"SEXP _FABLE_%s\(([^)]*)\)" % name, text)
        assert m and len([a for a in m.group(1).split(",") if a.strip()]) == arity, name
    rtext = open(os.path.join(ROOT, "rpkg", "R", "RcppExports.R")).read()
    for name, arity in want.items():
        m = re.search(r"^\s*%s\s*=\s*c\(([^)]*)\)" % name, rtext, flags=re.M)
        assert m and len(m.group(1).split(",")) == arity, name
