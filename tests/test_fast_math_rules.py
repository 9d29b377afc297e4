"""The two transcendental schemes of the evaluation epilogues (csrc/common.cuh), restated in numpy by tools/check_fast_exp.py, held to
the error figures the header states -- on the CPU, against extended precision -- and the literal 1/c_j, ln c_j table of
log1p_nonneg_fast compared with its definition.  The reference evaluates exp (-chi2/2) and pow (1 + chi2/nu, kappa) with libm
(ncm_stats_dist_kernel_gauss.c:246-333, ncm_stats_dist_kernel_st.c:239-243, 295-386); the tolerance of the path is 1e-10 relative on the
log-densities (BASELINE.json north_star), these rules sit five orders of magnitude inside it."""
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

LD = np.longdouble
pytestmark = pytest.mark.skipif(np.finfo(LD).eps >= np.finfo(np.float64).eps, reason="no extended precision on this platform")


def test_exp_nonpos_fast_relative_error():
    from check_fast_exp import exp_fast

    rs = np.random.default_rng(0)
    x = -np.abs(rs.normal(size=1_000_000)) * rs.choice([1e-6, 1e-3, 0.01, 1, 10, 100, 300], size=1_000_000)
    x = np.concatenate([x[x > -700], [0.0, -1e-300, -0.5 * np.log(2.0), -np.log(2.0), -699.9]])
    ref = np.exp(x.astype(LD))
    rel = np.abs((exp_fast(x).astype(LD) - ref) / ref)
    assert float(rel.max()) <= 1.0e-15, float(rel.max())                   # header: <= 9e-16 (with FMAs; this restatement has none)
    assert exp_fast(np.array([0.0]))[0] == 1.0


def test_log1p_nonneg_fast_absolute_error():
    from check_fast_exp import log1p_fast

    rs = np.random.default_rng(1)
    x = np.abs(rs.normal(size=1_000_000)) * rs.choice([1e-9, 1e-6, 1e-3, 0.1, 1, 10, 1e3, 1e6, 1e12, 1e18], size=1_000_000)
    ref = np.log1p(x.astype(LD))
    err = np.abs(log1p_fast(x).astype(LD) - ref).astype(np.float64)
    # what enters the density is kappa * log1p: an ABSOLUTE error; one ulp of the result at most on top of 1e-15
    assert np.all(err <= 1.0e-15 + np.spacing(np.asarray(ref, dtype=np.float64))), float(err.max())
    assert float(err[x < 1.0].max()) <= 1.0e-15
    assert abs(log1p_fast(np.array([0.0]))[0]) <= 1.0e-15                  # not exactly 0: the scheme bounds the absolute error only


def test_log_table_literals_are_their_definition():
    """NCM_LOG_TAB[2 j], [2 j + 1] = 1 / c_j, ln c_j with c_j = 1 + (j + 1/2) / 32, correctly rounded."""
    src = open(os.path.join(ROOT, "numcosmo_b200", "csrc", "common.cuh")).read()
    body = src[src.index("NCM_LOG_TAB[64] = {"):]
    body = body[body.index("{") + 1:body.index("};")]
    tab = np.array([float(t) for t in re.findall(r"[-+]?\d\.\d+e[-+]\d+", body)])
    assert tab.size == 64
    c = LD(1.0) + (np.arange(32).astype(LD) + LD(0.5)) / LD(32.0)
    assert np.array_equal(tab[0::2], (LD(1.0) / c).astype(np.float64))
    assert np.array_equal(tab[1::2], np.log(c).astype(np.float64))
