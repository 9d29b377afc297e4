"""Robust covariance types, CPU side: ncm_stats_vec.c:1821-2072 restated in the oracle and in the host library.

GSL's gsl_stats_Qn_from_sorted_data is a third-party routine absent from /root/reference and from this image: its finite-sample factors
are "parity unpinned".  What IS checked: the order statistic itself (two independent algorithms, bit for bit), the estimator's defining
properties (consistency at the normal, scale equivariance, 50 % breakdown) and the OGK matrix on clean and contaminated samples.
"""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import mvnd_problem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_dp = C.POINTER(C.c_double)


def _host_Qn(x):
    L = C.CDLL(os.path.join(ROOT, "numcosmo_b200", "lib", "libncm_stats_dist_b200.so"))
    L.ncm_b200_test_Qn.restype = C.c_double
    L.ncm_b200_test_Qn.argtypes = [_dp, C.c_int]
    x = np.ascontiguousarray(x, dtype=np.float64)
    return L.ncm_b200_test_Qn(x.ctypes.data_as(_dp), len(x))


def test_Qn_host_bisection_equals_oracle_full_sort_bit_for_bit(oracle):
    rng = np.random.default_rng(0)
    makers = [lambda n: rng.standard_normal(n) * 3, lambda n: np.round(rng.standard_normal(n), 1), lambda n: np.ones(n),
              lambda n: rng.standard_cauchy(n), lambda n: np.arange(n, dtype=float), lambda n: 1e-300 * rng.standard_normal(n)]
    for n in (2, 3, 4, 5, 8, 12, 13, 64, 101, 1000, 2001):
        for mk in makers:
            x = mk(n)
            assert _host_Qn(x) == oracle.stats_Qn(x)


def test_Qn_order_statistic_and_properties(oracle):
    rng = np.random.default_rng(1)
    # k-th smallest pairwise difference, k = h (h - 1) / 2, h = n / 2 + 1 (Rousseeuw & Croux 1993), times 2.21914 d_n
    for n in (13, 40, 77):
        x = rng.standard_normal(n)
        diffs = np.sort(np.abs(x[:, None] - x[None, :])[np.triu_indices(n, 1)])
        h = n // 2 + 1
        q0 = diffs[h * (h - 1) // 2 - 1]
        dn = (1.60188 + (-2.1284 - 5.172 / n) / n) if n % 2 else (3.67561 + (1.9654 + (6.987 - 77.0 / n) / n) / n)
        assert abs(oracle.stats_Qn(x) / (2.21914 * q0 / (dn / n + 1.0)) - 1) < 1e-14
    x = 3.0 * rng.standard_normal(3000)
    assert abs(oracle.stats_Qn(x) / 3.0 - 1) < 0.05                       # consistent for sigma at the normal
    assert abs(oracle.stats_Qn(5.0 * x + 7.0) / oracle.stats_Qn(x) - 5.0) < 1e-9   # location invariant, scale equivariant
    y = x.copy()
    y[: len(y) // 3] = 1e6 * rng.standard_normal(len(y) // 3)           # a third of gross outliers barely moves it
    assert oracle.stats_Qn(y) < 3.0 * oracle.stats_Qn(x)


@pytest.mark.parametrize("sd_s", ["kde", "vkde"])
def test_oracle_robust_covariances(oracle, sd_s):
    d, n = 4, 600
    mu, cov, X, m2lnL = mvnd_problem(oracle, d, n, seed=9)
    sdt = oracle.SD_KDE if sd_s == "kde" else oracle.SD_VKDE
    plain = oracle.StatsDist(sdt, oracle.KERNEL_GAUSS, d)
    plain.add_obs_matrix(X)
    assert plain.prepare() == 0
    S0 = plain.peek_full_cov().copy()
    # ROBUST_DIAG: the squared Q_n of each coordinate on the diagonal, nothing else (ncm_stats_vec.c:1858-1876)
    o = oracle.StatsDist(sdt, oracle.KERNEL_GAUSS, d)
    o.set_cov_type(oracle.COV_ROBUST_DIAG)
    o.add_obs_matrix(X)
    assert o.prepare_interp(m2lnL) == 0
    Cd = o.peek_full_cov()
    assert np.array_equal(np.diag(Cd), np.array([oracle.stats_Qn(X[:, i]) ** 2 for i in range(d)])) and np.count_nonzero(Cd - np.diag(np.diag(Cd))) == 0
    assert np.allclose(np.diag(Cd), np.diag(S0), rtol=0.2)
    U = o.peek_full_cov_decomp()
    assert np.allclose(np.triu(U), np.diag(np.sqrt(np.diag(Cd))), rtol=1e-14, atol=0)
    # ROBUST (OGK): close to the sample covariance on clean data, and still close to it with 10 % gross outliers, which wreck the sample one
    o = oracle.StatsDist(sdt, oracle.KERNEL_GAUSS, d)
    o.set_cov_type(oracle.COV_ROBUST)
    o.add_obs_matrix(X)
    assert o.prepare_interp(m2lnL) == 0
    Cr = o.peek_full_cov().copy()
    sc = np.sqrt(np.outer(np.diag(S0), np.diag(S0)))
    assert np.max(np.abs(Cr - S0) / sc) < 0.2 and np.allclose(Cr, Cr.T) and np.all(np.linalg.eigvalsh(Cr) > 0)
    Xc = X.copy()
    Xc[::10] += 50.0 * np.sqrt(np.diag(S0)) * np.random.default_rng(2).standard_normal((len(Xc[::10]), d))
    o2 = oracle.StatsDist(sdt, oracle.KERNEL_GAUSS, d)
    o2.set_cov_type(oracle.COV_ROBUST)
    o2.add_obs_matrix(Xc)
    assert o2.prepare() == 0
    p2 = oracle.StatsDist(sdt, oracle.KERNEL_GAUSS, d)
    p2.add_obs_matrix(Xc)
    assert p2.prepare() == 0
    assert np.max(np.abs(o2.peek_full_cov() - S0) / sc) < 1.0 < 5.0 < np.max(np.abs(p2.peek_full_cov() - S0) / sc)
    if sd_s == "vkde":
        # per-centre factors come from the same estimator over the k nearest neighbours (vkde.c:467-472)
        assert o.peek_cov_array().shape == (n, d, d) and np.all(np.isfinite(o.peek_lnnorms()))


def test_oracle_robust_needs_four_points(oracle):
    d = 2
    X = np.random.default_rng(0).standard_normal((40, d))
    o = oracle.StatsDist(oracle.SD_VKDE, oracle.KERNEL_GAUSS, d)
    o.set_cov_type(oracle.COV_ROBUST_DIAG)
    o.set_local_frac(0.06)          # k = 2 neighbours < 4 (ncm_stats_vec.c:1842-1844)
    o.add_obs_matrix(X)
    assert o.prepare() == -7
