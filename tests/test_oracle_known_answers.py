"""Pins the CPU oracle (and the host-side kernel/RNG mirror) on the reference's own known answers:
closed forms of tests/c/ncm/stats/test_ncm_stats_dist_kernel.c:179-428 and MT19937's published vectors.
Runs without a GPU."""
import math

import numpy as np
import pytest
from scipy import special

from helpers import mvnd_problem


def _kernels(oracle, d, nu):
    return [oracle.Kernel(oracle.KERNEL_GAUSS, d), oracle.Kernel(oracle.KERNEL_ST, d, nu)]


def _host_kernels(d, nu):
    from numcosmo_b200 import stats_dist as S

    return [S.StatsDistKernelGauss(d), S.StatsDistKernelST(d, nu)]


def test_mt19937_known_answers(oracle):
    r = oracle.RNG(5489)
    first = r.get()
    for _ in range(9998):
        r.get()
    assert first == 3499211612 and r.get() == 4123659995   # Matsumoto & Nishimura reference output
    # gsl_rng_set (r, 0) uses 4357
    a, b = oracle.RNG(0), oracle.RNG(4357)
    assert [a.get() for _ in range(5)] == [b.get() for _ in range(5)]


def test_host_rng_equals_oracle_stream(oracle):
    from numcosmo_b200 import stats_dist as S

    a, b = S.RNG(20260721), oracle.RNG(20260721)
    for _ in range(100):
        assert a.gen_ulong() == b.get()
        assert a.uniform01_gen() == b.uniform()
        assert a.uniform01_pos_gen() == b.uniform_pos()
        assert a.uniform_gen(-2.0, 3.0) == b.flat(-2.0, 3.0)
        assert a.ugaussian_gen() == b.gaussian(1.0)
        assert a.gaussian_gen(1.5, 0.3) == b.gaussian(0.3) + 1.5
        assert a.chisq_gen(3.0) == b.chisq(3.0)
        assert a.chisq_gen(1.0) == b.chisq(1.0)
        assert a.beta_gen(30.0, 30.0) == b.beta(30.0, 30.0)


def test_rng_moments(oracle):
    r = oracle.RNG(11)
    g = np.array([r.gaussian(2.0) for _ in range(40000)])
    z = np.array([r.gaussian_ziggurat(1.0) for _ in range(40000)])
    c = np.array([r.chisq(3.0) for _ in range(40000)])
    assert abs(g.mean()) < 0.05 and abs(g.std() - 2.0) < 0.05
    assert abs(z.mean()) < 0.03 and abs(z.std() - 1.0) < 0.03 and abs((z**4).mean() - 3.0) < 0.2
    assert abs(c.mean() - 3.0) < 0.06 and abs(c.var() - 6.0) < 0.4


@pytest.mark.parametrize("d", [1, 2, 3, 10, 30])
@pytest.mark.parametrize("nu", [1.0, 3.0, 4.5])
def test_bandwidth_and_lnnorm_closed_forms(oracle, d, nu):
    """test_ncm_stats_dist_kernel.c:179-258"""
    n = 1000.0
    rng = oracle.RNG(d)
    cov = oracle.fill_rand_cov(d, 0.5, 2.0, 30.0, rng) if d > 1 else np.array([[1.7]])
    U = np.linalg.cholesky(cov).T
    lndet = np.linalg.slogdet(cov)[1]
    for kset in (_kernels(oracle, d, nu), _host_kernels(d, nu)):
        kg, ks = kset
        assert math.isclose(kg.get_rot_bandwidth(n), (4.0 / (n * (d + 2.0))) ** (1.0 / (d + 4.0)), rel_tol=1e-15)
        nuc = max(nu, 3.0)
        ref = (16.0 * (nuc - 2) ** 2 * (1.0 + d + nuc) * (3.0 + d + nuc) /
               ((2.0 + d) * (d + nuc) * (2.0 + d + nuc) * (d + 2.0 * nuc) * (2.0 + d + 2.0 * nuc) * n)) ** (1.0 / (d + 4.0))
        assert math.isclose(ks.get_rot_bandwidth(n), ref, rel_tol=1e-14)
        assert math.isclose(kg.get_lnnorm(U), 0.5 * (d * math.log(2 * math.pi) + lndet), rel_tol=1e-13, abs_tol=1e-13)
        ref_st = special.gammaln(nu / 2) - special.gammaln((nu + d) / 2) + 0.5 * d * math.log(math.pi * nu) + 0.5 * lndet
        assert math.isclose(ks.get_lnnorm(U), ref_st, rel_tol=1e-13, abs_tol=1e-13)


@pytest.mark.parametrize("d,nu", [(2, 1.0), (5, 3.0), (10, 3.0)])
def test_eval_unnorm_and_gamma_lambda(oracle, d, nu):
    """test_ncm_stats_dist_kernel.c:260-428 incl. the stride-2 vector and the skip-the-max summation"""
    rs = np.random.default_rng(d)
    n = 200
    chi2 = rs.chisquare(d, size=n)
    w = rs.uniform(size=n)
    lnn = rs.normal(size=n)
    for kset in (_kernels(oracle, d, nu), _host_kernels(d, nu)):
        kg, ks = kset
        assert np.allclose(kg.eval_unnorm_vec(chi2), np.exp(-0.5 * chi2), rtol=1e-15)
        assert np.allclose(ks.eval_unnorm_vec(chi2), (1.0 + chi2 / nu) ** (-0.5 * (nu + d)), rtol=1e-14)
        assert math.isclose(ks.eval_unnorm(1.3), (1.0 + 1.3 / nu) ** (-0.5 * (nu + d)), rel_tol=1e-15)
        for k, lnK in ((kg, -0.5 * chi2), (ks, -0.5 * (nu + d) * np.log1p(chi2 / nu))):
            lnt = lnK - lnn + np.log(w)
            im = int(np.argmax(lnt))
            lam = np.sum(np.exp(np.delete(lnt, im) - lnt[im]))
            g, l = k.eval_sum0_gamma_lambda(chi2, w, lnn)
            assert math.isclose(g, lnt[im], rel_tol=1e-14) and math.isclose(l, lam, rel_tol=1e-13)
            lnt1 = lnK + np.log(w)
            im = int(np.argmax(lnt1))
            lam = np.sum(np.exp(np.delete(lnt1, im) - lnt1[im]))
            g, l = k.eval_sum1_gamma_lambda(chi2, w, 0.7)
            assert math.isclose(g, lnt1[im] - 0.7, rel_tol=1e-14) and math.isclose(l, lam, rel_tol=1e-13)
    # stride-2 input (oracle API)
    kg = oracle.Kernel(oracle.KERNEL_GAUSS, d)
    assert np.allclose(kg.eval_unnorm_vec(chi2, stride=2), np.exp(-0.5 * chi2[::2]), rtol=1e-15)


def test_kernel_sample_moments(oracle):
    """test_ncm_stats_dist_kernel.c:430-495: sample mean within 20 %"""
    from numcosmo_b200 import stats_dist as S

    d = 3
    rng0 = oracle.RNG(3)
    cov = oracle.fill_rand_cov(d, 0.5, 1.0, 30.0, rng0)
    U = np.linalg.cholesky(cov).T
    mu = np.array([1.0, -2.0, 3.0])
    for ko, kh in zip(_kernels(oracle, d, 5.0), _host_kernels(d, 5.0)):
        r1, r2 = oracle.RNG(9), S.RNG(9)
        xs = np.array([ko.sample(U, 0.7, mu, r1) for _ in range(4000)])
        xh = np.array([kh.sample(U, 0.7, mu, r2) for _ in range(4000)])
        assert np.max(np.abs(xs - xh)) < 1e-12
        assert np.all(np.abs(xs.mean(axis=0) / mu - 1) < 0.2)


def test_oracle_matches_independent_numpy_restatement(oracle):
    """SURVEY.md Appendix C formulas evaluated with numpy/scipy vs the C oracle"""
    import scipy.linalg as sl

    d, n = 4, 300
    mu, cov, X, m2lnL = mvnd_problem(oracle, d, n, seed=8)
    for sdt in (oracle.SD_KDE, oracle.SD_VKDE):
        for kk, nu in ((oracle.KERNEL_GAUSS, 3.0), (oracle.KERNEL_ST, 3.0)):
            sd = oracle.StatsDist(sdt, kk, d, nu)
            sd.add_obs_matrix(X)
            assert sd.prepare_interp(m2lnL) == 0
            w, h = sd.peek_weights(), sd.get_href()
            Us = sd.peek_cov_array()
            lnn = np.array([sd.get_lnnorm(i) for i in range(n)])
            Q = X[:5] + 0.01
            got = sd.eval_m2lnp_batch(Q, 2)
            for x, g in zip(Q, got):
                lnt = np.empty(n)
                for i in range(n):
                    y = sl.solve_triangular(Us[i], x - X[i], trans="T", lower=False)
                    chi2 = y @ y / h**2
                    lnk = -0.5 * chi2 if kk == oracle.KERNEL_GAUSS else -0.5 * (nu + d) * np.log1p(chi2 / nu)
                    lnt[i] = lnk - lnn[i] + np.log(w[i])
                assert math.isclose(g, -2 * special.logsumexp(lnt), rel_tol=1e-12)
            # IM definition and the NNLS optimality conditions (KKT) of the weights before shrink
            IM = sd.peek_IM()
            f = np.exp(-0.5 * (m2lnL - m2lnL.min()))
            IMu = sd.compute_IM() / f[:, None]
            assert np.allclose(IM, IMu, rtol=1e-13)
            x, rnorm, st = oracle.nnls_solve(IM, np.ones(n))
            assert np.all(x >= 0)
            grad = IM.T @ (np.ones(n) - IM @ x)
            assert np.all(grad[x == 0] <= 1e-6 * max(1.0, np.abs(grad).max()))
            assert math.isclose(rnorm**2, sd.get_rnorm(), rel_tol=1e-10, abs_tol=1e-25)
            assert np.allclose((1 - 0.01) * x / x.sum() + 0.01 / n, w, rtol=1e-10, atol=1e-16)


def test_nnls_against_scipy_lawson_hanson(oracle):
    """different algorithm, same optimum (cross-check only)"""
    from scipy.optimize import nnls

    rs = np.random.default_rng(1)
    for m, n in [(60, 30), (100, 100), (200, 50)]:
        A = np.abs(rs.standard_normal((m, n)))
        f = A @ np.maximum(rs.standard_normal(n), 0) + 0.05 * rs.standard_normal(m)
        x, rnorm, st = oracle.nnls_solve(A, f)
        xs, rs_ = nnls(A, f)
        assert math.isclose(rnorm, rs_, rel_tol=1e-8)
        assert np.allclose(x, xs, atol=1e-7 * max(1.0, np.abs(xs).max()))


def test_gsl_subset_selection(oracle):
    import ctypes as C

    rs = np.random.default_rng(2)
    v = rs.standard_normal(50)
    L = oracle.lib()
    for k in (1, 5, 50):
        p = (C.c_int * k)()
        L.orc_sort_smallest_index(p, k, v.ctypes.data_as(C.POINTER(C.c_double)), 1, 50)
        assert list(p) == list(np.argsort(v, kind="stable")[:k])
        L.orc_sort_largest_index(p, k, v.ctypes.data_as(C.POINTER(C.c_double)), 1, 50)
        assert list(p) == list(np.argsort(-v, kind="stable")[:k])


def test_oracle_error_paths(oracle):
    sd = oracle.StatsDist(oracle.SD_KDE, oracle.KERNEL_GAUSS, 3)
    for i in range(3):
        sd.add_obs(np.arange(3.0) + i)
    assert sd.prepare() == -1        # "the sample is too small" (ncm_stats_dist.c:748-749)
    sd = oracle.StatsDist(oracle.SD_VKDE, oracle.KERNEL_GAUSS, 3)
    rs = np.random.default_rng(0)
    for i in range(30):
        sd.add_obs(rs.standard_normal(3))
    assert sd.prepare() == -4        # "Too few observations" (ncm_stats_dist_vkde.c:505-510)


def test_oracle_range_guard_heuristic(oracle):
    """ncm_stats_dist.c:906-946: too many points outside the 144.2 range -> 0.9 / 0.1 heuristic weights"""
    d, n = 2, 100
    mu, cov, X, m2lnL = mvnd_problem(oracle, d, n, seed=4)
    m = m2lnL.copy()
    m[: n - 10] += 1000.0    # only 10 points inside the range
    sd = oracle.StatsDist(oracle.SD_KDE, oracle.KERNEL_GAUSS, d)
    sd.add_obs_matrix(X)
    assert sd.prepare_interp(m) == 0
    w = sd.peek_weights()
    assert math.isclose(w.sum(), 1.0, rel_tol=1e-12)
    assert np.allclose(w[n - 10:], 0.9 / 10) and np.allclose(w[: n - 10], 0.1 / (n - 10))
