"""Robust covariance types through the GPU path against the CPU oracle: the estimators run on the host
(numcosmo_b200/host/robust.cc), everything downstream of the factors -- upload, interpolation matrix, NNLS, batched evaluation,
sampling -- is the same GPU path as for the sample covariance (ncm_stats_dist_kde.c:423-441, ncm_stats_dist_vkde.c:467-472)."""
import numpy as np
import pytest

from helpers import assert_weights_parity, mvnd_problem, rel_err, support

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("sd_s,k_s,nu,d,n", [("kde", "gauss", 3.0, 3, 500), ("kde", "st", 3.0, 6, 600), ("vkde", "gauss", 3.0, 4, 400), ("vkde", "st", 1.0, 5, 600)])
@pytest.mark.parametrize("cov_type", ["ROBUST_DIAG", "ROBUST"])
def test_robust_cov_prepare_interp_eval(oracle, sd_s, k_s, nu, d, n, cov_type):
    from numcosmo_b200 import stats_dist as S

    mu, cov, X, m2lnL = mvnd_problem(oracle, d, n, seed=600 + d)
    X = X.copy()
    X[::25] += 20.0 * np.sqrt(np.diag(cov)) * np.random.default_rng(d).standard_normal((len(X[::25]), d))   # 4 % outliers
    kern = S.StatsDistKernelGauss(d) if k_s == "gauss" else S.StatsDistKernelST(d, nu)
    sd = (S.StatsDistKDE if sd_s == "kde" else S.StatsDistVKDE)(kern, S.StatsDistCV.NONE)
    o = oracle.StatsDist(oracle.SD_KDE if sd_s == "kde" else oracle.SD_VKDE, oracle.KERNEL_GAUSS if k_s == "gauss" else oracle.KERNEL_ST, d, nu)
    sd.set_cov_type(getattr(S.StatsDistKDECovType, cov_type))
    o.set_cov_type(getattr(oracle, "COV_" + cov_type))
    for x in X:
        sd.add_obs(x)
    o.add_obs_matrix(X)
    sd.set_use_threads(True)
    o.set_use_threads(True)
    if sd_s == "vkde":
        sd.set_local_frac(0.1)
        o.set_local_frac(0.1)
    m2 = np.einsum("ij,jk,ik->i", X - mu, np.linalg.inv(cov), X - mu)
    m2 = np.minimum(m2, m2.min() + 100.0)          # keep the outliers inside the dynamic-range guard
    sd.prepare_interp(m2)
    assert o.prepare_interp(m2) == 0
    C, Co = sd.peek_full_cov(), o.peek_full_cov()
    # ROBUST_DIAG is an order statistic: bit-identical; OGK goes through an eigen-decomposition (Jacobi here, dsyevr in the reference)
    if cov_type == "ROBUST_DIAG":
        assert np.array_equal(np.triu(C), np.triu(Co))
    else:
        assert np.max(np.abs(np.triu(C) - np.triu(Co))) < 1e-10 * np.abs(Co).max()
    assert np.max(np.abs(np.triu(sd.peek_full_cov_decomp()) - np.triu(o.peek_full_cov_decomp()))) < 1e-10 * np.abs(o.peek_full_cov_decomp()).max()
    for i in (0, n // 3, n - 1):
        Ui, Uo = np.triu(sd.peek_cov_decomp(i)), np.triu(o.peek_cov_decomp(i))
        assert np.max(np.abs(Ui - Uo)) < 1e-9 * np.abs(Uo).max()
        assert abs(sd.get_lnnorm(i) - o.get_lnnorm(i)) < 1e-8
    w, wo = sd.peek_weights(), o.peek_weights()
    st, so = sd.nnls_stats(), o.nnls_stats()
    Q = np.vstack([X[1:40] + 0.002, mu + 2.0 * (X[41:80] - mu)])
    # the factors differ at the 1e-10 level (OGK through Jacobi vs dsyevr), so the two NNLS problems are not bit-identical inputs:
    # same passive set required, weights to the conditioning-limited bound plus the propagated factor difference
    if cov_type == "ROBUST_DIAG":
        bound = assert_weights_parity(w, wo, st, so, o.peek_IM(), what=f"{sd_s}-{k_s} d={d} {cov_type}", rnorm2=(sd.get_rnorm(), o.get_rnorm()))
        if bound is not None:
            assert rel_err(sd.eval_m2lnp_array(Q), o.eval_m2lnp_batch(Q, 4)) <= max(1e-10, bound)
    elif so["n_lu"] == 0:
        assert st["n_lu"] == 0 and np.array_equal(support(w, n), support(wo, n)), (st, so)
        assert np.max(np.abs(w - wo)) / wo.max() < 1e-6
    # with the oracle's weights on both sides the densities agree to the kernel-evaluation bar
    o.set_weights(w)
    assert rel_err(sd.eval_m2lnp_array(Q), o.eval_m2lnp_batch(Q, 4)) < 1e-8
    # proposals from the robust factors: same stream, same points
    rg, ro = S.RNG(3), oracle.RNG(3)
    for _ in range(20):
        assert np.max(np.abs(sd.sample(rg) - o.sample(ro))) < 1e-9 * np.abs(X).max()


def test_robust_cov_too_few_neighbours_is_an_error():
    from numcosmo_b200 import stats_dist as S

    d = 2
    X = np.random.default_rng(0).standard_normal((40, d))
    sd = S.StatsDistVKDE(S.StatsDistKernelGauss(d), S.StatsDistCV.NONE)
    sd.set_cov_type(S.StatsDistKDECovType.ROBUST_DIAG)
    sd.set_local_frac(0.06)
    for x in X:
        sd.add_obs(x)
    with pytest.raises(Exception, match="too few points"):
        sd.prepare()
