"""ncm_sd_gpu_sample_philox: the counter-based throughput mode of the proposal draws.  It is NOT stream-compatible
with the reference generator (ncm_stats_dist.c:1565-1627 draws serially from one MT19937), so the test pins what such a mode has to
guarantee instead: the law of the draws (kernel index ~ weights; given the kernel, centre + href U_i^T z for Gauss and the same times
sqrt(nu / chi2_nu) for Student-t, ncm_stats_dist_kernel_gauss.c:335-355, ncm_stats_dist_kernel_st.c:388-414), and reproducibility per
(seed, offset, row)."""
import numpy as np
import pytest
from scipy import stats

from helpers import make_sd, mvnd_problem, upload_from_oracle

pytestmark = pytest.mark.gpu


def _ctx(oracle, gpu_ctx, kernel, nu, d, n, seed):
    from numcosmo_b200 import capi

    mu, cov, X, m2lnL = mvnd_problem(oracle, d, n, seed=seed)
    sd = make_sd(oracle, oracle.SD_VKDE, kernel, nu, X, m2lnp=m2lnL, local_frac=0.2)
    rs = np.random.default_rng(seed)
    w = rs.random(n) ** 3                         # very uneven weights, some exactly zero
    w[rs.integers(0, n, n // 10)] = 0.0
    w /= w.sum()
    href = upload_from_oracle(gpu_ctx, capi, oracle, sd, oracle.SD_VKDE, kernel, nu, X, weights=w)
    U = np.triu(sd.peek_cov_array())              # [n, d, d] upper factors, cov_i = U_i^T U_i
    return X, U.reshape(n, d, d), w, href


def test_philox_reproducible_per_seed_offset_row(oracle, gpu_ctx):
    X, U, w, href = _ctx(oracle, gpu_ctx, oracle.KERNEL_GAUSS, 3.0, 4, 200, seed=11)
    a, ka = gpu_ctx.sample_philox(5000, seed=42, offset=0)
    b, kb = gpu_ctx.sample_philox(5000, seed=42, offset=0)
    assert np.array_equal(a, b) and np.array_equal(ka, kb)
    # a row's stream depends on (seed, row, offset) only: a shorter batch is a prefix of the longer one
    c, kc = gpu_ctx.sample_philox(1234, seed=42, offset=0)
    assert np.array_equal(c, a[:1234]) and np.array_equal(kc, ka[:1234])
    for seed, offset in ((43, 0), (42, 1)):
        e, ke = gpu_ctx.sample_philox(5000, seed=seed, offset=offset)
        assert np.mean(ke == ka) < 0.2 and not np.any(np.all(e == a, axis=1))
    assert np.all(np.isfinite(a))


@pytest.mark.parametrize("kernel_s,nu,d", [("gauss", 3.0, 3), ("st", 3.0, 4), ("st", 1.0, 2)])
def test_philox_law(oracle, gpu_ctx, kernel_s, nu, d):
    kernel = oracle.KERNEL_GAUSS if kernel_s == "gauss" else oracle.KERNEL_ST
    n, q = 64, 400000
    X, U, w, href = _ctx(oracle, gpu_ctx, kernel, nu, d, n, seed=20 + d)
    S, k = gpu_ctx.sample_philox(q, seed=7, offset=3)
    assert k.min() >= 0 and k.max() < n
    # kernel choice ~ weights (chi-square goodness of fit over the kernels that can be drawn; zero-weight kernels never are)
    cnt = np.bincount(k, minlength=n)
    assert np.all(cnt[w == 0.0] == 0)
    pos = w > 0
    chi2 = np.sum((cnt[pos] - q * w[pos]) ** 2 / (q * w[pos]))
    assert stats.chi2.sf(chi2, pos.sum() - 1) > 1e-4, chi2
    # whiten every draw with its own kernel: y = U_i^-T (x - c_i) / href  ->  z (Gauss)  or  z sqrt(nu / chi2_nu) (Student-t)
    Y = np.empty_like(S)
    for i in np.flatnonzero(cnt):
        sel = k == i
        Y[sel] = np.linalg.solve(U[i].T, (S[sel] - X[i]).T).T / href
    r2 = np.einsum("ij,ij->i", Y, Y)
    if kernel_s == "gauss":
        assert stats.kstest(r2, stats.chi2(d).cdf).pvalue > 1e-4
        assert stats.kstest(Y[:, 0], stats.norm.cdf).pvalue > 1e-4 and stats.kstest(Y[:, d - 1], stats.norm.cdf).pvalue > 1e-4
        C = Y.T @ Y / q
        assert np.max(np.abs(C - np.eye(d))) < 6.0 / np.sqrt(q) * 2
    else:
        # |y|^2 / d ~ F(d, nu); a single coordinate ~ Student-t(nu)
        assert stats.kstest(r2 / d, stats.f(d, nu).cdf).pvalue > 1e-4
        assert stats.kstest(Y[:, 0], stats.t(nu).cdf).pvalue > 1e-4
        # directions are isotropic: the sign pattern and the normalised first coordinate do not depend on the radial scale
        u0 = Y[:, 0] / np.sqrt(r2)
        assert abs(np.mean(u0)) < 5.0 / np.sqrt(q * d) and abs(np.mean(u0**2) - 1.0 / d) < 5.0 / np.sqrt(q)


def test_philox_bad_arguments(oracle, gpu_ctx):
    from numcosmo_b200 import capi

    X, U, w, href = _ctx(oracle, gpu_ctx, oracle.KERNEL_GAUSS, 3.0, 3, 50, seed=5)
    with pytest.raises(capi.GpuError):
        gpu_ctx.sample_philox(0, seed=1)
