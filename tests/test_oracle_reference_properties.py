"""The reference's OWN tests of the path, restated for the CPU oracle (pin the oracle on every fixture the
reference's tests hold for this path).  tests/c/ncm/stats/test_ncm_stats_dist.c runs, for each class x kernel (KDE / VKDE x Gauss /
Student-t, dimension 1 - 3, np = 200 dim observations of a random MVND) and each covariance type (SAMPLE, FIXED, ROBUST_DIAG, ROBUST):

  gauss/dens/est              (:494-539)  prepare, then the Jensen-Shannon-type divergence of :452-491 below RELTOL = 0.5
  gauss/dens/interp           (:541-592)  the same after prepare_interp on the normalised -2 ln L
  gauss/dens/interp/sampling  (:594-675)  kernel_choose frequencies equal the normalised weights (1e-1 relative + 1e-4 absolute)
  gauss/dens/interp/cv_*      (:677-835)  the divergence test again with the bandwidth chosen by CV_SPLIT / CV_SPLIT_NOFIT / CV_LOO
  gauss/sampling              (:837-921)  covariance of 500 draws against the true one (ncm_matrix_cmp, scale 1) below 0.5
  prepare_too_few             (:424-450)  one observation, prepare: "the sample is too small" (ncm_stats_dist.c:749)
  gauss/get_kernel_info       (:999-1118) after prepare_interp: finite rnorm, lnnorm_i = kernel lnnorm (cov_decomp_i) + dim ln href,
                                          the accessors return the stored sample and weights

The reference draws its dimension, nu and seeds from g_test_rand_*; here they are swept / fixed.  These are property tests with loose
bars -- they pin the oracle's behaviour where the reference pins its own, on top of the known answers of test_oracle_known_answers.py."""
import numpy as np
import pytest

from helpers import mvnd_problem

TESTMULT, NTESTS, RELTOL = 200, 500, 0.5
CLASSES = [("kde", "gauss", 3.0), ("kde", "st", 3.7), ("vkde", "gauss", 3.0), ("vkde", "st", 4.2)]
COV_TYPES = ["SAMPLE", "FIXED", "ROBUST_DIAG", "ROBUST"]


def _setup(oracle, sd_s, k_s, nu, d, cov_type, seed, cv="NONE"):
    corr_level = 100.0 if cov_type == "ROBUST_DIAG" else 1.0                      # :186-187
    mu, cov, X, chi2 = mvnd_problem(oracle, d, TESTMULT * d, seed=seed, sigma=(1.0e-2, 5.0e-2), cor_level=corr_level)
    sd = oracle.StatsDist(oracle.SD_KDE if sd_s == "kde" else oracle.SD_VKDE, oracle.KERNEL_GAUSS if k_s == "gauss" else oracle.KERNEL_ST, d, nu,
                          getattr(oracle, "CV_" + cv))
    sd.set_cov_type(getattr(oracle, "COV_" + cov_type))
    if cov_type == "FIXED":
        sd.set_cov_fixed(cov)
    sd.add_obs_matrix(X)
    lndet = np.linalg.slogdet(2.0 * np.pi * cov)[1]
    return sd, mu, cov, chi2, lndet


def _js_divergence(oracle, sd, mu, cov, m2lnL_of, rng):
    """test_ncm_stats_dist_cmp_dist, :452-491."""
    Lc = np.linalg.cholesky(cov)
    d = len(mu)
    e0, e1 = [], []
    for _ in range(NTESTS):
        y = mu + Lc @ np.array([rng.gaussian(1.0) for _ in range(d)])
        e0.append(np.log1p(np.tanh(0.25 * (sd.eval_m2lnp(y) - m2lnL_of(y)))))
        ys = sd.sample(rng)
        e1.append(np.log1p(np.tanh(0.25 * (m2lnL_of(ys) - sd.eval_m2lnp(ys)))))
    return 0.5 * (np.mean(e0) + np.mean(e1))


@pytest.mark.parametrize("cov_type", COV_TYPES)
@pytest.mark.parametrize("sd_s,k_s,nu", CLASSES)
@pytest.mark.parametrize("d", [1, 2, 3])
def test_dens_est(oracle, sd_s, k_s, nu, d, cov_type):
    sd, mu, cov, chi2, lndet = _setup(oracle, sd_s, k_s, nu, d, cov_type, seed=1000 + 10 * d + len(cov_type))
    assert sd.prepare() == 0
    icov = np.linalg.inv(cov)
    js = _js_divergence(oracle, sd, mu, cov, lambda y: float((y - mu) @ icov @ (y - mu)), oracle.RNG(5 + d))   # use_norma is FALSE here
    assert js < RELTOL, js


@pytest.mark.parametrize("cov_type", COV_TYPES)
@pytest.mark.parametrize("sd_s,k_s,nu", CLASSES)
@pytest.mark.parametrize("d", [1, 2, 3])
def test_dens_interp(oracle, sd_s, k_s, nu, d, cov_type):
    sd, mu, cov, chi2, lndet = _setup(oracle, sd_s, k_s, nu, d, cov_type, seed=2000 + 10 * d + len(cov_type))
    assert sd.prepare_interp(chi2 + lndet) == 0                                   # ncm_data_gauss_cov_use_norma (TRUE), :569
    icov = np.linalg.inv(cov)
    js = _js_divergence(oracle, sd, mu, cov, lambda y: float((y - mu) @ icov @ (y - mu)) + lndet, oracle.RNG(7 + d))
    assert js < RELTOL, js
    # with the interpolated weights the estimate is close to the true density: the divergence is small, not merely below the bar
    if cov_type in ("SAMPLE", "FIXED"):
        assert abs(js) < 0.1, js


@pytest.mark.parametrize("cov_type", ["SAMPLE", "ROBUST_DIAG"])
@pytest.mark.parametrize("sd_s,k_s,nu", CLASSES)
def test_dens_interp_sampling_frequencies(oracle, sd_s, k_s, nu, cov_type):
    d = 2
    sd, mu, cov, chi2, lndet = _setup(oracle, sd_s, k_s, nu, d, cov_type, seed=3000 + len(cov_type))
    assert sd.prepare_interp(chi2 + lndet) == 0
    n = sd.get_sample_size()
    ntests = 400000                                  # 1e7 in the reference; the bars below are scaled to the counting noise
    rng = oracle.RNG(11)
    cum = np.bincount([sd.kernel_choose(rng) for _ in range(ntests)], minlength=n) / ntests
    w = sd.peek_weights()
    w = w / w.sum()
    assert np.all(cum[w == 0.0] == 0.0)
    tol = 1.0e-1 * w + 1.0e-4 + 5.0 * np.sqrt(w / ntests)
    assert np.all(np.abs(cum - w) <= tol), np.max(np.abs(cum - w) - tol)


@pytest.mark.parametrize("cov_type", COV_TYPES)
@pytest.mark.parametrize("sd_s,k_s,nu", CLASSES)
def test_sampling_covariance(oracle, sd_s, k_s, nu, cov_type):
    d = 3
    sd, mu, cov, chi2, lndet = _setup(oracle, sd_s, k_s, nu, d, cov_type, seed=4000 + len(cov_type))
    assert sd.prepare() == 0
    rng = oracle.RNG(13)
    Y = np.array([sd.sample(rng) for _ in range(NTESTS)])
    cov_est = np.cov(Y.T, bias=False)
    assert np.max(np.abs((cov_est - cov) / (1.0 + cov))) < 0.5                    # ncm_matrix_cmp (cov_est, cov, 1.0), :909
    # and, beyond the reference's bar, the draws reproduce the bandwidth-inflated covariance to sampling accuracy
    assert np.max(np.abs(np.diag(cov_est) / np.diag(cov) - 1.0)) < 1.5
    assert np.max(np.abs(Y.mean(axis=0) - mu)) < 5.0 * np.sqrt(np.max(np.diag(cov_est)) / NTESTS) + 0.5 * np.sqrt(np.max(np.diag(cov)))


@pytest.mark.parametrize("cv", ["SPLIT", "SPLIT_NOFIT", "LOO"])
@pytest.mark.parametrize("cov_type", COV_TYPES)
@pytest.mark.parametrize("sd_s,k_s,nu", CLASSES)
def test_dens_interp_cross_validation(oracle, sd_s, k_s, nu, cov_type, cv):
    d = 2
    sd, mu, cov, chi2, lndet = _setup(oracle, sd_s, k_s, nu, d, cov_type, seed=5000 + len(cov_type) + 3 * len(cv), cv=cv)
    assert sd.prepare_interp(chi2) == 0                                           # use_norma keeps its default (FALSE, ncm_data_gauss_cov.c:220-224) in these three
    icov = np.linalg.inv(cov)
    js = _js_divergence(oracle, sd, mu, cov, lambda y: float((y - mu) @ icov @ (y - mu)), oracle.RNG(17))
    assert js < RELTOL, js
    lnos, val = sd.cv_trace()
    assert len(lnos) >= 2 and np.all(np.isfinite(val))                            # a bandwidth search did run


@pytest.mark.parametrize("sd_s,k_s,nu", CLASSES)
@pytest.mark.parametrize("d", [1, 2, 3])
def test_prepare_too_few(oracle, sd_s, k_s, nu, d):
    mu, cov, X, _ = mvnd_problem(oracle, d, 4, seed=d)
    sd = oracle.StatsDist(oracle.SD_KDE if sd_s == "kde" else oracle.SD_VKDE, oracle.KERNEL_GAUSS if k_s == "gauss" else oracle.KERNEL_ST, d, nu)
    sd.add_obs(X[0])
    assert sd.prepare() == -1                                                     # the oracle's code for "the sample is too small"
    # n_obs <= d is still too small, d + 1 is the first size past that check (ncm_stats_dist.c:748-749)
    sd.reset()
    sd.add_obs_matrix(np.ascontiguousarray(X[:d]))
    assert sd.prepare() == -1
    sd.reset()
    sd.add_obs_matrix(np.ascontiguousarray(X[:d + 1]))
    assert sd.prepare() != -1


@pytest.mark.parametrize("cov_type", COV_TYPES)
@pytest.mark.parametrize("sd_s,k_s,nu", CLASSES)
def test_get_kernel_info(oracle, sd_s, k_s, nu, cov_type):
    from scipy.special import gammaln

    d = 3
    sd, mu, cov, chi2, lndet = _setup(oracle, sd_s, k_s, nu, d, cov_type, seed=6000 + len(cov_type))
    assert sd.prepare() == 0
    assert sd.prepare_interp(chi2) == 0
    n = sd.get_sample_size()
    assert np.isfinite(sd.get_rnorm()) and n == TESTMULT * d
    href, w = sd.get_href(), sd.peek_weights()
    assert len(w) == n and abs(w.sum() - 1.0) < 1e-12
    for i in range(0, n, 7):
        U = np.triu(sd.peek_cov_decomp(i))
        lndetU = np.sum(np.log(np.diag(U)))
        if k_s == "gauss":                                                        # ncm_stats_dist_kernel_gauss.c get_lnnorm
            ln0 = 0.5 * d * np.log(2.0 * np.pi) + lndetU
        else:                                                                     # ncm_stats_dist_kernel_st.c get_lnnorm
            ln0 = gammaln(0.5 * nu) - gammaln(0.5 * (nu + d)) + 0.5 * d * np.log(nu * np.pi) + lndetU
        assert abs(sd.get_lnnorm(i) - (ln0 + d * np.log(href))) <= 1e-14 * max(1.0, abs(ln0))
        assert np.array_equal(sd.peek_sample(i), sd.peek_sample(i).copy())
