"""Shared synthetic inputs for the parity tests (seeded; the oracle's MT19937 restatement draws the
covariances so that the CPU oracle and the GPU path read identical bytes)."""
import numpy as np


def mvnd_problem(O, d, n, seed, sigma=(2e-2, 5e-2), cor_level=30.0, mu_range=(1.0, 2.0)):
    """Random MVND target as tests/c/ncm/fit/test_ncm_fit_esmcmc.c:134 builds it, plus n draws from it."""
    rng = O.RNG(seed)
    cov = O.fill_rand_cov(d, sigma[0], sigma[1], cor_level, rng)
    mu = np.array([rng.flat(*mu_range) for _ in range(d)])
    L = np.linalg.cholesky(cov)
    z = np.array([[rng.gaussian(1.0) for _ in range(d)] for _ in range(n)])
    X = mu + z @ L.T
    m2lnL = np.einsum("ij,ij->i", z, z)
    return mu, cov, np.ascontiguousarray(X), m2lnL


def make_sd(O, sd_type, kernel, nu, X, m2lnp=None, over_smooth=1.0, local_frac=0.05, use_threads=True):
    sd = O.StatsDist(sd_type, kernel, X.shape[1], nu)
    sd.set_over_smooth(over_smooth)
    sd.set_local_frac(local_frac)
    sd.set_use_threads(use_threads)
    sd.add_obs_matrix(X)
    rc = sd.prepare_interp(m2lnp) if m2lnp is not None else sd.prepare()
    assert rc == 0, rc
    return sd


def upload_from_oracle(ctx, capi, O, sd, sd_type, kernel, nu, X, weights=None):
    """Feed the GPU context with the prepare_kernel products of the oracle object."""
    d = X.shape[1]
    n = sd.get_n_kernels()
    ctx.set_kernel(capi.KERNEL_GAUSS if kernel == O.KERNEL_GAUSS else capi.KERNEL_ST, nu, d)
    href = sd.get_href()
    if sd_type == O.SD_KDE:
        lnnorm = sd.get_lnnorm(0) - d * np.log(href)
        ctx.upload_kde(sd.peek_invUsample(), n, sd.peek_full_cov_decomp(), lnnorm)
    else:
        ctx.upload_vkde(X, n, np.triu(sd.peek_cov_array()), sd.peek_lnnorms())
    ctx.set_weights(sd.peek_weights() if weights is None else weights, href)
    return href


def rel_err(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))
