"""GPU parity of the VKDE prepare_kernel: kNN in whitened space,
local covariance of the raw neighbours in ascending-distance order, upper Cholesky factor — through the C ABI
(ncm_sd_gpu_vkde_prepare / ncm_sd_gpu_vkde_finish) against the CPU oracle
(_ncm_stats_dist_vkde_build_cov_array_kdtree, ncm_stats_dist_vkde.c:362-496)."""
import numpy as np
import pytest

from helpers import make_sd, mvnd_problem, rel_err

pytestmark = pytest.mark.gpu

# d, n, local_frac, kernel
CASES = [(2, 200, 0.05, "st"), (10, 600, 0.05, "gauss"), (10, 2048, 0.05, "gauss"), (20, 640, 0.1, "st"), (30, 700, 0.08, "gauss"), (3, 333, 0.5, "gauss")]


@pytest.mark.parametrize("d,n,local_frac,k_s", CASES)
def test_vkde_prepare_matches_oracle(oracle, gpu_ctx, d, n, local_frac, k_s):
    from numcosmo_b200 import capi

    kernel = oracle.KERNEL_GAUSS if k_s == "gauss" else oracle.KERNEL_ST
    mu, cov, X, m2lnL = mvnd_problem(oracle, d, n, seed=300 + d)
    sd = make_sd(oracle, oracle.SD_VKDE, kernel, 3.0, X, local_frac=local_frac)
    k = int(max(local_frac * n, 2.0))
    gpu_ctx.set_kernel(capi.KERNEL_GAUSS if k_s == "gauss" else capi.KERNEL_ST, 3.0, d)
    U, fail = gpu_ctx.vkde_prepare(X, sd.peek_invUsample(), n, k)
    assert not fail.any()
    U_ref = np.triu(sd.peek_cov_array())
    # factors: the oracle's dpotrf and the device's row-ordered Cholesky differ only in summation order
    scale = np.abs(U_ref).max(axis=(1, 2), keepdims=True)
    assert np.max(np.abs(np.triu(U) - U_ref) / scale) < 1e-11
    # finish with the oracle's lnnorms, then the evaluation must agree to the north-star tolerance
    gpu_ctx.vkde_finish(sd.peek_lnnorms())
    gpu_ctx.set_weights(sd.peek_weights(), sd.get_href())
    Q = np.vstack([X[:64] + 0.01, mu + 2.5 * (X[64:128] - mu)])
    assert rel_err(gpu_ctx.eval_m2lnp(Q), sd.eval_m2lnp_batch(Q, 4)) < 1e-10


def test_vkde_prepare_bit_identical_to_host_mirror(oracle):
    """The device prepare_kernel and the host mirror (NCM_B200_HOST_PREPARE_KERNEL path) issue the same IEEE
    operations in the same order: factors must be bit-identical, hence so is everything downstream."""
    from numcosmo_b200 import stats_dist as S

    d, n = 10, 700
    mu, cov, X, m2lnL = mvnd_problem(oracle, d, n, seed=77)
    outs = []
    for host in (False, True):
        S.lib().ncm_b200_set_host_prepare_kernel(int(host))
        sd = S.StatsDistVKDE(S.StatsDistKernelGauss(d), S.StatsDistCV.NONE)
        for x in X:
            sd.add_obs(x)
        sd.prepare_interp(m2lnL)
        outs.append((np.array([sd.peek_cov_decomp(i) for i in range(n)]), sd.peek_weights().copy(), sd.eval_m2lnp_array(X[:100] + 0.02)))
    S.lib().ncm_b200_set_host_prepare_kernel(0)
    assert np.array_equal(np.triu(outs[0][0]), np.triu(outs[1][0]))
    assert np.array_equal(outs[0][1], outs[1][1])
    assert np.array_equal(outs[0][2], outs[1][2])


def test_vkde_prepare_degenerate_falls_back(oracle):
    """k <= d neighbours give a singular local covariance: the device flags it, the host applies the reference's
    nearPD / diagonal fallback (ncm_stats_dist_kde.c:344-367) and the object stays usable."""
    from numcosmo_b200 import stats_dist as S

    d, n = 6, 60   # k = max(0.05 * 60, 2) = 3 < d
    mu, cov, X, m2lnL = mvnd_problem(oracle, d, n, seed=5)
    sd = S.StatsDistVKDE(S.StatsDistKernelGauss(d), S.StatsDistCV.NONE)
    for x in X:
        sd.add_obs(x)
    sd.prepare()
    out = sd.eval_m2lnp_array(X[:10])
    assert np.all(np.isfinite(out))
    # the host nearPD runs Higham's iteration over a cyclic-Jacobi eigen-solver where ncm_matrix_nearPD
    # (ncm_matrix.c:1248-1343) calls dsyevr.  A rank-k covariance (k = 3 of d = 6) has d - k eigenvalues that are zero up to rounding:
    # which of them come out negative, and hence the repaired values min_pos * eps along those directions, is rounding-driven in the
    # reference as well.  Stated bound: the repaired covariance U^T U -- everything the estimate determines -- agrees with the oracle's to
    # 1e-10 of its largest entry; the factor itself and ln |U| are compared only through that product.
    o = oracle.StatsDist(oracle.SD_VKDE, oracle.KERNEL_GAUSS, d, 3.0)
    o.add_obs_matrix(X)
    assert o.prepare() == 0
    worst = 0.0
    for i in range(n):
        Ug, Uo = np.triu(sd.peek_cov_decomp(i)), np.triu(o.peek_cov_decomp(i))
        Cg, Co = Ug.T @ Ug, Uo.T @ Uo
        worst = max(worst, np.max(np.abs(Cg - Co)) / np.abs(Co).max())
        assert np.all(np.isfinite(Ug)) and np.all(np.diag(Ug) > 0.0)
    print(f"nearPD (Jacobi) vs oracle (dsyevr), {n} rank-3 covariances in d = 6: max |U^T U - ref| / max |ref| = {worst:.2e}")
    assert worst < 1e-10


def test_vkde_prepare_beyond_shared_memory_capacity(oracle, gpu_ctx):
    """n_obs = 30000: one row of distances (240 KB) no longer fits in shared memory, the selection streams the
    precomputed distance matrix instead; same neighbours, same factors as the oracle on a sub-sample of the centres."""
    from numcosmo_b200 import capi

    d, n, k = 3, 30000, 300
    rs = np.random.default_rng(12)
    Z = rs.normal(size=(n, d))                       # already "whitened": raw = whitened here
    gpu_ctx.set_kernel(capi.KERNEL_GAUSS, 3.0, d)
    U, fail = gpu_ctx.vkde_prepare(Z, Z, n, k)
    assert not fail.any()
    # reference for a few centres: exact kNN by (distance, index), NcmStatsVec covariance, Cholesky
    for c in (0, 1, 14999, 29999):
        dist = ((Z - Z[c]) ** 2) @ np.ones(d)        # not bit-identical to the sequential sum; ties are measure-zero here
        order = np.lexsort((np.arange(n), dist))[:k]
        cov = np.cov(Z[order].T, bias=False)
        Uref = np.linalg.cholesky(cov).T
        assert np.max(np.abs(np.triu(U[c]) - Uref)) < 1e-10 * np.abs(Uref).max()
