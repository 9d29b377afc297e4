"""Edge cases of the C ABI (include/ncm_sd_gpu.h) against the CPU oracle: empty and single-row batches, ragged sizes
around every tile boundary, the largest supported dimension, far-away and coincident query points, one-hot and
exact-zero weights, argument errors.  The reference exercises the same situations through
tests/c/ncm/stats/test_ncm_stats_dist.c (dimension / sample-size sweeps, :451-491; errors, :1119-1190)."""
import numpy as np
import pytest

from helpers import make_sd, mvnd_problem, rel_err, upload_from_oracle

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _mk(oracle, gpu_ctx, sd_s, k_s, d, n, nu=3.0, seed=0, weights=None, local_frac=0.05):
    from numcosmo_b200 import capi

    sd_type = oracle.SD_KDE if sd_s == "kde" else oracle.SD_VKDE
    kernel = oracle.KERNEL_GAUSS if k_s == "gauss" else oracle.KERNEL_ST
    mu, cov, X, m2lnL = mvnd_problem(oracle, d, n, seed=700 + seed + d)
    sd = make_sd(oracle, sd_type, kernel, nu, X, local_frac=local_frac)
    if weights is not None:
        sd.set_weights(weights)
    upload_from_oracle(gpu_ctx, capi, oracle, sd, sd_type, kernel, nu, X, weights=weights)
    return sd, mu, X


@pytest.mark.parametrize("sd_s", ["kde", "vkde"])
def test_empty_and_single_query(oracle, gpu_ctx, sd_s):
    sd, mu, X = _mk(oracle, gpu_ctx, sd_s, "gauss", 5, 150)
    out = gpu_ctx.eval_m2lnp(np.empty((0, 5)))
    assert out.shape == (0,)
    one = gpu_ctx.eval_m2lnp(X[3] + 0.01)
    assert rel_err(one, [sd.eval_m2lnp(X[3] + 0.01)]) < TOL
    dens = gpu_ctx.eval(X[:9])
    assert rel_err(dens, sd.eval_batch(X[:9], 1)) < TOL


@pytest.mark.parametrize("sd_s,k_s", [("kde", "gauss"), ("kde", "st"), ("vkde", "gauss"), ("vkde", "st")])
@pytest.mark.parametrize("n,q", [(33, 1), (63, 127), (64, 128), (65, 129), (127, 255), (257, 513)])
def test_ragged_sizes_around_tile_boundaries(oracle, gpu_ctx, sd_s, k_s, n, q):
    d = 3
    sd, mu, X = _mk(oracle, gpu_ctx, sd_s, k_s, d, n, seed=n, local_frac=0.2)
    rs = np.random.default_rng(n + q)
    Q = mu + 1.5 * rs.standard_normal((q, d)) * np.std(X, axis=0)
    assert rel_err(gpu_ctx.eval_m2lnp(Q), sd.eval_m2lnp_batch(Q, 2)) < TOL


@pytest.mark.parametrize("sd_s,k_s", [("kde", "gauss"), ("vkde", "gauss"), ("vkde", "st")])
def test_largest_dimension(oracle, gpu_ctx, sd_s, k_s):
    d, n = 32, 400
    sd, mu, X = _mk(oracle, gpu_ctx, sd_s, k_s, d, n, local_frac=0.2)
    Q = np.vstack([X[:40] + 1e-3, mu + 2.0 * (X[40:80] - mu)])
    assert rel_err(gpu_ctx.eval_m2lnp(Q), sd.eval_m2lnp_batch(Q, 2)) < TOL


@pytest.mark.parametrize("sd_s,d,n", [("kde", 4, 200), ("vkde", 4, 200), ("vkde", 24, 300)])
@pytest.mark.parametrize("nu", [1.0, 3.0, 2.5])
def test_student_t_linear_domain_and_its_repair_pass(oracle, gpu_ctx, sd_s, d, n, nu):
    """Student-t with integer nu: (1 + chi2/nu)^(-(nu + d)/2) as an integer power of rsqrt, summed in the linear domain (csrc/common.cuh
    st_pow_u / lin_push) -- against the oracle's log-sum-exp of kappa log1p (ncm_stats_dist_kernel_st.c:295-386).  A query so far away
    that every term underflows in the linear domain (chi2 ~ 1e150) must come out of the log-domain repair pass, finite and equal to the
    reference; nu = 2.5 (no integer power) keeps the log-domain path throughout.  d = 24: the tensor-core VKDE kernel."""
    sd, mu, X = _mk(oracle, gpu_ctx, sd_s, "st", d, n, nu=nu, local_frac=0.2)
    span = np.std(X, axis=0)
    e0 = np.zeros(d)
    e0[0] = 1.0
    Q = np.vstack([X[:50] + 1e-3 * span, mu + 3.0 * (X[50:100] - mu), mu + 1e3 * span])
    got, exp = gpu_ctx.eval_m2lnp(Q), sd.eval_m2lnp_batch(Q, 2)
    assert rel_err(got, exp) < TOL
    dens, dens_o = gpu_ctx.eval(Q[:100]), sd.eval_batch(Q[:100], 2)
    assert rel_err(dens[dens_o > 1e-290], dens_o[dens_o > 1e-290]) < 1e-9
    Qfar = np.vstack([Q[:7], mu + 1e75 * span * e0, mu - 1e60 * span])
    got, exp = gpu_ctx.eval_m2lnp(Qfar), sd.eval_m2lnp_batch(Qfar, 1)
    assert np.all(np.isfinite(got)), got
    assert rel_err(got, exp) < TOL
    # and the next well-behaved batch is served by the fast path again, same answers
    assert rel_err(gpu_ctx.eval_m2lnp(Q), sd.eval_m2lnp_batch(Q, 2)) < TOL


@pytest.mark.parametrize("sd_s,k_s", [("kde", "gauss"), ("kde", "st"), ("vkde", "gauss"), ("vkde", "st")])
def test_far_and_coincident_queries(oracle, gpu_ctx, sd_s, k_s):
    """Queries on top of a centre (chi2 = 0 for one pair) and hundreds of bandwidths away (every exp underflows
    relative to nothing: the log-sum-exp must stay finite and agree with the reference's gamma + log1p(lambda))."""
    d, n = 4, 200
    sd, mu, X = _mk(oracle, gpu_ctx, sd_s, k_s, d, n, local_frac=0.1)
    span = np.std(X, axis=0)
    Q = np.vstack([X[:5], mu + 60.0 * span, mu - 300.0 * span, mu + np.array([1e3, 0, 0, 0]) * span])
    got, exp = gpu_ctx.eval_m2lnp(Q), sd.eval_m2lnp_batch(Q, 1)
    assert np.all(np.isfinite(got))
    assert rel_err(got, exp) < TOL


def test_one_hot_and_zero_weights(oracle, gpu_ctx):
    d, n = 6, 129
    w = np.zeros(n)
    w[77] = 1.0
    for sd_s in ("kde", "vkde"):
        sd, mu, X = _mk(oracle, gpu_ctx, sd_s, "gauss", d, n, weights=w, local_frac=0.2)
        Q = X[70:90] + 0.01
        assert rel_err(gpu_ctx.eval_m2lnp(Q), sd.eval_m2lnp_batch(Q, 1)) < TOL


def test_argument_errors(gpu_ctx):
    from numcosmo_b200 import capi

    c = capi.Context(0)
    try:
        with pytest.raises(capi.GpuError):
            c.eval_m2lnp(np.zeros((3, 2)))            # nothing uploaded
        c.set_kernel(capi.KERNEL_GAUSS, 3.0, 2)
        with pytest.raises(capi.GpuError):
            c.upload_vkde(np.zeros((4, 2)), 5, np.zeros((5, 2, 2)), np.zeros(5))   # more kernels than observations
        with pytest.raises(capi.GpuError):
            c.set_kernel(capi.KERNEL_ST, 3.0, 33)     # beyond NCM_SD_GPU_MAX_DIM
        with pytest.raises(capi.GpuError):
            c.nnls_solve()                             # no interpolation matrix yet
    finally:
        c.close()
