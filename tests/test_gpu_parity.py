"""GPU parity: the CUDA path through the C ABI against the CPU oracle on identical seeded inputs.

Tolerance (BASELINE.json north_star): relative 1e-10 on log-densities and on interpolation
weights; the weights bar is conditioning-limited (hard part b) and is
checked as stated in test_nnls_* below.
"""
import numpy as np
import pytest

from helpers import assert_weights_parity, make_sd, mvnd_problem, rel_err, upload_from_oracle

pytestmark = pytest.mark.gpu

TOL = 1e-10

CASES = [
    # sd_type, kernel, nu, d, n
    ("vkde", "gauss", 3.0, 10, 600),
    ("vkde", "st", 3.0, 10, 600),
    ("vkde", "st", 1.0, 2, 200),
    ("vkde", "gauss", 3.0, 30, 700),
    ("vkde", "st", 3.0, 20, 640),
    ("vkde", "gauss", 3.0, 3, 333),
    ("kde", "gauss", 3.0, 10, 600),
    ("kde", "st", 3.0, 10, 600),
    ("kde", "st", 1.0, 2, 200),
    ("kde", "gauss", 3.0, 30, 700),
    ("kde", "st", 3.0, 20, 641),
    ("kde", "gauss", 3.0, 6, 129),
]


def _types(oracle, sd_s, k_s):
    return (oracle.SD_KDE if sd_s == "kde" else oracle.SD_VKDE, oracle.KERNEL_GAUSS if k_s == "gauss" else oracle.KERNEL_ST)


@pytest.mark.parametrize("sd_s,k_s,nu,d,n", CASES)
def test_eval_m2lnp_parity(oracle, gpu_ctx, sd_s, k_s, nu, d, n):
    from numcosmo_b200 import capi

    sd_type, kernel = _types(oracle, sd_s, k_s)
    mu, cov, X, m2lnL = mvnd_problem(oracle, d, n, seed=100 + d)
    sd = make_sd(oracle, sd_type, kernel, nu, X)
    # Dirichlet-like weights with exact zeros to exercise ln 0
    rs = np.random.default_rng(7)
    w = rs.uniform(size=n)
    w[rs.uniform(size=n) < 0.1] = 0.0
    w /= w.sum()
    href = upload_from_oracle(gpu_ctx, capi, oracle, sd, sd_type, kernel, nu, X, weights=w)
    Q = np.vstack([X[:50] + 0.01, mu + 3.0 * (X[50:120] - mu), X[:7]])
    sd.set_weights(w)
    got = gpu_ctx.eval_m2lnp(Q)
    exp = sd.eval_m2lnp_batch(Q, 4)
    assert rel_err(got, exp) < TOL
    # uniform weights (what prepare() sets)
    w1 = np.full(n, 1.0 / n)
    sd.set_weights(w1)
    gpu_ctx.set_weights(w1, href)
    got = gpu_ctx.eval_m2lnp(Q)
    exp = sd.eval_m2lnp_batch(Q, 4)
    assert rel_err(got, exp) < TOL
    p = gpu_ctx.eval(Q)
    pe = sd.eval_batch(Q, 4)
    ok = pe > 1e-290
    assert rel_err(p[ok], pe[ok]) < 1e-9


@pytest.mark.parametrize("sd_s,k_s,nu,d,n", CASES)
def test_im_and_weights_parity(oracle, gpu_ctx, sd_s, k_s, nu, d, n):
    from numcosmo_b200 import capi

    sd_type, kernel = _types(oracle, sd_s, k_s)
    mu, cov, X, m2lnL = mvnd_problem(oracle, d, n, seed=200 + d)
    sd = make_sd(oracle, sd_type, kernel, nu, X, m2lnp=m2lnL)
    href = upload_from_oracle(gpu_ctx, capi, oracle, sd, sd_type, kernel, nu, X)
    f = np.exp(-0.5 * (m2lnL - m2lnL.min()))
    IM = gpu_ctx.compute_IM(1.0 / f, fetch=True, nrows=n)
    IM_o = sd.peek_IM()
    scale = np.abs(IM_o).max()
    assert np.max(np.abs(IM - IM_o)) / scale < 1e-12
    assert rel_err(IM[IM_o > 1e-200 * scale], IM_o[IM_o > 1e-200 * scale]) < 1e-9

    x, rnorm, st = gpu_ctx.nnls_solve()
    so = sd.nnls_stats()
    w_o = sd.peek_weights()
    shrink = 0.01
    w = (1.0 - shrink) * x / x.sum() + shrink / n
    # residual norm: rnorm^2 as ncm_stats_dist_get_rnorm returns it
    assert abs(rnorm**2 - sd.get_rnorm()) <= 1e-8 * max(sd.get_rnorm(), 1e-20) + 1e-18
    # weights: same passive set, no fallback on either side, 1e-10 of the largest weight wherever cond(M[P,P]) allows it
    bound = assert_weights_parity(w, w_o, st, so, IM_o, shrink, f"{sd_s}-{k_s} d={d}")
    # downstream densities with the GPU weights vs the oracle with its own weights
    gpu_ctx.set_weights(w, href)
    Q = np.vstack([X[:64] + 0.003, mu + 1.5 * (X[64:128] - mu)])
    got = gpu_ctx.eval_m2lnp(Q)
    exp = sd.eval_m2lnp_batch(Q, 4)
    assert bound is None or rel_err(got, exp) <= max(1e-10, bound)


def test_nnls_generic_parity(oracle, gpu_ctx):
    rs = np.random.default_rng(3)
    for (m, n) in [(300, 200), (257, 129), (64, 64), (500, 37)]:
        A = np.abs(rs.standard_normal((m, n))) + 0.1 * np.eye(m, n)
        xt = np.maximum(rs.standard_normal(n), 0.0)
        f = A @ xt + 0.01 * rs.standard_normal(m)
        x, rnorm, st = gpu_ctx.nnls_solve_host(A, f)
        xo, rno, so = oracle.nnls_solve(A, f)
        assert st["n_passive"] == so["n_passive"]
        assert np.max(np.abs(x - xo)) / np.abs(xo).max() < 1e-10
        assert abs(rnorm - rno) / rno < 1e-10
        assert np.all(x >= 0.0)


def test_sample_apply_parity(oracle, gpu_ctx):
    from numcosmo_b200 import capi

    d, n = 10, 400
    for kernel, nu in ((oracle.KERNEL_GAUSS, 3.0), (oracle.KERNEL_ST, 3.0)):
        mu, cov, X, m2lnL = mvnd_problem(oracle, d, n, seed=5)
        sd = make_sd(oracle, oracle.SD_VKDE, kernel, nu, X, m2lnp=m2lnL)
        href = upload_from_oracle(gpu_ctx, capi, oracle, sd, oracle.SD_VKDE, kernel, nu, X)
        # replay the oracle's stream on the host: index, normals, chi2 -- then only the affine map runs on the GPU
        r1, r2 = oracle.RNG(99), oracle.RNG(99)
        q = 300
        kidx, Z, S, Xo = [], [], [], []
        for _ in range(q):
            Xo.append(sd.sample(r1))
            kidx.append(sd.kernel_choose(r2))
            Z.append([r2.gaussian(1.0) for _ in range(d)])
            S.append(np.sqrt(nu / r2.chisq(nu)) if kernel == oracle.KERNEL_ST else 1.0)
        Xg = gpu_ctx.sample_apply(np.array(kidx), np.array(Z), np.array(S) if kernel == oracle.KERNEL_ST else None)
        assert np.max(np.abs(Xg - np.array(Xo))) / np.abs(np.array(Xo)).max() < 1e-13


def test_empty_and_errors(gpu_ctx):
    from numcosmo_b200 import capi

    ctx = capi.Context(0)
    with pytest.raises(capi.GpuError):
        ctx.eval_m2lnp(np.zeros((3, 2)))          # nothing uploaded
    with pytest.raises(capi.GpuError):
        ctx.set_kernel(capi.KERNEL_GAUSS, 1.0, 64)  # dimension out of range
    ctx.close()


@pytest.mark.parametrize("d,n,k_s", [(21, 640, "gauss"), (24, 700, "st"), (27, 555, "st"), (30, 700, "gauss"), (32, 650, "gauss")])
def test_vkde_tensor_core_path(oracle, gpu_ctx, d, n, k_s):
    """d >= 21 is served by the DMMA kernel (vkde_mma.cu, explicit inverses gated on the condition number):
    eval and IM against the oracle at the north-star tolerance, ragged sizes, exact-zero weights."""
    from numcosmo_b200 import capi

    kernel = oracle.KERNEL_GAUSS if k_s == "gauss" else oracle.KERNEL_ST
    mu, cov, X, m2lnL = mvnd_problem(oracle, d, n, seed=500 + d)
    sd = make_sd(oracle, oracle.SD_VKDE, kernel, 3.0, X, local_frac=0.15)
    rs = np.random.default_rng(d)
    w = rs.uniform(size=n)
    w[rs.uniform(size=n) < 0.1] = 0.0
    w /= w.sum()
    href = upload_from_oracle(gpu_ctx, capi, oracle, sd, oracle.SD_VKDE, kernel, 3.0, X, weights=w)
    uses_mma, cond = gpu_ctx.vkde_path()
    assert uses_mma and cond < 1e5, (uses_mma, cond)
    sd.set_weights(w)
    Q = np.vstack([X[:77] + 0.01, mu + 3.0 * (X[77:200] - mu), X[:5]])
    assert rel_err(gpu_ctx.eval_m2lnp(Q), sd.eval_m2lnp_batch(Q, 4)) < TOL
    IM = gpu_ctx.compute_IM(None, fetch=True, nrows=n)
    IM_ref = sd.compute_IM()
    scale = np.abs(IM_ref).max()
    assert np.max(np.abs(IM - IM_ref)) < 1e-12 * scale
    big = np.abs(IM_ref) > 1e-200 * scale
    assert np.max(np.abs(IM[big] - IM_ref[big]) / np.abs(IM_ref[big])) < 1e-9
