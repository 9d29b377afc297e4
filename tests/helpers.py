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


EPS = np.finfo(float).eps


def support(w, n, shrink=0.01):
    """Passive set of the NNLS solution behind normalised + shrunk weights: w_i = (1 - s) x_i / sum x + s / n > s / n  <=>  x_i > 0
    (ncm_stats_dist.c:1087-1093)."""
    return w > (shrink / n) * (1.0 + 1e-9)


def weight_bound(IM, passive):
    """Conditioning-limited bound on the relative (to the largest) error of NNLS weights obtained through the normal
    equations: the forward error of a backward-stable solve of M[P,P] x = b[P] is ~ cond_2(M[P,P]) eps, for each of the two
    implementations (CPU dposv and the device factorisation).  cond_2 = lambda_max / lambda_min by power / inverse iteration
    on a Cholesky factor (an SVD of a 16384-column matrix would take longer than the test).  Returns (cond, bound): the
    north-star bar (1e-10) wherever the conditioning allows it, never looser than 1e-6 of the largest weight."""
    import scipy.linalg as sl

    A = np.ascontiguousarray(IM[:, passive])
    M = A.T @ A
    try:
        c = sl.cho_factor(M, lower=True, check_finite=False)
    except np.linalg.LinAlgError:
        return np.inf, 1e-6
    rs = np.random.default_rng(0)
    v = rs.standard_normal(M.shape[0])
    u = v.copy()
    lmax = lmin_inv = 1.0
    for _ in range(30):
        v = M @ (v / np.linalg.norm(v))
        lmax = np.linalg.norm(v)
        u = sl.cho_solve(c, u / np.linalg.norm(u), check_finite=False)
        lmin_inv = np.linalg.norm(u)
    cond_M = lmax * lmin_inv
    return cond_M, max(1e-10, min(4.0 * cond_M * EPS, 1e-6))


def assert_weights_parity(w, wo, st, so, IM, shrink=0.01, what="", rnorm2=None):
    """Unconditional parity of the interpolation weights.

    Positive-definite regime (no passive-set system left dposv on either side): the two NNLS runs end on the SAME passive set and the
    weights agree to the conditioning-limited bound of that set's normal matrix; returns that bound.

    Fallback regime (the oracle, i.e. the reference algorithm, itself had to leave dposv for dsysv / dgels, ncm_nnls.c:573-638): the
    systems are singular to working precision, the solution vector is rounding-driven on the CPU too and no two LAPACK builds agree on
    it.  What is determined is the interpolation itself -- the fitted values IM w at the centres and the residual norm: those are
    asserted, the GPU path must have taken the same kind of fallback, and None is returned (callers then compare the evaluation at
    fixed weights instead of the weights)."""
    n = len(wo)
    fallback = so["n_lu"] > 0 or so["n_qr"] > 0
    if not fallback:
        assert st["n_lu"] == 0 and st["n_qr"] == 0, f"{what}: the GPU path left dposv where the oracle did not ({st} vs {so})"
        pg, po = support(w, n, shrink), support(wo, n, shrink)
        assert np.array_equal(pg, po), f"{what}: passive sets differ in {np.count_nonzero(pg != po)} of {n} indices ({st} vs {so})"
        # (a passive x_i that is tiny against sum x leaves w_i == shrink / n in floating point: the support seen through the weights can
        # be smaller than the passive set, never larger)
        assert st["n_passive"] == so["n_passive"] and int(po.sum()) <= so["n_passive"], (what, st, so)
        cond_M, bound = weight_bound(IM, po)
        err = np.max(np.abs(w - wo)) / wo.max()
        assert err <= bound, f"{what}: weights differ by {err:.2e} of the largest (bound {bound:.2e}, cond(M[P,P]) = {cond_M:.2e})"
        return bound
    assert st["n_lu"] > 0, f"{what}: the oracle took the dsysv fallback {so['n_lu']} times, the GPU path never ({st})"
    assert np.all(np.isfinite(w)) and abs(w.sum() - 1.0) < 1e-12 and w.min() >= (shrink / n) * (1.0 - 1e-12)
    fit_g, fit_o = IM @ w, IM @ wo
    dfit = np.linalg.norm(fit_g - fit_o) / np.linalg.norm(fit_o)
    print(f"{what}: fallback regime (oracle {so['n_lu']} lu / {so['n_qr']} qr of {so['n_chol']}, gpu {st['n_lu']} / {st['n_qr']} of {st['n_chol']}): "
          f"|P| {st['n_passive']} vs {so['n_passive']}, fitted values differ by {dfit:.2e}" + (f", rnorm^2 {rnorm2[0]:.6e} vs {rnorm2[1]:.6e}" if rnorm2 else ""))
    assert dfit <= 1e-2, f"{what}: fitted values at the centres differ by {dfit:.2e}"
    if rnorm2 is not None:   # the residual norm of the NNLS optimum is unique
        assert abs(np.sqrt(rnorm2[0]) - np.sqrt(rnorm2[1])) <= 1e-6 * np.sqrt(IM.shape[0]), (what, rnorm2)
    return None
