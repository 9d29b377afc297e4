"""Cross-validation bandwidth selection on the GPU path against the CPU oracle.

ncm_stats_dist.c:484-701 (objectives + simplex driver), :703-789 (prepare), :1018-1072 (CV_SPLIT).  Every objective
evaluation is a GPU pass: a batched eval of the held-out points (CV_SPLIT_NOFIT), interpolation matrices (CV_LOO), or
IM + NNLS + eval of all observations (CV_SPLIT); n_obs > n_kernels exercises the rectangular IM / NNLS.  The optimiser
traces (every trial ln over_smooth and its objective) are compared, not just the end result.
"""
import numpy as np
import pytest

from helpers import assert_weights_parity, mvnd_problem, rel_err

pytestmark = pytest.mark.gpu

CASES = [("kde", "gauss", 3.0, 3), ("kde", "st", 3.0, 4), ("vkde", "gauss", 3.0, 5), ("vkde", "st", 1.0, 2), ("vkde", "st", 3.0, 10)]


def _mk(oracle, sd_s, k_s, nu, d, cv_name):
    from numcosmo_b200 import stats_dist as S

    kern = S.StatsDistKernelGauss(d) if k_s == "gauss" else S.StatsDistKernelST(d, nu)
    cv = getattr(S.StatsDistCV, cv_name)
    sd = S.StatsDistKDE(kern, cv) if sd_s == "kde" else S.StatsDistVKDE(kern, cv)
    o = oracle.StatsDist(oracle.SD_KDE if sd_s == "kde" else oracle.SD_VKDE, oracle.KERNEL_GAUSS if k_s == "gauss" else oracle.KERNEL_ST, d, nu,
                         getattr(oracle, "CV_" + cv_name))
    return sd, o


def _feed(sd, o, X):
    for x in X:
        sd.add_obs(x)
    o.add_obs_matrix(X)
    sd.set_use_threads(True)
    o.set_use_threads(True)


def _check_densities(sd, o, X, mu, bound):
    """bound: what assert_weights_parity returned -- the conditioning-limited bound, or None in the fallback regime (then the two sides
    are compared at the SAME weights: the evaluation proper)."""
    Q = np.vstack([X[:40] + 0.002, mu + 2.0 * (X[40:80] - mu)])
    if bound is None:
        o.set_weights(sd.peek_weights())
        bound = 1e-10
    assert rel_err(sd.eval_m2lnp_array(Q), o.eval_m2lnp_batch(Q, 4)) <= max(1e-10, bound)


@pytest.mark.parametrize("sd_s,k_s,nu,d", CASES)
def test_cv_split_nofit(oracle, sd_s, k_s, nu, d):
    n = 700
    mu, cov, X, m2lnL = mvnd_problem(oracle, d, n, seed=500 + d)
    sd, o = _mk(oracle, sd_s, k_s, nu, d, "SPLIT_NOFIT")
    sd.set_split_frac(0.6)
    o.set_split_frac(0.6)
    _feed(sd, o, X)
    sd.prepare_interp(m2lnL)
    assert o.prepare_interp(m2lnL) == 0
    assert sd.get_n_kernels() == o.get_n_kernels() == int(np.ceil(0.6 * n)) and sd.get_sample_size() == n
    (lg, vg), (lo, vo) = sd.cv_trace(), o.cv_trace()
    assert len(lg) == len(lo) and np.array_equal(lg, lo)          # identical trial bandwidths, step by step
    assert rel_err(vg, vo) < 1e-10                                # the objective: sum of held-out -2 ln p
    assert sd.get_over_smooth() == o.get_over_smooth() and abs(sd.get_href() / o.get_href() - 1) < 1e-14
    w, wo = sd.peek_weights(), o.peek_weights()
    assert abs(w.sum() - 1) < 1e-12
    st, so = sd.nnls_stats(), o.nnls_stats()
    bound = assert_weights_parity(w, wo, st, so, o.peek_IM(), what=f"{sd_s}-{k_s} d={d}")
    _check_densities(sd, o, X, mu, bound)


# the Monte-Carlo integral of p^2 needs ~ var(p) / (1e-4 mean(p)^2) draws per objective evaluation: minutes on the CPU oracle at d = 10
LOO_CASES = CASES[:4] + [("vkde", "st", 3.0, 4)]


@pytest.mark.parametrize("sd_s,k_s,nu,d", LOO_CASES)
def test_cv_loo(oracle, sd_s, k_s, nu, d):
    n = 300
    mu, cov, X, m2lnL = mvnd_problem(oracle, d, n, seed=520 + d)
    sd, o = _mk(oracle, sd_s, k_s, nu, d, "LOO")
    _feed(sd, o, X)
    sd.prepare_interp(m2lnL)
    assert o.prepare_interp(m2lnL) == 0
    assert sd.get_n_kernels() == o.get_n_kernels() == n
    (lg, vg), (lo, vo) = sd.cv_trace(), o.cv_trace()
    assert len(lg) == len(lo) and np.array_equal(lg, lo)
    # KDE-Gauss: closed form from two IMs; otherwise IM row sums + the Monte-Carlo mean of p over antithetic pairs drawn from
    # the fixed stream (seed 0): the same number of draws and the same points, so the objective agrees to rounding
    assert np.max(np.abs(vg - vo)) < 1e-9 * np.max(np.abs(vo))
    assert sd.get_over_smooth() == o.get_over_smooth()
    w, wo = sd.peek_weights(), o.peek_weights()
    st, so = sd.nnls_stats(), o.nnls_stats()
    bound = assert_weights_parity(w, wo, st, so, o.peek_IM(), what=f"{sd_s}-{k_s} d={d}")
    _check_densities(sd, o, X, mu, bound)
    # the reference leaves wcum as built (from uniform weights) by the first kernel_choose of the Monte-Carlo loop for every class /
    # kernel pair but KDE-Gauss: the next draws follow it -- both sides must agree on the kernel indices
    from numcosmo_b200 import stats_dist as S

    rg, ro = S.RNG(5), oracle.RNG(5)
    assert [sd.kernel_choose(rg) for _ in range(50)] == [o.kernel_choose(ro) for _ in range(50)]


@pytest.mark.parametrize("sd_s,k_s,nu,d", CASES)
def test_cv_split(oracle, sd_s, k_s, nu, d):
    n = 500
    mu, cov, X, m2lnL = mvnd_problem(oracle, d, n, seed=540 + d)
    sd, o = _mk(oracle, sd_s, k_s, nu, d, "SPLIT")
    _feed(sd, o, X)
    sd.prepare_interp(m2lnL)
    assert o.prepare_interp(m2lnL) == 0
    assert sd.get_n_kernels() == o.get_n_kernels() == n // 2
    (lg, vg), (lo, vo) = sd.cv_trace(), o.cv_trace()
    # 1 + 10 random tries (host MT19937 seeded 0, bit-identical stream), then the levmar evaluations
    assert np.array_equal(lg[:11], lo[:11])
    assert rel_err(vg[:11], vo[:11]) < 1e-7                       # NNLS rnorm of a rectangular (n_obs x n_kernels) system
    k = min(len(lg), len(lo))
    print(f"CV_SPLIT {sd_s}-{k_s} d={d}: {len(lg)} / {len(lo)} objective evaluations, max |d ln os| = {np.max(np.abs(lg[:k] - lo[:k])):.3e}, "
          f"over_smooth {sd.get_over_smooth():.12g} vs {o.get_over_smooth():.12g}")
    # the fit is a chain of NNLS solves on ill-conditioned Gram matrices: the bandwidth it lands on agrees to the conditioning of those
    # solves, not to rounding
    assert abs(sd.get_over_smooth() / o.get_over_smooth() - 1) < 1e-4
    assert abs(sd.get_rnorm() - o.get_rnorm()) <= 1e-3 * max(o.get_rnorm(), 1e-20)
    w = sd.peek_weights()
    assert abs(w.sum() - 1) < 1e-12 and w.min() > 0


def test_cv_split_repeat_is_deterministic_and_advances_the_object_rng(oracle):
    """Two identically fed objects give identical results (test_ncm_stats_dist.c:922-996); a second prepare_interp on the SAME object
    continues the object's RNG stream (ncm_stats_dist.c:178, :1047), on both sides alike."""
    d, n = 3, 300
    mu, cov, X, m2lnL = mvnd_problem(oracle, d, n, seed=77)
    a, oa = _mk(oracle, "vkde", "gauss", 3.0, d, "SPLIT")
    b, _ = _mk(oracle, "vkde", "gauss", 3.0, d, "SPLIT")
    _feed(a, oa, X)
    for x in X:
        b.add_obs(x)
    b.set_use_threads(True)
    a.prepare_interp(m2lnL)
    b.prepare_interp(m2lnL)
    assert oa.prepare_interp(m2lnL) == 0
    assert np.array_equal(a.cv_trace()[0], b.cv_trace()[0]) and np.array_equal(a.peek_weights(), b.peek_weights())
    first = a.cv_trace()[0].copy()
    a.set_over_smooth(1.0)
    oa.set_over_smooth(1.0)
    a.prepare_interp(m2lnL)
    assert oa.prepare_interp(m2lnL) == 0
    second, second_o = a.cv_trace()[0], oa.cv_trace()[0]
    assert not np.array_equal(first[:11], second[:11]) and np.array_equal(second[:11], second_o[:11])


def test_cv_split_nofit_dynamic_range_guard(oracle):
    """ADVICE r01 (stats_dist.cc:721): the dynamic-range guard (ncm_stats_dist.c:906-982) with n_obs > n_kernels.  40 % of the
    observations lie more than 4 |ln eps| above the minimum: the guard sorts ALL n_obs values, drops the out-of-range observations and
    recurses on the remaining ones (a smaller split), identically on both sides, without writing past the cut vector."""
    d, n = 3, 500
    mu, cov, X, m2lnL = mvnd_problem(oracle, d, n, seed=91)
    m2 = m2lnL.copy()
    out = np.random.default_rng(4).permutation(n)[: int(0.4 * n)]
    m2[out] += 400.0
    sd, o = _mk(oracle, "vkde", "gauss", 3.0, d, "SPLIT_NOFIT")
    sd.set_split_frac(0.7)
    o.set_split_frac(0.7)
    _feed(sd, o, X)
    sd.prepare_interp(m2)
    assert o.prepare_interp(m2) == 0
    n_in = n - len(out)
    assert sd.get_sample_size() == o.get_n_obs() == n_in                       # the sample array itself was cut (:967-972)
    assert sd.get_n_kernels() == o.get_n_kernels() == int(np.ceil(0.7 * n_in))
    w, wo = sd.peek_weights(), o.peek_weights()
    assert len(w) == len(wo) and abs(w.sum() - 1) < 1e-12
    bound = assert_weights_parity(w, wo, sd.nnls_stats(), o.nnls_stats(), o.peek_IM(), what="guard")
    _check_densities(sd, o, X, mu, bound)
    # fewer than half of the observations in range: the 90 % / 10 % weights of :934-946, no NNLS at all
    sd2, o2 = _mk(oracle, "vkde", "gauss", 3.0, d, "SPLIT_NOFIT")
    sd2.set_split_frac(0.7)
    o2.set_split_frac(0.7)
    _feed(sd2, o2, X)
    m3 = m2lnL.copy()
    m3[np.random.default_rng(5).permutation(n)[: int(0.7 * n)]] += 400.0
    sd2.prepare_interp(m3)
    assert o2.prepare_interp(m3) == 0
    assert np.array_equal(sd2.peek_weights(), o2.peek_weights())
