"""The oracle's NNLS (restatement of ncm_nnls_solve: block pivoting on the normal equations, ncm_nnls.c:767-871) against the REFERENCE'S
OWN Lawson-Hanson solver -- numcosmo/external/misc/nnls.c (nnls_c, what ncm_nnls.c:873-935 calls for its alternative method), compiled
where it lies into oracle/_ref/libnnls_ref.so by oracle/Makefile.  Two different active-set algorithms; the NNLS minimiser of a full-column-
rank system is unique, so they must land on the same point."""
import numpy as np
import pytest

from helpers import make_sd, mvnd_problem


def _ref(oracle, A, f):
    out = oracle.ref_nnls_lh(A, f)
    if out is None:
        pytest.skip("oracle/_ref/libnnls_ref.so absent (built by `make -C oracle ref` where /root/reference exists)")
    x, rnorm, mode = out
    assert mode == 1
    return x, rnorm


@pytest.mark.parametrize("m,n,seed", [(60, 20, 0), (200, 80, 1), (300, 300, 2), (500, 120, 3), (64, 64, 4)])
def test_block_pivoting_equals_lawson_hanson_on_generic_systems(oracle, m, n, seed):
    rng = np.random.default_rng(seed)
    A = np.ascontiguousarray(np.abs(rng.standard_normal((m, n))) + 0.1 * rng.standard_normal((m, n)))
    f = rng.uniform(0.5, 1.5, m)
    x, rnorm = oracle.nnls_solve(A, f)[:2]
    xr, rr = _ref(oracle, A, f)
    assert np.all(x >= 0) and np.all(xr >= 0)
    assert abs(rnorm - rr) <= 1e-12 * rr
    assert np.max(np.abs(x - xr)) <= 1e-10 * np.max(xr)
    assert np.array_equal(x > 0, xr > 0)          # the same passive set


def test_unconstrained_optimum_inside_the_cone_and_fully_clamped_system(oracle):
    rng = np.random.default_rng(7)
    A = np.ascontiguousarray(rng.standard_normal((80, 10)))
    xt = rng.uniform(0.5, 2.0, 10)
    x, rnorm = oracle.nnls_solve(A, A @ xt)[:2]
    xr, rr = _ref(oracle, A, A @ xt)
    assert np.allclose(x, xt, rtol=1e-10) and np.allclose(xr, xt, rtol=1e-10) and rnorm < 1e-10 and rr < 1e-10
    x, rnorm = oracle.nnls_solve(A, -(A @ xt))[:2]     # every unconstrained coefficient negative: the minimiser is the origin
    xr, rr = _ref(oracle, A, -(A @ xt))
    assert not np.any(x) and not np.any(xr) and abs(rnorm - rr) <= 1e-13 * rr


@pytest.mark.parametrize("sd_s,k_s,d,n", [("kde", "gauss", 3, 150), ("vkde", "st", 4, 160)])
def test_interpolation_weight_systems(oracle, sd_s, k_s, d, n):
    """The systems prepare_interp actually solves: IM / f rows against ones (ncm_stats_dist.c:791-804, 1077-1079).  Kernel Gram matrices are
    badly conditioned, so the two algorithms are compared on what is well determined: the residual norm and the fitted values IM x."""
    mu, cov, X, m2lnL = mvnd_problem(oracle, d, n, seed=30 + d)
    sd = make_sd(oracle, oracle.SD_KDE if sd_s == "kde" else oracle.SD_VKDE, oracle.KERNEL_GAUSS if k_s == "gauss" else oracle.KERNEL_ST, 3.0, X,
                 m2lnp=m2lnL, use_threads=False)
    IM = sd.compute_IM() / np.exp(-0.5 * (m2lnL - m2lnL.min()))[:, None]
    A = np.ascontiguousarray(IM)
    f = np.ones(n)
    x, rnorm = oracle.nnls_solve(A, f)[:2]
    xr, rr = _ref(oracle, A, f)
    assert np.all(x >= 0) and np.all(xr >= 0)
    assert abs(rnorm - rr) <= 1e-6 * max(rr, 1e-12) + 1e-9
    assert np.max(np.abs(A @ x - A @ xr)) <= 1e-6
