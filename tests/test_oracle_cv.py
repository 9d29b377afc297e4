"""Cross-validation bandwidth selection, CPU side.

* the oracle's restatement of levmar's dlevmar_dif against the REFERENCE'S OWN levmar sources, compiled where they lie
  into oracle/_ref/liblevmar_ref.so (oracle/Makefile, target "ref"): bit-identical parameters and info[] arrays;
* the product's one-parameter optimisers (numcosmo_b200/host/optim.cc) against both: identical trial-point sequences;
* the nmsimplex2 restatement on closed-form minima (GSL itself is absent: "parity unpinned" for its exact trial sequence);
* the oracle's CV_SPLIT / CV_SPLIT_NOFIT / CV_LOO modes: structural checks of ncm_stats_dist.c:703-789, 1018-1072.
No GPU compute here: the host library is only used through its optimiser test faces.
"""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import mvnd_problem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_dp = C.POINTER(C.c_double)


def _host_lib():
    L = C.CDLL(os.path.join(ROOT, "numcosmo_b200", "lib", "libncm_stats_dist_b200.so"))
    L.ncm_b200_test_simplex1.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int, _dp, _dp]
    L.ncm_b200_test_lm1_dif.argtypes = [C.c_void_p, C.c_void_p, _dp, _dp, C.c_int, C.c_int, _dp, _dp]
    return L


_F1 = C.CFUNCTYPE(C.c_double, C.c_double, C.c_void_p)
_FL = C.CFUNCTYPE(None, C.c_double, _dp, C.c_int, C.c_void_p)


def _host_simplex(L, f, x0, step, tol=1e-3, max_iter=1000):
    trace = []

    def g(x, _):
        trace.append(x)
        return float(f(x))

    cb = _F1(g)
    xb, fb = C.c_double(), C.c_double()
    it = L.ncm_b200_test_simplex1(C.cast(cb, C.c_void_p), None, x0, step, tol, max_iter, C.byref(xb), C.byref(fb))
    return it, xb.value, fb.value, trace


def _host_lm(L, f, p0, n, x=None, opts=None, itmax=1000):
    def g(p, hx, nn, _):
        v = f(np.array([p]))
        for k in range(nn):
            hx[k] = v[k]

    cb = _FL(g)
    p, info = C.c_double(p0), np.zeros(10)
    o = np.ascontiguousarray(opts, dtype=np.float64)
    xx = None if x is None else np.ascontiguousarray(x, dtype=np.float64)
    it = L.ncm_b200_test_lm1_dif(C.cast(cb, C.c_void_p), None, C.byref(p), None if xx is None else xx.ctypes.data_as(_dp), n, itmax,
                                 o.ctypes.data_as(_dp), info.ctypes.data_as(_dp))
    return p.value, info, it


def _need_ref(oracle):
    if oracle.ref_levmar() is None:
        pytest.skip("oracle/_ref/liblevmar_ref.so absent (built by `make -C oracle ref` where /root/reference exists)")


# ---------------------------------------------------------------------------------------------------------------
# levmar: restatement == the reference's own build
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 7, 8, 9, 64, 1024, 1025, 3000])
@pytest.mark.parametrize("delta", [1e-6, -1e-6])
def test_lm_dif_one_parameter_bit_identical_to_reference_levmar(oracle, n, delta):
    """m = 1 is the only case NcmStatsDist uses (ncm_stats_dist.c:1062); n * m > 1024 switches levmar to its blocked J^T J."""
    _need_ref(oracle)
    L = _host_lib()
    rng = np.random.default_rng(n)
    t = np.linspace(0, 1, n) if n > 1 else np.array([0.5])
    y = np.exp(-1.3 * t) + 1e-2 * rng.standard_normal(n)
    opts = [1e-3, 1e-7, 1e-7, 1e-10, delta]   # the options of ncm_stats_dist.c:1035-1039
    for x in (None, y):
        f = (lambda p: np.exp(p[0] * t) - y) if x is None else (lambda p: np.exp(p[0] * t))
        p_ref, info_ref, it_ref = oracle.lm_dif(f, [0.5], n, x=x, opts=opts, reference=True)
        p_orc, info_orc, it_orc = oracle.lm_dif(f, [0.5], n, x=x, opts=opts)
        p_host, info_host, it_host = _host_lm(L, f, 0.5, n, x=x, opts=opts)
        assert p_orc[0] == p_ref[0] and it_orc == it_ref and np.array_equal(info_orc, info_ref)
        assert p_host == p_ref[0] and it_host == it_ref and np.array_equal(info_host, info_ref)
        assert abs(p_ref[0] + 1.3) < 0.05 and info_ref[6] in (1, 2, 6)


def test_lm_dif_several_parameters_bit_identical_to_reference_levmar(oracle):
    _need_ref(oracle)
    n = 50
    t = np.linspace(0, 2, n)
    y = 2.0 * np.exp(-0.7 * t) + 0.3
    f = lambda p: p[0] * np.exp(p[1] * t) + p[2] - y
    a = oracle.lm_dif(f, [1.0, 0.0, 0.0], n)
    b = oracle.lm_dif(f, [1.0, 0.0, 0.0], n, reference=True)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2]
    assert np.allclose(a[0], [2.0, -0.7, 0.3], atol=1e-6)
    # blocked J^T J (n m > 1024), explicit target, central differences
    n = 1500
    t = np.linspace(0, 2, n)
    x = 2.0 * np.exp(-0.7 * t) + 0.01 * np.random.default_rng(3).standard_normal(n)
    g = lambda p: p[0] * np.exp(p[1] * t)
    o = [1e-3, 1e-12, 1e-12, 1e-12, -1e-6]
    a = oracle.lm_dif(g, [1.0, 0.0], n, x=x, opts=o)
    b = oracle.lm_dif(g, [1.0, 0.0], n, x=x, opts=o, reference=True)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2]


def test_lm_dif_failure_modes_match_reference(oracle):
    """non-finite residuals (stop 7 -> LM_ERROR) and an already-converged start (stop 6)."""
    _need_ref(oracle)
    L = _host_lib()
    opts = [1e-3, 1e-7, 1e-7, 1e-10, 1e-6]
    n = 20
    t = np.linspace(0, 1, n)
    bad = lambda p: np.full(n, np.nan)
    zero = lambda p: np.zeros(n)
    for f in (bad, zero):
        r = oracle.lm_dif(f, [0.3], n, opts=opts, reference=True)
        o = oracle.lm_dif(f, [0.3], n, opts=opts)
        h = _host_lm(L, f, 0.3, n, opts=opts)
        assert r[2] == o[2] == h[2]
        assert r[1][6] == o[1][6] == h[1][6]
    assert oracle.lm_dif(bad, [0.3], n, opts=opts)[2] == -1


# ---------------------------------------------------------------------------------------------------------------
# nmsimplex2
# ---------------------------------------------------------------------------------------------------------------
def test_nmsimplex2_closed_form_minima(oracle):
    x, fv, size, it = oracle.nmsimplex2_minimize(lambda x: (x[0] - 0.3) ** 2 + 1, [0.0], [0.1])
    assert abs(x[0] - 0.3) < 2e-3 and abs(fv - 1) < 1e-5 and size < 1e-3 and 0 < it < 50
    x, fv, size, it = oracle.nmsimplex2_minimize(lambda x: (1 - x[0]) ** 2 + 100 * (x[1] - x[0] ** 2) ** 2, [-1.2, 1.0], [0.1, 0.1], 1e-8, 5000)
    assert np.allclose(x, [1, 1], atol=1e-6) and fv < 1e-12 and size < 1e-8
    x, fv, size, it = oracle.nmsimplex2_minimize(lambda x: np.sum((x - np.arange(4)) ** 2), np.zeros(4), np.full(4, 0.5), 1e-7, 5000)
    assert np.allclose(x, np.arange(4), atol=1e-5)
    # first step from (0, 0.1) on a decreasing function is the reflection-then-expansion pair 0.2, 0.3 (coefficients -1, -2)
    seen = []
    oracle.nmsimplex2_minimize(lambda x: seen.append(float(x[0])) or -x[0], [0.0], [0.1], 1e-3, 1)
    assert np.allclose(seen, [0.0, 0.1, 0.2, 0.3], atol=1e-15)
    # on an increasing one the highest corner (0.1) mirrors to -0.1, lower than the lowest, so the expansion -0.2 is tried and kept
    seen = []
    x, fv, size, it = oracle.nmsimplex2_minimize(lambda x: seen.append(float(x[0])) or x[0], [0.0], [0.1], 1e-3, 1)
    assert np.allclose(seen, [0.0, 0.1, -0.1, -0.2], atol=1e-15) and abs(x[0] + 0.2) < 1e-15


def test_host_simplex_visits_the_same_points_as_the_general_restatement(oracle):
    L = _host_lib()
    fs = [lambda x: (x - 0.3) ** 2 + 1, lambda x: np.cosh(x - 2.0) + 0.1 * np.sin(5 * x), lambda x: abs(x + 1.7) ** 1.5 + x,
          lambda x: -np.exp(-(x - 0.5) ** 2 / 0.02), lambda x: (x * x - 2) ** 2, lambda x: float("nan") if x > 0.35 else (x - 1) ** 2,
          lambda x: 1.0]
    for f in fs:
        for x0, st in ((0.0, 0.1), (1.0, 0.1), (-3.0, 0.5), (0.2, -0.1)):
            it_h, x_h, f_h, tr_h = _host_simplex(L, f, x0, st)
            tr_o = []
            x_o, f_o, size_o, it_o = oracle.nmsimplex2_minimize(lambda v: tr_o.append(float(v[0])) or f(v[0]), [x0], [st])
            assert it_h == it_o and tr_h == tr_o
            if it_h > 0:
                assert x_h == x_o[0] and (f_h == f_o or (f_h != f_h and f_o != f_o))


# ---------------------------------------------------------------------------------------------------------------
# the oracle's CV modes
# ---------------------------------------------------------------------------------------------------------------
CASES = [("kde", "gauss", 3.0), ("kde", "st", 3.0), ("vkde", "gauss", 3.0), ("vkde", "st", 1.0)]


def _orc(oracle, sd_s, k_s, nu, d, cv):
    return oracle.StatsDist(oracle.SD_KDE if sd_s == "kde" else oracle.SD_VKDE, oracle.KERNEL_GAUSS if k_s == "gauss" else oracle.KERNEL_ST, d, nu, cv)


@pytest.mark.parametrize("sd_s,k_s,nu", CASES)
def test_oracle_cv_split_nofit(oracle, sd_s, k_s, nu):
    d, n = 3, 240
    mu, cov, X, m2lnL = mvnd_problem(oracle, d, n, seed=41)
    o = _orc(oracle, sd_s, k_s, nu, d, oracle.CV_SPLIT_NOFIT)
    o.set_split_frac(0.6)
    o.add_obs_matrix(X)
    assert o.prepare_interp(m2lnL) == 0
    assert o.get_n_obs() == n and o.get_n_kernels() == int(np.ceil(n * 0.6))       # ncm_stats_dist.c:743-744
    lnos, val = o.cv_trace()
    assert len(lnos) >= 4 and lnos[0] == 0.0 and abs(lnos[1] - 0.1) < 1e-15          # ln over_smooth = 0, step 0.1 (:662-663)
    assert val.min() <= val[0] and np.all(np.isfinite(val))
    # the object keeps the bandwidth of the LAST trial point, not of the best corner (:660-701 never restores it)
    assert abs(o.get_over_smooth() / np.exp(lnos[-1]) - 1) < 4e-16
    # the simplex stopped at size < 1e-3 around the minimum of the sampled objective
    assert abs(lnos[-1] - lnos[np.argmin(val)]) < 5e-3
    w = o.peek_weights()
    assert len(w) == o.get_n_kernels() and abs(w.sum() - 1) < 1e-12 and w.min() >= 0.01 / len(w) * (1 - 1e-12)
    # the objective itself: -2 ln L of the held-out points under uniform weights at that bandwidth
    chk = _orc(oracle, sd_s, k_s, nu, d, oracle.CV_NONE)
    chk.add_obs_matrix(X[: o.get_n_kernels()])
    k = int(np.argmin(val))
    chk.set_over_smooth(np.exp(lnos[k]))
    assert chk.prepare() == 0
    if sd_s == "kde":   # same kernels, same covariance (first n_kernels points), same bandwidth rule
        got = chk.eval_m2lnp_batch(X[o.get_n_kernels():], 2).sum()
        assert abs(got / val[k] - 1) < 1e-12


@pytest.mark.parametrize("sd_s,k_s,nu", CASES)
def test_oracle_cv_loo(oracle, sd_s, k_s, nu):
    d, n = 2, 160
    mu, cov, X, m2lnL = mvnd_problem(oracle, d, n, seed=43)
    o = _orc(oracle, sd_s, k_s, nu, d, oracle.CV_LOO)
    o.add_obs_matrix(X)
    assert o.prepare_interp(m2lnL) == 0
    assert o.get_n_obs() == n and o.get_n_kernels() == n
    lnos, val = o.cv_trace()
    assert len(lnos) >= 4 and np.all(np.isfinite(val)) and val.min() <= val[0]
    assert abs(o.get_over_smooth() / np.exp(lnos[-1]) - 1) < 4e-16
    if sd_s == "kde" and k_s == "gauss":
        # closed form of _ncm_stats_dist_amise_kde_gauss (:513-558) from the two interpolation matrices
        k = int(np.argmin(val))
        chk = _orc(oracle, sd_s, k_s, nu, d, oracle.CV_NONE)
        chk.add_obs_matrix(X)
        chk.set_over_smooth(np.exp(lnos[k]) * np.sqrt(2.0))
        assert chk.prepare() == 0
        IM2 = chk.compute_IM()
        chk.set_over_smooth(np.exp(lnos[k]))
        assert chk.prepare() == 0
        IM1 = chk.compute_IM()
        amise = IM2.sum() / n**2 - 2.0 * (IM1.sum() - np.trace(IM1)) / (n * (n - 1))
        assert abs(amise / val[k] - 1) < 1e-10


@pytest.mark.parametrize("sd_s,k_s,nu", CASES)
def test_oracle_cv_split_uses_reference_levmar(oracle, sd_s, k_s, nu):
    """CV_SPLIT through the reference's own dlevmar_dif and through the restatement: identical optimiser traces."""
    d, n = 3, 200
    mu, cov, X, m2lnL = mvnd_problem(oracle, d, n, seed=47)
    out = []
    modes = [False] + ([True] if oracle.ref_levmar() is not None else [])
    try:
        for use_ref in modes:
            oracle.use_ref_levmar(use_ref)
            o = _orc(oracle, sd_s, k_s, nu, d, oracle.CV_SPLIT)
            o.add_obs_matrix(X)
            assert o.prepare_interp(m2lnL) == 0
            out.append((o.cv_trace(), o.get_over_smooth(), o.peek_weights().copy(), o.get_rnorm()))
    finally:
        oracle.use_ref_levmar(False)
    (lnos, val), os_fit, w, rn = out[0]
    assert o.get_n_kernels() == n // 2 and o.get_n_obs() == n
    assert len(lnos) >= 13 and lnos[0] == 0.0             # 1 + 10 random tries + the levmar evaluations (:1041-1064)
    # the ten tries are Gaussian steps of width 0.5 around the running best, drawn from the object's own MT19937 seeded 0 (:178)
    rng = oracle.RNG(0)
    best, best_r = 0.0, val[0]
    for k in range(1, 11):
        assert lnos[k] == rng.gaussian(0.5) + best
        if val[k] < best_r:
            best, best_r = lnos[k], val[k]
    assert lnos[11] == best                                 # levmar starts from the best try
    assert abs(lnos[12] - (best + max(1e-4 * abs(best), 1e-6))) < 1e-15   # forward-difference step of misc_core.c:153-161
    assert abs(w.sum() - 1) < 1e-12 and rn >= 0
    if len(out) == 2:
        (lnos_r, val_r), os_r, w_r, rn_r = out[1]
        assert np.array_equal(lnos, lnos_r) and np.array_equal(val, val_r) and os_fit == os_r and np.array_equal(w, w_r) and rn == rn_r
