"""Size-independent properties at the BASELINE.json sizes (65536 proposals x 65536 centres), where the CPU oracle cannot
run: (a) the two independent evaluation paths agree — a VKDE whose factors are all the global factor IS the KDE;
(b) additivity of the density over a partition of the centres (log-sum-exp merge); (c) batch invariance: a point gets
the same value whatever batch (and centre split) it is evaluated in; (d) interpolation-matrix rows equal single-kernel
densities.  Together with the oracle parity at small sizes these pin the full-size results."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N = Q = 65536


def _problem(d, seed):
    rs = np.random.default_rng(seed)
    sig = rs.uniform(2e-2, 5e-2, size=d)
    R = rs.normal(size=(d, d)) / np.sqrt(d)
    cov = (0.7 * np.eye(d) + 0.3 * (R @ R.T)) * np.outer(sig, sig)
    U = np.ascontiguousarray(np.linalg.cholesky(cov).T)
    mu = rs.uniform(1.0, 2.0, size=d)
    C = np.ascontiguousarray(mu + rs.normal(size=(N, d)) @ U)
    X = np.ascontiguousarray(mu + rs.normal(size=(Q, d)) @ U)
    w = rs.uniform(size=N)
    w[rs.uniform(size=N) < 0.1] = 0.0
    w /= w.sum()
    return U, C, X, w


def _lnnorm_gauss(U):
    d = U.shape[0]
    return 0.5 * (d * math.log(2.0 * math.pi) + 2.0 * np.log(np.diag(U)).sum())


@pytest.mark.parametrize("d", [10, 30])
def test_kde_equals_vkde_with_global_factors_full_size(d):
    from numcosmo_b200 import capi

    U, C, X, w = _problem(d, 40 + d)
    href = 0.9
    kde, vk = capi.Context(0), capi.Context(0)
    try:
        kde.set_kernel(capi.KERNEL_GAUSS, 3.0, d)
        kde.upload_kde(np.linalg.solve(U.T, C.T).T, N, U, _lnnorm_gauss(U))
        kde.set_weights(w, href)
        vk.set_kernel(capi.KERNEL_GAUSS, 3.0, d)
        U_all = np.broadcast_to(U, (N, d, d)).copy()
        vk.upload_vkde(C, N, U_all, np.full(N, _lnnorm_gauss(U)))
        del U_all
        vk.set_weights(w, href)
        a, b = kde.eval_m2lnp(X), vk.eval_m2lnp(X)
        assert np.all(np.isfinite(a)) and np.all(np.isfinite(b))
        # KDE forms chi2 as |a|^2 + |b|^2 - 2 a.b on whitened centred points, VKDE by a per-centre solve: different rounding
        assert np.max(np.abs(a - b) / np.abs(a)) < 1e-10
        # (c) batch invariance: 1000 scattered points evaluated alone
        idx = np.random.default_rng(1).choice(Q, 1000, replace=False)
        assert np.max(np.abs(vk.eval_m2lnp(X[idx]) - b[idx]) / np.abs(b[idx])) < 1e-12
        assert np.max(np.abs(kde.eval_m2lnp(X[idx]) - a[idx]) / np.abs(a[idx])) < 1e-12
        # (b) additivity over a partition of the centres
        sel = np.arange(N) % 3 == 0
        wa, wb = np.where(sel, w, 0.0), np.where(sel, 0.0, w)
        sub = X[idx]
        p = np.exp(-0.5 * vk.eval_m2lnp(sub))
        vk.set_weights(wa, href)
        pa = np.exp(-0.5 * vk.eval_m2lnp(sub))
        vk.set_weights(wb, href)
        pb = np.exp(-0.5 * vk.eval_m2lnp(sub))
        assert np.max(np.abs(pa + pb - p) / p) < 1e-11
    finally:
        kde.close()
        vk.close()


def test_im_rows_are_single_kernel_densities():
    """IM[i, j] = K_j(x_i) / norm_j: the same number as the density of a one-hot weight vector (N = 16384, d = 20)."""
    from numcosmo_b200 import capi

    d, n = 20, 16384
    rs = np.random.default_rng(3)
    U, C, X, w = _problem(d, 77)
    C = C[:n]
    T = np.triu(rs.normal(size=(n, d, d)) * (0.15 / np.sqrt(d)), 1)
    T[:, np.arange(d), np.arange(d)] = rs.uniform(0.3, 0.6, size=(n, d))
    U_all = np.triu(T @ U)
    lnn = 0.5 * (d * math.log(2.0 * math.pi) + 2.0 * np.log(np.abs(U_all[:, np.arange(d), np.arange(d)])).sum(axis=1))
    c = capi.Context(0)
    try:
        c.set_kernel(capi.KERNEL_GAUSS, 3.0, d)
        c.upload_vkde(C, n, U_all, lnn)
        c.set_weights(np.full(n, 1.0 / n), 1.0)
        IM = c.compute_IM(None, fetch=True, nrows=n)
        for j in (0, 777, n - 1):
            e = np.zeros(n)
            e[j] = 1.0
            c.set_weights(e, 1.0)
            rows = np.r_[0:64, j, n - 64:n]
            dens = c.eval(C[rows])
            ref = IM[rows, j]
            big = ref > 1e-250
            assert np.max(np.abs(dens[big] - ref[big]) / ref[big]) < 1e-11
    finally:
        c.close()
