"""The Householder rule of csrc/qr_ls.cu against LAPACK (CPU only): qr_ls_kernel's reflectors restated in numpy (dlarfg: beta =
-sign(alpha) hypot(alpha, |x|), tau = (beta - alpha) / beta, v = x / (alpha - beta); applied column by column to the trailing columns
and the right-hand side, dgeqr2 order; back substitution on R) must give LAPACK dgeqrf's R and dgels' solution.  The reference reaches
dgels from _ncm_nnls_solve_normal_QR, ncm_nnls.c:608-638."""
import numpy as np
import pytest
from scipy.linalg import lapack


def qr_ls(A, f):
    """As qr_ls_kernel: returns (R upper triangle incl. diagonal, x)."""
    Q = np.hstack([A.astype(float).copy(), f[:, None].copy()])
    m, n = A.shape
    for k in range(n):
        ck = Q[:, k]
        alpha, xnorm = ck[k], np.sqrt(np.sum(ck[k + 1:] ** 2))
        tau, scal, beta = 0.0, 0.0, alpha
        if xnorm != 0.0:
            beta = -np.copysign(np.hypot(alpha, xnorm), alpha)
            tau = (beta - alpha) / beta
            scal = 1.0 / (alpha - beta)
        if tau != 0.0:
            v = ck[k + 1:] * scal
            for j in range(k + 1, n + 1):
                w = Q[k, j] + v @ Q[k + 1:, j]
                Q[k + 1:, j] -= tau * w * v
                Q[k, j] -= tau * w
        Q[k, k] = beta
    R, c = np.triu(Q[:n, :n]), Q[:n, n].copy()
    x = np.zeros(n)
    for k in range(n - 1, -1, -1):
        x[k] = c[k] / R[k, k]
        c[:k] -= x[k] * R[:k, k]
    return R, x


@pytest.mark.parametrize("m,n", [(5, 5), (40, 17), (120, 60), (300, 200)])
def test_reflectors_and_solution_equal_lapack(m, n):
    rs = np.random.default_rng(m + n)
    A, f = rs.standard_normal((m, n)), rs.standard_normal(m)
    R, x = qr_ls(A, f)
    qr, tau, work, info = lapack.dgeqrf(A)
    assert info == 0
    Rref = np.triu(qr[:n, :n])
    assert np.max(np.abs(R - Rref)) <= 1e-12 * np.abs(Rref).max()     # same sign convention, same values
    lqr, xs, info = lapack.dgels(A, f)
    assert info == 0
    assert np.max(np.abs(x - xs[:n])) <= 1e-10 * np.abs(xs[:n]).max()
