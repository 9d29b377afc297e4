"""The pivoting rule of csrc/ldl_bk.cu against LAPACK itself (CPU only).

The device kernel cannot be run without a GPU, but its decisions can: `bk_pivots` below is the kernel's pivot search restated line by
line in numpy (same comparisons, same alpha, same interchanges, same rank-1 / rank-2 updates), and LAPACK's dsytrf -- which for
n < 64 runs the unblocked dsytf2 the kernel follows -- must return the very same IPIV on indefinite, zero-diagonal and
tiny-diagonal matrices.  (The reference reaches dsytrf through ncm_lapack_dsysv, ncm_nnls.c:573-606.)"""
import numpy as np
import pytest
from scipy.linalg import lapack

ALPHA = (1.0 + np.sqrt(17.0)) / 8.0


def bk_pivots(S):
    """Lower Bunch-Kaufman as in bk_factor_solve (csrc/ldl_bk.cu); returns LAPACK-style 1-based IPIV and the factored matrix."""
    A = np.tril(S).copy()
    n = A.shape[0]
    ipiv = np.zeros(n, dtype=int)
    k = 0
    while k < n:
        absakk = abs(A[k, k])
        if k + 1 < n:
            col = np.abs(A[k + 1:, k])
            imax = k + 1 + int(np.argmax(col))          # first maximum, as idamax
            colmax = col[imax - k - 1]
        else:
            imax, colmax = -1, 0.0
        kp, kstep = k, 1
        if max(absakk, colmax) == 0.0:
            pass                                        # singular column: INFO, no elimination
        elif not (absakk >= ALPHA * colmax):
            rowmax = 0.0
            if imax > k:
                rowmax = np.max(np.abs(A[imax, k:imax]))
            if imax + 1 < n:
                rowmax = max(rowmax, np.max(np.abs(A[imax + 1:, imax])))
            if absakk >= ALPHA * colmax * (colmax / rowmax):
                kp = k
            elif abs(A[imax, imax]) >= ALPHA * rowmax:
                kp = imax
            else:
                kp, kstep = imax, 2
        kk = k + kstep - 1
        if kp != kk:
            A[kp + 1:, [kk, kp]] = A[kp + 1:, [kp, kk]]
            tmp = A[kk + 1:kp, kk].copy()
            A[kk + 1:kp, kk] = A[kp, kk + 1:kp]
            A[kp, kk + 1:kp] = tmp
            A[kk, kk], A[kp, kp] = A[kp, kp], A[kk, kk]
            if kstep == 2:
                A[k + 1, k], A[kp, k] = A[kp, k], A[k + 1, k]
        if max(absakk, colmax) != 0.0:
            if kstep == 1:
                if k < n - 1:
                    r1 = 1.0 / A[k, k]
                    x = A[k + 1:, k].copy()
                    A[k + 1:, k + 1:] -= np.tril(np.outer(x, r1 * x))
                    A[k + 1:, k] = x * r1
            elif k < n - 2:
                d21 = A[k + 1, k]
                d11, d22 = A[k + 1, k + 1] / d21, A[k, k] / d21
                t = 1.0 / (d11 * d22 - 1.0)
                d21 = t / d21
                a0, a1 = A[k + 2:, k].copy(), A[k + 2:, k + 1].copy()
                wk, wkp1 = d21 * (d11 * a0 - a1), d21 * (d22 * a1 - a0)
                A[k + 2:, k + 2:] -= np.tril(np.outer(a0, wk) + np.outer(a1, wkp1))
                A[k + 2:, k], A[k + 2:, k + 1] = wk, wkp1
        if kstep == 1:
            ipiv[k] = kp + 1
        else:
            ipiv[k] = ipiv[k + 1] = -(kp + 1)
        k += kstep
    return ipiv, A


def _matrix(rs, n, kind):
    S = rs.standard_normal((n, n))
    S = S + S.T
    if kind == "zero_diag":
        np.fill_diagonal(S, 0.0)
    elif kind == "small_diag":
        np.fill_diagonal(S, 1e-3 * rs.standard_normal(n))
    elif kind == "gram_singular":          # what the NNLS meets: a Gram matrix singular to working precision
        B = np.abs(rs.standard_normal((n + 5, n // 2)))
        A = np.hstack([B, B @ np.abs(rs.standard_normal((n // 2, n - n // 2))) / n])
        S = A.T @ A
    return S


@pytest.mark.parametrize("kind", ["indefinite", "zero_diag", "small_diag", "gram_singular"])
@pytest.mark.parametrize("n", [2, 3, 8, 21, 40, 60])
def test_pivot_sequence_equals_lapack_dsytrf(kind, n):
    rs = np.random.default_rng(100 * n + len(kind))
    S = _matrix(rs, n, kind)
    ipiv, L = bk_pivots(S)
    ldu, ipiv_ref, info = lapack.dsytrf(S, lower=1)
    assert info == 0 or kind == "gram_singular"
    if kind == "gram_singular":
        # singular to working precision: comparisons are decided by rounding after the first n / 2 pivots; the leading, well-determined
        # part of the sequence must still agree
        m = n // 4
        assert np.array_equal(ipiv[:m], ipiv_ref[:m]), (ipiv, ipiv_ref)
        return
    assert np.array_equal(ipiv, ipiv_ref), (ipiv, ipiv_ref)
    # and the factors agree to rounding (same operations up to the order inside the updates)
    assert np.max(np.abs(np.tril(L) - np.tril(ldu))) <= 1e-10 * np.max(np.abs(ldu))
