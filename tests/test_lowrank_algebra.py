"""The algebra of csrc/lowrank.cu restated in numpy (CPU only): a passive set P = (B \\ D) u A solved through the factor of M[B,B].

    T = W^T [M_BA  E_D  b_B],  W = U^-1,  M_BB = U^T U
    H = T_R^T T_R - diag(M_AA, 0)            (k x k, quasi-definite: negative definite on A, positive definite on D)
    H z = T_R^T t_b - [b_A; 0]               z = [x_A; mu]
    x_B = W (t_b - T_R z)                    (x_D = 0 enforced by the multipliers mu)

and H = L J L^T, J = diag(-I_A, +I_D), without pivoting.  Checked against a direct solve of M[P,P] x = b[P] (what the reference
does, ncm_nnls.c:544-568 + ncm_matrix.c:1199-1210), for pure removals (the case of every NNLS that starts from the full set),
pure additions and mixed sets; and the append-only extension used when the removed set only grows."""
import numpy as np
import pytest


def ljl_factor(H, na):
    """L J L^T without pivoting (as lr_small_kernel): returns L (lower) with J = diag(-1 ... -1, +1 ... +1)."""
    k = H.shape[0]
    L = np.zeros_like(H)
    J = np.array([-1.0] * na + [1.0] * (k - na))
    A = H.copy()
    for c in range(k):
        p = J[c] * A[c, c]
        assert p > 0.0, (c, p)                      # the kernel reports info = c + 1 here
        L[c, c] = np.sqrt(p)
        L[c + 1:, c] = J[c] * A[c + 1:, c] / L[c, c]
        A[c + 1:, c + 1:] -= J[c] * np.outer(L[c + 1:, c], L[c + 1:, c])
    return L, J


def lowrank_solve(M, b, B, P):
    B, P = list(B), list(P)
    A = [i for i in P if i not in B]
    D = [q for q, i in enumerate(B) if i not in P]      # positions in B
    U = np.linalg.cholesky(M[np.ix_(B, B)]).T
    W = np.linalg.inv(U)
    ED = np.zeros((len(B), len(D)))
    ED[D, range(len(D))] = 1.0
    bB = b[B].copy()
    V = np.hstack([M[np.ix_(B, A)], ED, bB[:, None]])
    T = W.T @ V
    TR, tb = T[:, :-1], T[:, -1]
    na, k = len(A), len(A) + len(D)
    H = TR.T @ TR
    H[:na, :na] -= M[np.ix_(A, A)]
    rhs = TR.T @ tb
    rhs[:na] -= b[A]
    L, J = ljl_factor(H, na)
    z = np.linalg.solve(L.T, J * np.linalg.solve(L, rhs))
    xB = W @ (tb - TR @ z)
    x = np.zeros(M.shape[0])
    x[B] = xB
    x[A] = z[:na]
    return x, xB[D], L


@pytest.mark.parametrize("nrem,nadd", [(7, 0), (0, 5), (6, 4), (1, 1), (25, 0)])
def test_bordered_system_equals_direct_solve(nrem, nadd):
    rs = np.random.default_rng(10 * nrem + nadd)
    n = 90
    G = np.abs(rs.standard_normal((n + 20, n))) + np.eye(n + 20, n)
    M, b = G.T @ G, G.T @ np.ones(n + 20)
    full = rs.permutation(n)
    B = np.sort(full[: n - nadd])                       # base set; the last nadd indices are outside it
    rem = rs.permutation(len(B))[:nrem]
    P = np.sort(np.concatenate([np.delete(B, rem), full[n - nadd:]]))
    x, xD, _ = lowrank_solve(M, b, B, P)
    xr = np.zeros(n)
    xr[P] = np.linalg.solve(M[np.ix_(P, P)], b[P])
    assert xD.size == 0 or np.max(np.abs(xD)) <= 1e-9 * np.abs(xr).max()   # the constraint x_D = 0 holds
    assert np.max(np.abs(x - xr)) <= 1e-9 * np.abs(xr).max()
    outside = np.setdiff1d(np.arange(n), P)
    assert np.all(x[outside] == 0.0) or np.max(np.abs(x[outside])) <= 1e-9 * np.abs(xr).max()


def test_append_only_extension_of_the_small_factor():
    """Removed set D1 then D2 = D1 + new members appended at the end: the factor of H(D1) is the leading block of the factor of H(D2)
    (what lr_small_kernel reuses when kold > 0)."""
    rs = np.random.default_rng(3)
    n = 70
    G = np.abs(rs.standard_normal((n + 10, n))) + np.eye(n + 10, n)
    M, b = G.T @ G, G.T @ np.ones(n + 10)
    B = np.arange(n)
    D1 = [5, 17, 40, 41, 60]
    D2 = D1 + [3, 22, 66]                                # appended, not sorted
    U = np.linalg.cholesky(M).T
    W = np.linalg.inv(U)

    def small(D):
        TR = W.T[:, D]                                   # W^T E_D: a gather of rows of W
        return ljl_factor(TR.T @ TR, 0)[0]

    L1, L2 = small(D1), small(D2)
    assert np.max(np.abs(L2[: len(D1), : len(D1)] - L1)) <= 1e-13 * np.abs(L1).max()
    # and the solution through the extended factor is the direct one
    P = np.setdiff1d(B, D2)
    x, _, _ = lowrank_solve(M, b, B, P)
    xr = np.zeros(n)
    xr[P] = np.linalg.solve(M[np.ix_(P, P)], b[P])
    assert np.max(np.abs(x - xr)) <= 1e-9 * np.abs(xr).max()


def trinv_recursive_doubling(U, blk):
    """lowrank.cu trinv_upper_on: level 0 inverts the blk x blk diagonal blocks (trinv_diag_kernel, 64 on the device); level s = blk,
    2 blk, ... takes every pair of adjacent s-blocks (r0 = 2 p s, r1 = r0 + s, the second one min (s, n - r1) wide, absent when
    r1 >= n) and fills  W12 = - W11 (U12 W22)  (trinv_step1 / step2, or step12 in one launch for s <= 128)."""
    n = U.shape[0]
    W = np.zeros_like(U)
    for k0 in range(0, n, blk):
        k1 = min(k0 + blk, n)
        W[k0:k1, k0:k1] = np.linalg.solve(U[k0:k1, k0:k1], np.eye(k1 - k0))
    s = blk
    while s < n:
        for r0 in range(0, n, 2 * s):
            r1 = r0 + s
            m2 = min(s, n - r1)
            if m2 <= 0:
                continue
            S12 = U[r0:r1, r1:r1 + m2] @ W[r1:r1 + m2, r1:r1 + m2]
            W[r0:r1, r1:r1 + m2] = -W[r0:r1, r0:r1] @ S12
        s *= 2
    return W


@pytest.mark.parametrize("n,blk", [(1, 4), (4, 4), (5, 4), (8, 4), (9, 4), (13, 4), (64, 8), (100, 8), (129, 16), (200, 64)])
def test_triangular_inverse_by_recursive_doubling(n, blk):
    rs = np.random.default_rng(n)
    B = rs.standard_normal((n + 5, n))
    U = np.linalg.cholesky(B.T @ B + 0.5 * np.eye(n)).T
    W = trinv_recursive_doubling(U, blk)
    assert np.array_equal(W, np.triu(W))
    assert np.max(np.abs(W @ U - np.eye(n))) <= 1e-12 * np.linalg.cond(U)
