"""The reference's fallback chain for passive-set systems that are not positive definite:
dposv -> dsysv -> dgels, ncm_nnls.c:573-638, 655-666.  csrc/ldl_bk.cu (Bunch-Kaufman L D L^T, dsytf2 / dsytrs order) and csrc/qr_ls.cu
(Householder least squares, dgeqr2 order) against LAPACK, and the NNLS that reaches them against the CPU oracle on deliberately
rank-deficient Gram matrices.

What can agree: on a system that is singular to working precision the solution is rounding-driven on the CPU as well (two LAPACK
builds do not agree on it), so the parity quantities there are the residual of the normal equations, the fitted values A x and the
NNLS residual norm -- not x."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _sym_indefinite(rs, n, kind):
    if kind == "random":                       # eigenvalues of both signs, well conditioned
        Q, _ = np.linalg.qr(rs.standard_normal((n, n)))
        lam = rs.uniform(0.5, 2.0, n) * rs.choice([-1.0, 1.0], n)
        return (Q * lam) @ Q.T
    if kind == "zero_diag":                    # forces 2 x 2 pivot blocks
        S = rs.standard_normal((n, n))
        S = S + S.T
        np.fill_diagonal(S, 0.0)
        return S
    if kind == "small_diag":                   # forces interchanges (|a_kk| << column maximum)
        S = rs.standard_normal((n, n))
        S = S + S.T
        np.fill_diagonal(S, 1e-3 * rs.standard_normal(n))
        return S
    raise ValueError(kind)


@pytest.mark.parametrize("n", [1, 2, 3, 7, 50, 174, 232, 233, 500, 1200])
@pytest.mark.parametrize("kind", ["random", "zero_diag", "small_diag"])
def test_dsysv_upper_against_lapack(gpu_ctx, n, kind):
    """n <= 232: one CTA, packed triangle in shared memory; above: the cooperative grid in global memory."""
    import torch

    if n == 1 and kind == "zero_diag":
        pytest.skip("a 1 x 1 zero matrix is the singular case, tested below")
    rs = np.random.default_rng(7 * n + len(kind))
    S = _sym_indefinite(rs, n, kind)
    b = rs.standard_normal(n)
    ld = (n + 7) // 8 * 8
    dM = torch.full((n, ld), float("nan"), dtype=torch.float64, device="cuda")      # the lower triangle must never be read
    iu = np.triu_indices(n)
    Mh = np.full((n, ld), np.nan)
    Mh[iu] = S[iu]
    dM.copy_(torch.from_numpy(Mh))
    db = torch.from_numpy(b).cuda()
    torch.cuda.synchronize()
    info = gpu_ctx.dsysv_upper_dev(n, dM.data_ptr(), ld, db.data_ptr())
    assert info == 0
    x = db.cpu().numpy()
    xr = np.linalg.solve(S, b)
    cond = np.linalg.cond(S)
    assert np.max(np.abs(x - xr)) <= 64 * n * cond * np.finfo(float).eps * np.abs(xr).max(), (cond, np.max(np.abs(x - xr)) / np.abs(xr).max())
    # backward error of a stable factorisation: |S x - b| ~ eps |S| |x|
    assert np.max(np.abs(S @ x - b)) <= 64 * n * np.finfo(float).eps * (np.abs(S) @ np.abs(x)).max()
    # deterministic: a second call on the same data gives the same bits
    dM.copy_(torch.from_numpy(Mh))
    db2 = torch.from_numpy(b).cuda()
    torch.cuda.synchronize()
    assert gpu_ctx.dsysv_upper_dev(n, dM.data_ptr(), ld, db2.data_ptr()) == 0
    assert np.array_equal(db2.cpu().numpy(), x)


def test_dsysv_pivot_sequence_is_lapacks(gpu_ctx):
    """Same pivoting rule as dsytf2 'L': on a positive definite matrix no interchange happens and L D L^T reduces to the Cholesky
    factor up to the diagonal scaling; on the classic [[eps, 1], [1, eps]] block a 2 x 2 pivot is taken (a 1 x 1 pivot on eps would
    lose the solution)."""
    import torch

    rs = np.random.default_rng(2)
    n = 40
    B = rs.standard_normal((n + 5, n))
    S = B.T @ B
    ld = 40
    for Smat in (S, np.kron(np.eye(n // 2), np.array([[1e-18, 1.0], [1.0, 1e-18]])) + 1e-3 * S):
        b = rs.standard_normal(n)
        dM = torch.from_numpy(np.ascontiguousarray(np.triu(Smat))).cuda()
        db = torch.from_numpy(b).cuda()
        torch.cuda.synchronize()
        assert gpu_ctx.dsysv_upper_dev(n, dM.data_ptr(), ld, db.data_ptr()) == 0
        xr = np.linalg.solve(Smat, b)
        assert np.max(np.abs(db.cpu().numpy() - xr)) <= 1e-9 * np.abs(xr).max()


@pytest.mark.parametrize("n", [1, 5, 100, 300])
def test_dsysv_exactly_singular_reports_info(gpu_ctx, n):
    """dsytf2 sets INFO = k when the k-th pivot column is exactly zero; dsysv then returns without solving (ncm_nnls.c:604 goes on to dgels)."""
    import torch

    rs = np.random.default_rng(n)
    S = np.zeros((n, n))
    r = max(0, n - 3)                              # a positive definite leading block, then exact zeros
    if r > 0:
        B = rs.standard_normal((r + 3, r))
        S[:r, :r] = B.T @ B
    ld = (n + 7) // 8 * 8
    Mh = np.zeros((n, ld))
    Mh[:, :n] = np.triu(S)
    dM = torch.from_numpy(Mh).cuda()
    db = torch.from_numpy(rs.standard_normal(n)).cuda()
    torch.cuda.synchronize()
    info = gpu_ctx.dsysv_upper_dev(n, dM.data_ptr(), ld, db.data_ptr())
    assert info == r + 1


@pytest.mark.parametrize("m,ncols,n", [(5, 5, 5), (60, 40, 17), (300, 300, 300), (700, 512, 200), (1500, 900, 640)])
def test_dgels_cols_against_lapack(gpu_ctx, m, ncols, n):
    import torch

    rs = np.random.default_rng(m + n)
    A = rs.standard_normal((m, ncols))
    f = rs.standard_normal(m)
    idx = np.sort(rs.permutation(ncols)[:n]).astype(np.int32)
    lda = (ncols + 7) // 8 * 8
    Ah = np.zeros((m, lda))
    Ah[:, :ncols] = A
    dA, dF, dIdx = torch.from_numpy(Ah).cuda(), torch.from_numpy(f).cuda(), torch.from_numpy(idx).cuda()
    dX = torch.zeros(n, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    info = gpu_ctx.dgels_cols_dev(m, n, dA.data_ptr(), lda, dIdx.data_ptr(), dF.data_ptr(), dX.data_ptr())
    assert info == 0
    xr = np.linalg.lstsq(A[:, idx], f, rcond=None)[0]
    cond = np.linalg.cond(A[:, idx])
    assert np.max(np.abs(dX.cpu().numpy() - xr)) <= 64 * n * cond**2 * np.finfo(float).eps * max(np.abs(xr).max(), 1e-300)
    # an exactly zero column: R gets an exactly zero diagonal entry -> info, as dtrtrs inside dgels
    Ah[:, idx[n // 2]] = 0.0
    dA.copy_(torch.from_numpy(Ah))
    torch.cuda.synchronize()
    info = gpu_ctx.dgels_cols_dev(m, n, dA.data_ptr(), lda, dIdx.data_ptr(), dF.data_ptr(), dX.data_ptr())
    assert info == n // 2 + 1


def _rank_deficient(rs, m, n, r, noise):
    """m x n non-negative matrix of numerical rank r: n - r columns are combinations of the others plus `noise` relative perturbation."""
    B = np.abs(rs.standard_normal((m, r))) + 0.05
    C = np.abs(rs.standard_normal((r, n - r))) / r
    A = np.hstack([B, B @ C * (1.0 + noise * rs.standard_normal((m, n - r)))])
    return np.ascontiguousarray(A[:, rs.permutation(n)])


@pytest.mark.parametrize("m,n,r,noise", [(120, 80, 60, 0.0), (300, 200, 150, 1e-15), (300, 200, 199, 1e-12), (900, 600, 500, 1e-14)])
def test_nnls_rank_deficient_gram_follows_the_fallback_chain(oracle, gpu_ctx, m, n, r, noise):
    """M = A^T A is singular to working precision: dposv fails on both sides, both go through the symmetric-indefinite solve.  The
    minimiser of |A x - f| over x >= 0 is not unique, its residual norm and (to the conditioning of the active columns) its fitted
    values are: those are compared, x is checked for feasibility and optimality (KKT) only."""
    rs = np.random.default_rng(m + n + r)
    A = _rank_deficient(rs, m, n, r, noise)
    xt = np.maximum(rs.standard_normal(n), 0.0)
    f = A @ xt + 1e-3 * rs.standard_normal(m)
    x, rnorm, st = gpu_ctx.nnls_solve_host(A, f)
    xo, rno, so = oracle.nnls_solve(A, f)
    print(f"rank-deficient NNLS m={m} n={n} r={r}: gpu {st['n_chol']} chol / {st['n_lu']} lu / {st['n_qr']} qr, oracle {so['n_chol']} / {so['n_lu']} / {so['n_qr']}; "
          f"rnorm {rnorm:.12e} vs {rno:.12e}; |P| {st['n_passive']} vs {so['n_passive']}")
    assert so["n_lu"] > 0 and st["n_lu"] > 0                       # the fallback was exercised on both sides
    assert np.all(np.isfinite(x)) and np.all(x >= 0.0)
    assert abs(rnorm - rno) <= 1e-6 * rno
    assert np.linalg.norm(A @ x - A @ xo) <= 1e-5 * np.linalg.norm(A @ xo)
    # KKT of the optimum: gradient A^T (A x - f) >= 0 up to rounding, ~ 0 on the support
    g = A.T @ (A @ x - f)
    scale = np.linalg.norm(A, axis=0) * np.linalg.norm(f)
    assert np.all(g >= -1e-6 * scale)
    assert np.all(np.abs(g[x > 1e-8 * x.max()]) <= 1e-6 * scale[x > 1e-8 * x.max()])


def test_nnls_exact_duplicate_columns(oracle, gpu_ctx):
    """Two walkers on the same point give two identical IM columns: M is exactly singular.  The chain still returns a finite, feasible,
    optimal answer (whichever of dsysv / dgels ends up solving the systems)."""
    rs = np.random.default_rng(12)
    m, n = 150, 100
    A = np.abs(rs.standard_normal((m, n))) + 0.1 * np.eye(m, n)
    A[:, 40] = A[:, 7]
    A[:, 41] = A[:, 7]
    f = A @ np.maximum(rs.standard_normal(n), 0.0)
    x, rnorm, st = gpu_ctx.nnls_solve_host(A, f)
    xo, rno, so = oracle.nnls_solve(A, f)
    print(f"duplicate columns: gpu {st}, oracle {so}, rnorm {rnorm:.3e} vs {rno:.3e}")
    assert np.all(np.isfinite(x)) and np.all(x >= 0.0)
    assert abs(rnorm - rno) <= 1e-8 * np.linalg.norm(f)
    assert np.linalg.norm(A @ x - A @ xo) <= 1e-8 * np.linalg.norm(f)
