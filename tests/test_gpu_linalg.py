"""FP64 building blocks exported by the C ABI (DMMA SYRK, blocked Cholesky) against numpy."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("m,n", [(64, 64), (300, 130), (1000, 257), (128, 1024), (2048, 2048)])
def test_dsyrk_ata(gpu_ctx, m, n):
    import torch

    rs = np.random.default_rng(m + n)
    A = rs.standard_normal((m, n))
    lda = (n + 7) // 8 * 8
    dA = torch.zeros((m, lda), dtype=torch.float64, device="cuda")
    dA[:, :n] = torch.from_numpy(A).cuda()
    dM = torch.zeros((n, lda), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    gpu_ctx.dsyrk_ata_dev(m, n, dA.data_ptr(), lda, dM.data_ptr(), lda)
    gpu_ctx.synchronize()
    M = dM.cpu().numpy()[:, :n]
    ref = A.T @ A
    iu = np.triu_indices(n)
    assert np.max(np.abs(M[iu] - ref[iu])) < 1e-12 * np.abs(ref).max()


@pytest.mark.parametrize("n", [1, 5, 64, 65, 200, 1000, 2048])
def test_dpotrf_upper(gpu_ctx, n):
    import torch

    rs = np.random.default_rng(n)
    B = rs.standard_normal((n + 10, n))
    S = B.T @ B + 0.1 * np.eye(n)
    ld = (n + 7) // 8 * 8
    dM = torch.zeros((n, ld), dtype=torch.float64, device="cuda")
    dM[:, :n] = torch.from_numpy(np.triu(S)).cuda()
    torch.cuda.synchronize()
    info = gpu_ctx.dpotrf_upper_dev(n, dM.data_ptr(), ld)
    assert info == 0
    U = np.triu(dM.cpu().numpy()[:, :n])
    Uref = np.linalg.cholesky(S).T
    assert np.max(np.abs(U - Uref)) < 1e-10 * np.abs(Uref).max()
    # not positive definite -> 1-based index of the failing pivot
    S2 = S.copy()
    k = n // 2
    S2[k, k] = -1.0
    dM[:, :n] = torch.from_numpy(np.triu(S2)).cuda()
    torch.cuda.synchronize()
    info = gpu_ctx.dpotrf_upper_dev(n, dM.data_ptr(), ld)
    assert info == k + 1


@pytest.mark.parametrize("n", [1, 5, 63, 64, 65, 128, 200, 1000, 1899, 2048, 3001, 4096, 4200])
def test_dposv_upper(gpu_ctx, n):
    """dposv 'U' (ncm_matrix_cholesky_solve, ncm_matrix.c:1199-1210): factor + forward + back substitution.
    n <= 4096 runs the single-launch kernel (chol_fused.cu), n = 4200 the launch-per-step path (chol.cu)."""
    import torch

    rs = np.random.default_rng(1000 + n)
    B = rs.standard_normal((n + 10, n))
    S = B.T @ B + 0.1 * np.eye(n)
    b = rs.standard_normal(n)
    ld = (n + 7) // 8 * 8
    dM = torch.full((n, ld), float("nan"), dtype=torch.float64, device="cuda")   # the lower triangle must never be read
    dM[:, :n] = torch.from_numpy(np.triu(S) + np.tril(np.full((n, n), np.nan), -1)).cuda()
    dB = torch.from_numpy(b).cuda()
    torch.cuda.synchronize()
    info = gpu_ctx.dposv_upper_dev(n, dM.data_ptr(), ld, dB.data_ptr())
    assert info == 0
    x = dB.cpu().numpy()
    xref = np.linalg.solve(S, b)
    assert np.max(np.abs(x - xref)) < 1e-9 * np.abs(xref).max()
    U = np.triu(dM.cpu().numpy()[:, :n])
    Uref = np.linalg.cholesky(S).T
    assert np.max(np.abs(U - Uref)) < 1e-10 * np.abs(Uref).max()
    # repeat on the same buffers (flag epochs) and check determinism
    dM[:, :n] = torch.from_numpy(np.triu(S)).cuda()
    dB.copy_(torch.from_numpy(b).cuda())
    torch.cuda.synchronize()
    assert gpu_ctx.dposv_upper_dev(n, dM.data_ptr(), ld, dB.data_ptr()) == 0
    assert np.array_equal(dB.cpu().numpy(), x)


@pytest.mark.parametrize("n", [4500, 6501, 12345])
def test_dposv_upper_lookahead_path(gpu_ctx, n):
    """n > 4096: the launch-per-step blocked Cholesky with look-ahead (chol.cu): panel groups of 128 / 256 / 512 columns on a
    high-priority stream, trailing-update tails on the context stream.  Sizes off every tile multiple; solved twice on the same
    buffers (stream fork / join must leave nothing in flight)."""
    import torch

    rs = np.random.default_rng(n)
    B = rs.standard_normal((n + 10, n))
    S = B.T @ B + 0.1 * np.eye(n)
    del B
    b = rs.standard_normal(n)
    ld = (n + 7) // 8 * 8
    dS = torch.from_numpy(np.triu(S)).cuda()
    dM = torch.full((n, ld), float("nan"), dtype=torch.float64, device="cuda")
    dM[:, :n] = dS
    dM[:, :n] += torch.tril(torch.full((n, n), float("nan"), dtype=torch.float64, device="cuda"), -1)   # the lower triangle must never be read
    dB = torch.from_numpy(b).cuda()
    torch.cuda.synchronize()
    assert gpu_ctx.dposv_upper_dev(n, dM.data_ptr(), ld, dB.data_ptr()) == 0
    x = dB.cpu().numpy()
    U = torch.triu(dM[:, :n])
    # residuals instead of a host factorisation at this size: U^T U = S and S x = b
    R = U.T @ U
    assert float(torch.max(torch.abs(torch.triu(R) - dS))) < 1e-11 * float(torch.max(torch.abs(dS)))
    Sfull = torch.from_numpy(S).cuda()
    assert float(torch.max(torch.abs(Sfull @ dB - torch.from_numpy(b).cuda()))) < 1e-8 * max(1.0, float(np.abs(x).max()) * float(torch.max(torch.abs(dS))))
    dM[:, :n] = dS
    dB.copy_(torch.from_numpy(b).cuda())
    torch.cuda.synchronize()
    assert gpu_ctx.dposv_upper_dev(n, dM.data_ptr(), ld, dB.data_ptr()) == 0
    assert np.array_equal(dB.cpu().numpy(), x)
