"""Low-rank reuse of the passive-set factor inside the NNLS (csrc/lowrank.cu): the triangular inverse by recursive doubling
against numpy, and NNLS runs with many consecutive passive sets against the CPU oracle, which calls dposv on every one of
them (ncm_nnls.c:655-666, 728-751): same passive set, same solution."""
import numpy as np
import pytest

from helpers import make_sd, mvnd_problem, upload_from_oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1, 7, 64, 65, 129, 200, 640, 1000, 1901, 2048, 3000])
def test_dtrtri_upper(gpu_ctx, n):
    import torch

    rs = np.random.default_rng(n)
    B = rs.standard_normal((n + 10, n))
    U = np.linalg.cholesky(B.T @ B + 0.5 * np.eye(n)).T
    ld = (n + 7) // 8 * 8
    dU = torch.full((n, ld), float("nan"), dtype=torch.float64, device="cuda")   # the lower triangle must never be read
    iu = np.triu_indices(n)
    Uh = np.full((n, ld), np.nan)
    Uh[iu] = U[iu]
    dU.copy_(torch.from_numpy(Uh))
    dW = torch.full((n, ld), float("nan"), dtype=torch.float64, device="cuda")
    dS = torch.full((n, ld), float("nan"), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    gpu_ctx.dtrtri_upper_dev(n, dU.data_ptr(), ld, dW.data_ptr(), dS.data_ptr())
    gpu_ctx.synchronize()
    W = dW.cpu().numpy()[:, :n]
    assert np.all(np.tril(W, -1) == 0.0)
    Wref = np.linalg.inv(U)
    assert np.max(np.abs(W - Wref)) <= 1e-12 * np.abs(Wref).max() * max(1.0, np.linalg.cond(U) / 100)
    assert np.max(np.abs(W @ U - np.eye(n))) < 1e-11


@pytest.mark.parametrize("m,n,frac_zero,noise", [(1400, 1200, 0.3, 1e-3), (900, 900, 0.5, 1e-2), (2500, 2048, 0.1, 1e-3), (700, 640, 0.6, 1e-1)])
def test_nnls_many_passive_sets(oracle, gpu_ctx, m, n, frac_zero, noise):
    """Generic systems whose solution sits on the boundary for a large share of the unknowns: the block-pivoting loop visits many
    neighbouring passive sets, most of them served by low-rank modification on the device."""
    rs = np.random.default_rng(m + n)
    A = np.abs(rs.standard_normal((m, n))) + 0.5 * np.eye(m, n)
    xt = np.maximum(rs.standard_normal(n) + (0.5 - frac_zero) * 2.0, 0.0)
    f = A @ xt + noise * rs.standard_normal(m)
    x, rnorm, st = gpu_ctx.nnls_solve_host(A, f)
    xo, rno, so = oracle.nnls_solve(A, f)
    print(f"({m} x {n}): gpu {st}, oracle {so}")
    assert np.array_equal(x > 0, xo > 0), np.count_nonzero((x > 0) != (xo > 0))
    assert st["n_passive"] == so["n_passive"]
    assert st["n_chol"] + st["n_lowrank"] - st["n_lowrank_fallback"] == so["n_chol"]     # the same systems, one by one
    assert st["n_lowrank_fallback"] == 0 and (st["n_lowrank"] > 0 or n < 1024)   # below 512 unknowns every system is factorised afresh
    assert np.max(np.abs(x - xo)) / np.abs(xo).max() < 1e-10
    assert abs(rnorm - rno) / rno < 1e-10
    assert np.all(x >= 0.0)


@pytest.mark.parametrize("sd_s,k_s,d,n,same_path", [("vkde", "gauss", 10, 2048, True), ("vkde", "st", 5, 1500, True), ("vkde", "gauss", 4, 1500, True),
                                                         ("kde", "gauss", 3, 1024, False), ("vkde", "gauss", 4, 777, True)])
def test_kernel_gram_weights_with_reuse(oracle, gpu_ctx, sd_s, k_s, d, n, same_path):
    """The systems prepare_interp solves (kernel Gram matrices, d small => many active-set iterations).  same_path: the device visits
    the very sequence of passive sets the oracle does.  The KDE d = 3 case does not, with or without the low-rank solves
    (tools/nnls_trace_compare.py: the full 1024-set system is conditioned ~1e8 / eps beyond what fixes the sign of one coefficient, so the
    device's dposv and OpenBLAS's already part at the second system, 721 vs 722 unknowns); the two paths meet again at the same final set."""
    from numcosmo_b200 import capi

    sd_type = oracle.SD_KDE if sd_s == "kde" else oracle.SD_VKDE
    kernel = oracle.KERNEL_GAUSS if k_s == "gauss" else oracle.KERNEL_ST
    mu, cov, X, m2lnL = mvnd_problem(oracle, d, n, seed=900 + d)
    sd = make_sd(oracle, sd_type, kernel, 3.0, X, m2lnp=m2lnL)
    upload_from_oracle(gpu_ctx, capi, oracle, sd, sd_type, kernel, 3.0, X)
    f = np.exp(-0.5 * (m2lnL - m2lnL.min()))
    gpu_ctx.compute_IM(1.0 / f, fetch=False, nrows=n)
    x, rnorm, st = gpu_ctx.nnls_solve()
    so = sd.nnls_stats()
    w_o = sd.peek_weights()
    w = (1.0 - 0.01) * x / x.sum() + 0.01 / n
    print(f"{sd_s}-{k_s} d={d} n={n}: gpu {st}, oracle {so}")
    assert 0 <= st["n_lowrank_nested"] <= st["n_lowrank"]   # solves whose k x k factor was extended from the previous, nested one
    assert so["n_lu"] == 0 and st["n_lu"] == 0
    assert np.array_equal(x > 0, w_o > (0.01 / n) * (1 + 1e-9))
    if same_path:
        assert st["n_chol"] + st["n_lowrank"] - st["n_lowrank_fallback"] == so["n_chol"]
    assert st["n_lowrank"] > 0 or n < 1024
    assert abs(rnorm**2 - sd.get_rnorm()) <= 1e-8 * sd.get_rnorm()
    assert np.max(np.abs(w - w_o)) / w_o.max() < (1e-8 if same_path else 1e-6)
