#!/usr/bin/env python
"""Stress of the single-launch Cholesky solve: random orders 1..4096 (with and without right-hand side), residual check,
repeat-call determinism; any dependency that was never satisfied surfaces as NCM_SD_GPU_ECUDA (spin limit), never as a hang."""
import sys, os, time
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from numcosmo_b200 import capi

ctx = capi.Context(0)
rs = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
count = int(sys.argv[2]) if len(sys.argv) > 2 else 200
worst = 0.0
t0 = time.time()
for it in range(count):
    n = int(rs.choice([rs.integers(1, 130), rs.integers(1, 1100), rs.integers(1, 4097)]))
    ld = (n + 7) // 8 * 8
    B = torch.randn((n + 8, n), dtype=torch.float64, device="cuda")
    S = B.T @ B + 0.05 * torch.eye(n, dtype=torch.float64, device="cuda")
    M = torch.full((n, ld), float("nan"), dtype=torch.float64, device="cuda")
    M[:, :n] = torch.triu(S) + torch.tril(torch.full((n, n), float("nan"), dtype=torch.float64, device="cuda"), -1)
    b = torch.randn(n, dtype=torch.float64, device="cuda")
    x = b.clone()
    torch.cuda.synchronize()
    if it % 5 == 4:
        assert ctx.dpotrf_upper_dev(n, M.data_ptr(), ld) == 0
        U = torch.triu(M[:, :n])
        err = float((U.T @ U - S).abs().max() / S.abs().max())
    else:
        assert ctx.dposv_upper_dev(n, M.data_ptr(), ld, x.data_ptr()) == 0
        err = float((S @ x - b).abs().max() / (S.abs().max() * x.abs().max() + b.abs().max()))
    worst = max(worst, err)
    assert err < 1e-11, (n, err)
print({"solves": count, "worst_residual": worst, "seconds": round(time.time() - t0, 1)})
