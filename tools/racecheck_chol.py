#!/usr/bin/env python
"""A handful of small single-launch Cholesky solves, meant to run under `compute-sanitizer --tool racecheck`
(shared-memory hazards of the warp-level data flow in chol_fused.cu)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from numcosmo_b200 import capi

ctx = capi.Context(0)
for n in (int(a) for a in sys.argv[1:]) if len(sys.argv) > 1 else (64, 130, 200, 520):
    B = torch.randn((n + 8, n), dtype=torch.float64, device="cuda")
    S = B.T @ B + 0.05 * torch.eye(n, dtype=torch.float64, device="cuda")
    ld = (n + 7) // 8 * 8
    M = torch.zeros((n, ld), dtype=torch.float64, device="cuda")
    M[:, :n] = torch.triu(S)
    b = torch.randn(n, dtype=torch.float64, device="cuda")
    x = b.clone()
    torch.cuda.synchronize()
    assert ctx.dposv_upper_dev(n, M.data_ptr(), ld, x.data_ptr()) == 0
    print(n, float((S @ x - b).abs().max()))
