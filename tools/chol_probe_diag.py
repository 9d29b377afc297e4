#!/usr/bin/env python
"""clock64 stamps of the eight warps inside diag_factor (chol_fused.cu built with -DNCM_FUSED_PROBE): n = 64, one diagonal block."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from numcosmo_b200 import capi

n = 64
ctx = capi.Context(0)
L = capi.load()
L.ncm_sd_gpu_chol_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
nsm, cap = 148, 2048
tr = torch.zeros((nsm, cap, 2), dtype=torch.int64, device="cuda")
A = torch.randn((n + 10, n), dtype=torch.float64, device="cuda")
M = A.T @ A + 0.1 * torch.eye(n, dtype=torch.float64, device="cuda")
for it in range(4):
    W = M.clone()
    tr.zero_()
    torch.cuda.synchronize()
    L.ncm_sd_gpu_chol_trace(ctx._h, C.c_void_p(tr.data_ptr()) if it == 3 else None, cap)
    assert ctx.dpotrf_upper_dev(n, W.data_ptr(), n) == 0
t = tr.cpu().numpy()
t0 = min(int(t[140 + w, 1, 0]) for w in range(8) if t[140 + w, 0, 0] > 0)
for w in range(8):
    cnt = int(t[140 + w, 0, 0])
    print("warp", w, " ".join(f"{int(t[140 + w, i, 1])}:{int(t[140 + w, i, 0]) - t0}" for i in range(1, cnt + 1)))
