#!/usr/bin/env python
"""CUDA-event timings of the FP64 building blocks: DMMA SYRK and the blocked Cholesky (factor only)."""
import json
import sys, os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from numcosmo_b200 import capi

ctx = capi.Context(0)
st = torch.cuda.ExternalStream(ctx.stream)
sizes = [int(a) for a in sys.argv[1:]] or [1024, 2048, 4096, 8192, 16384]
for n in sizes:
    ld = n
    A = torch.randn((n, ld), dtype=torch.float64, device="cuda")
    M = torch.empty((n, ld), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); ctx.dsyrk_ata_dev(n, n, A.data_ptr(), ld, M.data_ptr(), ld); e1.record(st); ctx.synchronize()
        best = min(best, e0.elapsed_time(e1))
    syrk = {"n": n, "syrk_ms": best, "syrk_tflops": n**3 / best / 1e9}
    M2 = M + n * torch.eye(n, dtype=torch.float64, device="cuda")
    best = 1e9
    for _ in range(3):
        W = M2.clone(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); info = ctx.dpotrf_upper_dev(n, W.data_ptr(), ld); e1.record(st); ctx.synchronize()
        assert info == 0
        best = min(best, e0.elapsed_time(e1))
    syrk.update({"potrf_ms": best, "potrf_tflops": n**3 / 3 / best / 1e9})
    rhs = torch.randn(n, dtype=torch.float64, device="cuda")
    best = 1e9
    for _ in range(4):
        W = M2.clone(); r = rhs.clone(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); info = ctx.dposv_upper_dev(n, W.data_ptr(), ld, r.data_ptr()); e1.record(st); ctx.synchronize()
        assert info == 0
        best = min(best, e0.elapsed_time(e1))
    syrk.update({"posv_ms": best, "posv_tflops": n**3 / 3 / best / 1e9})
    print(json.dumps(syrk), flush=True)
    del A, M, M2, W
