#!/usr/bin/env python
"""Device time of the VKDE prepare_kernel stages (kNN sort, covariance + Cholesky) through the C ABI."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from numcosmo_b200 import capi

cases = [(2048, 10, 102), (16384, 20, 819), (16384, 30, 819)] if len(sys.argv) < 4 else [tuple(int(a) for a in sys.argv[1:4])]
ctx = capi.Context(0)
for n, d, k in cases:
    rs = np.random.default_rng(n + d)
    X = rs.normal(size=(n, d))
    ctx.set_kernel(capi.KERNEL_GAUSS, 3.0, d)
    ctx.vkde_prepare(X, X, n, k)
    ctx.enable_timers(True)
    ctx.reset_timers()
    t0 = time.perf_counter()
    U, fail = ctx.vkde_prepare(X, X, n, k)
    wall = time.perf_counter() - t0
    tm, _ = ctx.get_timers()
    ctx.enable_timers(False)
    print(json.dumps({"n": n, "d": d, "k": k, "prep_ms": tm["prep"], "h2d_ms": tm["h2d"], "d2h_ms": tm["d2h"], "wall_ms": wall * 1e3, "fail": int(fail.sum())}), flush=True)
