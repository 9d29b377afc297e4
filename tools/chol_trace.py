#!/usr/bin/env python
"""Event trace of the single-launch Cholesky (chol_fused.cu): where does a phase spend its time?"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from numcosmo_b200 import capi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
ctx = capi.Context(0)
L = capi.load()
L.ncm_sd_gpu_chol_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
nsm, cap = 148, 2048
tr = torch.zeros((nsm, cap, 2), dtype=torch.int64, device="cuda")
A = torch.randn((n + 10, n), dtype=torch.float64, device="cuda")
M = A.T @ A + 0.1 * torch.eye(n, dtype=torch.float64, device="cuda")
rhs = torch.randn(n, dtype=torch.float64, device="cuda")
for it in range(3):
    W, r = M.clone(), rhs.clone()
    tr.zero_()
    torch.cuda.synchronize()
    L.ncm_sd_gpu_chol_trace(ctx._h, C.c_void_p(tr.data_ptr()) if it == 2 else None, cap)
    assert ctx.dposv_upper_dev(n, W.data_ptr(), n, r.data_ptr()) == 0
t = tr.cpu().numpy()
ev = []
for b in range(nsm):
    cnt = int(t[b, 0, 0])
    for i in range(1, cnt + 1):
        code = int(t[b, i, 1])
        ev.append((int(t[b, i, 0]), b, code >> 24, (code >> 12) & 0xFFF, code & 0xFFF))
ev.sort()
t0 = ev[0][0]
names = {1: "diag", 2: "panel", 3: "update", 4: "bsolve", 5: "bprod"}
fine = [e for e in ev if e[2] in (6, 7)]
ev = [e for e in ev if e[2] not in (6, 7)]
# durations per type: start -> (after waits) -> end
open_ev = {}
stats = {}
for ts, b, ty, x, y in ev:
    base = ty & 7
    key = (b, base, x, y)
    if ty < 8:
        open_ev[key] = [ts, ts]
    elif ty & 16:
        if key in open_ev:
            open_ev[key][1] = ts
    else:
        s0, s1 = open_ev.pop(key)
        st = stats.setdefault(names[base], {"n": 0, "wait_us": 0.0, "work_us": 0.0})
        st["n"] += 1
        st["wait_us"] += (s1 - s0) / 1e3
        st["work_us"] += (ts - s1) / 1e3
for k, v in stats.items():
    print(json.dumps({"op": k, "n": v["n"], "avg_wait_us": v["wait_us"] / v["n"], "avg_work_us": v["work_us"] / v["n"]}))
# critical chain: time of diag end per k
dend = {x: ts for ts, b, ty, x, y in ev if ty == 9}
ks = sorted(dend)
print(json.dumps({"total_us": (ev[-1][0] - t0) / 1e3, "factor_us": (dend[ks[-1]] - t0) / 1e3,
                  "diag_end_us": [round((dend[k] - t0) / 1e3, 1) for k in ks]}))
# detail of phases 10 and 11 on the critical path
for ts, b, ty, x, y in ev:
    if (ty & 7) in (1, 2, 3) and ((ty & 7) == 1 and x in (10, 11) or (ty & 7) == 2 and x == 10 and y == 11 or (ty & 7) == 3 and x == 11 and y == 11):
        print(names[ty & 7], "start" if ty < 8 else ("ready" if ty & 16 else "end"), x, y, "cta", b, round((ts - t0) / 1e3, 2))

# the spine's own timeline over phases 20 and 21: 6/0 after diag_factor, 6/1 after diag_store, 7/0 after panel_solve, 7/1 after tile_store
sp = [(ts, ty, x, y) for ts, b, ty, x, y in sorted([e for e in ev if e[1] == 0] + [e for e in fine if e[1] == 0]) if x in (20, 21) or (ty & 7) == 3 and x in (21, 22)]
if sp:
    tb = sp[0][0]
    print("spine:", " ".join(f"{ty}/{x}/{y}@{(ts - tb) / 1e3:.2f}" for ts, ty, x, y in sp))
# fine-grained: sub-block boundaries inside the diag of k = 10 (cta of diag 10) and the panel (10, 11)
d10 = [e for e in ev if e[2] == 1 and e[3] == 10][0]
p10 = [e for e in ev if e[2] == 2 and e[3] == 10 and e[4] == 11][0]
for nm, ty, start in (("diag10", 6, d10), ("panel10_11", 7, p10)):
    end = [e for e in ev if e[1] == start[1] and e[0] > start[0] and (e[2] & 8) and not (e[2] & 16)][0]
    pts = [(e[0] - start[0]) / 1e3 for e in fine if e[1] == start[1] and e[2] == ty and start[0] <= e[0] <= end[0]]
    if not pts:
        continue
    print(nm, "cta", start[1], "total", (end[0] - start[0]) / 1e3, [round(x, 2) for x in pts])
