#!/usr/bin/env python
"""Debugging aid: runs one prepare_interp-style NNLS on the GPU and on the CPU oracle with the per-system traces on
(NCM_SD_GPU_NNLS_TRACE / ORC_NNLS_TRACE) and reports where the two sequences of passive-set sizes part.
usage: python tools/nnls_trace_compare.py kde|vkde gauss|st d n [seed]"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import sys, numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
from oracle import ncm_oracle as O
from numcosmo_b200 import capi
from helpers import make_sd, mvnd_problem, upload_from_oracle
sd_s, k_s, d, n, seed, which = %(args)r
sd_type = O.SD_KDE if sd_s == "kde" else O.SD_VKDE
kernel = O.KERNEL_GAUSS if k_s == "gauss" else O.KERNEL_ST
mu, cov, X, m2lnL = mvnd_problem(O, d, n, seed=seed)
if which == "oracle":
    sd = make_sd(O, sd_type, kernel, 3.0, X, m2lnp=m2lnL)
    print("STATS", sd.nnls_stats(), file=sys.stderr)
else:
    import os
    tr = os.environ.pop("ORC_NNLS_TRACE", None)
    sd = make_sd(O, sd_type, kernel, 3.0, X)
    ctx = capi.Context(0)
    upload_from_oracle(ctx, capi, O, sd, sd_type, kernel, 3.0, X)
    f = np.exp(-0.5 * (m2lnL - m2lnL.min()))
    ctx.compute_IM(1.0 / f, fetch=False, nrows=n)
    x, rnorm, st = ctx.nnls_solve()
    print("STATS", st, file=sys.stderr)
'''


def run(args, which, env_extra):
    env = dict(os.environ, **env_extra)
    p = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT, "args": tuple(args) + (which,)}], env=env, capture_output=True, text=True)
    if p.returncode != 0:
        print(p.stderr[-2000:])
        raise SystemExit(1)
    return p.stderr.splitlines()


def main():
    sd_s, k_s, d, n = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    seed = int(sys.argv[5]) if len(sys.argv) > 5 else 900 + d
    args = (sd_s, k_s, d, n, seed)
    o = run(args, "oracle", {"ORC_NNLS_TRACE": "1"})
    so = [int(m.group(1)) for l in o for m in [re.search(r"orc_nnls: chol \|P\| = (\d+)", l)] if m]
    for label, env in (("reuse", {"NCM_SD_GPU_NNLS_TRACE": "1"}), ("fresh", {"NCM_SD_GPU_NNLS_TRACE": "1", "NCM_SD_GPU_NNLS_REUSE": "0"})):
        g = run(args, "gpu", env)
        lines = [l for l in g if l.startswith("gpu_nnls:")]
        sg = [int(re.search(r"\|P\| = (\d+)", l).group(1)) for l in lines]
        # a fallback prints a lowrank line and then the chol line of the same system: drop the first of the pair
        keep, seq = [], []
        for i, l in enumerate(lines):
            fb = "lowrank" in l and (float(re.search(r"corr = (\S+)", l).group(1)) > 1e-7 or " info = 0" not in l)
            if not fb:
                keep.append(l)
                seq.append(sg[i])
        first = next((i for i, (a, b) in enumerate(zip(seq, so)) if a != b), None)
        print(f"[{label}] gpu systems {len(seq)} oracle {len(so)}; first divergence at system {first}", [l for l in g if l.startswith('STATS')])
        if first is not None:
            for l in keep[max(0, first - 3):first + 2]:
                print("   ", l)
            print("    oracle:", so[max(0, first - 3):first + 2])
        corr = [float(re.search(r"corr = (\S+)", l).group(1)) for l in lines if "lowrank" in l]
        if corr:
            import numpy as np
            print(f"    low-rank corrections: median {np.median(corr):.2e} max {np.max(corr):.2e}")


if __name__ == "__main__":
    main()
