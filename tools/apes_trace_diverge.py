#!/usr/bin/env python
"""Debugging aid: runs the configs[0] chain (APES, VKDE object, Cauchy kernel, 2-D Rosenbrock, 400 walkers) on the CPU oracle and on
the GPU path with the per-system NNLS traces on, and reports where the two sequences of passive-set systems part.
usage: python tools/apes_trace_diverge.py [iters] [outdir]"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
from oracle import ncm_oracle as oracle
which, iters = %(which)r, %(iters)d
W, d = 400, 2
lb, ub = np.array([-200.0, -400.0]), np.array([200.0, 800.0])
tgt = oracle.Target(oracle.TARGET_ROSENBROCK, d, lb, ub)
r = oracle.RNG(1234)
theta = np.ascontiguousarray(np.array([[r.gaussian(1.0) for _ in range(d)] for _ in range(W)]) * [1.0, 2.0] + [0.5, 1.0])
m2lnL0 = np.array([tgt.m2lnL(x) for x in theta])
th, ml = theta.copy(), m2lnL0.copy()
if which == "oracle":
    ao = oracle.APES(W, d, oracle.SD_VKDE, oracle.KERNEL_ST, 1.0, over_smooth=1.1, use_interp=True, use_threads=True)
    acc = ao.run(tgt, th, ml, iters, oracle.RNG(4321), nthreads=4)
else:
    from numcosmo_b200 import stats_dist as S
    ag = S.FitESMCMCWalkerAPES(W, d, S.FitESMCMCWalkerAPESMethod.VKDE, S.FitESMCMCWalkerAPESKType.CAUCHY, 1.1, True)
    ag.set_use_threads(True)
    acc, _ = ag.run("rosenbrock", lb, ub, th, ml, iters, S.RNG(4321))
np.save(%(out)r, np.asarray(acc))
'''


def run(which, iters, out, env_extra):
    env = dict(os.environ, **env_extra)
    p = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT, "which": which, "iters": iters, "out": out}], env=env, capture_output=True, text=True)
    if p.returncode != 0:
        print(p.stderr[-3000:])
        raise SystemExit(1)
    return p.stderr.splitlines()


def main():
    import numpy as np

    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    outdir = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out")
    os.makedirs(outdir, exist_ok=True)
    o = run("oracle", iters, os.path.join(outdir, "trace_acc_oracle.npy"), {"ORC_NNLS_TRACE": "1"})
    g = run("gpu", iters, os.path.join(outdir, "trace_acc_gpu.npy"), {"NCM_SD_GPU_NNLS_TRACE": "1"})
    open(os.path.join(outdir, "trace_oracle.txt"), "w").write("\n".join(o))
    open(os.path.join(outdir, "trace_gpu.txt"), "w").write("\n".join(g))
    so = [l for l in o if l.startswith("orc_nnls:")]
    sg = [l for l in g if l.startswith("gpu_nnls:") and "lowrank" not in l and l.rstrip().endswith("shift = 0")]   # retries of a system: extra lines
    key = lambda l: tuple(int(x) for x in re.search(r"\|P\| = (\d+) info = (-?\d+)", l).groups())
    ko, kg = [key(l) for l in so], [key(l) for l in sg]
    first = next((i for i, (a, b) in enumerate(zip(ko, kg)) if a != b), None)
    print(f"oracle systems {len(ko)}, gpu systems {len(kg)}, first divergence at system {first}")
    if first is not None:
        for i in range(max(0, first - 6), min(len(ko), len(kg), first + 4)):
            print(f"  {i}: oracle {so[i]!r:60s} gpu {sg[i]!r}")
    ao, ag = np.load(os.path.join(outdir, "trace_acc_oracle.npy")), np.load(os.path.join(outdir, "trace_acc_gpu.npy"))
    diff = np.argwhere(ao != ag)
    print("accepted sequences:", "identical" if diff.size == 0 else f"first divergence at (iter, walker) = {diff[0]}")


if __name__ == "__main__":
    main()
