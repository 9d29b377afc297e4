#!/usr/bin/env python
"""Generates tests/golden/apes_path_v1.npz: seeded inputs and the CPU oracle's outputs for the hot path.

The reference (NumCosmo v0.27.0) cannot be built or imported in the development container (no GLib / GSL / meson,
SURVEY.md section 8c) and holds no golden vectors for this path, so these fixtures are outputs of the line-by-line C
restatement under oracle/ (pinned on the reference's closed-form known answers, tests/test_oracle_known_answers.py).
They freeze that oracle: any later change of the oracle or of the CUDA path that moves a number shows up as a diff
against a committed file.  Regenerate with:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import make_sd, mvnd_problem  # noqa: E402
from oracle import ncm_oracle as O  # noqa: E402

CASES = [
    # name, sd, kernel, nu, d, n, local_frac
    ("vkde_gauss_d10", O.SD_VKDE, O.KERNEL_GAUSS, 3.0, 10, 300, 0.05),
    ("vkde_st3_d4", O.SD_VKDE, O.KERNEL_ST, 3.0, 4, 257, 0.05),
    ("vkde_cauchy_d2", O.SD_VKDE, O.KERNEL_ST, 1.0, 2, 200, 0.05),
    ("vkde_gauss_d24", O.SD_VKDE, O.KERNEL_GAUSS, 3.0, 24, 320, 0.12),
    ("kde_gauss_d6", O.SD_KDE, O.KERNEL_GAUSS, 3.0, 6, 310, 0.05),
    ("kde_st3_d10", O.SD_KDE, O.KERNEL_ST, 3.0, 10, 300, 0.05),
]


def main():
    O.lib().orc_set_blas_threads(1)   # fixed summation order inside OpenBLAS: the file must not depend on the core count
    out = {}
    for name, sd_type, kernel, nu, d, n, lf in CASES:
        mu, cov, X, m2lnL = mvnd_problem(O, d, n, seed=900 + d)
        sd = make_sd(O, sd_type, kernel, nu, X, m2lnp=m2lnL, local_frac=lf, use_threads=False)
        Q = np.vstack([X[:40] + 0.01, mu + 2.5 * (X[40:90] - mu)])
        out[f"{name}/X"] = X
        out[f"{name}/m2lnL"] = m2lnL
        out[f"{name}/Q"] = Q
        out[f"{name}/weights"] = sd.peek_weights().copy()
        out[f"{name}/m2lnp"] = sd.eval_m2lnp_batch(Q, 1)
        out[f"{name}/IM_sub"] = sd.compute_IM()[::7, ::5].copy()   # every 7th row, 5th column: keeps the fixture small
        out[f"{name}/href"] = np.array([sd.get_href()])
        out[f"{name}/rnorm"] = np.array([sd.get_rnorm()])
        out[f"{name}/meta"] = np.array([sd_type, kernel, nu, d, n, lf], dtype=np.float64)
    # APES accepted-sample sequences for a fixed stream
    for name, kernel, nu, d, W, iters, seed in (("apes_mvnd_gauss_d5", O.KERNEL_GAUSS, 3.0, 5, 300, 4, 21), ("apes_mvnd_st3_d3", O.KERNEL_ST, 3.0, 3, 200, 4, 22)):
        mu, cov, X, m2lnL = mvnd_problem(O, d, W, seed=950 + d)
        lb, ub = np.full(d, -50.0), np.full(d, 50.0)
        tgt = O.Target(O.TARGET_MVND, d, lb, ub, mu=mu, cov=cov)
        th, ml = X.copy(), m2lnL.copy()
        acc = O.APES(W, d, O.SD_VKDE, kernel, nu, use_threads=False).run(tgt, th, ml, iters, O.RNG(seed), nthreads=1)
        out[f"{name}/X"] = X
        out[f"{name}/m2lnL"] = m2lnL
        out[f"{name}/mu"] = mu
        out[f"{name}/U"] = np.asarray(tgt.U)
        out[f"{name}/accepted"] = np.asarray(acc).astype(np.uint8)
        out[f"{name}/theta_final"] = th
        out[f"{name}/meta"] = np.array([kernel, nu, d, W, iters, seed], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "apes_path_v1.npz"), **out)
    print("wrote", os.path.join(HERE, "apes_path_v1.npz"), sum(v.nbytes for v in out.values()), "bytes raw")


if __name__ == "__main__":
    main()
