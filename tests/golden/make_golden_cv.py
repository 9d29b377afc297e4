#!/usr/bin/env python
"""Generates tests/golden/cv_robust_v1.npz: seeded inputs and the CPU oracle's outputs for the cross-validation modes and the robust
covariance types.  The CV_SPLIT cases are produced with the oracle routed through the REFERENCE'S OWN
dlevmar_dif (oracle/_ref/liblevmar_ref.so, compiled in place from numcosmo/external/levmar); the restatement reproduces them bit for bit
(tests/test_golden_cv.py).  Regenerate with:  python tests/golden/make_golden_cv.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import mvnd_problem  # noqa: E402
from oracle import ncm_oracle as O  # noqa: E402

CASES = [
    # name, sd, kernel, nu, d, n, cv, cov_type, split_frac, local_frac
    ("cv_split_vkde_gauss_d3", O.SD_VKDE, O.KERNEL_GAUSS, 3.0, 3, 200, O.CV_SPLIT, O.COV_SAMPLE, 0.5, 0.05),
    ("cv_split_kde_st3_d5", O.SD_KDE, O.KERNEL_ST, 3.0, 5, 260, O.CV_SPLIT, O.COV_SAMPLE, 0.5, 0.05),
    ("cv_nofit_kde_st3_d4", O.SD_KDE, O.KERNEL_ST, 3.0, 4, 240, O.CV_SPLIT_NOFIT, O.COV_SAMPLE, 0.6, 0.05),
    ("cv_nofit_vkde_cauchy_d2", O.SD_VKDE, O.KERNEL_ST, 1.0, 2, 220, O.CV_SPLIT_NOFIT, O.COV_SAMPLE, 0.7, 0.05),
    ("cv_loo_kde_gauss_d2", O.SD_KDE, O.KERNEL_GAUSS, 3.0, 2, 160, O.CV_LOO, O.COV_SAMPLE, 0.5, 0.05),
    ("cv_loo_vkde_gauss_d3", O.SD_VKDE, O.KERNEL_GAUSS, 3.0, 3, 150, O.CV_LOO, O.COV_SAMPLE, 0.5, 0.08),
    ("robust_diag_kde_gauss_d4", O.SD_KDE, O.KERNEL_GAUSS, 3.0, 4, 300, O.CV_NONE, O.COV_ROBUST_DIAG, 0.5, 0.05),
    ("robust_ogk_kde_st3_d4", O.SD_KDE, O.KERNEL_ST, 3.0, 4, 300, O.CV_NONE, O.COV_ROBUST, 0.5, 0.05),
    ("robust_ogk_vkde_gauss_d3", O.SD_VKDE, O.KERNEL_GAUSS, 3.0, 3, 200, O.CV_NONE, O.COV_ROBUST, 0.5, 0.1),
]


def run_case(sd_type, kernel, nu, d, cv, cov_type, split_frac, local_frac, X, m2lnL):
    sd = O.StatsDist(int(sd_type), int(kernel), int(d), float(nu), int(cv))
    sd.set_cov_type(int(cov_type))
    sd.set_split_frac(float(split_frac))
    sd.set_local_frac(float(local_frac))
    sd.set_use_threads(False)
    sd.add_obs_matrix(X)
    assert sd.prepare_interp(m2lnL) == 0
    return sd


def main():
    O.lib().orc_set_blas_threads(1)
    have_ref = O.use_ref_levmar(True)
    out = {}
    try:
        for name, sd_type, kernel, nu, d, n, cv, cov_type, sf, lf in CASES:
            mu, cov, X, m2lnL = mvnd_problem(O, d, n, seed=700 + d + n)
            if cov_type != O.COV_SAMPLE:   # outliers are what the robust types are for
                X = X.copy()
                X[::20] += 25.0 * np.sqrt(np.diag(cov)) * np.random.default_rng(n).standard_normal((len(X[::20]), d))
                m2lnL = np.minimum(np.einsum("ij,jk,ik->i", X - mu, np.linalg.inv(cov), X - mu), 100.0)
            sd = run_case(sd_type, kernel, nu, d, cv, cov_type, sf, lf, X, m2lnL)
            Q = np.vstack([X[:30] + 0.01, mu + 2.5 * (X[30:60] - mu)])
            lnos, val = sd.cv_trace()
            out[f"{name}/X"] = X
            out[f"{name}/m2lnL"] = m2lnL
            out[f"{name}/Q"] = Q
            out[f"{name}/lnos"] = lnos
            out[f"{name}/val"] = val
            out[f"{name}/over_smooth"] = np.array([sd.get_over_smooth()])
            out[f"{name}/weights"] = sd.peek_weights().copy()
            out[f"{name}/m2lnp"] = sd.eval_m2lnp_batch(Q, 1)
            out[f"{name}/cov"] = np.triu(sd.peek_full_cov())
            out[f"{name}/meta"] = np.array([sd_type, kernel, nu, d, n, cv, cov_type, sf, lf, float(have_ref and cv == O.CV_SPLIT)], dtype=np.float64)
    finally:
        O.use_ref_levmar(False)
    path = os.path.join(HERE, "cv_robust_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, sum(v.nbytes for v in out.values()), "bytes raw; CV_SPLIT through the reference's levmar:", have_ref)


if __name__ == "__main__":
    main()
