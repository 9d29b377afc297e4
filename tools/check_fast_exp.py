#!/usr/bin/env python
"""Accuracy of the exp_nonpos_fast scheme of numcosmo_b200/csrc/common.cuh, restated in numpy (no FMA): k = rint(x log2 e),
r = (x - k ln2) / 4, degree-9 Taylor, two squarings.  Prints the maximum and mean relative error against numpy's exp."""
import math

import numpy as np


def exp_fast(x):
    x = np.asarray(x, dtype=np.float64)
    magic = 6755399441055744.0
    t = x * 1.4426950408889634 + magic
    kf = t - magic
    k = kf.astype(np.int64)
    r = x + kf * (-6.93147180369123816490e-01)
    r = r + kf * (-1.90821492927058770002e-10)
    r = r * 0.25
    c = [1.0 / math.factorial(i) for i in range(10)]
    p = np.full_like(r, c[9])
    for i in range(8, -1, -1):
        p = p * r + c[i]
    p = p * p
    p = p * p
    return np.ldexp(p, k)


if __name__ == "__main__":
    rs = np.random.default_rng(0)
    x = -np.abs(rs.normal(size=4_000_000)) * rs.choice([1e-3, 0.01, 1, 10, 100, 300], size=4_000_000)
    x = x[x > -700]
    ref = np.exp(x)
    e = np.abs(exp_fast(x) - ref) / ref
    print({"max_rel_err": float(e.max()), "mean_rel_err": float(e.mean()), "n": int(x.size)})


def log1p_fast(x):
    """log1p_nonneg_fast of common.cuh restated in numpy."""
    c = 1.0 + (np.arange(32) + 0.5) / 32
    invc, lnc = 1.0 / c, np.log(c)
    u = 1.0 + np.asarray(x, dtype=np.float64)
    bits = u.view(np.int64)
    e = ((bits >> 52) & 0x7FF) - 1023
    j = (bits >> 47) & 31
    m = ((bits & ((1 << 52) - 1)) | (1023 << 52)).view(np.float64)
    r = m * invc[j] - 1.0
    p = np.full_like(r, 1.0 / 7)
    for k in (6, 5, 4, 3, 2):
        p = p * r + ((-1.0) ** (k + 1)) / k
    p = (p * r + 1.0) * r
    return (e * 6.93147180369123816490e-01 + lnc[j]) + (p + e * 1.90821492927058770002e-10)


if __name__ == "__main__":
    rs = np.random.default_rng(1)
    x = np.abs(rs.normal(size=3_000_000)) * rs.choice([1e-6, 1e-3, 0.1, 1, 10, 1e3, 1e6, 1e12, 1e18], size=3_000_000)
    ae = np.abs(log1p_fast(x) - np.log1p(x))
    print({"log1p_max_abs_err": float(ae.max()), "log1p_max_abs_err_x_below_1": float(ae[x < 1].max())})
