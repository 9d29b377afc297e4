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
