"""The Student-t kernel with an integer number of degrees of freedom on the device (csrc/common.cuh st_pow_u):
(1 + chi2/nu)^(-(nu + d)/2) = r^m, r = rsqrt (1 + chi2/nu), m = nu + d, by straight-line binary powering -- six squarings, the factors
picked by the bits of m, multiplied as (lo * mid) * hi.  Restated here in IEEE double with numpy and compared with the reference's
pow (1 + chi2/nu, kappa) (ncm_stats_dist_kernel_st.c:239-243) evaluated in extended precision: the rounding error must stay within
(1.25 m + 2) units of 2^-52 with a correctly rounded r (measured: 1.08 - 1.17 m) and within (2.25 m + 2) with r one ulp off either way,
which models the device rsqrt (measured: 1.95 - 2.06 m; 1.6e-14 at m = 35) -- i.e. four orders of magnitude inside the 1e-10 relative tolerance of BASELINE.json's north_star for every (nu, d) the path accepts (m < 128)."""
import numpy as np
import pytest

ULP = 2.0 ** -52


def st_pow_u(u, m, r_ulps=0):
    b1 = 1.0 / np.sqrt(1.0 + u)
    for _ in range(abs(r_ulps)):
        b1 = np.nextafter(b1, np.inf if r_ulps > 0 else -np.inf)
    b2 = b1 * b1
    b4 = b2 * b2
    b8 = b4 * b4
    b16 = b8 * b8
    b32 = b16 * b16
    b64 = b32 * b32
    one = np.ones_like(b1)
    lo = (b1 if m & 1 else one) * (b2 if m & 2 else one)
    mid = (b4 if m & 4 else one) * (b8 if m & 8 else one)
    hi = (b16 if m & 16 else one) * (b32 if m & 32 else one) * (b64 if m & 64 else one)
    return (lo * mid) * hi


@pytest.mark.skipif(np.finfo(np.longdouble).eps >= np.finfo(np.float64).eps, reason="no extended precision on this platform")
@pytest.mark.parametrize("m", [2, 3, 4, 5, 7, 11, 13, 23, 33, 35, 63, 64, 65, 95, 127])
def test_integer_power_within_stated_ulps(m):
    rs = np.random.default_rng(m)
    u = np.concatenate([[0.0], 10.0 ** rs.uniform(-12, 4, 20000), rs.uniform(0, 30, 20000)])
    ref = (np.longdouble(1.0) + u.astype(np.longdouble)) ** (np.longdouble(-0.5) * m)
    for r_ulps, bound in ((0, 1.25 * m + 2), (1, 2.25 * m + 2), (-1, 2.25 * m + 2)):
        got = st_pow_u(u, m, r_ulps)
        ok = ref > np.longdouble(1e-290)                                   # above the subnormal range (the linear-domain sum carries a scale)
        rel = np.abs((got.astype(np.longdouble) - ref) / ref)[ok]
        assert float(rel.max()) <= bound * ULP, (m, r_ulps, float(rel.max()) / ULP)
        assert float(rel.max()) < 1e-13
    assert st_pow_u(np.array([0.0]), m)[0] == 1.0                          # chi2 = 0: exactly the kernel's maximum


def test_bit_selection_covers_every_exponent():
    """Every m in 1 .. 127 is the sum of the selected powers: the product of the selected factors of r = 1/2 is exactly 2^-m."""
    for m in range(1, 128):
        assert st_pow_u(np.array([3.0]), m)[0] == 2.0 ** -m
