"""The accessors of the host mirror on CPU (no device call is made: the GPU context is only created by prepare): the reference's
`sanity` case of tests/c/ncm/stats/test_ncm_stats_dist.c:371-421 and the constructor block of
tests/c/ncm/fit/test_ncm_fit_esmcmc.c:129-227, with the range checks of the setters (ncm_stats_dist.c:1307-1308, 1342-1343;
ncm_fit_esmcmc_walker_apes.c:1178-1227), which the reference enforces with g_assert / g_error and the mirror reports as NcmError."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def S():
    from numcosmo_b200 import stats_dist as S

    return S


@pytest.mark.parametrize("cls,kern", [("StatsDistKDE", "gauss"), ("StatsDistKDE", "st"), ("StatsDistVKDE", "gauss"), ("StatsDistVKDE", "st")])
def test_stats_dist_sanity(S, cls, kern):
    d = 3
    k = S.StatsDistKernelGauss(d) if kern == "gauss" else S.StatsDistKernelST(d, 3.7)
    sd = getattr(S, cls)(k, S.StatsDistCV.NONE)
    sd.set_use_threads(True)
    assert sd.get_use_threads()
    sd.set_use_threads(False)
    assert not sd.get_use_threads()
    sd.set_over_smooth(1.2)
    assert sd.get_over_smooth() == 1.2
    sd.set_split_frac(0.2)
    assert sd.get_split_frac() == 0.2
    sd.set_cv_type(S.StatsDistCV.NONE)
    assert sd.get_cv_type() == S.StatsDistCV.NONE
    sd.set_cv_type(S.StatsDistCV.SPLIT)
    assert sd.get_cv_type() == S.StatsDistCV.SPLIT
    assert sd.get_dim() == d and sd.peek_kernel() is k and k.get_dim() == d
    # the function-level asserts: 0.01 <= split_frac <= 1, 0 <= shrink <= 1; the value is left alone on failure
    for bad in (0.009, 1.01, float("nan")):
        with pytest.raises(S.NcmError, match="split_frac"):
            sd.set_split_frac(bad)
    sd.set_split_frac(0.01)
    sd.set_split_frac(1.0)
    sd.set_split_frac(0.2)
    for bad in (-1e-3, 1.5):
        with pytest.raises(S.NcmError, match="shrink"):
            sd.set_shrink(bad)
    sd.set_shrink(0.0)
    sd.set_shrink(1.0)
    assert sd.get_shrink() == 1.0 and sd.get_split_frac() == 0.2


@pytest.mark.parametrize("case", range(6))
def test_apes_constructors_and_setters(S, case):
    M, K = S.FitESMCMCWalkerAPESMethod, S.FitESMCMCWalkerAPESKType
    d = 1 + case % 3
    W = 100 * d
    if case == 0:                                                   # ncm_fit_esmcmc_walker_apes_new: VKDE, Cauchy, over_smooth 1, interpolation
        ap = S.FitESMCMCWalkerAPES(W, d)
        method, k_type = M.VKDE, K.CAUCHY
    else:
        method, k_type = [(M.VKDE, K.ST3), (M.VKDE, K.GAUSS), (M.KDE, K.CAUCHY), (M.KDE, K.ST3), (M.KDE, K.GAUSS)][case - 1]
        ap = S.FitESMCMCWalkerAPES.new_full(W, d, method, k_type, 1.0, True)
    assert ap.get_method() == method and ap.get_k_type() == k_type and ap.interp() and ap.get_over_smooth() == 1.0
    # defaults of the property specs (walker_apes.c:358-475)
    assert ap.get_shrink() == 0.01 and ap.get_random_walk_prob() == 0.02 and not ap.get_use_threads()
    ap.set_over_smooth(1.01)
    assert ap.get_over_smooth() == 1.01
    sd0, sd1 = ap.peek_sds()
    assert sd0.get_over_smooth() == 1.01 and sd1.get_over_smooth() == 1.01     # forwarded at once (:1162-1166)
    assert sd0.get_dim() == d and sd1.get_dim() == d
    ap.use_interp(False)
    assert not ap.interp()
    ap.set_use_threads(True)
    assert ap.get_use_threads() and sd0.get_use_threads() and sd1.get_use_threads()     # forwarded (:1349-1361)
    sd1.set_use_threads(False)
    with pytest.raises(S.NcmError, match="use_threads0 == use_threads1"):               # the getter's own asserts (:1378-1391)
        ap.get_use_threads()
    sd1.set_use_threads(True)
    # range checks, :1178-1227
    for bad in (-0.1, 1.1):
        with pytest.raises(S.NcmError, match="invalid shrink"):
            ap.set_shrink(bad)
        with pytest.raises(S.NcmError, match="invalid probability"):
            ap.set_random_walk_prob(bad)
    for bad in (0.0, -2.0):
        with pytest.raises(S.NcmError, match="invalid scale"):
            ap.set_random_walk_scale(bad)
    ap.set_shrink(0.3)
    ap.set_random_walk_prob(0.5)
    ap.set_random_walk_scale(2.0)
    assert ap.get_shrink() == 0.3 and ap.get_random_walk_prob() == 0.5 and ap.get_random_walk_scale() == 2.0
    # the stored shrink does not reach the objects built in set_sys (:1178-1187 only stores)
    assert sd0.get_shrink() == 0.01 and sd1.get_shrink() == 0.01


def test_apes_vkde_needs_enough_walkers_per_block(S):
    """_ncm_fit_esmcmc_walker_apes_vkde_check_sizes: local_frac * nwalkers / 2 < 2 is refused for METHOD_VKDE only (walker_apes.c:490-507, 563-572)."""
    M, K = S.FitESMCMCWalkerAPESMethod, S.FitESMCMCWalkerAPESKType
    with pytest.raises(S.NcmError, match="too low"):
        S.FitESMCMCWalkerAPES.new_full(40, 2, M.VKDE, K.CAUCHY, 1.0, True)
    S.FitESMCMCWalkerAPES.new_full(40, 2, M.KDE, K.CAUCHY, 1.0, True)
    S.FitESMCMCWalkerAPES.new_full(80, 2, M.VKDE, K.CAUCHY, 1.0, True)


def test_precondition_failures_before_any_device_call(S):
    """g_assert / g_error sites of the classes that are decided on the host: they must fire here exactly as on a GPU box (no device is
    touched before them), with the reference's messages."""
    d = 3
    kern = S.StatsDistKernelGauss(d)
    sd = S.StatsDistKDE(kern, S.StatsDistCV.NONE)
    with pytest.raises(S.NcmError, match=r"ncm_vector_len \(y\) == d"):                  # ncm_stats_dist.c add_obs
        sd.add_obs(np.zeros(d + 1))
    for i in range(d):
        sd.add_obs(np.arange(float(d)) + i)
    with pytest.raises(S.NcmError, match="the sample is too small"):                     # ncm_stats_dist.c:748-749, n_obs <= d
        sd.prepare()
    with pytest.raises(S.NcmError, match="cov_fixed"):                                   # ncm_stats_dist_kde.c:831-832
        sd.set_cov_fixed(np.eye(d + 1))
    sd.set_cov_type(S.StatsDistKDECovType.FIXED)
    with pytest.raises(S.NcmError, match="not positive definite"):                       # :843
        sd.set_cov_fixed(np.diag([1.0, -1.0, 1.0]))
    sd.set_cov_fixed(np.eye(d))
    v = S.StatsDistVKDE(kern, S.StatsDistCV.NONE)
    for bad in (0.0009, 1.01):                                                           # ncm_stats_dist_vkde.c:808-809
        with pytest.raises(S.NcmError, match="local_frac"):
            v.set_local_frac(bad)
    v.set_local_frac(0.001)
    v.set_local_frac(1.0)
    assert v.get_local_frac() == 1.0
    v.set_local_frac(0.05)
    rs = np.random.default_rng(0)
    for _ in range(30):
        v.add_obs(rs.standard_normal(d))
    with pytest.raises(S.NcmError, match="Too few observations"):                        # :505-510, local_frac n_obs = 1.5 < 2
        v.prepare()
    with pytest.raises(S.NcmError, match="outside"):                                     # the one limit of this implementation: d <= 32
        S.StatsDistKernelGauss(33)
    with pytest.raises(S.NcmError, match="outside"):
        S.StatsDistKernelST(0, 3.0)


def test_apes_set_sys_rebuilds_only_on_a_changed_configuration(S):
    """_ncm_fit_esmcmc_walker_apes_set_sys (walker_apes.c:509-598): the two objects are rebuilt when size, dimension, method or kernel
    type change -- and then start from their own defaults again, the walker holds no local fraction of its own -- and are left alone
    otherwise; set_local_frac is a VKDE-method call that re-checks the block size (:1428-1439); the enum setters refuse values past
    their _LEN (:1110-1144)."""
    M, K = S.FitESMCMCWalkerAPESMethod, S.FitESMCMCWalkerAPESKType
    ap = S.FitESMCMCWalkerAPES.new_full(200, 2, M.VKDE, K.CAUCHY, 1.3, True)
    ap.set_local_frac(0.2)
    sd0, sd1 = ap.peek_sds()
    assert sd0.get_local_frac() == 0.2 and sd1.get_local_frac() == 0.2
    # same configuration: nothing is rebuilt, the objects keep what they were given
    ap.set_method(M.VKDE)
    ap.set_k_type(K.CAUCHY)
    sd0, sd1 = ap.peek_sds()
    assert sd0.get_local_frac() == 0.2 and sd0.get_over_smooth() == 1.3
    # a new kernel type: fresh objects, the local fraction is the class default again, over_smooth is the walker's
    ap.set_k_type(K.ST3)
    sd0, sd1 = ap.peek_sds()
    assert sd0.get_local_frac() == 0.05 and sd1.get_local_frac() == 0.05 and sd0.get_over_smooth() == 1.3
    assert ap.get_k_type() == K.ST3
    # too small a fraction for the block: 0.01 * 100 = 1 < 2
    with pytest.raises(S.NcmError, match="too low"):
        ap.set_local_frac(0.01)
    with pytest.raises(S.NcmError, match="local_frac"):                                  # the object's own range assert, nothing forwarded
        ap.set_local_frac(2.0)
    ap.set_method(M.KDE)
    with pytest.raises(S.NcmError, match="non-VKDE"):
        ap.set_local_frac(0.2)
    with pytest.raises(S.NcmError, match="invalid method"):
        ap.set_method(2)
    with pytest.raises(S.NcmError, match="invalid method"):
        ap.set_k_type(3)
    assert ap.get_method() == M.KDE and ap.get_k_type() == K.ST3
