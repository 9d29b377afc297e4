"""The reference's end-to-end test of the APES walker, restated for the CPU oracle: tests/c/ncm/fit/test_ncm_fit_esmcmc.c.

  new_apes (:129-274)   a random MVND posterior of dimension 1 - 3 (ncm_data_gauss_cov_mvnd_new_full (dim, 2e-2, 5e-2, 30, 1, 2)),
                        100 dim walkers started from a Gaussian with the Fisher covariance at the best fit, over_smooth set to 1.01,
                        walker = APES default (VKDE, Cauchy, interpolation) or new_full (VKDE | KDE, Cauchy | ST3 | Gauss, 1.0, TRUE)
  run      (:433-510)   dim * [15000, 20000) / nrun_div iterations (nrun_div = 1000 for the APES cases), burn-in trimmed, then
                        - ncm_fit_esmcmc_validate: -2 ln L recomputed on the chain agrees with the stored one
                        - variances of the chain against the posterior's: ncm_matrix_cmp_diag (cat, data, 0) < 0.25
                        - correlations against the posterior's: ncm_matrix_cmp (cor_cat, cor_data, 1) < 0.25
                        retried with twice the iterations (up to 15 times) while the two bars are not met; the Gauss kernel cases are
                        skipped by the reference ("APES-Move:*:Gauss walker not supported", :443-448) and are not asserted here either.

The ensemble loop is the oracle's (orc_apes_run follows ncm_fit_esmcmc.c:2235-2288); the reference draws dim / seeds from g_test_rand_*,
here they are swept.  Note that the reference's set_sys builds NcmStatsDistVKDE objects for METHOD_KDE as well
(ncm_fit_esmcmc_walker_apes.c:563-572), so its cases 3 / 4 repeat 0 / 1; the oracle's "kde" arm below runs the true KDE class through the
same ensemble loop, which the reference's suite never does.  This pins the WHOLE oracle path (prepare_kernel, interpolation matrix, NNLS weights, proposal draws, acceptance)
on the one property the reference's own suite demands of it: the chain samples the posterior."""
import numpy as np
import pytest

from helpers import mvnd_problem

TOL = 2.5e-1                     # TEST_NCM_FIT_ESMCMC_TOL, :430
MAX_TRIES = 15                   # :438
CASES = [("vkde", 1.0), ("vkde", 3.0), ("kde", 1.0), ("kde", 3.0)]     # cases 0/3 (Cauchy), 1/4 (ST3); 2/5 (Gauss) skipped as in the reference


def _cmp(a, b, scale):           # ncm_matrix_cmp, ncm_matrix.c:914-940
    return np.max(np.abs((a - b) / (scale + b)))


def _cov2cor(c):
    s = np.sqrt(np.diag(c))
    return c / np.outer(s, s)


@pytest.mark.parametrize("d", [1, 2, 3])
@pytest.mark.parametrize("sd_s,nu", CASES)
def test_esmcmc_run_recovers_posterior_covariance(oracle, sd_s, nu, d):
    O = oracle
    W = 100 * d
    seed = 1000 * d + int(10 * nu) + (7 if sd_s == "kde" else 0)
    mu, cov, X, m2lnL = mvnd_problem(O, d, W, seed=seed)            # initial ensemble ~ N (best fit, Fisher covariance)
    tgt = O.Target(O.TARGET_MVND, d, np.full(d, -50.0), np.full(d, 50.0), mu=mu, cov=cov)
    ap = O.APES(W, d, O.SD_KDE if sd_s == "kde" else O.SD_VKDE, O.KERNEL_ST, nu, over_smooth=1.01, use_interp=True, use_threads=False)
    rng = O.RNG(seed + 1)
    th, ml = X.copy(), np.array([tgt.m2lnL(x) for x in X])
    run = d * 17
    chain, nacc, ntot = [], 0, 0
    ok = False
    for _ in range(MAX_TRIES):
        for _ in range(run):
            acc = ap.run(tgt, th, ml, 1, rng, nthreads=1)
            nacc += int(acc.sum())
            ntot += W
            chain.append(th.copy())
        # validate (:468): the stored -2 ln L is the target's at the stored point
        assert np.allclose(ml, [tgt.m2lnL(x) for x in th], rtol=1e-13, atol=1e-13)
        C = np.concatenate(chain[len(chain) // 5:])                 # burn-in: first fifth dropped
        cat = np.atleast_2d(np.cov(C.T))
        ok = _cmp(np.diag(cat), np.diag(cov), 0.0) < TOL and _cmp(_cov2cor(cat), _cov2cor(cov), 1.0) < TOL
        if ok:
            break
        run *= 2
    print(f"{sd_s} nu={nu} d={d}: {len(chain)} iterations, acceptance {nacc / ntot:.3f}, "
          f"var {_cmp(np.diag(cat), np.diag(cov), 0.0):.3f}, cor {_cmp(_cov2cor(cat), _cov2cor(cov), 1.0):.3f}")
    assert ok
    assert np.max(np.abs(C.mean(axis=0) - mu) / np.sqrt(np.diag(cov))) < 0.25
    assert nacc / ntot > 0.2                                        # an independence sampler with a fitted proposal: far from stuck


@pytest.mark.parametrize("sd_s,nu", CASES)
def test_esmcmc_run_forgets_a_displaced_start(oracle, sd_s, nu):
    """Harder than the reference's start (which already samples the posterior): the initial ensemble is displaced by two standard
    deviations and three times too wide.  The same two bars must be met after burn-in -- the property that makes the walker usable,
    and the one a wrong weight solve or a wrong proposal density in the acceptance ratio would break."""
    O = oracle
    d, W = 2, 200
    seed = 77 + int(10 * nu) + (7 if sd_s == "kde" else 0)
    mu, cov, X, _ = mvnd_problem(O, d, W, seed=seed)
    X = np.ascontiguousarray(mu + 2.0 * np.sqrt(np.diag(cov)) + 3.0 * (X - mu))
    tgt = O.Target(O.TARGET_MVND, d, np.full(d, -50.0), np.full(d, 50.0), mu=mu, cov=cov)
    ap = O.APES(W, d, O.SD_KDE if sd_s == "kde" else O.SD_VKDE, O.KERNEL_ST, nu, over_smooth=1.01, use_interp=True, use_threads=False)
    rng = O.RNG(seed + 1)
    th, ml = X.copy(), np.array([tgt.m2lnL(x) for x in X])
    chain = []
    for _ in range(120):
        ap.run(tgt, th, ml, 1, rng, nthreads=1)
        chain.append(th.copy())
    C = np.concatenate(chain[40:])
    cat = np.cov(C.T)
    v, c = _cmp(np.diag(cat), np.diag(cov), 0.0), _cmp(_cov2cor(cat), _cov2cor(cov), 1.0)
    print(f"{sd_s} nu={nu}: displaced start, var {v:.3f}, cor {c:.3f}, mean {np.max(np.abs(C.mean(axis=0) - mu) / np.sqrt(np.diag(cov))):.3f} sigma")
    assert v < TOL and c < TOL
    assert np.max(np.abs(C.mean(axis=0) - mu) / np.sqrt(np.diag(cov))) < 0.25


@pytest.mark.gpu
@pytest.mark.parametrize("d", [2, 3])
@pytest.mark.parametrize("method,k_type", [("VKDE", "CAUCHY"), ("VKDE", "ST3"), ("KDE", "CAUCHY"), ("KDE", "ST3")])
def test_gpu_esmcmc_run_recovers_posterior_covariance(oracle, method, k_type, d):
    """The same test through the product (host mirror of NcmFitESMCMCWalkerAPES over the CUDA library): cases 0, 1, 3, 4 of
    test_ncm_fit_esmcmc.c:155-183, including the reference's retry rule.  The oracle is used for the problem set-up only."""
    from numcosmo_b200 import stats_dist as S

    W = 100 * d
    seed = 500 * d + len(method) + len(k_type)
    mu, cov, X, _ = mvnd_problem(oracle, d, W, seed=seed)
    lb, ub = np.full(d, -50.0), np.full(d, 50.0)
    U = np.ascontiguousarray(np.linalg.cholesky(cov).T)
    ap = S.FitESMCMCWalkerAPES.new_full(W, d, getattr(S.FitESMCMCWalkerAPESMethod, method), getattr(S.FitESMCMCWalkerAPESKType, k_type), 1.0, True)
    ap.set_over_smooth(1.01)                                        # :202-203
    rng = S.RNG(seed + 1)
    th = X.copy()
    z = np.linalg.solve(U.T, (th - mu).T)
    ml = np.ascontiguousarray(np.einsum("ij,ij->j", z, z))
    run = d * 17
    chain, nacc, ntot, ok = [], 0, 0, False
    for _ in range(MAX_TRIES):
        for _ in range(run):
            acc, _t = ap.run("mvnd", lb, ub, th, ml, 1, rng, target_args=(mu, U))
            nacc += int(acc.sum())
            ntot += W
            chain.append(th.copy())
        z = np.linalg.solve(U.T, (th - mu).T)
        assert np.allclose(ml, np.einsum("ij,ij->j", z, z), rtol=1e-9, atol=1e-9)      # validate, :468
        C = np.concatenate(chain[len(chain) // 5:])
        cat = np.cov(C.T)
        v, c = _cmp(np.diag(cat), np.diag(cov), 0.0), _cmp(_cov2cor(cat), _cov2cor(cov), 1.0)
        ok = v < TOL and c < TOL
        if ok or len(chain) > 2000:
            break
        run *= 2
    print(f"gpu {method} {k_type} d={d}: {len(chain)} iterations, acceptance {nacc / ntot:.3f}, var {v:.3f}, cor {c:.3f}")
    assert ok
    assert nacc / ntot > 0.2


@pytest.mark.parametrize("sd_s,nu", CASES)
def test_serial_and_threaded_runs_draw_the_same_sequence(oracle, sd_s, nu):
    """The pattern of test_ncm_fit_esmcmc.c:888-971 (parity/serial_vs_threaded: fixed seed, dim 2, 60 walkers, 5 iterations, every catalog
    entry compared with ==), which the reference runs on the Stretch walker, applied to APES on the oracle: use_threads changes who
    evaluates a point, not what is computed, so positions, -2 ln L and accept flags must be identical bit for bit."""
    O = oracle
    d, W, iters = 2, 60, 5
    mu, cov, X, _ = mvnd_problem(O, d, W, seed=20260721 % 100000, sigma=(1.0e-2, 2.0e-2), cor_level=0.3, mu_range=(-1.0, 1.0))
    tgt = O.Target(O.TARGET_MVND, d, np.full(d, -50.0), np.full(d, 50.0), mu=mu, cov=cov)
    ml0 = np.array([tgt.m2lnL(x) for x in X])
    runs = []
    for use_threads, nthreads in ((False, 1), (True, 4)):
        ap = O.APES(W, d, O.SD_KDE if sd_s == "kde" else O.SD_VKDE, O.KERNEL_ST, nu, use_interp=True, use_threads=use_threads, local_frac=0.1)
        th, ml = X.copy(), ml0.copy()
        acc = ap.run(tgt, th, ml, iters, O.RNG(20260721), nthreads=nthreads)
        runs.append((th, ml, np.asarray(acc).copy()))
    assert np.array_equal(runs[0][0], runs[1][0]) and np.array_equal(runs[0][1], runs[1][1]) and np.array_equal(runs[0][2], runs[1][2])
    assert runs[0][2].any()


def _var_m2lnL_run(run_iter, W, iters, burn):
    vals = []
    for it in range(iters):
        ml = run_iter()
        if it >= burn:
            vals.append(ml.copy())
    return float(np.var(np.concatenate(vals), ddof=1))


@pytest.mark.parametrize("exploration", [0, 10])
@pytest.mark.parametrize("d", [1, 2, 3])
@pytest.mark.parametrize("sd_s,nu", CASES)
def test_variance_of_m2lnL_after_burnin_and_exploration(oracle, sd_s, nu, d, exploration):
    """run/burnin (:563-614) and run/exploration (:617-655): 100 iterations -- the first with the proposal density left out of the
    acceptance for `exploration` setup calls (two per iteration, ncm_fit_esmcmc_walker_apes.c:815-816, 901-919) -- then
    var (-2 ln L) over the trimmed catalog equals 2 dim (a chi-square with dim degrees of freedom) within 0.6."""
    O = oracle
    W = 100 * d
    seed = 300 * d + int(10 * nu) + (7 if sd_s == "kde" else 0) + exploration
    mu, cov, X, _ = mvnd_problem(O, d, W, seed=seed)
    tgt = O.Target(O.TARGET_MVND, d, np.full(d, -50.0), np.full(d, 50.0), mu=mu, cov=cov)
    ap = O.APES(W, d, O.SD_KDE if sd_s == "kde" else O.SD_VKDE, O.KERNEL_ST, nu, over_smooth=1.01, use_interp=True, use_threads=False)
    ap.set_exploration(exploration)
    rng = O.RNG(seed + 1)
    th, ml = X.copy(), np.array([tgt.m2lnL(x) for x in X])

    def one():
        ap.run(tgt, th, ml, 1, rng, nthreads=1)
        return ml

    v = _var_m2lnL_run(one, W, 100, 20)
    print(f"{sd_s} nu={nu} d={d} exploration={exploration}: var(-2 ln L) = {v:.3f} (2 dim = {2 * d})")
    assert abs(v / (2.0 * d) - 1.0) < 0.6


def test_exploration_leaves_the_proposal_density_out_for_that_many_setups(oracle):
    """While exploring, acceptance is min (1, L*/L).  The counter is decremented at the END of every setup call (one per half-ensemble,
    ncm_fit_esmcmc_walker_apes.c:815-816), i.e. before the acceptance of that half reads it (:901-919): `exploration` = n leaves the
    proposal density out of n - 1 half-steps.  With n = 3 the whole first iteration accepts every proposal that improves the
    likelihood; with n = 2 only its first half does."""
    O = oracle
    d, W = 2, 200
    mu, cov, X, _ = mvnd_problem(O, d, W, seed=5)
    tgt = O.Target(O.TARGET_MVND, d, np.full(d, -50.0), np.full(d, 50.0), mu=mu, cov=cov)
    ml0 = np.array([tgt.m2lnL(x) for x in X])
    ap = O.APES(W, d, O.SD_VKDE, O.KERNEL_ST, 1.0, use_interp=True, use_threads=False)
    ap.set_exploration(3)                                           # both halves of the first iteration
    th, ml, rng = X.copy(), ml0.copy(), O.RNG(9)
    before = ml.copy()
    acc = ap.run(tgt, th, ml, 1, rng, nthreads=1)[0].astype(bool)
    star = np.array([tgt.m2lnL(x) for x in ap.peek_thetastar()])
    assert np.all(acc[star <= before])                              # an improving proposal is always taken while exploring
    # without exploration the same stream takes a different decision somewhere (the proposal density matters)
    ap2 = O.APES(W, d, O.SD_VKDE, O.KERNEL_ST, 1.0, use_interp=True, use_threads=False)
    th2, ml2 = X.copy(), ml0.copy()
    acc2 = ap2.run(tgt, th2, ml2, 1, O.RNG(9), nthreads=1)[0].astype(bool)
    assert np.array_equal(ap2.peek_thetastar()[: W // 2], ap.peek_thetastar()[: W // 2])   # first half: same proposals, drawn before any decision
    assert not np.array_equal(acc, acc2)
    # n = 2: the first half explores, the second does not -- from there on the run IS the ordinary one only if the first halves agreed,
    # so compare the decisions of the first half with the exploring run and check that the second half's differ from it somewhere
    ap3 = O.APES(W, d, O.SD_VKDE, O.KERNEL_ST, 1.0, use_interp=True, use_threads=False)
    ap3.set_exploration(2)
    th3, ml3 = X.copy(), ml0.copy()
    acc3 = ap3.run(tgt, th3, ml3, 1, O.RNG(9), nthreads=1)[0].astype(bool)
    assert np.array_equal(acc3[: W // 2], acc[: W // 2])
    star3 = np.array([tgt.m2lnL(x) for x in ap3.peek_thetastar()])
    assert not np.all(acc3[W // 2:][star3[W // 2:] <= before[W // 2:]])            # an improving proposal refused: the density is back


@pytest.mark.gpu
@pytest.mark.parametrize("exploration", [3, 10])
def test_gpu_exploration_phase_follows_the_oracle(oracle, exploration):
    """ncm_fit_esmcmc_walker_apes_set_exploration through the product: same generator stream, same accepted sequence and positions as
    the oracle across the exploring half-steps and the return to the ordinary acceptance."""
    from numcosmo_b200 import stats_dist as S

    d, W, iters = 3, 300, 7
    mu, cov, X, _ = mvnd_problem(oracle, d, W, seed=321)
    lb, ub = np.full(d, -50.0), np.full(d, 50.0)
    tgt = oracle.Target(oracle.TARGET_MVND, d, lb, ub, mu=mu, cov=cov)
    ml0 = np.array([tgt.m2lnL(x) for x in X])
    ao = oracle.APES(W, d, oracle.SD_VKDE, oracle.KERNEL_ST, 3.0, over_smooth=1.0, use_interp=True, use_threads=True)
    ao.set_exploration(exploration)
    th_o, ml_o = X.copy(), ml0.copy()
    acc_o = ao.run(tgt, th_o, ml_o, iters, oracle.RNG(77), nthreads=4)
    assert ao.fallback_counts()[1:] == (0, 0)                      # every weight solve stayed in dposv: the bit-level regime
    ag = S.FitESMCMCWalkerAPES(W, d, S.FitESMCMCWalkerAPESMethod.VKDE, S.FitESMCMCWalkerAPESKType.ST3, 1.0, True)
    ag.set_use_threads(True)
    ag.set_exploration(exploration)
    th_g, ml_g = X.copy(), ml0.copy()
    acc_g, _ = ag.run("mvnd", lb, ub, th_g, ml_g, iters, S.RNG(77), target_args=(mu, tgt.U))
    diff = np.argwhere(np.asarray(acc_o) != np.asarray(acc_g))
    assert diff.size == 0, f"first divergence at (iter, walker) = {diff[0]}"
    assert np.max(np.abs(th_g - th_o)) <= 1e-9 * np.abs(th_o).max()


VARIANTS = [("no_interp", dict(use_interp=False)), ("random_walk_half", dict(random_walk_prob=0.5)), ("no_random_walk", dict(random_walk_prob=0.0)),
            ("shrink_10pc", dict(shrink=0.1))]


@pytest.mark.parametrize("name,kw", VARIANTS)
@pytest.mark.parametrize("sd_s,nu", CASES)
def test_walker_options_still_sample_the_posterior(oracle, sd_s, nu, name, kw):
    """The options of the walker that change the proposal -- uniform weights instead of the interpolated ones (use_interp FALSE: prepare,
    not prepare_interp, walker_apes.c:789-802), the random-walk mixture (random_walk_prob, :700-737, whose density enters the transition
    probability :822-860), the shrink floor of the weights -- must leave the chain's law alone: same two bars as the run test."""
    O = oracle
    d, W = 2, 200
    seed = 900 + int(10 * nu) + (7 if sd_s == "kde" else 0) + len(name)
    mu, cov, X, _ = mvnd_problem(O, d, W, seed=seed)
    tgt = O.Target(O.TARGET_MVND, d, np.full(d, -50.0), np.full(d, 50.0), mu=mu, cov=cov)
    args = dict(over_smooth=1.01, use_interp=True, use_threads=False)
    args.update(kw)
    ap = O.APES(W, d, O.SD_KDE if sd_s == "kde" else O.SD_VKDE, O.KERNEL_ST, nu, **args)
    rng = O.RNG(seed + 1)
    th, ml = X.copy(), np.array([tgt.m2lnL(x) for x in X])
    chain, run, ok = [], 40, False
    for _ in range(6):
        for _ in range(run):
            ap.run(tgt, th, ml, 1, rng, nthreads=1)
            chain.append(th.copy())
        C = np.concatenate(chain[len(chain) // 5:])
        cat = np.cov(C.T)
        v, c = _cmp(np.diag(cat), np.diag(cov), 0.0), _cmp(_cov2cor(cat), _cov2cor(cov), 1.0)
        ok = v < TOL and c < TOL
        if ok:
            break
        run *= 2
    print(f"{sd_s} nu={nu} {name}: {len(chain)} iterations, var {v:.3f}, cor {c:.3f}")
    assert ok


@pytest.mark.gpu
@pytest.mark.parametrize("name,kw", VARIANTS[:3])
def test_gpu_walker_options_follow_the_oracle(oracle, name, kw):
    """The same options through the product's setters: identical accepted sequence and positions as the oracle.  (The shrink setter
    of the walker only stores its value once the object is constructed, walker_apes.c:1178-1187 -- the objects got theirs in set_sys
    -- so there is nothing to follow for it.)"""
    from numcosmo_b200 import stats_dist as S

    d, W, iters = 3, 300, 5
    mu, cov, X, _ = mvnd_problem(oracle, d, W, seed=123 + len(name))
    lb, ub = np.full(d, -50.0), np.full(d, 50.0)
    tgt = oracle.Target(oracle.TARGET_MVND, d, lb, ub, mu=mu, cov=cov)
    ml0 = np.array([tgt.m2lnL(x) for x in X])
    args = dict(over_smooth=1.0, use_interp=True, use_threads=True)
    args.update(kw)
    ao = oracle.APES(W, d, oracle.SD_VKDE, oracle.KERNEL_ST, 3.0, **args)
    th_o, ml_o = X.copy(), ml0.copy()
    acc_o = ao.run(tgt, th_o, ml_o, iters, oracle.RNG(55), nthreads=4)
    assert ao.fallback_counts()[1:] == (0, 0)                      # every weight solve stayed in dposv: the bit-level regime
    ag = S.FitESMCMCWalkerAPES(W, d, S.FitESMCMCWalkerAPESMethod.VKDE, S.FitESMCMCWalkerAPESKType.ST3, 1.0, True)
    ag.set_use_threads(True)
    if "use_interp" in kw:
        ag.use_interp(kw["use_interp"])
    if "random_walk_prob" in kw:
        ag.set_random_walk_prob(kw["random_walk_prob"])
    th_g, ml_g = X.copy(), ml0.copy()
    acc_g, _ = ag.run("mvnd", lb, ub, th_g, ml_g, iters, S.RNG(55), target_args=(mu, tgt.U))
    diff = np.argwhere(np.asarray(acc_o) != np.asarray(acc_g))
    assert diff.size == 0, f"{name}: first divergence at (iter, walker) = {diff[0]}"
    assert np.max(np.abs(th_g - th_o)) <= 1e-9 * np.abs(th_o).max()
