"""Committed golden vectors (tests/golden/apes_path_v1.npz, written by tests/golden/make_golden.py from the CPU oracle):
 * not gpu: the oracle still reproduces them bit for bit (single-threaded, so the summation order is fixed);
 * gpu:     the CUDA path through the host mirror and the C ABI matches them at the north-star tolerance
            (rel 1e-10 on m2lnp; IM 1e-12 of its maximum; identical accepted-sample sequence for the fixed stream).
"""
import os

import numpy as np
import pytest

from helpers import make_sd, rel_err, support, weight_bound

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "apes_path_v1.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _cases(g, prefix):
    return sorted({k.split("/")[0] for k in g.files if k.startswith(prefix)})


def test_golden_file_is_small_and_complete(gold):
    assert os.path.getsize(GOLD) < 2_000_000
    assert len(_cases(gold, "vkde_")) == 4 and len(_cases(gold, "kde_")) == 2 and len(_cases(gold, "apes_")) == 2


def test_oracle_reproduces_golden(oracle, gold):
    oracle.lib().orc_set_blas_threads(1)
    for name in _cases(gold, "vkde_") + _cases(gold, "kde_"):
        sd_type, kernel, nu, d, n, lf = gold[f"{name}/meta"]
        sd = make_sd(oracle, int(sd_type), int(kernel), float(nu), gold[f"{name}/X"], m2lnp=gold[f"{name}/m2lnL"], local_frac=float(lf),
                     use_threads=False)
        assert np.array_equal(sd.peek_weights(), gold[f"{name}/weights"]), name
        assert np.array_equal(sd.eval_m2lnp_batch(gold[f"{name}/Q"], 1), gold[f"{name}/m2lnp"]), name
        assert np.array_equal(sd.compute_IM()[::7, ::5], gold[f"{name}/IM_sub"]), name
        assert sd.get_href() == gold[f"{name}/href"][0] and sd.get_rnorm() == gold[f"{name}/rnorm"][0]


def test_oracle_apes_reproduces_golden(oracle, gold):
    O = oracle
    O.lib().orc_set_blas_threads(1)
    for name in _cases(gold, "apes_"):
        kernel, nu, d, W, iters, seed = gold[f"{name}/meta"]
        d, W = int(d), int(W)
        lb, ub = np.full(d, -50.0), np.full(d, 50.0)
        U = gold[f"{name}/U"]
        tgt = O.Target(O.TARGET_MVND, d, lb, ub, mu=gold[f"{name}/mu"], cov=U.T @ U)
        th, ml = gold[f"{name}/X"].copy(), gold[f"{name}/m2lnL"].copy()
        acc = O.APES(W, d, O.SD_VKDE, int(kernel), float(nu), use_threads=False).run(tgt, th, ml, int(iters), O.RNG(int(seed)), nthreads=1)
        assert np.array_equal(np.asarray(acc).astype(np.uint8), gold[f"{name}/accepted"]), name


@pytest.mark.gpu
def test_gpu_matches_golden(gold):
    from numcosmo_b200 import stats_dist as S

    for name in _cases(gold, "vkde_") + _cases(gold, "kde_"):
        sd_type, kernel, nu, d, n, lf = gold[f"{name}/meta"]
        d = int(d)
        kern = S.StatsDistKernelGauss(d) if int(kernel) == 0 else S.StatsDistKernelST(d, float(nu))
        sd = (S.StatsDistKDE if name.startswith("kde_") else S.StatsDistVKDE)(kern, S.StatsDistCV.NONE)
        if not name.startswith("kde_"):
            sd.set_local_frac(float(lf))
        for x in gold[f"{name}/X"]:
            sd.add_obs(x)
        sd.prepare_interp(gold[f"{name}/m2lnL"])
        # weights: same passive set as the golden run, weights to the conditioning-limited bound of M[P,P] (computed here from the
        # row-scaled interpolation matrix of the GPU path), and the densities they produce to the same bound
        from numcosmo_b200 import capi

        w, wg = sd.peek_weights(), gold[f"{name}/weights"]
        st = sd.nnls_stats()
        assert st["n_lu"] == 0, (name, st)
        assert np.array_equal(support(w, int(n)), support(wg, int(n))), f"{name}: passive set differs from the golden one"
        cb = capi.Context.borrowed(S.lib().ncm_stats_dist_b200_peek_ctx(sd._h))
        cb.n_kernels = cb.n_obs = int(n)
        cb.d = d
        m2 = gold[f"{name}/m2lnL"]
        IMs = cb.compute_IM(np.exp(0.5 * (m2 - m2.min())), fetch=True, nrows=int(n))
        cond_M, bound = weight_bound(IMs, support(wg, int(n)))
        err = np.max(np.abs(w - wg)) / wg.max()
        assert err <= bound, f"{name}: weights differ by {err:.2e} (bound {bound:.2e}, cond {cond_M:.2e})"
        assert rel_err(sd.eval_m2lnp_array(gold[f"{name}/Q"]), gold[f"{name}/m2lnp"]) <= max(1e-10, bound), name
        # the evaluation proper at the north-star tolerance: same weights on both sides
        sd.prepare()
        ctx_w = gold[f"{name}/weights"]
        c = capi.Context.borrowed(S.lib().ncm_stats_dist_b200_peek_ctx(sd._h))
        c.n_kernels = c.n_obs = int(n)
        c.d = d
        c.set_weights(ctx_w, float(gold[f"{name}/href"][0]))
        assert rel_err(c.eval_m2lnp(gold[f"{name}/Q"]), gold[f"{name}/m2lnp"]) < 1e-10, name
        IM = c.compute_IM(None, fetch=True, nrows=int(n))   # klass->compute_IM proper: no 1/f row scaling, as the golden entries
        ref = gold[f"{name}/IM_sub"]
        assert np.max(np.abs(IM[::7, ::5] - ref)) < 1e-12 * np.abs(ref).max(), name


@pytest.mark.gpu
def test_gpu_apes_matches_golden_sequence(gold):
    from numcosmo_b200 import stats_dist as S

    for name in _cases(gold, "apes_"):
        kernel, nu, d, W, iters, seed = gold[f"{name}/meta"]
        d, W = int(d), int(W)
        lb, ub = np.full(d, -50.0), np.full(d, 50.0)
        kt = S.FitESMCMCWalkerAPESKType.GAUSS if int(kernel) == 0 else S.FitESMCMCWalkerAPESKType.ST3
        ap = S.FitESMCMCWalkerAPES(W, d, S.FitESMCMCWalkerAPESMethod.VKDE, kt, 1.0, True)
        th, ml = gold[f"{name}/X"].copy(), gold[f"{name}/m2lnL"].copy()
        acc, _ = ap.run("mvnd", lb, ub, th, ml, int(iters), S.RNG(int(seed)), target_args=(gold[f"{name}/mu"], gold[f"{name}/U"]))
        acc = np.asarray(acc).astype(np.uint8)
        ref = gold[f"{name}/accepted"]
        if not np.array_equal(acc.ravel(), ref.ravel()):
            first = int(np.argmax(acc.ravel() != ref.ravel()))
            pytest.fail(f"{name}: accepted-sample sequence diverges from the golden one at flat index {first}")
        assert np.allclose(th, gold[f"{name}/theta_final"], rtol=0, atol=1e-12)
