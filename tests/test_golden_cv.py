"""Committed golden vectors for the cross-validation modes and the robust covariance types (tests/golden/cv_robust_v1.npz, written by
tests/golden/make_golden_cv.py; its CV_SPLIT cases went through the reference's own dlevmar_dif):
 * not gpu: the oracle -- with its levmar RESTATEMENT -- reproduces them bit for bit (single-threaded);
 * gpu:     the CUDA path through the host mirror reproduces the optimiser traces, bandwidths, covariances, weights and densities.
"""
import os

import numpy as np
import pytest

from helpers import rel_err

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cv_robust_v1.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _cases(g):
    return sorted({k.split("/")[0] for k in g.files})


def test_golden_cv_file_is_small_and_complete(gold):
    assert os.path.getsize(GOLD) < 500_000
    names = _cases(gold)
    assert sum(n.startswith("cv_split_") for n in names) == 2 and sum(n.startswith("cv_nofit_") for n in names) == 2
    assert sum(n.startswith("cv_loo_") for n in names) == 2 and sum(n.startswith("robust_") for n in names) == 3
    assert all(gold[f"{n}/meta"][9] == 1.0 for n in names if n.startswith("cv_split_"))   # produced through the reference's levmar


def test_oracle_reproduces_golden_cv(oracle, gold):
    O = oracle
    O.lib().orc_set_blas_threads(1)
    O.use_ref_levmar(False)   # the restatement must land on the numbers the reference's own levmar produced
    for name in _cases(gold):
        sd_type, kernel, nu, d, n, cv, cov_type, sf, lf, _ = gold[f"{name}/meta"]
        sd = O.StatsDist(int(sd_type), int(kernel), int(d), float(nu), int(cv))
        sd.set_cov_type(int(cov_type))
        sd.set_split_frac(float(sf))
        sd.set_local_frac(float(lf))
        sd.set_use_threads(False)
        sd.add_obs_matrix(gold[f"{name}/X"])
        assert sd.prepare_interp(gold[f"{name}/m2lnL"]) == 0, name
        lnos, val = sd.cv_trace()
        assert np.array_equal(lnos, gold[f"{name}/lnos"]) and np.array_equal(val, gold[f"{name}/val"]), name
        assert sd.get_over_smooth() == gold[f"{name}/over_smooth"][0], name
        assert np.array_equal(sd.peek_weights(), gold[f"{name}/weights"]), name
        assert np.array_equal(np.triu(sd.peek_full_cov()), gold[f"{name}/cov"]), name
        assert np.array_equal(sd.eval_m2lnp_batch(gold[f"{name}/Q"], 1), gold[f"{name}/m2lnp"]), name


@pytest.mark.gpu
def test_gpu_matches_golden_cv(gold):
    from numcosmo_b200 import stats_dist as S

    for name in _cases(gold):
        sd_type, kernel, nu, d, n, cv, cov_type, sf, lf, _ = gold[f"{name}/meta"]
        d = int(d)
        kern = S.StatsDistKernelGauss(d) if int(kernel) == 0 else S.StatsDistKernelST(d, float(nu))
        sd = (S.StatsDistKDE if int(sd_type) == 0 else S.StatsDistVKDE)(kern, S.StatsDistCV(int(cv)))
        sd.set_cov_type(S.StatsDistKDECovType(int(cov_type)))
        sd.set_split_frac(float(sf))
        if int(sd_type) == 1:
            sd.set_local_frac(float(lf))
        for x in gold[f"{name}/X"]:
            sd.add_obs(x)
        sd.prepare_interp(gold[f"{name}/m2lnL"])
        lnos, val = sd.cv_trace()
        g_lnos, g_val = gold[f"{name}/lnos"], gold[f"{name}/val"]
        if name.startswith("cv_split_"):
            assert np.array_equal(lnos[:11], g_lnos[:11]), name                 # the ten random tries: bit-identical stream
            assert len(lnos) == len(g_lnos) and np.max(np.abs(lnos - g_lnos)) < 1e-7, name
            assert abs(sd.get_over_smooth() / gold[f"{name}/over_smooth"][0] - 1) < 1e-6, name
        else:
            assert np.array_equal(lnos, g_lnos), name                            # simplex trial points: identical
            assert len(val) == 0 or np.max(np.abs(val - g_val)) < 1e-9 * np.max(np.abs(g_val)), name
            assert sd.get_over_smooth() == gold[f"{name}/over_smooth"][0], name
        C, Cg = np.triu(sd.peek_full_cov()), gold[f"{name}/cov"]
        assert np.max(np.abs(C - Cg)) < 1e-10 * np.abs(Cg).max(), name
        w, wg = sd.peek_weights(), gold[f"{name}/weights"]
        if np.count_nonzero(w > 0.011 / len(w)) == np.count_nonzero(wg > 0.011 / len(wg)) and sd.nnls_stats()["n_lu"] == 0:
            assert np.max(np.abs(w - wg)) / wg.max() < 1e-5, name
            assert rel_err(sd.eval_m2lnp_array(gold[f"{name}/Q"]), gold[f"{name}/m2lnp"]) < 1e-5, name
