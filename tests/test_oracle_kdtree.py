"""The oracle's exact-kNN restatement against the REFERENCE'S OWN kd-tree.

numcosmo/external/misc/kdtree.c + rb_knn_list.c compile from their own sources (oracle/Makefile target "ref" builds them where they lie
into oracle/_ref/libkdtree_ref.so; oracle/ref_kdtree_driver.c drives them the way ncm_stats_dist_vkde.c:374-389, 440-452 does).  The VKDE
local covariances are accumulated over the neighbours IN LIST ORDER, so the oracle (and the device kernel, which is bit-identical to the
host mirror of the same rule) must reproduce the neighbour set, the order and the tie-breaking: SURVEY.md section 8f-1 "parity risk"."""
import numpy as np
import pytest


def _compare(oracle, P, k, queries):
    ref = oracle.ref_kdtree_knn(P, queries, k)
    if ref is None:
        pytest.skip("oracle/_ref/libkdtree_ref.so absent (built by `make -C oracle ref` where /root/reference exists)")
    ri, rd = ref
    for a, q in enumerate(queries):
        bi, bd = oracle.knn_brute(P, q, k)
        assert np.array_equal(bi, ri[a]), f"query {q}: neighbour order differs from the reference kd-tree"
        assert np.array_equal(bd, rd[a])          # the same squared distances, bit for bit (kdtree.c:27-38 sums in index order)
        assert bd[0] == 0.0 and q in bi[bd == 0.0]  # the centre (or an exact duplicate with a lower index) comes first


@pytest.mark.parametrize("d,n,k", [(2, 400, 10), (5, 2000, 102), (10, 2048, 102), (30, 1500, 75)])
def test_knn_order_random_points(oracle, d, n, k):
    rng = np.random.default_rng(d)
    P = rng.standard_normal((n, d)) * rng.uniform(0.5, 2.0, size=d)
    _compare(oracle, P, k, rng.choice(n, size=100, replace=False))


def test_knn_order_with_exact_ties(oracle):
    # a lattice: many exactly equal distances, also across the k-th position; every point queried
    G = np.array([[i, j] for i in range(20) for j in range(20)], dtype=float)
    _compare(oracle, G, 25, np.arange(len(G)))
    G3 = np.array([[i, j, l] for i in range(7) for j in range(7) for l in range(7)], dtype=float)
    _compare(oracle, G3, 33, np.arange(len(G3)))
    # repeated observations (an ensemble with rejected proposals holds duplicates of walkers): zero distances beyond the centre itself
    rng = np.random.default_rng(1)
    D = np.repeat(rng.standard_normal((100, 3)), 4, axis=0)
    _compare(oracle, D, 10, np.arange(len(D)))
    D = D[rng.permutation(len(D))]
    _compare(oracle, D, 10, np.arange(len(D)))


def test_knn_k_equals_n_and_minimum_k(oracle):
    rng = np.random.default_rng(2)
    P = rng.standard_normal((64, 4))
    _compare(oracle, P, 64, np.arange(64))      # local_frac = 1: every observation is a neighbour
    _compare(oracle, P, 2, np.arange(64))       # the floor GSL_MAX (local_frac * n_obs, 2) of vkde.c:426
