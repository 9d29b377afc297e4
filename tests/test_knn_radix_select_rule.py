"""The neighbour selection of csrc/vkde_prep.cu knn_kernel restated in numpy: radix SELECT of the k-th smallest squared distance on the
bit patterns of the (non-negative) doubles -- digits of 11, 11, 11, 11, 11, 9 bits from the top, prefix / rank bookkeeping as in the
kernel -- then every key below the threshold plus, in index order, the first `rank` keys equal to it, then the (distance, index)
sort of the k winners.  The result must be the reference's neighbour list: the k nearest ordered by (distance, index), as the kd-tree's
red-black list orders them (numcosmo/external/misc/kdtree.c:192-321, rb_knn_list.c:31-40) -- including exact ties (duplicated
points), zero distances, denormals and infinities."""
import numpy as np
import pytest


def radix_select_knn(dist, k):
    keys = np.asarray(dist, dtype=np.float64).view(np.uint64)
    n = keys.size
    prefix, mask, rank, shift = np.uint64(0), np.uint64(0), k, 64
    for p in range(6):
        bits = 11 if p < 5 else 9
        shift -= bits
        nbins = 1 << bits
        sel = (keys & mask) == prefix
        digit = ((keys[sel] >> np.uint64(shift)) & np.uint64(nbins - 1)).astype(np.int64)
        hist = np.bincount(digit, minlength=nbins)
        cum = np.cumsum(hist)
        b = int(np.searchsorted(cum, rank, side="left"))           # first bin whose cumulative count reaches the rank
        rank -= int(cum[b] - hist[b])
        prefix |= np.uint64(b) << np.uint64(shift)
        mask |= np.uint64(nbins - 1) << np.uint64(shift)
    assert shift == 0
    T, need_eq = prefix, rank
    lt = keys < T
    eq = keys == T
    take_eq = eq & (np.cumsum(eq) - eq < need_eq)                  # equal keys with a smaller index already taken < need_eq
    idx = np.flatnonzero(lt | take_eq)                             # ordered compaction: index order
    assert idx.size == k
    order = np.lexsort((idx, dist[idx]))                           # bitonic sort of the k winners by (distance, index)
    return idx[order]


def reference_list(dist, k):
    return np.lexsort((np.arange(dist.size), dist))[:k]


@pytest.mark.parametrize("n,k", [(1, 1), (5, 5), (17, 3), (600, 30), (2048, 102), (5000, 250), (16384, 819)])
def test_generic_distances(n, k):
    rs = np.random.default_rng(n + k)
    d2 = (rs.standard_normal((n, 6)) ** 2).sum(axis=1)
    d2[rs.integers(0, n)] = 0.0                                    # the centre itself
    assert np.array_equal(radix_select_knn(d2, k), reference_list(d2, k))


@pytest.mark.parametrize("k", [1, 2, 7, 8, 9, 40, 64])
def test_ties_are_taken_in_index_order(k):
    """Duplicated points: many exactly equal distances straddle the k-th place."""
    rs = np.random.default_rng(k)
    base = rs.uniform(0.0, 4.0, 8)
    d2 = np.repeat(base, 8)[rs.permutation(64)]
    got = radix_select_knn(d2, k)
    assert np.array_equal(got, reference_list(d2, k))
    assert np.all(np.diff(d2[got]) >= 0)


def test_special_values():
    d2 = np.array([0.0, 5e-324, 2.2250738585072014e-308, 1.0, 1.0, np.nextafter(1.0, 2.0), 1e300, np.inf, np.inf, 0.0])
    for k in range(1, d2.size + 1):
        assert np.array_equal(radix_select_knn(d2, k), reference_list(d2, k))
    allsame = np.full(100, 3.5)
    assert np.array_equal(radix_select_knn(allsame, 37), np.arange(37))
