// Robust covariance estimators behind NCM_STATS_DIST_KDE_COV_TYPE_ROBUST_DIAG / _ROBUST:
// ncm_stats_vec_compute_cov_robust_diag / _ogk, numcosmo/ncm/stats/ncm_stats_vec.c:1821-2072, over GSL's
// gsl_stats_Qn_from_sorted_data (Rousseeuw & Croux's Q_n scale).  O(d^2 n log n) host work next to the O(N d^2)
// prepare_kernel it belongs to; the factors it produces feed the same GPU upload as the sample covariance.
//
// Q_n is the k-th smallest of the n (n - 1) / 2 pairwise differences (k = h (h - 1) / 2, h = n / 2 + 1).  Here the
// order statistic is located by bisection on the bit pattern of the candidate value (non-negative doubles order like
// their integer images), counting the pairs at or below a candidate with one two-pointer pass over the sorted data:
// exact, 63 passes of O(n), no n^2 storage.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include "internal.h"

namespace {

inline double from_bits(uint64_t b) {
  double v;
  memcpy(&v, &b, sizeof v);
  return v;
}
inline uint64_t to_bits(double v) {
  uint64_t b;
  memcpy(&b, &v, sizeof b);
  return b;
}

// number of pairs (i > j) with a[i] - a[j] <= t, a ascending
long long pairs_at_or_below(const std::vector<double> &a, double t) {
  const int n = (int) a.size();
  long long c = 0;
  int j = 0;
  for (int i = 1; i < n; i++) {
    while (j < i && a[i] - a[j] > t) j++;
    c += i - j;
  }
  return c;
}

// finite-sample factor d_n of gsl_stats_Qn_from_sorted_data (GSL >= 2.5)
double qn_dn(int n) {
  static const double small_n[13] = {1.0, 1.0, 0.399356, 0.99365, 0.51321, 0.84401, 0.61220, 0.85877, 0.66993, 0.87344, 0.72014, 0.88906, 0.75743};
  if (n <= 12) return small_n[n];
  const double dn = (n % 2 == 1) ? 1.60188 + (-2.1284 - 5.172 / n) / n : 3.67561 + (1.9654 + (6.987 - 77.0 / n) / n) / n;
  return 1.0 / (dn / (double) n + 1.0);
}

}   // namespace

// gsl_sort + gsl_stats_Qn_from_sorted_data; `data` is sorted in place
double ncm_b200_stats_Qn(std::vector<double> &data) {
  const int n = (int) data.size();
  if (n < 2) return 0.0;
  std::sort(data.begin(), data.end());
  const long long h = n / 2 + 1, k = h * (h - 1) / 2;
  uint64_t lo = 0, hi = to_bits(data[n - 1] - data[0]);
  while (lo < hi) {
    const uint64_t mid = lo + (hi - lo) / 2;
    if (pairs_at_or_below(data, from_bits(mid)) >= k)
      hi = mid;
    else
      lo = mid + 1;
  }
  return 2.21914 * qn_dn(n) * from_bits(lo);
}

// kind 0: ncm_stats_vec_compute_cov_robust_diag (ncm_stats_vec.c:1832-1882); kind 1: ..._robust_ogk (:1896-2072).
// rows[n] point to d-vectors.  Returns false (after ncm_b200_error) when there are too few points.
bool ncm_b200_cov_robust(int kind, const double *const *rows, int n, int d, double *cov) {
  if (n < 4) {
    ncm_b200_error("ncm_stats_vec_compute_cov_robust_diag: too few points to estimate the covariance [%d].", n);
    return false;
  }
  std::vector<double> data(n), sigma_x(d);
  std::fill(cov, cov + (size_t) d * d, 0.0);
  if (kind == 0) {
    for (int i = 0; i < d; i++) {
      for (int a = 0; a < n; a++) data[a] = rows[a][i];
      const double s = ncm_b200_stats_Qn(data);
      cov[i * d + i] = s * s;
    }
    return true;
  }
  // OGK: scale every coordinate by its Q_n, pairwise robust correlations from Q_n (y_i + y_j) and Q_n (y_i - y_j),
  // eigenvectors of that matrix, Q_n of the projections, back to the original scale
  std::vector<double> y((size_t) n * d);
  for (int a = 0; a < n; a++) memcpy(&y[(size_t) a * d], rows[a], sizeof(double) * d);
  for (int i = 0; i < d; i++) {
    for (int a = 0; a < n; a++) data[a] = y[(size_t) a * d + i];
    sigma_x[i]     = ncm_b200_stats_Qn(data);
    const double s = 1.0 / sigma_x[i];
    for (int a = 0; a < n; a++) y[(size_t) a * d + i] *= s;
  }
  std::vector<double> C((size_t) d * d, 0.0), w, V;
  for (int i = 0; i < d; i++) C[i * d + i] = 1.0;
  for (int i = 0; i < d; i++)
    for (int j = i + 1; j < d; j++) {
      for (int a = 0; a < n; a++) data[a] = y[(size_t) a * d + i] + y[(size_t) a * d + j];
      const double sp = ncm_b200_stats_Qn(data);
      for (int a = 0; a < n; a++) data[a] = y[(size_t) a * d + i] - y[(size_t) a * d + j];
      const double sm = ncm_b200_stats_Qn(data);
      C[i * d + j] = C[j * d + i] = 0.25 * (sp * sp - sm * sm);
    }
  ncm_b200_jacobi_eig(C, d, w, V);   // columns of V: the eigenvectors (the reference's dsyevr rows of E)
  std::vector<double> G((size_t) d * d);   // G[e][j] = sigma_z_e * V[j][e] * sigma_x_j
  for (int e = 0; e < d; e++) {
    for (int a = 0; a < n; a++) {
      double z = 0.0;
      for (int j = 0; j < d; j++) z += V[j * d + e] * y[(size_t) a * d + j];
      data[a] = z;
    }
    const double sigma_z = ncm_b200_stats_Qn(data);
    for (int j = 0; j < d; j++) G[e * d + j] = sigma_z * V[j * d + e] * sigma_x[j];
  }
  for (int i = 0; i < d; i++)
    for (int j = i; j < d; j++) {
      double s = 0.0;
      for (int e = 0; e < d; e++) s += G[e * d + i] * G[e * d + j];
      cov[i * d + j] = cov[j * d + i] = s;
    }
  return true;
}

extern "C" double ncm_b200_test_Qn(const double *x, int n) {
  std::vector<double> v(x, x + n);
  return ncm_b200_stats_Qn(v);
}
