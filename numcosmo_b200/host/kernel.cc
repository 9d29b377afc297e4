// NcmStatsDistKernel{Gauss,ST}: the O(1) / O(n) host-side pieces of the kernel interface
// (ncm_stats_dist_kernel_gauss.c:193-355, ncm_stats_dist_kernel_st.c:225-414).  The batched
// (point, centre) work never goes through these: it is fused into the CUDA kernels.
#include <cmath>
#include "internal.h"

namespace {
constexpr double LN2PI = 1.8378770664093454835606594728112352797227949472755668;
constexpr double LNPI  = 1.1447298858494001741434273513530587116472948129153115;
}

extern "C" {

NcmStatsDistKernelGauss *ncm_stats_dist_kernel_gauss_new(const guint dim) {
  if (dim < 1 || dim > NCM_SD_GPU_MAX_DIM) {
    ncm_b200_error("ncm_stats_dist_kernel_gauss_new: dimension %u outside [1, %d].", dim, NCM_SD_GPU_MAX_DIM);
    return nullptr;
  }
  return new NcmStatsDistKernel{NCM_SD_GPU_KERNEL_GAUSS, dim, 0.0, 1};
}
NcmStatsDistKernelST *ncm_stats_dist_kernel_st_new(const guint dim, const gdouble nu) {
  if (dim < 1 || dim > NCM_SD_GPU_MAX_DIM) {
    ncm_b200_error("ncm_stats_dist_kernel_st_new: dimension %u outside [1, %d].", dim, NCM_SD_GPU_MAX_DIM);
    return nullptr;
  }
  return new NcmStatsDistKernel{NCM_SD_GPU_KERNEL_ST, dim, nu, 1};
}
NcmStatsDistKernel *ncm_stats_dist_kernel_ref(NcmStatsDistKernel *sdk) {
  sdk->ref++;
  return sdk;
}
void ncm_stats_dist_kernel_free(NcmStatsDistKernel *sdk) {
  if (sdk != nullptr && --sdk->ref == 0) delete sdk;
}
void ncm_stats_dist_kernel_clear(NcmStatsDistKernel **sdk) {
  if (sdk != nullptr && *sdk != nullptr) {
    ncm_stats_dist_kernel_free(*sdk);
    *sdk = nullptr;
  }
}
NcmStatsDistKernelGauss *ncm_stats_dist_kernel_gauss_ref(NcmStatsDistKernelGauss *k) { return ncm_stats_dist_kernel_ref(k); }
void ncm_stats_dist_kernel_gauss_free(NcmStatsDistKernelGauss *k) { ncm_stats_dist_kernel_free(k); }
void ncm_stats_dist_kernel_gauss_clear(NcmStatsDistKernelGauss **k) { ncm_stats_dist_kernel_clear(k); }
NcmStatsDistKernelST *ncm_stats_dist_kernel_st_ref(NcmStatsDistKernelST *k) { return ncm_stats_dist_kernel_ref(k); }
void ncm_stats_dist_kernel_st_free(NcmStatsDistKernelST *k) { ncm_stats_dist_kernel_free(k); }
void ncm_stats_dist_kernel_st_clear(NcmStatsDistKernelST **k) { ncm_stats_dist_kernel_clear(k); }
void ncm_stats_dist_kernel_st_set_nu(NcmStatsDistKernelST *k, const gdouble nu) { k->nu = nu; }
gdouble ncm_stats_dist_kernel_st_get_nu(NcmStatsDistKernelST *k) { return k->nu; }

guint ncm_stats_dist_kernel_get_dim(NcmStatsDistKernel *sdk) { return sdk->d; }

gdouble ncm_stats_dist_kernel_get_rot_bandwidth(NcmStatsDistKernel *sdk, const gdouble n) {
  const double d = sdk->d;
  if (sdk->kind == NCM_SD_GPU_KERNEL_GAUSS) return pow(4.0 / (n * (d + 2.0)), 1.0 / (d + 4.0));
  const double nu = (sdk->nu >= 3.0) ? sdk->nu : 3.0;
  return pow(16.0 * ((nu - 2) * (nu - 2)) * (1.0 + d + nu) * (3.0 + d + nu) /
                 ((2.0 + d) * (d + nu) * (2.0 + d + nu) * (d + 2.0 * nu) * (2.0 + d + 2.0 * nu) * n),
             1.0 / (d + 4.0));
}

gdouble ncm_stats_dist_kernel_get_lnnorm(NcmStatsDistKernel *sdk, NcmMatrix *cov_decomp) {
  const double d     = sdk->d;
  const double lndet = ncm_b200_cholesky_lndet(cov_decomp->data, (int) sdk->d, (int) cov_decomp->tda);
  if (sdk->kind == NCM_SD_GPU_KERNEL_GAUSS) return 0.5 * (d * LN2PI + lndet);
  const double lg_lnnorm   = lgamma(sdk->nu / 2.0) - lgamma((sdk->nu + d) / 2.0);
  const double chol_lnnorm = 0.5 * lndet;
  const double nc_lnnorm   = (d / 2.0) * (LNPI + log(sdk->nu));
  return lg_lnnorm + nc_lnnorm + chol_lnnorm;
}

gdouble ncm_stats_dist_kernel_eval_unnorm(NcmStatsDistKernel *sdk, const gdouble chi2) {
  if (sdk->kind == NCM_SD_GPU_KERNEL_GAUSS) return exp(-0.5 * chi2);
  return pow(1.0 + chi2 / sdk->nu, -0.5 * (sdk->nu + sdk->d));
}

void ncm_stats_dist_kernel_eval_unnorm_vec(NcmStatsDistKernel *sdk, NcmVector *chi2, NcmVector *Ku) {
  const guint n = ncm_vector_len(chi2);
  if (ncm_vector_len(Ku) != n) {
    ncm_b200_error("ncm_stats_dist_kernel_eval_unnorm_vec: assertion failed (ncm_vector_len (Ku) == n)");
    return;
  }
  for (guint i = 0; i < n; i++) ncm_vector_set(Ku, i, ncm_stats_dist_kernel_eval_unnorm(sdk, ncm_vector_get(chi2, i)));
}

static void sum_gamma_lambda(NcmStatsDistKernel *sdk, NcmVector *chi2, NcmVector *weights, NcmVector *lnnorms, gdouble lnnorm, NcmVector *lnK,
                             gdouble *gamma, gdouble *lambda) {
  const guint n      = ncm_vector_len(chi2);
  const double kappa = -0.5 * (sdk->nu + sdk->d);
  double lnt_max     = -INFINITY;
  guint i_max        = 0;
  if (n != ncm_vector_len(weights) || n != ncm_vector_len(lnK) || (lnnorms != nullptr && n != ncm_vector_len(lnnorms)) ||
      ncm_vector_stride(chi2) != 1 || ncm_vector_stride(weights) != 1 || ncm_vector_stride(lnK) != 1 ||
      (lnnorms != nullptr && ncm_vector_stride(lnnorms) != 1)) {
    ncm_b200_error("ncm_stats_dist_kernel_eval_sum_gamma_lambda: assertion failed (vector lengths / strides)");
    return;
  }
  for (guint i = 0; i < n; i++) {
    const double chi2_i = chi2->data[i], w_i = weights->data[i];
    const double lnu_i  = lnnorms != nullptr ? lnnorms->data[i] : 0.0;
    double lnt_i        = (sdk->kind == NCM_SD_GPU_KERNEL_GAUSS) ? -0.5 * chi2_i : kappa * log1p(chi2_i / sdk->nu);
    lnt_i               = lnt_i - lnu_i + log(w_i);
    if (lnt_i > lnt_max) {
      i_max   = i;
      lnt_max = lnt_i;
    }
    lnK->data[i] = lnt_i;
  }
  lambda[0] = 0.0;
  for (guint i = 0; i < i_max; i++) lambda[0] += exp(lnK->data[i] - lnt_max);
  for (guint i = i_max + 1; i < n; i++) lambda[0] += exp(lnK->data[i] - lnt_max);
  gamma[0] = lnt_max - lnnorm;
}

void ncm_stats_dist_kernel_eval_sum0_gamma_lambda(NcmStatsDistKernel *sdk, NcmVector *chi2, NcmVector *weights, NcmVector *lnnorms,
                                                  NcmVector *lnK, gdouble *gamma, gdouble *lambda) {
  sum_gamma_lambda(sdk, chi2, weights, lnnorms, 0.0, lnK, gamma, lambda);
}

void ncm_stats_dist_kernel_eval_sum1_gamma_lambda(NcmStatsDistKernel *sdk, NcmVector *chi2, NcmVector *weights, gdouble lnnorm, NcmVector *lnK,
                                                  gdouble *gamma, gdouble *lambda) {
  sum_gamma_lambda(sdk, chi2, weights, nullptr, lnnorm, lnK, gamma, lambda);
}

// x = mu + s U^T (h z): d normals in index order, then (ST) one chi-square -- the draw order of
// _kernel_gauss.c:335-355 / _kernel_st.c:388-414.  The arithmetic lives in ncm_b200_kernel_sample_from (below, after the extern "C"
// block) so that proposals whose draws were generated ahead of time (apes.cc) go through the very same operations.
void ncm_stats_dist_kernel_sample(NcmStatsDistKernel *sdk, NcmMatrix *cov_decomp, const gdouble href, NcmVector *mu, NcmVector *y, NcmRNG *rng) {
  const int d = (int) sdk->d;
  double z[NCM_SD_GPU_MAX_DIM];
  for (int i = 0; i < d; i++) z[i] = ncm_rng_ugaussian_gen(rng);
  const double chisq = (sdk->kind == NCM_SD_GPU_KERNEL_ST) ? ncm_rng_chisq_gen(rng, sdk->nu) : 0.0;
  ncm_b200_kernel_sample_from(sdk, cov_decomp, href, mu, y, z, chisq);
}

}   // extern "C"

// the affine map of kernel->sample from already drawn unit normals z_raw[d] and (Student-t) chi-square
void ncm_b200_kernel_sample_from(NcmStatsDistKernel *sdk, NcmMatrix *cov_decomp, double href, NcmVector *mu, NcmVector *y, const double *z_raw, double chisq) {
  const int d = (int) sdk->d;
  double z[NCM_SD_GPU_MAX_DIM], t[NCM_SD_GPU_MAX_DIM];
  for (int i = 0; i < d; i++) z[i] = z_raw[i] * href;
  // dtrmv Upper/Trans: t_k = sum_{j <= k} U[j][k] z_j
  const double *U = cov_decomp->data;
  const int ld    = (int) cov_decomp->tda;
  for (int k = 0; k < d; k++) {
    double s = 0.0;
    for (int j = 0; j <= k; j++) s += U[j * ld + k] * z[j];
    t[k] = s;
  }
  double scale = 1.0;
  if (sdk->kind == NCM_SD_GPU_KERNEL_ST) scale = sqrt(sdk->nu / chisq);
  for (int k = 0; k < d; k++) ncm_vector_set(y, k, t[k] * scale + ncm_vector_get(mu, k));
}
