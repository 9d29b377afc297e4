// NcmStatsDist / NcmStatsDistKDE / NcmStatsDistVKDE host objects: same entry points as the reference,
// bodies rewired to the C ABI of libncm_sd_gpu (include/ncm_sd_gpu.h).
//
//   prepare_kernel   host, O(N d^2) (+ kNN for VKDE): ncm_stats_dist_kde.c:378-490, ncm_stats_dist_vkde.c:362-515
//                    -> ncm_sd_gpu_upload_kde / ncm_sd_gpu_upload_vkde
//   prepare          ncm_stats_dist.c:703-789                     -> ncm_sd_gpu_set_weights
//   prepare_interp   ncm_stats_dist.c:878-1094                    -> ncm_sd_gpu_compute_IM + ncm_sd_gpu_nnls_solve
//   cross-validation ncm_stats_dist.c:484-701, 806-876, 1018-1072: every objective evaluation is a GPU pass
//                    (batched eval of the held-out points / IM / IM + NNLS + eval), the one-parameter
//                    optimisers run on the host (host/optim.cc)
//   eval[_m2lnp]     ncm_stats_dist.c:1527-1554                   -> ncm_sd_gpu_eval[_m2lnp] (q = 1, or a whole batch
//                                                                    through the new ncm_stats_dist_eval_m2lnp_array)
//   kernel_choose / sample   ncm_stats_dist.c:1565-1627           host, serial RNG stream in reference order
// The peek_* accessors keep returning live host NcmMatrix / NcmVector mirrors.
#include <algorithm>
#include <chrono>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <mutex>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "internal.h"

namespace {

// GPU calls of ONE object are serialised by its own mutex (eval may be called concurrently from OpenMP threads, ncm_fit_esmcmc.c:2158);
// different objects own different contexts and streams and run side by side

double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

bool gpu_ok(NcmStatsDist *sd, int rc, const char *where) {
  if (rc == NCM_SD_GPU_OK) return true;
  ncm_b200_error("%s: GPU path failed (%d): %s", where, rc, ncm_sd_gpu_last_error(sd->gpu));
  return false;
}

bool ensure_gpu(NcmStatsDist *sd) {
  if (sd->gpu != nullptr) return true;
  const int rc = ncm_sd_gpu_ctx_new(&sd->gpu, ncm_b200_default_device());
  if (rc != NCM_SD_GPU_OK) {
    sd->gpu = nullptr;
    ncm_b200_error("numcosmo_b200: no usable sm_100 CUDA device (ncm_sd_gpu_ctx_new returned %d); there is no CPU fallback.", rc);
    return false;
  }
  return true;
}

NcmStatsDist *sd_new(int type, NcmStatsDistKernel *sdk, NcmStatsDistCV cv_type) {
  NcmStatsDist *sd   = new NcmStatsDist();
  sd->ref            = 1;
  sd->type           = type;
  sd->kernel         = ncm_stats_dist_kernel_ref(sdk);
  sd->d              = sdk->d;
  sd->over_smooth    = 1.0;
  sd->shrink         = 0.01;
  sd->split_frac     = 0.5;
  sd->local_frac     = 0.05;
  sd->cv_type        = cv_type;
  sd->use_threads    = FALSE;
  sd->print_fit      = FALSE;
  sd->use_rot_href   = FALSE;
  sd->cov_type       = NCM_STATS_DIST_KDE_COV_TYPE_SAMPLE;
  sd->nearPD_maxiter = 200;
  sd->cov_fixed      = nullptr;
  sd->n_obs = sd->n_kernels = 0;
  sd->href = sd->min_m2lnp = sd->max_m2lnp = sd->rnorm = 0.0;
  sd->weights = sd->wcum = nullptr;
  sd->wcum_ready = FALSE;
  sd->prepared   = false;
  sd->cov        = ncm_matrix_new(sd->d, sd->d);
  sd->cov_decomp = ncm_matrix_new(sd->d, sd->d);
  sd->kernel_lnnorm = 0.0;
  sd->gpu           = nullptr;
  memset(&sd->nnls_stats, 0, sizeof(sd->nnls_stats));
  sd->host_prepare_kernel_ms = 0.0;
  sd->resident               = false;
  sd->sample_view.pdata      = nullptr;
  sd->sample_view.len        = 0;
  sd->cv_rng                 = ncm_rng_seeded_new(nullptr, 0);
  return sd;
}

void clear_cov_array(NcmStatsDist *sd) {
  for (NcmMatrix *m : sd->cov_array) ncm_matrix_free(m);
  sd->cov_array.clear();
}

// online mean / covariance update of NcmStatsVec with NCM_STATS_VEC_COV, weight 1
// (ncm_stats_vec.c:510-551) and the covariance read-out (ncm_stats_vec.c:2375-2399)
struct StatsVec {
  int len;
  double weight = 0.0, weight2 = 0.0, bias_wt = 0.0;
  std::vector<double> mean, var, cov;
  explicit StatsVec(int n) : len(n), mean(n, 0.0), var(n, 0.0), cov((size_t) n * n, 0.0) {}
  void reset() {
    weight = weight2 = bias_wt = 0.0;
    std::fill(mean.begin(), mean.end(), 0.0);
    std::fill(var.begin(), var.end(), 0.0);
    std::fill(cov.begin(), cov.end(), 0.0);
  }
  void append(const double *x) {
    const double w = 1.0, curweight = weight + w;
    for (int i = 0; i < len; i++) {
      double mean_i        = mean[i];
      const double x_i     = x[i];
      const double delta_i = x_i - mean_i;
      const double R_i     = delta_i * w / curweight;
      const double dvar    = weight * delta_i * R_i;
      mean_i += R_i;
      mean[i] = mean_i;
      var[i] += dvar;
      for (int j = i + 1; j < len; j++) {
        const double dC_ij = w * (x_i - mean_i) * (x[j] - mean[j]);
        const double C_ij  = cov[(size_t) i * len + j] + dC_ij;
        cov[(size_t) i * len + j] = C_ij;
        cov[(size_t) j * len + i] = C_ij;
      }
    }
    weight = curweight;
    weight2 += w * w;
    bias_wt = 1.0 / (weight - weight2 / weight);
  }
  void get_cov(double *m) const {
    for (int i = 0; i < len * len; i++) m[i] = cov[i];
    for (int i = 0; i < len; i++) m[i * len + i] = var[i];
    for (int i = 0; i < len * len; i++) m[i] *= bias_wt;
  }
};

double sd_href(NcmStatsDist *sd) {
  const double base = sd->over_smooth * ncm_stats_dist_kernel_get_rot_bandwidth(sd->kernel, sd->n_kernels);
  if (sd->type == NCM_SD_GPU_KDE) return base;            // ncm_stats_dist.c:476-482
  return sd->use_rot_href ? base / sd->local_frac : sd->over_smooth;   // ncm_stats_dist_vkde.c:316-335
}

// ncm_stats_dist_kde.c:378-490
bool kde_prepare_kernel(NcmStatsDist *sd) {
  NcmB200ProfScope prof("kde_prepare_kernel(host cov+whiten)");
  const int d = (int) sd->d;
  StatsVec sv(d);
  for (guint i = 0; i < sd->n_kernels; i++) sv.append(((NcmVector *) sd->sample[i])->data);
  std::vector<double> cov((size_t) d * d);
  switch (sd->cov_type) {
    case NCM_STATS_DIST_KDE_COV_TYPE_SAMPLE:
      sv.get_cov(cov.data());
      ncm_b200_cholesky_decomp_fallback(sd->cov_decomp->data, cov.data(), d, (int) sd->nearPD_maxiter);
      memcpy(sd->cov->data, cov.data(), sizeof(double) * d * d);
      break;
    case NCM_STATS_DIST_KDE_COV_TYPE_FIXED:
      if (sd->cov_fixed == nullptr) {
        ncm_b200_error("_ncm_stats_dist_kde_prepare_kernel: cov_type is FIXED but a fixed covariance matrix was not provided, "
                       "use ncm_stats_dist_kde_set_cov_fixed to set one.");
        return false;
      }
      for (int i = 0; i < d; i++)
        for (int j = 0; j < d; j++) sd->cov->data[i * d + j] = ncm_matrix_get(sd->cov_fixed, i, j);
      break;   // cov_decomp was set by set_cov_fixed (kde.c:838-844)
    case NCM_STATS_DIST_KDE_COV_TYPE_ROBUST_DIAG:
    case NCM_STATS_DIST_KDE_COV_TYPE_ROBUST: {
      // ncm_stats_dist_kde.c:423-441 over ncm_stats_vec.c:1821-2072 (host/robust.cc)
      std::vector<const double *> rows(sd->n_kernels);
      for (guint i = 0; i < sd->n_kernels; i++) rows[i] = ((NcmVector *) sd->sample[i])->data;
      if (!ncm_b200_cov_robust(sd->cov_type == NCM_STATS_DIST_KDE_COV_TYPE_ROBUST ? 1 : 0, rows.data(), (int) sd->n_kernels, d, cov.data())) return false;
      ncm_b200_cholesky_decomp_fallback(sd->cov_decomp->data, cov.data(), d, (int) sd->nearPD_maxiter);
      memcpy(sd->cov->data, cov.data(), sizeof(double) * d * d);
      break;
    }
    default:
      ncm_b200_error("_ncm_stats_dist_kde_prepare_kernel: code should not be reached (unknown covariance type %d).", (int) sd->cov_type);
      return false;
  }
  sd->kernel_lnnorm = ncm_stats_dist_kernel_get_lnnorm(sd->kernel, sd->cov_decomp);
  sd->sample_matrix.resize((size_t) sd->n_obs * d);
  sd->invUsample.resize((size_t) sd->n_obs * d);
  for (guint i = 0; i < sd->n_obs; i++) memcpy(&sd->sample_matrix[(size_t) i * d], ((NcmVector *) sd->sample[i])->data, sizeof(double) * d);
  // invUsample = sample . U^-1  (dtrsm Right/Upper/NoTrans): row r solves z U = x, i.e. forward in k
  const double *U = sd->cov_decomp->data;
#pragma omp parallel for if (sd->use_threads)
  for (int r = 0; r < (int) sd->n_obs; r++) {
    const double *x = &sd->sample_matrix[(size_t) r * d];
    double *z       = &sd->invUsample[(size_t) r * d];
    for (int k = 0; k < d; k++) {
      double t = x[k];
      for (int j = 0; j < k; j++) t -= z[j] * U[j * d + k];
      z[k] = t / U[k * d + k];
    }
  }
  return true;
}

// ncm_stats_dist_vkde.c:362-496: exact k nearest neighbours in whitened space, ordered by
// (distance, index) as the kd-tree's red-black list does (rb_knn_list.c:31-40), local covariance of
// the raw neighbours accumulated in that order, Cholesky with the nearPD / diagonal fallback.
bool vkde_build_cov_array(NcmStatsDist *sd) {
  const int d = (int) sd->d, n_obs = (int) sd->n_obs, nk = (int) sd->n_kernels;
  NcmB200ProfScope prof_pre("vkde_build_cov_array(host setup)", true);
  const double kd = (sd->local_frac * n_obs > 2.0) ? sd->local_frac * n_obs : 2.0;   // GSL_MAX (local_frac * n_obs, 2)
  const size_t k  = (size_t) kd;
  const bool robust = sd->cov_type == NCM_STATS_DIST_KDE_COV_TYPE_ROBUST_DIAG || sd->cov_type == NCM_STATS_DIST_KDE_COV_TYPE_ROBUST;
  if (robust && std::min(k, (size_t) n_obs) < 4) {
    ncm_b200_error("ncm_stats_vec_compute_cov_robust_diag: too few points to estimate the covariance [%d].", (int) std::min(k, (size_t) n_obs));
    return false;
  }
  if ((int) sd->cov_array.size() != nk) {
    clear_cov_array(sd);
    sd->cov_slab.assign((size_t) nk * d * d, 0.0);
    sd->cov_array.resize(nk);
    for (int i = 0; i < nk; i++) {
      NcmMatrix *m = new NcmMatrix;
      m->data      = &sd->cov_slab[(size_t) i * d * d];
      m->nrows = m->ncols = m->tda = (guint) d;
      m->ref                       = 1;
      m->own                       = false;
      sd->cov_array[i]             = m;
    }
    sd->lnnorms.assign(nk, 0.0);
  }
  // one centre on the host: exact kNN, online covariance in neighbour order, Cholesky with the nearPD / diagonal fallback
  auto host_centre = [&](int i, StatsVec &sv, std::vector<std::pair<double, int>> &items, std::vector<double> &cov) {
    const double *target = &sd->invUsample[(size_t) i * d];
    for (int m = 0; m < n_obs; m++) {
      const double *c1 = &sd->invUsample[(size_t) m * d];
      double dist      = 0;
      for (int r = 0; r < d; r++) {
        const double df = c1[r] - target[r];
        dist += df * df;
      }
      items[m] = {dist, m};
    }
    const size_t kk = std::min(k, (size_t) n_obs);
    std::partial_sort(items.begin(), items.begin() + kk, items.end());
    if (robust) {
      // ncm_stats_dist_vkde.c:467-472: robust estimators over the neighbours in ascending-distance order
      std::vector<const double *> rows(kk);
      for (size_t j = 0; j < kk; j++) rows[j] = ((NcmVector *) sd->sample[items[j].second])->data;
      ncm_b200_cov_robust(sd->cov_type == NCM_STATS_DIST_KDE_COV_TYPE_ROBUST ? 1 : 0, rows.data(), (int) kk, d, cov.data());
    } else {
      sv.reset();
      for (size_t j = 0; j < kk; j++) sv.append(((NcmVector *) sd->sample[items[j].second])->data);
      sv.get_cov(cov.data());
    }
    ncm_b200_cholesky_decomp_fallback(&sd->cov_slab[(size_t) i * d * d], cov.data(), d, (int) sd->nearPD_maxiter);
  };

  // Device path: kNN + covariance + Cholesky in libncm_sd_gpu, bit-identical to host_centre;
  // only the matrices whose plain Cholesky fails come back to the host for the reference's fallback chain.
  // the robust estimators are sort-bound O(d^2 k log k) work per centre: they stay on the host, OpenMP over the centres
  prof_pre.stop();
  const bool host_only = ncm_b200_host_prepare_kernel() || robust;
  if (!host_only && n_obs <= 65536 && ensure_gpu(sd)) {
    std::vector<int> fail(nk, 0);
    NcmB200ProfScope prof_dev("vkde device path(total)");
    int rc;
    {
      NcmB200ProfScope prof("vkde_prepare(ABI)");
      std::lock_guard<std::mutex> lk(sd->gpu_mutex);
      rc = ncm_sd_gpu_set_kernel(sd->gpu, sd->kernel->kind, sd->kernel->nu, d);
      if (rc == NCM_SD_GPU_OK)
        rc = ncm_sd_gpu_vkde_prepare(sd->gpu, n_obs, nk, sd->sample_matrix.data(), d, sd->invUsample.data(), d, (int) std::min(k, (size_t) n_obs),
                                     sd->cov_slab.data(), fail.data());
    }
    if (!gpu_ok(sd, rc, "_ncm_stats_dist_vkde_build_cov_array_kdtree")) return false;
    std::vector<int> fixed_idx;
    for (int i = 0; i < nk; i++)
      if (fail[i]) fixed_idx.push_back(i);
    std::vector<double> fixed_U((size_t) fixed_idx.size() * d * d);
    if (!fixed_idx.empty()) {
      StatsVec sv(d);
      std::vector<std::pair<double, int>> items(n_obs);
      std::vector<double> cov((size_t) d * d);
      for (size_t f = 0; f < fixed_idx.size(); f++) {
        host_centre(fixed_idx[f], sv, items, cov);
        memcpy(&fixed_U[f * d * d], &sd->cov_slab[(size_t) fixed_idx[f] * d * d], sizeof(double) * d * d);
      }
    }
    {
      NcmB200ProfScope prof("lnnorms(host)");
#pragma omp parallel for if (sd->use_threads)
      for (int i = 0; i < nk; i++) sd->lnnorms[i] = ncm_stats_dist_kernel_get_lnnorm(sd->kernel, sd->cov_array[i]);
    }
    {
      NcmB200ProfScope prof("vkde_finish(ABI)");
      std::lock_guard<std::mutex> lk(sd->gpu_mutex);
      rc = ncm_sd_gpu_vkde_finish(sd->gpu, sd->lnnorms.data(), (int) fixed_idx.size(), fixed_idx.data(), fixed_U.data());
    }
    if (!gpu_ok(sd, rc, "_ncm_stats_dist_vkde_build_cov_array_kdtree")) return false;
    sd->resident = true;
    return true;
  }

#pragma omp parallel if (sd->use_threads)
  {
    StatsVec sv(d);
    std::vector<std::pair<double, int>> items(n_obs);
    std::vector<double> cov((size_t) d * d);
#pragma omp for schedule(dynamic, 1)
    for (int i = 0; i < nk; i++) {
      host_centre(i, sv, items, cov);
      sd->lnnorms[i] = ncm_stats_dist_kernel_get_lnnorm(sd->kernel, sd->cov_array[i]);
    }
  }
  sd->resident = false;
  return true;
}

bool upload(NcmStatsDist *sd) {
  if (!ensure_gpu(sd)) return false;
  if (sd->type == NCM_SD_GPU_VKDE && sd->resident) return true;   // ncm_sd_gpu_vkde_prepare / _finish left everything in HBM
  std::lock_guard<std::mutex> lk(sd->gpu_mutex);
  if (!gpu_ok(sd, ncm_sd_gpu_set_kernel(sd->gpu, sd->kernel->kind, sd->kernel->nu, (int) sd->d), "ncm_stats_dist_prepare_kernel")) return false;
  int rc;
  if (sd->type == NCM_SD_GPU_KDE)
    rc = ncm_sd_gpu_upload_kde(sd->gpu, (int) sd->n_obs, (int) sd->n_kernels, sd->invUsample.data(), (int) sd->d, sd->cov_decomp->data,
                               (int) sd->cov_decomp->tda, sd->kernel_lnnorm);
  else
    rc = ncm_sd_gpu_upload_vkde(sd->gpu, (int) sd->n_obs, (int) sd->n_kernels, sd->sample_matrix.data(), (int) sd->d, sd->cov_slab.data(),
                                sd->lnnorms.data());
  return gpu_ok(sd, rc, "ncm_stats_dist_prepare_kernel");
}

bool prepare_kernel(NcmStatsDist *sd) {
  const double t0 = now_ms();
  if (sd->type == NCM_SD_GPU_VKDE && sd->local_frac * sd->n_obs < 2) {
    ncm_b200_error("Too few observations.\n\tThe number of observations is too small to use the local covariance method.\n"
                   "\tThe local fraction = %f times number of observations %d is less than 2.",
                   sd->local_frac, sd->n_obs);
    return false;
  }
  sd->resident = false;
  if (!kde_prepare_kernel(sd)) return false;
  {
    NcmB200ProfScope prof("vkde_build_cov_array(total)");
    if (sd->type == NCM_SD_GPU_VKDE && !vkde_build_cov_array(sd)) return false;
  }
  sd->host_prepare_kernel_ms += now_ms() - t0;
  NcmB200ProfScope prof("upload");
  return upload(sd);
}

bool push_weights(NcmStatsDist *sd) {
  std::lock_guard<std::mutex> lk(sd->gpu_mutex);
  return gpu_ok(sd, ncm_sd_gpu_set_weights(sd->gpu, (int) sd->n_kernels, sd->weights->data, sd->href), "ncm_stats_dist_prepare");
}

// ---- cross-validation objectives ------------------------------------------------------
// Each one sets over_smooth / href as a side effect, exactly as the reference's callbacks do: after the simplex
// stops, the object keeps the bandwidth of the LAST point evaluated (ncm_stats_dist.c:660-701 never copies the
// best corner back).

void cv_trace_add(NcmStatsDist *sd, double lnos, double val) {
  sd->cv_trace.push_back(lnos);
  sd->cv_trace.push_back(val);
}

// _ncm_stats_dist_m2lnp, ncm_stats_dist.c:484-511: -2 ln L of the held-out observations [n_kernels, n_obs), one batched eval
double cv_obj_m2lnp(NcmStatsDist *sd, double lnos) {
  sd->over_smooth = exp(lnos);
  sd->href        = sd_href(sd);
  const int d = (int) sd->d, q = (int) (sd->n_obs - sd->n_kernels);
  std::vector<double> out((size_t) std::max(q, 1));
  if (!push_weights(sd)) return NAN;
  if (q > 0) {
    std::lock_guard<std::mutex> lk(sd->gpu_mutex);
    if (!gpu_ok(sd, ncm_sd_gpu_eval_m2lnp(sd->gpu, q, &sd->sample_matrix[(size_t) sd->n_kernels * d], d, out.data()), "_ncm_stats_dist_m2lnp")) return NAN;
  }
  double m2lnp = 0.0;
  for (int i = 0; i < q; i++) m2lnp += out[i];
  if (sd->print_fit) fprintf(stderr, "# over-smooth: % 22.15g, m2lnp = % 22.15g\n", sd->over_smooth, m2lnp);
  cv_trace_add(sd, lnos, m2lnp);
  return m2lnp;
}

bool cv_IM_to_host(NcmStatsDist *sd, const char *where) {
  sd->IM_host.resize((size_t) sd->n_obs * sd->n_kernels);
  std::lock_guard<std::mutex> lk(sd->gpu_mutex);
  if (!gpu_ok(sd, ncm_sd_gpu_set_href(sd->gpu, sd->href), where)) return false;
  return gpu_ok(sd, ncm_sd_gpu_compute_IM(sd->gpu, nullptr, sd->IM_host.data()), where);
}

// _ncm_stats_dist_amise_kde_gauss, ncm_stats_dist.c:513-558
double cv_obj_amise_kde_gauss(NcmStatsDist *sd, double lnos) {
  const guint nk = sd->n_kernels;
  double amise   = 0.0;
  sd->over_smooth = exp(lnos);
  sd->href        = sqrt(2.0) * sd_href(sd);
  if (!cv_IM_to_host(sd, "_ncm_stats_dist_amise_kde_gauss")) return NAN;
  const double *IM = sd->IM_host.data();
  for (guint i = 0; i < nk; i++)
    for (guint j = 0; j < nk; j++) amise += IM[(size_t) i * nk + j] / ((double) nk * (double) nk);
  sd->over_smooth = exp(lnos);
  sd->href        = sd_href(sd);
  if (!cv_IM_to_host(sd, "_ncm_stats_dist_amise_kde_gauss")) return NAN;
  IM               = sd->IM_host.data();
  const double nn1 = (double) (nk * (nk - 1));   // guint product, as in the reference
  for (guint i = 0; i < nk; i++) {
    for (guint j = 0; j < i; j++) amise -= 2.0 * IM[(size_t) i * nk + j] / nn1;
    for (guint j = i + 1; j < nk; j++) amise -= 2.0 * IM[(size_t) i * nk + j] / nn1;
  }
  if (sd->print_fit) fprintf(stderr, "# over-smooth: % 22.15g, amise = % 22.15g\n", sd->over_smooth, amise);
  cv_trace_add(sd, lnos, amise);
  return amise;
}

// _ncm_stats_dist_amise, ncm_stats_dist.c:562-658: leave-one-out term from the IM rows + Monte-Carlo integral of p^2 with
// the antithetic kernel pairs of ncm_stats_dist_sample2 (:1629-1651); the two densities of a pair are one q = 2 batch
double cv_obj_amise(NcmStatsDist *sd, double lnos) {
  const guint nk = sd->n_kernels;
  const int d    = (int) sd->d;
  double amise   = 0.0;
  sd->over_smooth = exp(lnos);
  sd->href        = sd_href(sd);
  if (!cv_IM_to_host(sd, "_ncm_stats_dist_amise")) return NAN;
  const double *IM = sd->IM_host.data();
  const double nn1 = (double) (nk * (nk - 1));
  std::vector<double> dens(nk);
  std::vector<size_t> sort(nk);
  for (guint i = 0; i < nk; i++) {
    double row_sum = 0.0;
    for (guint j = 0; j < i; j++) row_sum += IM[(size_t) i * nk + j];
    for (guint j = i + 1; j < nk; j++) row_sum += IM[(size_t) i * nk + j];
    amise -= 2.0 * row_sum / nn1;
    dens[i] = (row_sum + IM[(size_t) i * nk + i]) / nk;
    sort[i] = i;
  }
  // gsl_sort_index is a heapsort; equal keys (not expected for densities) are ordered by index here
  std::stable_sort(sort.begin(), sort.end(), [&](size_t a, size_t b) { return dens[a] < dens[b]; });
  if (!push_weights(sd)) return NAN;

  NcmRNG *rng = ncm_rng_seeded_new(nullptr, 0);
  StatsVec stats(2);
  // The antithetic pairs are drawn from a generator that lives only inside this objective (seeded 0, freed below), so pairs drawn beyond the
  // stopping point cost nothing but time: PAIRS of them are generated in stream order, evaluated in ONE batched GPU call, and then fed to
  // the running statistics one by one with the reference's stopping rule (100 pairs unconditionally, then until the standard error of the mean
  // falls below 1 %).  Same draws, same statistics, same stopping index as the pair-at-a-time loop.
  constexpr int PAIRS = 256;
  NcmMatrix *X = ncm_matrix_new(2 * PAIRS, sd->d);
  std::vector<double> pv(2 * PAIRS);
  double mean = 0.0;
  bool ok = true, stop = false;
  const unsigned long long max_pairs = 100ULL + 100000000ULL;
  for (unsigned long long t0 = 0; t0 < max_pairs && ok && !stop; t0 += PAIRS) {
    for (int b = 0; b < PAIRS; b++) {
      NcmVector x1{X->data + (size_t) (2 * b) * d, sd->d, 1, 1, false}, x2{X->data + (size_t) (2 * b + 1) * d, sd->d, 1, 1, false};
      const guint i   = ncm_stats_dist_kernel_choose(sd, rng);
      const guint o_i = (guint) sort[i];
      ncm_stats_dist_kernel_sample(sd->kernel, ncm_stats_dist_peek_cov_decomp(sd, o_i), sd->href, (NcmVector *) sd->sample[o_i], &x1, rng);
      const guint j   = (guint) sd->sample.size() - 1 - i;
      const guint o_j = (guint) sort[j];
      ncm_stats_dist_kernel_sample(sd->kernel, ncm_stats_dist_peek_cov_decomp(sd, o_j), sd->href, (NcmVector *) sd->sample[o_j], &x2, rng);
    }
    {
      std::lock_guard<std::mutex> lk(sd->gpu_mutex);
      ok = gpu_ok(sd, ncm_sd_gpu_eval(sd->gpu, 2 * PAIRS, X->data, d, pv.data()), "_ncm_stats_dist_amise");
    }
    for (int b = 0; b < PAIRS && ok; b++) {
      const unsigned long long t = t0 + (unsigned long long) b;
      stats.append(&pv[2 * b]);
      if (t >= 100) {
        const double it  = (double) (t - 100);
        mean             = 0.5 * (stats.mean[0] + stats.mean[1]);
        const double var = 0.25 * (stats.var[0] * stats.bias_wt + stats.var[1] * stats.bias_wt + 2.0 * (stats.cov[0 * 2 + 1] * stats.bias_wt));
        const double msd = sqrt(var / (it + 101.0)) / mean;
        if (msd < 1.0e-2) {
          stop = true;
          break;
        }
      }
    }
  }
  amise += mean;
  ncm_matrix_free(X);
  ncm_rng_free(rng);
  if (!ok) return NAN;
  if (sd->print_fit) fprintf(stderr, "# over-smooth: % 22.15g, amise = % 22.15g\n", sd->over_smooth, amise);
  cv_trace_add(sd, lnos, amise);
  return amise;
}

// _ncm_stats_dist_minimize_obj, ncm_stats_dist.c:660-701
void cv_minimize_obj(NcmStatsDist *sd, double (*objective)(NcmStatsDist *, double)) {
  std::function<double(double)> f = [&](double lnos) { return objective(sd, lnos); };
  double x_best = 0.0, f_best = 0.0;
  const int iter = ncm_b200_simplex1_minimize(f, log(sd->over_smooth), 0.1, 1.0e-3, 1000, &x_best, &f_best);
  if (sd->print_fit) printf("# iter: %d, over-smooth: % 22.15g, m2lnp = % 22.15g\n", iter, sd->over_smooth, f_best);
}

// ncm_stats_dist.c:703-789
bool do_prepare(NcmStatsDist *sd) {
  sd->cv_trace.clear();
  switch (sd->cv_type) {
    case NCM_STATS_DIST_CV_LOO:
    case NCM_STATS_DIST_CV_NONE:
      sd->n_obs     = (guint) sd->sample.size();
      sd->n_kernels = (guint) sd->sample.size();
      break;
    case NCM_STATS_DIST_CV_SPLIT:
    case NCM_STATS_DIST_CV_SPLIT_NOFIT:
      sd->n_obs     = (guint) sd->sample.size();
      sd->n_kernels = (guint) ceil(sd->sample.size() * sd->split_frac);
      break;
    default:
      ncm_b200_error("_ncm_stats_dist_prepare: code should not be reached (unknown cross-validation type %d).", (int) sd->cv_type);
      return false;
  }
  if (sd->n_obs <= sd->d) {
    ncm_b200_error("_ncm_stats_dist_prepare: the sample is too small.");
    return false;
  }
  if (!prepare_kernel(sd)) return false;
  if ((sd->weights == nullptr) || (sd->n_kernels != ncm_vector_len(sd->weights))) {
    ncm_vector_clear(&sd->weights);
    ncm_vector_clear(&sd->wcum);
    sd->weights = ncm_vector_new(sd->n_kernels);
    sd->wcum    = ncm_vector_new(sd->n_kernels + 1);
  }
  sd->href = sd_href(sd);
  ncm_vector_set_all(sd->weights, 1.0 / (1.0 * sd->n_kernels));
  sd->wcum_ready = FALSE;
  sd->prepared   = true;
  switch (sd->cv_type) {
    case NCM_STATS_DIST_CV_NONE:
    case NCM_STATS_DIST_CV_SPLIT:
      break;
    case NCM_STATS_DIST_CV_SPLIT_NOFIT:
      cv_minimize_obj(sd, &cv_obj_m2lnp);
      break;
    case NCM_STATS_DIST_CV_LOO:
      if (sd->type == NCM_SD_GPU_KDE && sd->kernel->kind == NCM_SD_GPU_KERNEL_GAUSS)
        cv_minimize_obj(sd, &cv_obj_amise_kde_gauss);
      else
        cv_minimize_obj(sd, &cv_obj_amise);
      break;
    default:
      break;
  }
  return !ncm_b200_error_pending();
}

}   // namespace

extern "C" {

NcmStatsDistKDE *ncm_stats_dist_kde_new(NcmStatsDistKernel *sdk, NcmStatsDistCV CV_type) { return sd_new(NCM_SD_GPU_KDE, sdk, CV_type); }
NcmStatsDistVKDE *ncm_stats_dist_vkde_new(NcmStatsDistKernel *sdk, NcmStatsDistCV CV_type) { return sd_new(NCM_SD_GPU_VKDE, sdk, CV_type); }

NcmStatsDist *ncm_stats_dist_ref(NcmStatsDist *sd) {
  sd->ref++;
  return sd;
}
void ncm_stats_dist_free(NcmStatsDist *sd) {
  if (sd == nullptr || --sd->ref > 0) return;
  ncm_stats_dist_reset(sd);
  for (void *p : sd->obs_pool) ncm_vector_free((NcmVector *) p);
  sd->obs_pool.clear();
  clear_cov_array(sd);
  ncm_stats_dist_kernel_free(sd->kernel);
  ncm_matrix_clear(&sd->cov_fixed);
  ncm_matrix_clear(&sd->cov);
  ncm_matrix_clear(&sd->cov_decomp);
  ncm_vector_clear(&sd->weights);
  ncm_vector_clear(&sd->wcum);
  if (sd->gpu != nullptr) ncm_sd_gpu_ctx_free(sd->gpu);
  ncm_rng_clear(&sd->cv_rng);
  delete sd;
}
void ncm_stats_dist_clear(NcmStatsDist **sd) {
  if (sd != nullptr && *sd != nullptr) {
    ncm_stats_dist_free(*sd);
    *sd = nullptr;
  }
}
NcmStatsDistKDE *ncm_stats_dist_kde_ref(NcmStatsDistKDE *s) { return ncm_stats_dist_ref(s); }
void ncm_stats_dist_kde_free(NcmStatsDistKDE *s) { ncm_stats_dist_free(s); }
void ncm_stats_dist_kde_clear(NcmStatsDistKDE **s) { ncm_stats_dist_clear(s); }
NcmStatsDistVKDE *ncm_stats_dist_vkde_ref(NcmStatsDistVKDE *s) { return ncm_stats_dist_ref(s); }
void ncm_stats_dist_vkde_free(NcmStatsDistVKDE *s) { ncm_stats_dist_free(s); }
void ncm_stats_dist_vkde_clear(NcmStatsDistVKDE **s) { ncm_stats_dist_clear(s); }

void ncm_stats_dist_set_kernel(NcmStatsDist *sd, NcmStatsDistKernel *sdk) {
  ncm_stats_dist_kernel_ref(sdk);
  ncm_stats_dist_kernel_free(sd->kernel);
  sd->kernel = sdk;
  if (sd->d != sdk->d) {
    sd->d = sdk->d;
    ncm_matrix_clear(&sd->cov);
    ncm_matrix_clear(&sd->cov_decomp);
    sd->cov        = ncm_matrix_new(sd->d, sd->d);
    sd->cov_decomp = ncm_matrix_new(sd->d, sd->d);
    clear_cov_array(sd);
  }
  sd->prepared = false;
}
NcmStatsDistKernel *ncm_stats_dist_peek_kernel(NcmStatsDist *sd) { return sd->kernel; }
NcmStatsDistKernel *ncm_stats_dist_get_kernel(NcmStatsDist *sd) { return ncm_stats_dist_kernel_ref(sd->kernel); }

guint ncm_stats_dist_get_dim(NcmStatsDist *sd) { return sd->d; }
guint ncm_stats_dist_get_sample_size(NcmStatsDist *sd) { return (guint) sd->sample.size(); }
guint ncm_stats_dist_get_n_kernels(NcmStatsDist *sd) { return sd->n_kernels; }
gdouble ncm_stats_dist_get_href(NcmStatsDist *sd) { return sd_href(sd); }

void ncm_stats_dist_set_over_smooth(NcmStatsDist *sd, const gdouble os) { sd->over_smooth = os; }
gdouble ncm_stats_dist_get_over_smooth(NcmStatsDist *sd) { return sd->over_smooth; }
// the function's own asserts (ncm_stats_dist.c:1307-1308); the narrower [0.10, 0.95] is the range of the GObject property (:430-434)
void ncm_stats_dist_set_split_frac(NcmStatsDist *sd, const gdouble f) {
  if (!(f >= 0.01)) {
    ncm_b200_error("ncm_stats_dist_set_split_frac: assertion failed (split_frac >= 0.01): (%g >= 0.01)", f);
    return;
  }
  if (!(f <= 1.0)) {
    ncm_b200_error("ncm_stats_dist_set_split_frac: assertion failed (split_frac <= 1.0): (%g <= 1.0)", f);
    return;
  }
  sd->split_frac = f;
}
gdouble ncm_stats_dist_get_split_frac(NcmStatsDist *sd) { return sd->split_frac; }
// ncm_stats_dist.c:1342-1343
void ncm_stats_dist_set_shrink(NcmStatsDist *sd, const gdouble s) {
  if (!(s >= 0.0 && s <= 1.0)) {
    ncm_b200_error("ncm_stats_dist_set_shrink: assertion failed (0.0 <= shrink <= 1.0): (%g)", s);
    return;
  }
  sd->shrink = s;
}
gdouble ncm_stats_dist_get_shrink(NcmStatsDist *sd) { return sd->shrink; }
void ncm_stats_dist_set_print_fit(NcmStatsDist *sd, const gboolean p) { sd->print_fit = p; }
gboolean ncm_stats_dist_get_print_fit(NcmStatsDist *sd) { return sd->print_fit; }
void ncm_stats_dist_set_cv_type(NcmStatsDist *sd, const NcmStatsDistCV cv) { sd->cv_type = cv; }
NcmStatsDistCV ncm_stats_dist_get_cv_type(NcmStatsDist *sd) { return sd->cv_type; }
void ncm_stats_dist_set_use_threads(NcmStatsDist *sd, const gboolean u) { sd->use_threads = u; }
gboolean ncm_stats_dist_get_use_threads(NcmStatsDist *sd) { return sd->use_threads; }

void ncm_stats_dist_kde_set_nearPD_maxiter(NcmStatsDistKDE *sd, const guint maxiter) { sd->nearPD_maxiter = maxiter; }
guint ncm_stats_dist_kde_get_nearPD_maxiter(NcmStatsDistKDE *sd) { return sd->nearPD_maxiter; }
// ncm_stats_dist_kde.c:784-797: switching to FIXED with a matrix already set factors it here
void ncm_stats_dist_kde_set_cov_type(NcmStatsDistKDE *sd, NcmStatsDistKDECovType t) {
  sd->cov_type = t;
  if (sd->cov_type == NCM_STATS_DIST_KDE_COV_TYPE_FIXED && sd->cov_fixed != nullptr) {
    memcpy(sd->cov_decomp->data, sd->cov_fixed->data, sizeof(double) * sd->d * sd->d);
    if (ncm_b200_cholesky_upper(sd->cov_decomp->data, (int) sd->d, (int) sd->d) != 0)
      ncm_b200_error("ncm_stats_dist_kde_set_cov_fixed: matrix cov_fixed is not positive definite.");
  }
}
NcmStatsDistKDECovType ncm_stats_dist_kde_get_cov_type(NcmStatsDistKDE *sd) { return sd->cov_type; }
// ncm_stats_dist_kde.c:825-845
void ncm_stats_dist_kde_set_cov_fixed(NcmStatsDistKDE *sd, NcmMatrix *cov_fixed) {
  if (ncm_matrix_ncols(cov_fixed) != sd->d || ncm_matrix_nrows(cov_fixed) != sd->d) {
    ncm_b200_error("ncm_stats_dist_kde_set_cov_fixed: assertion failed (ncm_matrix_ncols (cov_fixed) == d)");
    return;
  }
  ncm_matrix_clear(&sd->cov_fixed);
  sd->cov_fixed = ncm_matrix_dup(cov_fixed);
  if (sd->cov_type == NCM_STATS_DIST_KDE_COV_TYPE_FIXED) {
    memcpy(sd->cov_decomp->data, sd->cov_fixed->data, sizeof(double) * sd->d * sd->d);
    if (ncm_b200_cholesky_upper(sd->cov_decomp->data, (int) sd->d, (int) sd->d) != 0)
      ncm_b200_error("ncm_stats_dist_kde_set_cov_fixed: matrix cov_fixed is not positive definite.");
  }
}
NcmMatrix *ncm_stats_dist_kde_peek_cov_fixed(NcmStatsDistKDE *sd) { return sd->cov_fixed; }

void ncm_stats_dist_vkde_set_local_frac(NcmStatsDistVKDE *sd, const gdouble lf) {
  if (!(lf >= 0.001 && lf <= 1.0)) {
    ncm_b200_error("ncm_stats_dist_vkde_set_local_frac: assertion failed (0.001 <= local_frac <= 1.0)");
    return;
  }
  sd->local_frac = lf;
}
gdouble ncm_stats_dist_vkde_get_local_frac(NcmStatsDistVKDE *sd) { return sd->local_frac; }
void ncm_stats_dist_vkde_set_use_rot_href(NcmStatsDistVKDE *sd, const gboolean u) { sd->use_rot_href = u; }
gboolean ncm_stats_dist_vkde_get_use_rot_href(NcmStatsDistVKDE *sd) { return sd->use_rot_href; }

// ncm_stats_dist.c:1681-1686: the observation is copied
void ncm_stats_dist_add_obs(NcmStatsDist *sd, NcmVector *y) {
  if (ncm_vector_len(y) != sd->d) {
    ncm_b200_error("ncm_stats_dist_add_obs: assertion failed (ncm_vector_len (y) == d)");
    return;
  }
  // observations are copied (ncm_stats_dist.c:1681-1686); APES resets and refills the object twice per iteration, so the copies released
  // by the last reset are reused instead of going through the allocator 2 x n_obs times
  if (!sd->obs_pool.empty()) {
    NcmVector *v = (NcmVector *) sd->obs_pool.back();
    sd->obs_pool.pop_back();
    for (guint i = 0; i < sd->d; i++) v->data[i] = y->data[(size_t) i * y->stride];
    sd->sample.push_back(v);
    return;
  }
  sd->sample.push_back(ncm_vector_dup(y));
}

void ncm_stats_dist_reset(NcmStatsDist *sd) {
  for (void *p : sd->sample) {
    NcmVector *v = (NcmVector *) p;
    if (v->ref == 1 && v->own && v->len == sd->d && v->stride == 1 && sd->obs_pool.size() < 131072)
      sd->obs_pool.push_back(v);   // nobody else holds it: keep the storage for the next add_obs
    else
      ncm_vector_free(v);
  }
  sd->sample.clear();
  sd->prepared = false;
}

GPtrArray *ncm_stats_dist_peek_sample_array(NcmStatsDist *sd) {
  sd->sample_view.pdata = sd->sample.data();
  sd->sample_view.len   = (guint) sd->sample.size();
  return &sd->sample_view;
}

void ncm_stats_dist_prepare_kernel(NcmStatsDist *sd, GPtrArray *sample_array) {
  // the reference's vfunc receives the object's own sample_array (ncm_stats_dist.c:751)
  if (sample_array != nullptr && sample_array->pdata != sd->sample.data()) {
    ncm_b200_error("ncm_stats_dist_prepare_kernel: only the object's own sample array is supported.");
    return;
  }
  sd->n_obs = sd->n_kernels = (guint) sd->sample.size();
  prepare_kernel(sd);
}

void ncm_stats_dist_prepare(NcmStatsDist *sd) {
  ncm_b200_error_clear();
  if (!do_prepare(sd)) return;
  push_weights(sd);
}

// ncm_stats_dist.c:878-1094 (CV_NONE)
void ncm_stats_dist_prepare_interp(NcmStatsDist *sd, NcmVector *m2lnp) {
  if (ncm_b200_prepare_interp_begin(sd, m2lnp)) ncm_b200_prepare_interp_finish(sd, m2lnp);
}

}   // extern "C"

bool ncm_b200_prepare_interp_begin(NcmStatsDist *sd, NcmVector *m2lnp) {
  ncm_b200_error_clear();
  if (!do_prepare(sd)) return false;
  if (ncm_vector_len(m2lnp) != sd->n_obs) {
    ncm_b200_error("_ncm_stats_dist_prepare_interp: assertion failed (ncm_vector_len (m2lnp) == n_obs): (%u == %u)", ncm_vector_len(m2lnp), sd->n_obs);
    return false;
  }
  return true;
}

void ncm_b200_prepare_interp_finish(NcmStatsDist *sd, NcmVector *m2lnp) {
  const double dbl_limit = 2.0;
  const double range_max = -2.0 * dbl_limit * log(DBL_EPSILON);
  sd->min_m2lnp          = INFINITY;
  sd->max_m2lnp          = -INFINITY;
  for (guint i = 0; i < sd->n_kernels; i++) {
    const double v = ncm_vector_get(m2lnp, i);
    sd->min_m2lnp  = std::min(sd->min_m2lnp, v);
    sd->max_m2lnp  = std::max(sd->max_m2lnp, v);
  }
  if (sd->max_m2lnp - sd->min_m2lnp > range_max) {
    // dynamic-range guard, ncm_stats_dist.c:906-982
    // gsl_sort_index runs over the whole vector (n_obs entries, :912-915); the scan below reads the first n_kernels of them
    std::vector<size_t> sort(sd->n_obs);
    for (guint i = 0; i < sd->n_obs; i++) sort[i] = i;
    std::stable_sort(sort.begin(), sort.end(), [&](size_t a, size_t b) { return ncm_vector_get(m2lnp, (guint) a) < ncm_vector_get(m2lnp, (guint) b); });
    guint n_cut = 0;
    for (guint i = 0; i < sd->n_kernels; i++) {
      if (ncm_vector_get(m2lnp, (guint) sort[i]) - sd->min_m2lnp > range_max) {
        n_cut = i;
        break;
      }
    }
    if (n_cut < (guint) (0.5 * sd->n_obs)) {
      ncm_vector_set_all(sd->weights, 0.1 / (sd->n_kernels - n_cut));
      for (guint i = 0; i < n_cut; i++)
        if (sort[i] < sd->n_kernels) ncm_vector_set(sd->weights, (guint) sort[i], 0.9 / n_cut);   // held-out observations (CV_SPLIT) carry no weight
      push_weights(sd);
      return;   // returns before the shrink normalisation, as the reference does (:934-946)
    }
    // the reference asserts j == n_cut after the copy (:965); with n_obs > n_kernels (CV_SPLIT) the two counts can differ: count first,
    // fail as the reference does instead of writing past m2lnp_cut
    guint n_in = 0;
    for (guint i = 0; i < sd->n_obs; i++)
      if (ncm_vector_get(m2lnp, i) - sd->min_m2lnp <= range_max) n_in++;
    if (n_in != n_cut) {
      ncm_b200_error("_ncm_stats_dist_prepare_interp: assertion failed (j == n_cut): (%u == %u)", n_in, n_cut);
      return;
    }
    NcmVector *m2lnp_cut = ncm_vector_new(n_cut);
    std::vector<void *> keep;
    guint j = 0;
    for (guint i = 0; i < sd->n_obs; i++) {
      const double v = ncm_vector_get(m2lnp, i);
      if (v - sd->min_m2lnp <= range_max) {
        ncm_vector_set(m2lnp_cut, j++, v);
        keep.push_back(sd->sample[i]);
      } else {
        ncm_vector_free((NcmVector *) sd->sample[i]);
      }
    }
    sd->sample.swap(keep);   // the reference mutates sample_array too (:967-972)
    ncm_stats_dist_prepare_interp(sd, m2lnp_cut);
    ncm_vector_free(m2lnp_cut);
    return;
  }

  ncm_vector_set_all(sd->weights, 0.0);
  std::vector<double> inv_f(sd->n_obs);
  for (guint i = 0; i < sd->n_obs; i++) {
    const double f_i = exp(-0.5 * (ncm_vector_get(m2lnp, i) - sd->min_m2lnp));   // :1006-1011
    inv_f[i]         = 1.0 / f_i;                                                 // row scaling of :791-804
  }
  if (sd->n_kernels > 20000) fprintf(stderr, "_ncm_stats_dist_prepare_interp: very large system n = %u!\n", sd->n_kernels);
  // _ncm_stats_dist_compute_IM_full + NCM_NNLS_SOLVE at the current href; the raw solution lands in sd->weights
  auto IM_nnls = [&](double *rnorm_out) -> bool {
    NcmB200ProfScope prof("IM+NNLS(ABI)");
    std::lock_guard<std::mutex> lk(sd->gpu_mutex);
    // compute_IM needs the bandwidth (weights are irrelevant for IM)
    if (!gpu_ok(sd, ncm_sd_gpu_set_href(sd->gpu, sd->href), "prepare_interp")) return false;
    if (!gpu_ok(sd, ncm_sd_gpu_compute_IM(sd->gpu, inv_f.data(), nullptr), "_ncm_stats_dist_compute_IM_full")) return false;
    return gpu_ok(sd, ncm_sd_gpu_nnls_solve(sd->gpu, DBL_EPSILON, sd->weights->data, rnorm_out, &sd->nnls_stats), "ncm_nnls_solve");
  };
  if (sd->cv_type == NCM_STATS_DIST_CV_SPLIT) {
    // ncm_stats_dist.c:1018-1072: ten Gaussian tries around ln over_smooth, then a one-parameter levmar fit of the
    // interpolation residuals expm1 (-(m2lnp_interp - m2lnp_target) / 2) over ALL observations
    const double opts[5] = {1e-3 /* LM_INIT_MU */, 1.0e-7, 1.0e-7, 1.0e-10, 1e-6 /* LM_DIFF_DELTA */};
    double info[10], rnorm0 = 0.0;
    double ln_os = log(sd->over_smooth);
    if (!IM_nnls(&rnorm0)) return;
    cv_trace_add(sd, ln_os, rnorm0);
    for (int i = 0; i < 10; i++) {
      const double ln_os_try = ncm_rng_gaussian_gen(sd->cv_rng, ln_os, 0.5);
      double rnorm_try       = 0.0;
      sd->over_smooth        = exp(ln_os_try);
      sd->href               = sd_href(sd);
      if (!IM_nnls(&rnorm_try)) return;
      cv_trace_add(sd, ln_os_try, rnorm_try);
      if (rnorm_try < rnorm0) {
        ln_os  = ln_os_try;
        rnorm0 = rnorm_try;
      }
    }
    bool ok = true;
    std::vector<double> m2lnpi(sd->n_obs);
    // _ncm_stats_dist_prepare_interp_fit_nnls_f, ncm_stats_dist.c:815-851 (eval_m2lnp reads the raw NNLS solution)
    std::function<void(double, double *)> fit_f = [&](double p, double *hx) {
      double rnorm    = 0.0;
      sd->over_smooth = exp(p);
      sd->href        = sd_href(sd);
      ok              = ok && IM_nnls(&rnorm) && push_weights(sd);
      if (ok) {
        std::lock_guard<std::mutex> lk(sd->gpu_mutex);
        ok = gpu_ok(sd, ncm_sd_gpu_eval_m2lnp(sd->gpu, (int) sd->n_obs, sd->sample_matrix.data(), (int) sd->d, m2lnpi.data()),
                    "_ncm_stats_dist_prepare_interp_fit_nnls_f");
      }
      for (guint i = 0; i < sd->n_obs; i++) {
        const double m2lnpt_i = ncm_vector_get(m2lnp, i) - sd->min_m2lnp;
        hx[i]                 = ok ? expm1(-0.5 * (m2lnpi[i] - m2lnpt_i)) : NAN;
      }
      if (sd->print_fit) fprintf(stderr, "# over-smooth: % 22.15g, rnorm = % 22.15g\n", sd->over_smooth, rnorm);
      cv_trace_add(sd, p, rnorm);
    };
    ncm_b200_lm1_dif(fit_f, &ln_os, nullptr, (int) sd->n_obs, 10000, opts, info);
    if (!ok) return;
    sd->over_smooth = exp(ln_os);
    sd->href        = sd_href(sd);
    if (!IM_nnls(&sd->rnorm)) return;
  } else {
    if (!IM_nnls(&sd->rnorm)) return;
  }
  {
    // ncm_stats_dist.c:1087-1093
    double total_weight = 0.0;
    for (guint i = 0; i < sd->n_kernels; i++) total_weight += sd->weights->data[i];
    if (!(total_weight > 0.0)) {
      ncm_b200_error("_ncm_stats_dist_prepare_interp: assertion failed (total_weight > 0.0)");
      return;
    }
    const double s = (1.0 - sd->shrink) / total_weight, c = sd->shrink / sd->n_kernels;
    for (guint i = 0; i < sd->n_kernels; i++) sd->weights->data[i] *= s;
    for (guint i = 0; i < sd->n_kernels; i++) sd->weights->data[i] += c;
  }
  // the reference does not invalidate wcum here: _ncm_stats_dist_prepare did, and only CV_LOO's sample2 rebuilds it in between
  push_weights(sd);
}

extern "C" {

static bool check_prepared(NcmStatsDist *sd, const char *where) {
  if (!sd->prepared || sd->gpu == nullptr) {
    ncm_b200_error("%s: object not prepared, call ncm_stats_dist_prepare or ncm_stats_dist_prepare_interp first.", where);
    return false;
  }
  return true;
}

gdouble ncm_stats_dist_eval(NcmStatsDist *sd, NcmVector *x) {
  if (!check_prepared(sd, "ncm_stats_dist_eval")) return NAN;
  double xx[NCM_SD_GPU_MAX_DIM], out = NAN;
  for (guint k = 0; k < sd->d; k++) xx[k] = ncm_vector_get(x, k);
  std::lock_guard<std::mutex> lk(sd->gpu_mutex);
  gpu_ok(sd, ncm_sd_gpu_eval(sd->gpu, 1, xx, (int) sd->d, &out), "ncm_stats_dist_eval");
  return out;
}

gdouble ncm_stats_dist_eval_m2lnp(NcmStatsDist *sd, NcmVector *x) {
  if (!check_prepared(sd, "ncm_stats_dist_eval_m2lnp")) return NAN;
  double xx[NCM_SD_GPU_MAX_DIM], out = NAN;
  for (guint k = 0; k < sd->d; k++) xx[k] = ncm_vector_get(x, k);
  std::lock_guard<std::mutex> lk(sd->gpu_mutex);
  gpu_ok(sd, ncm_sd_gpu_eval_m2lnp(sd->gpu, 1, xx, (int) sd->d, &out), "ncm_stats_dist_eval_m2lnp");
  return out;
}

static void eval_array(NcmStatsDist *sd, NcmMatrix *X, NcmVector *out, bool density) {
  if (!check_prepared(sd, "ncm_stats_dist_eval_m2lnp_array")) return;
  if (ncm_matrix_ncols(X) != sd->d || ncm_vector_len(out) != ncm_matrix_nrows(X) || ncm_vector_stride(out) != 1) {
    ncm_b200_error("ncm_stats_dist_eval_m2lnp_array: assertion failed (X is q x d, out has q contiguous entries)");
    return;
  }
  std::lock_guard<std::mutex> lk(sd->gpu_mutex);
  const int rc = density ? ncm_sd_gpu_eval(sd->gpu, (int) X->nrows, X->data, (int) X->tda, out->data)
                         : ncm_sd_gpu_eval_m2lnp(sd->gpu, (int) X->nrows, X->data, (int) X->tda, out->data);
  gpu_ok(sd, rc, "ncm_stats_dist_eval_m2lnp_array");
}
void ncm_stats_dist_eval_m2lnp_array(NcmStatsDist *sd, NcmMatrix *X, NcmVector *out) { eval_array(sd, X, out, false); }
void ncm_stats_dist_eval_array(NcmStatsDist *sd, NcmMatrix *X, NcmVector *out) { eval_array(sd, X, out, true); }

// ncm_stats_dist.c:1565-1606
guint ncm_stats_dist_kernel_choose(NcmStatsDist *sd, NcmRNG *rng) {
  const double p = ncm_rng_uniform_gen(rng, 0.0, 1.0);
  return ncm_b200_kernel_choose_p(sd, p);
}

}   // extern "C"

// the cumulative-weight table and the bisection of kernel_choose for a uniform p already drawn.  (In the reference the table is built before the
// draw; the order is immaterial, the table does not consume the stream.)
guint ncm_b200_kernel_choose_p(NcmStatsDist *sd, double p) {
  if (!sd->wcum_ready) {
    double cum        = 0.0;
    sd->wcum->data[0] = cum;
    for (guint i = 0; i < sd->n_kernels; i++) {
      cum += sd->weights->data[i];
      sd->wcum->data[i + 1] = cum;
    }
    const double s = 1.0 / cum;
    for (guint i = 0; i < sd->n_kernels + 1; i++) sd->wcum->data[i] *= s;
    sd->wcum_ready = TRUE;
  }
  gint ilo = 0, ihi = (gint) sd->n_kernels;
  while (ihi > ilo + 1) {
    const gint mi = (ihi + ilo) / 2;
    if (sd->wcum->data[mi] > p)
      ihi = mi;
    else
      ilo = mi;
  }
  return (guint) ilo;
}

extern "C" {

// ncm_stats_dist.c:1618-1627
void ncm_stats_dist_sample(NcmStatsDist *sd, NcmVector *x, NcmRNG *rng) {
  if (!check_prepared(sd, "ncm_stats_dist_sample")) return;
  const guint i    = ncm_stats_dist_kernel_choose(sd, rng);
  NcmVector *x_i   = (NcmVector *) sd->sample[i];
  NcmMatrix *cov_U = ncm_stats_dist_peek_cov_decomp(sd, i);
  ncm_stats_dist_kernel_sample(sd->kernel, cov_U, sd->href, x_i, x, rng);
}

gdouble ncm_stats_dist_get_rnorm(NcmStatsDist *sd) { return sd->rnorm * sd->rnorm; }   // sic, ncm_stats_dist.c:1664-1669

NcmMatrix *ncm_stats_dist_peek_cov_decomp(NcmStatsDist *sd, guint i) {
  if (sd->type == NCM_SD_GPU_KDE) return sd->cov_decomp;
  if (i >= sd->cov_array.size()) {
    ncm_b200_error("_ncm_stats_dist_vkde_peek_cov_decomp: assertion failed (i < self->cov_array->len)");
    return nullptr;
  }
  return sd->cov_array[i];
}
NcmMatrix *ncm_stats_dist_peek_full_cov_decomp(NcmStatsDist *sd) { return sd->cov_decomp; }
NcmMatrix *ncm_stats_dist_peek_full_cov(NcmStatsDist *sd) { return sd->cov; }
gdouble ncm_stats_dist_get_lnnorm(NcmStatsDist *sd, guint i) {
  if (sd->type == NCM_SD_GPU_KDE) return sd->kernel_lnnorm + sd->d * log(sd->href);
  if (i >= sd->lnnorms.size()) {
    ncm_b200_error("_ncm_stats_dist_vkde_get_lnnorm: assertion failed (i < self->cov_array->len)");
    return NAN;
  }
  return sd->lnnorms[i] + sd->d * log(sd->href);
}
NcmVector *ncm_stats_dist_peek_weights(NcmStatsDist *sd) { return sd->weights; }

// ncm_stats_dist.c:1804-1828
void ncm_stats_dist_get_Ki(NcmStatsDist *sd, const guint i, NcmVector **y_i, NcmMatrix **cov_i, gdouble *n_i, gdouble *w_i) {
  if (i >= sd->sample.size()) {
    ncm_b200_error("ncm_stats_dist_get_Ki: assertion failed (i < ncm_stats_dist_get_sample_size (sd))");
    return;
  }
  NcmMatrix *cd       = ncm_stats_dist_peek_cov_decomp(sd, i);
  const double lnnorm = ncm_stats_dist_get_lnnorm(sd, i);
  const double href   = sd_href(sd);
  const int d         = (int) sd->d;
  y_i[0]              = ncm_vector_dup((NcmVector *) sd->sample[i]);
  cov_i[0]            = ncm_matrix_new(sd->d, sd->d);
  n_i[0]              = exp(lnnorm);
  w_i[0]              = ncm_vector_get(sd->weights, i);
  // ncm_matrix_triang_to_sym (cov_decomp, 'U', TRUE, cov_i): cov = U^T U, then scaled by href^2
  for (int a = 0; a < d; a++)
    for (int b = 0; b < d; b++) {
      double s = 0.0;
      for (int k = 0; k <= std::min(a, b); k++) s += ncm_matrix_get(cd, k, a) * ncm_matrix_get(cd, k, b);
      ncm_matrix_set(cov_i[0], a, b, s * href * href);
    }
}

void ncm_stats_dist_b200_get_nnls_lowrank_stats(NcmStatsDist *sd, gint *n_lowrank, gint *n_fallback, gint *n_trinv, gint *max_k) {
  if (n_lowrank) *n_lowrank = sd->nnls_stats.n_lowrank;
  if (n_fallback) *n_fallback = sd->nnls_stats.n_lowrank_fallback;
  if (n_trinv) *n_trinv = sd->nnls_stats.n_trinv;
  if (max_k) *max_k = sd->nnls_stats.max_lowrank_k;
}

void ncm_stats_dist_b200_get_nnls_fallback_stats(NcmStatsDist *sd, gint *n_lu, gint *n_qr) {
  if (n_lu) *n_lu = sd->nnls_stats.n_lu;
  if (n_qr) *n_qr = sd->nnls_stats.n_qr;
}
void ncm_stats_dist_b200_get_nnls_stats(NcmStatsDist *sd, gint *n_chol, gint *n_lu, gint *n_outer, gint *n_passive) {
  if (n_chol) *n_chol = sd->nnls_stats.n_chol;
  if (n_lu) *n_lu = sd->nnls_stats.n_lu;
  if (n_outer) *n_outer = sd->nnls_stats.n_outer;
  if (n_passive) *n_passive = sd->nnls_stats.n_passive;
}
gint ncm_stats_dist_b200_get_cv_trace(NcmStatsDist *sd, gdouble *lnos, gdouble *val, gint cap) {
  const gint n = (gint) (sd->cv_trace.size() / 2);
  for (gint i = 0; i < n && i < cap; i++) {
    lnos[i] = sd->cv_trace[2 * i];
    val[i]  = sd->cv_trace[2 * i + 1];
  }
  return n;
}
void ncm_stats_dist_b200_get_timers(NcmStatsDist *sd, gdouble *ms7, long long *n_launches, gdouble *host_prepare_kernel_ms) {
  if (sd->gpu != nullptr) ncm_sd_gpu_get_timers(sd->gpu, ms7, n_launches);
  if (host_prepare_kernel_ms) *host_prepare_kernel_ms = sd->host_prepare_kernel_ms;
}
void ncm_stats_dist_b200_enable_timers(NcmStatsDist *sd, gboolean on) {
  if (ensure_gpu(sd)) ncm_sd_gpu_enable_timers(sd->gpu, on);
}
// Multi-rank (SPMD) mode: every rank of the job makes the same calls on the same host data; behind them the interpolation-matrix rows and
// the query rows of this object are sharded over the ranks and exchanged with NCCL (ncm_sd_gpu_set_auto_shard).
gint ncm_stats_dist_b200_comm_unique_id(gchar id_out[128]) { return ncm_sd_gpu_comm_unique_id(id_out); }

gboolean ncm_stats_dist_b200_comm_init(NcmStatsDist *sd, gint nranks, gint rank, const gchar id[128]) {
  if (!ensure_gpu(sd)) return FALSE;
  std::lock_guard<std::mutex> lk(sd->gpu_mutex);
  if (!gpu_ok(sd, ncm_sd_gpu_comm_init(sd->gpu, nranks, rank, id), "ncm_stats_dist_b200_comm_init")) return FALSE;
  return gpu_ok(sd, ncm_sd_gpu_set_auto_shard(sd->gpu, 1), "ncm_stats_dist_b200_comm_init");
}

void *ncm_stats_dist_b200_peek_ctx(NcmStatsDist *sd) {
  ensure_gpu(sd);
  return sd->gpu;
}

}   // extern "C"
