// Internal definitions of libncm_stats_dist_b200 (host-side mirror of the reference interface).
#pragma once
#include <cstddef>
#include <cstdlib>
#include <functional>
#include <string>
#include <mutex>
#include <vector>
#include "../../include/ncm_sd_gpu.h"
#include "../../include/ncm_stats_dist_b200.h"

// page-locked host array (cudaHostAlloc through the C ABI, plain malloc when no device is usable): the VKDE factor
// slab comes back from the device every prepare_kernel (118 MB at N = 16384, d = 30) and pageable memory halves that copy
// NCM_B200_PROFILE_HOST=1: wall-clock of the host-side stages of the end-to-end path, printed at exit (debugging aid)
struct NcmB200HostProf {
  const char *name;
  double ms = 0.0;
  long long calls = 0;
};
NcmB200HostProf *ncm_b200_prof_slot(const char *name);
bool ncm_b200_prof_on();
double ncm_b200_now_ms();
struct NcmB200ProfScope {
  NcmB200HostProf *s;
  double t0;
  explicit NcmB200ProfScope(const char *name, bool = false) : s(ncm_b200_prof_on() ? ncm_b200_prof_slot(name) : nullptr), t0(s ? ncm_b200_now_ms() : 0.0) {}
  void stop() {
    if (s) {
      s->ms += ncm_b200_now_ms() - t0;
      s->calls++;
      s = nullptr;
    }
  }
  ~NcmB200ProfScope() { stop(); }
};

struct NcmB200PinnedVec {
  double *p = nullptr;
  size_t n = 0;
  bool pinned = false;
  NcmB200PinnedVec() = default;
  NcmB200PinnedVec(const NcmB200PinnedVec &) = delete;
  NcmB200PinnedVec &operator=(const NcmB200PinnedVec &) = delete;
  ~NcmB200PinnedVec() { release(); }
  void release() {
    if (p != nullptr) {
      if (pinned) ncm_sd_gpu_host_free(p); else free(p);
    }
    p = nullptr;
    n = 0;
  }
  void resize(size_t count) {   // contents are unspecified after a change of size
    if (count == n) return;
    release();
    void *q = nullptr;
    if (count > 0 && ncm_sd_gpu_host_alloc(&q, count * sizeof(double)) == NCM_SD_GPU_OK && q != nullptr) {
      p = static_cast<double *>(q);
      pinned = true;
    } else if (count > 0) {
      p = static_cast<double *>(malloc(count * sizeof(double)));
      pinned = false;
    }
    n = count;
  }
  void assign(size_t count, double v) {
    if (count != n) {
      release();
      void *q = nullptr;
      if (count > 0 && ncm_sd_gpu_host_alloc(&q, count * sizeof(double)) == NCM_SD_GPU_OK && q != nullptr) {
        p = static_cast<double *>(q);
        pinned = true;
      } else if (count > 0) {
        p = static_cast<double *>(malloc(count * sizeof(double)));
        pinned = false;
      }
      n = count;
    }
    for (size_t i = 0; i < n; i++) p[i] = v;
  }
  double *data() { return p; }
  double &operator[](size_t i) { return p[i]; }
  size_t size() const { return n; }
};

struct _NcmVector {
  double *data;
  guint len;
  guint stride;
  int ref;
  bool own;
};

struct _NcmMatrix {
  double *data;
  guint nrows, ncols, tda;
  int ref;
  bool own;
};

struct _NcmRNG {
  unsigned long mt[624];
  int mti;
  unsigned long seed;
};

struct _NcmStatsDistKernel {
  int kind;   // NCM_SD_GPU_KERNEL_GAUSS / ST
  guint d;
  double nu;
  int ref;
};

struct _NcmStatsDist {
  int ref;
  std::mutex gpu_mutex;   // serialises the GPU calls of this object (its context, stream and buffers)
  int type;   // NCM_SD_GPU_KDE / VKDE
  NcmStatsDistKernel *kernel;
  guint d;
  // sample_array: GPtrArray of NcmVector (add_obs dups, ncm_stats_dist.c:1681-1686)
  std::vector<void *> sample;
  std::vector<void *> obs_pool;   // observation copies released by reset, reused by add_obs
  GPtrArray sample_view;
  // properties
  double over_smooth, shrink, split_frac, local_frac;
  NcmStatsDistCV cv_type;
  gboolean use_threads, print_fit, use_rot_href;
  NcmStatsDistKDECovType cov_type;
  guint nearPD_maxiter;
  NcmMatrix *cov_fixed;
  // state (NcmStatsDistPrivate, ncm_stats_dist_private.h:39-76)
  guint n_obs, n_kernels;
  double href, min_m2lnp, max_m2lnp, rnorm;
  NcmVector *weights, *wcum;
  gboolean wcum_ready;
  bool prepared;
  // KDE (NcmStatsDistKDEPrivate)
  NcmMatrix *cov, *cov_decomp;
  double kernel_lnnorm;
  NcmB200PinnedVec sample_matrix, invUsample;     // [n_obs x d], page-locked: uploaded at every prepare_kernel
  // VKDE (NcmStatsDistVKDEPrivate): cov_array as live NcmMatrix objects over one slab
  NcmB200PinnedVec cov_slab;                       // [n_kernels x d x d], page-locked
  std::vector<NcmMatrix *> cov_array;
  std::vector<double> lnnorms;
  // GPU
  ncm_sd_gpu_ctx *gpu;
  ncm_sd_gpu_nnls_stats nnls_stats;
  double host_prepare_kernel_ms;
  bool resident;   // prepare_kernel ran on the device: points, factors and records are already in HBM (no upload)
  // cross-validation (ncm_stats_dist.c:175-178): self->rng seeded 0, used by the CV_SPLIT random tries
  NcmRNG *cv_rng;
  std::vector<double> cv_trace;   // (ln over_smooth, objective) per objective evaluation of the last prepare / prepare_interp
  std::vector<double> IM_host;    // host copy of the interpolation matrix for the CV_LOO objectives
};

void ncm_b200_error(const char *fmt, ...);
bool ncm_b200_error_pending();   // true when a handler swallowed an error since the last clear
void ncm_b200_error_clear();

// dense helpers (host side only works on d x d objects and O(N d^2) preparation)
int ncm_b200_cholesky_upper(double *a, int n, int ld);                 // A = U^T U in place (upper); 0 or 1-based failing pivot
int ncm_b200_nearPD_upper(double *a, int n, int maxiter);              // Higham nearPD + Cholesky, ncm_matrix.c:1248-1343
double ncm_b200_cholesky_lndet(const double *U, int n, int ld);        // ncm_matrix.c:1157-1185
void ncm_b200_cholesky_decomp_fallback(double *cov_decomp, const double *cov, int d, int maxiter);   // kde.c:344-367

// one-parameter optimisers of the cross-validation modes (host/optim.cc)
int ncm_b200_simplex1_minimize(const std::function<double(double)> &f, double x0, double step, double size_tol, int max_iter, double *x_best,
                               double *f_best);
int ncm_b200_lm1_dif(const std::function<void(double, double *)> &func, double *p_io, const double *x, int n, int itmax, const double opts[5],
                     double info[10]);

// robust covariance estimators (host/robust.cc) and the symmetric eigen-solver they share with nearPD (host/shim.cc)
double ncm_b200_stats_Qn(std::vector<double> &data);
bool ncm_b200_cov_robust(int kind, const double *const *rows, int n, int d, double *cov);
void ncm_b200_jacobi_eig(std::vector<double> &A, int n, std::vector<double> &w, std::vector<double> &V);

// pieces of kernel->sample / kernel_choose / prepare_interp that apes.cc drives separately (proposal draws generated ahead of the weights)
void ncm_b200_kernel_sample_from(NcmStatsDistKernel *sdk, NcmMatrix *cov_decomp, double href, NcmVector *mu, NcmVector *y, const double *z_raw, double chisq);
guint ncm_b200_kernel_choose_p(NcmStatsDist *sd, double p);                      // ncm_stats_dist_kernel_choose with the uniform already drawn
bool ncm_b200_prepare_interp_begin(NcmStatsDist *sd, NcmVector *m2lnp);          // _ncm_stats_dist_prepare + the length check
void ncm_b200_prepare_interp_finish(NcmStatsDist *sd, NcmVector *m2lnp);         // range guard, IM, NNLS, normalisation

int ncm_b200_default_device();
bool ncm_b200_host_prepare_kernel();   // debugging / parity switch: run the VKDE prepare_kernel loop on the host
