// NcmVector / NcmMatrix surface, error channel and the small dense helpers (d x d) of the host mirror.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cfloat>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "internal.h"

static NcmB200ErrorHandler g_handler   = nullptr;
static void *g_handler_data            = nullptr;
static thread_local bool g_err_pending = false;
static int g_device                    = -1;

extern "C" void ncm_b200_set_error_handler(NcmB200ErrorHandler handler, void *user_data) {
  g_handler      = handler;
  g_handler_data = user_data;
}

extern "C" void ncm_b200_set_device(gint device) { g_device = device; }

// OpenMP threads used by the host-side prepare_kernel (several ranks share one host)
extern "C" void ncm_b200_set_num_threads(gint n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void) n;
#endif
}

static int g_host_prepare = -1;
extern "C" void ncm_b200_set_host_prepare_kernel(gboolean on) { g_host_prepare = on ? 1 : 0; }
bool ncm_b200_host_prepare_kernel() {
  if (g_host_prepare < 0) g_host_prepare = getenv("NCM_B200_HOST_PREPARE_KERNEL") != nullptr ? 1 : 0;
  return g_host_prepare == 1;
}

int ncm_b200_default_device() {
  if (g_device >= 0) return g_device;
  const char *e = getenv("NCM_SD_GPU_DEVICE");
  return e != nullptr ? atoi(e) : 0;
}

// g_error equivalent: message to stderr + abort, or the installed handler
void ncm_b200_error(const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (g_handler != nullptr) {
    g_err_pending = true;
    g_handler(buf, g_handler_data);
    return;
  }
  fprintf(stderr, "\n** (numcosmo_b200) ERROR **: %s\n", buf);
  fflush(stderr);
  abort();
}
bool ncm_b200_error_pending() { return g_err_pending; }
void ncm_b200_error_clear() { g_err_pending = false; }

extern "C" {

NcmVector *ncm_vector_new(const guint n) {
  NcmVector *v = new NcmVector;
  v->data      = (double *) calloc(n > 0 ? n : 1, sizeof(double));
  v->len       = n;
  v->stride    = 1;
  v->ref       = 1;
  v->own       = true;
  return v;
}
NcmVector *ncm_vector_new_data_static(gdouble *d, const guint size, const guint stride) {
  NcmVector *v = new NcmVector;
  v->data      = d;
  v->len       = size;
  v->stride    = stride;
  v->ref       = 1;
  v->own       = false;
  return v;
}
NcmVector *ncm_vector_ref(NcmVector *cv) {
  cv->ref++;
  return cv;
}
NcmVector *ncm_vector_dup(const NcmVector *cv) {
  NcmVector *v = ncm_vector_new(cv->len);
  for (guint i = 0; i < cv->len; i++) v->data[i] = cv->data[(size_t) i * cv->stride];
  return v;
}
void ncm_vector_free(NcmVector *cv) {
  if (cv == nullptr) return;
  if (--cv->ref == 0) {
    if (cv->own) free(cv->data);
    delete cv;
  }
}
void ncm_vector_clear(NcmVector **cv) {
  if (cv != nullptr && *cv != nullptr) {
    ncm_vector_free(*cv);
    *cv = nullptr;
  }
}
guint ncm_vector_len(const NcmVector *cv) { return cv->len; }
guint ncm_vector_stride(const NcmVector *cv) { return cv->stride; }
gdouble *ncm_vector_data(NcmVector *cv) { return cv->data; }
gdouble ncm_vector_get(const NcmVector *cv, const guint i) { return cv->data[(size_t) i * cv->stride]; }
void ncm_vector_set(NcmVector *cv, const guint i, const gdouble val) { cv->data[(size_t) i * cv->stride] = val; }
void ncm_vector_set_all(NcmVector *cv, const gdouble val) {
  for (guint i = 0; i < cv->len; i++) cv->data[(size_t) i * cv->stride] = val;
}
void ncm_vector_memcpy(NcmVector *cv1, const NcmVector *cv2) {
  for (guint i = 0; i < cv1->len; i++) cv1->data[(size_t) i * cv1->stride] = cv2->data[(size_t) i * cv2->stride];
}

NcmMatrix *ncm_matrix_new(const guint nrows, const guint ncols) {
  NcmMatrix *m = new NcmMatrix;
  m->data      = (double *) calloc((size_t) nrows * ncols > 0 ? (size_t) nrows * ncols : 1, sizeof(double));
  m->nrows     = nrows;
  m->ncols     = ncols;
  m->tda       = ncols;
  m->ref       = 1;
  m->own       = true;
  return m;
}
NcmMatrix *ncm_matrix_ref(NcmMatrix *cm) {
  cm->ref++;
  return cm;
}
NcmMatrix *ncm_matrix_dup(const NcmMatrix *cm) {
  NcmMatrix *m = ncm_matrix_new(cm->nrows, cm->ncols);
  for (guint i = 0; i < cm->nrows; i++) memcpy(m->data + (size_t) i * m->tda, cm->data + (size_t) i * cm->tda, sizeof(double) * cm->ncols);
  return m;
}
void ncm_matrix_free(NcmMatrix *cm) {
  if (cm == nullptr) return;
  if (--cm->ref == 0) {
    if (cm->own) free(cm->data);
    delete cm;
  }
}
void ncm_matrix_clear(NcmMatrix **cm) {
  if (cm != nullptr && *cm != nullptr) {
    ncm_matrix_free(*cm);
    *cm = nullptr;
  }
}
guint ncm_matrix_nrows(const NcmMatrix *cm) { return cm->nrows; }
guint ncm_matrix_ncols(const NcmMatrix *cm) { return cm->ncols; }
guint ncm_matrix_tda(const NcmMatrix *cm) { return cm->tda; }
gdouble *ncm_matrix_data(NcmMatrix *cm) { return cm->data; }
gdouble ncm_matrix_get(const NcmMatrix *cm, const guint i, const guint j) { return cm->data[(size_t) i * cm->tda + j]; }
void ncm_matrix_set(NcmMatrix *cm, const guint i, const guint j, const gdouble val) { cm->data[(size_t) i * cm->tda + j] = val; }

}   // extern "C"

// ---- dense helpers ------------------------------------------------------------------------------------

// A = U^T U, U upper triangular, row-major, only the upper triangle is read/written
// (what ncm_matrix_cholesky_decomp (cm, 'U') returns, ncm_matrix.c:1124-1130).  Row-by-row
// (Cholesky-Crout by rows of U): U_ii = sqrt(a_ii - sum_k U_ki^2), U_ij = (a_ij - sum_k U_ki U_kj) / U_ii.
int ncm_b200_cholesky_upper(double *a, int n, int ld) {
  for (int i = 0; i < n; i++) {
    double s = a[i * ld + i];
    for (int k = 0; k < i; k++) s -= a[k * ld + i] * a[k * ld + i];
    if (!(s > 0.0) || !std::isfinite(s)) return i + 1;
    const double uii = sqrt(s);
    a[i * ld + i]    = uii;
    for (int j = i + 1; j < n; j++) {
      double t = a[i * ld + j];
      for (int k = 0; k < i; k++) t -= a[k * ld + i] * a[k * ld + j];
      a[i * ld + j] = t / uii;
    }
  }
  return 0;
}

// ncm_matrix.c:1157-1185
double ncm_b200_cholesky_lndet(const double *U, int n, int ld) {
  const double lb = 1.0e-200, ub = 1.0e+200;
  double detL   = 1.0;
  long exponent = 0;
  for (int i = 0; i < n; i++) {
    const double Lii   = fabs(U[i * ld + i]);
    const double ndetL = detL * Lii;
    if ((ndetL < lb) || (ndetL > ub)) {
      int e = 0;
      detL  = frexp(ndetL, &e);
      exponent += e;
    } else {
      detL = ndetL;
    }
  }
  return 2.0 * (log(detL) + exponent * M_LN2);
}

// cyclic Jacobi eigen-decomposition of a symmetric matrix (n <= 32): A = V diag(w) V^T, V columns
void ncm_b200_jacobi_eig(std::vector<double> &A, int n, std::vector<double> &w, std::vector<double> &V) {
  V.assign((size_t) n * n, 0.0);
  for (int i = 0; i < n; i++) V[i * n + i] = 1.0;
  for (int sweep = 0; sweep < 64; sweep++) {
    double off = 0.0;
    for (int i = 0; i < n; i++)
      for (int j = i + 1; j < n; j++) off += A[i * n + j] * A[i * n + j];
    if (off < 1e-300) break;
    for (int p = 0; p < n; p++) {
      for (int q = p + 1; q < n; q++) {
        const double apq = A[p * n + q];
        if (fabs(apq) < 1e-300) continue;
        const double theta = (A[q * n + q] - A[p * n + p]) / (2.0 * apq);
        const double t     = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; k++) {
          const double akp = A[k * n + p], akq = A[k * n + q];
          A[k * n + p] = c * akp - s * akq;
          A[k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; k++) {
          const double apk = A[p * n + k], aqk = A[q * n + k];
          A[p * n + k] = c * apk - s * aqk;
          A[q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; k++) {
          const double vkp = V[k * n + p], vkq = V[k * n + q];
          V[k * n + p] = c * vkp - s * vkq;
          V[k * n + q] = s * vkp + c * vkq;
        }
      }
    }
  }
  w.resize(n);
  for (int i = 0; i < n; i++) w[i] = A[i * n + i];
}

// Higham (2002) nearest positive-definite iteration as ncm_matrix_nearPD (ncm_matrix.c:1248-1343)
// runs it (UL = 'U', cholesky_decomp = TRUE); the reference's dsyevr is replaced by a Jacobi sweep
// (the matrices are d x d).  Returns 0 when a Cholesky factor was found and stored in a.
int ncm_b200_nearPD_upper(double *a, int n, int maxiter) {
  std::vector<double> cm((size_t) n * n), D_S((size_t) n * n, 0.0), R((size_t) n * n), diag(n), w, V, X((size_t) n * n);
  for (int i = 0; i < n; i++)
    for (int j = i; j < n; j++) cm[i * n + j] = cm[j * n + i] = a[i * n + j];
  for (int i = 0; i < n; i++) diag[i] = cm[i * n + i];
  int ret = 1;
  for (int iter = 0;; iter++) {
    for (int i = 0; i < n * n; i++) cm[i] -= D_S[i];
    R = cm;
    X = cm;
    ncm_b200_jacobi_eig(X, n, w, V);
    double min_pos = INFINITY;
    for (int i = 0; i < n; i++)
      if (w[i] > 0.0 && w[i] < min_pos) min_pos = w[i];
    if (!std::isfinite(min_pos)) return 1;   // negative semi-definite
    for (int i = 0; i < n; i++)
      if (w[i] < 0.0) w[i] = min_pos * DBL_EPSILON;
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) {
        double s = 0.0;
        for (int k = 0; k < n; k++) s += V[i * n + k] * w[k] * V[j * n + k];
        cm[i * n + j] = s;
      }
    for (int i = 0; i < n * n; i++) D_S[i] = cm[i] - R[i];
    for (int i = 0; i < n; i++) cm[i * n + i] = diag[i];
    R   = cm;
    ret = ncm_b200_cholesky_upper(R.data(), n, n);
    if (ret == 0) break;
    if (iter > maxiter) break;
  }
  memcpy(a, R.data(), sizeof(double) * n * n);
  return ret;
}

// _cholesky_decomp of kde.c:344-367 / vkde.c:337-360: dpotrf -> nearPD -> diagonal
void ncm_b200_cholesky_decomp_fallback(double *cov_decomp, const double *cov, int d, int maxiter) {
  memcpy(cov_decomp, cov, sizeof(double) * d * d);
  if (ncm_b200_cholesky_upper(cov_decomp, d, d) != 0) {
    memcpy(cov_decomp, cov, sizeof(double) * d * d);
    if (ncm_b200_nearPD_upper(cov_decomp, d, maxiter) != 0) {
      memset(cov_decomp, 0, sizeof(double) * d * d);
      for (int i = 0; i < d; i++) cov_decomp[i * d + i] = cov[i * d + i];
      ncm_b200_cholesky_upper(cov_decomp, d, d);
    }
  }
}


// ---- NCM_B200_PROFILE_HOST ---------------------------------------------------------------------------------------------------
#include <chrono>
#include <mutex>
namespace {
NcmB200HostProf g_prof[64];
int g_nprof = 0;
std::mutex g_prof_mutex;
void prof_dump() {
  for (int i = 0; i < g_nprof; i++)
    fprintf(stderr, "host_prof: %-34s %10.3f ms %8lld calls %9.2f us/call\n", g_prof[i].name, g_prof[i].ms, g_prof[i].calls,
            g_prof[i].calls ? 1e3 * g_prof[i].ms / g_prof[i].calls : 0.0);
}
}   // namespace
bool ncm_b200_prof_on() {
  static const bool on = getenv("NCM_B200_PROFILE_HOST") != nullptr;
  return on;
}
double ncm_b200_now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
NcmB200HostProf *ncm_b200_prof_slot(const char *name) {
  std::lock_guard<std::mutex> lk(g_prof_mutex);
  for (int i = 0; i < g_nprof; i++)
    if (g_prof[i].name == name || strcmp(g_prof[i].name, name) == 0) return &g_prof[i];
  if (g_nprof == 0) atexit(prof_dump);
  if (g_nprof >= 64) return &g_prof[63];
  g_prof[g_nprof].name = name;
  return &g_prof[g_nprof++];
}
