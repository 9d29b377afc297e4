/*
 * oracle/orc_stats_dist.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Restatement of NcmStatsDist / NcmStatsDistKDE / NcmStatsDistVKDE:
 *   numcosmo/ncm/stats/ncm_stats_dist.c:476-482, 703-804, 878-1094, 1548-1627, 1664-1669
 *   numcosmo/ncm/stats/ncm_stats_dist_kde.c:344-367, 378-681
 *   numcosmo/ncm/stats/ncm_stats_dist_vkde.c:316-723
 *   numcosmo/ncm/stats/ncm_stats_vec.c:510-551, 2375-2399 (online covariance)
 *   numcosmo/ncm/algebra/ncm_matrix.c:1124-1130 (dpotrf), 1248-1343 (nearPD)
 *   numcosmo/external/misc/kdtree.c:192-321 + rb_knn_list.c:31-40 (exact kNN, ordered by
 *   (distance, index); restated as a brute-force selection, which returns the same set
 *   in the same order).
 *
 * Cross-validation modes: ncm_stats_dist.c:484-701 (objectives, simplex driver),
 * :703-789 (prepare), :806-876 and :1018-1072 (CV_SPLIT: random tries + levmar fit of ln over_smooth).
 * Robust covariance types (section 8f-4): ncm_stats_vec.c:1821-2072 (Q_n scale per coordinate; OGK).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "ncm_oracle.h"
#include "orc_blas.h"

struct orc_sd
{
  int type;
  orc_kernel kernel;
  int d;
  /* sample array (GPtrArray of NcmVector dup's, ncm_stats_dist.c:1681-1686) */
  double **sample;
  int n_sample, cap_sample;
  /* properties (defaults ncm_stats_dist.c:389-450, kde.c:290-310, vkde.c:276-290) */
  double over_smooth, shrink, split_frac, local_frac;
  int cv_type, use_threads, use_rot_href, cov_type, nearPD_maxiter;
  double *cov_fixed;
  /* state */
  int n_obs, n_kernels;
  double href;
  double *weights, *wcum;
  int weights_len, wcum_ready;
  double min_m2lnp, max_m2lnp, rnorm;
  double *IM, *f;
  int alloc_n_obs, alloc_n_kernels;
  /* KDE */
  double *cov, *cov_decomp;
  double kernel_lnnorm;
  double *sample_matrix, *invUsample;
  int sm_rows;
  /* VKDE */
  double *cov_array, *lnnorms;
  int cov_array_len;
  orc_nnls_stats nnls_stats;
  double timers[3];
  /* cross-validation: self->fmin (nmsimplex2, 1 parameter), self->rng (seeded 0), ncm_stats_dist.c:175-178 */
  orc_nmsimplex2 *fmin;
  orc_rng cv_rng;
  const double *cv_m2lnp; /* the target of the CV_SPLIT fit (NcmStatsDistEval.m2lnp) */
  double *cv_trace;       /* (ln over_smooth, objective) per objective evaluation */
  int cv_trace_len, cv_trace_cap;
};

/* The reference environment pins an OpenMP build of OpenBLAS (environment.yml), where BLAS calls made
 * from inside an OpenMP region run single-threaded.  SciPy's OpenBLAS is a pthreads build: the same
 * nesting serialises on its global lock.  Around the OpenMP regions the oracle therefore switches
 * OpenBLAS to one thread and restores the previous count afterwards. */
static int
blas_enter_omp (void)
{
  const int prev = scipy_openblas_get_num_threads ();

  scipy_openblas_set_num_threads (1);

  return prev;
}

static void
blas_leave_omp (int prev)
{
  scipy_openblas_set_num_threads (prev);
}

/* gsl_blas_dtrsv (CblasUpper, CblasTrans, CblasNonUnit, U, x) for the d x d factors of the per-point
 * evaluation loops, written out as the forward substitution it is.  OpenBLAS takes a global buffer lock
 * in every level-2 call, which under the walkers-parallel OpenMP loop made these d <= 30 calls ~20x
 * slower than the arithmetic; inlining gives the CPU baseline its best case.  daxpy / ddot likewise. */
static inline void
trsv_upper_trans (const int d, const double *U, const int ld, double *x)
{
  int k, j;

  for (k = 0; k < d; k++)
  {
    double t = x[k];

    for (j = 0; j < k; j++)
      t -= U[j * ld + k] * x[j];

    x[k] = t / U[k * ld + k];
  }
}

static double
now_s (void)
{
  struct timespec ts;

  clock_gettime (CLOCK_MONOTONIC, &ts);

  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

orc_sd *
orc_sd_new (int type, int kernel_kind, double nu, int d, int cv_type)
{
  orc_sd *sd = (orc_sd *) calloc (1, sizeof (orc_sd));

  sd->type           = type;
  sd->kernel.kind    = kernel_kind;
  sd->kernel.nu      = nu;
  sd->kernel.d       = d;
  sd->d              = d;
  sd->cv_type        = cv_type;
  sd->over_smooth    = 1.0;
  sd->shrink         = 0.01;
  sd->split_frac     = 0.5;
  sd->local_frac     = 0.05;
  sd->use_threads    = 0;
  sd->use_rot_href   = 0;
  sd->cov_type       = ORC_COV_SAMPLE;
  sd->nearPD_maxiter = 200;
  sd->cov            = (double *) calloc ((size_t) d * d, sizeof (double));
  sd->cov_decomp     = (double *) calloc ((size_t) d * d, sizeof (double));
  sd->fmin           = orc_nmsimplex2_new (1);
  orc_rng_set (&sd->cv_rng, 0);

  return sd;
}

void
orc_sd_free (orc_sd *sd)
{
  orc_sd_reset (sd);
  free (sd->sample);
  free (sd->cov_fixed);
  free (sd->weights);
  free (sd->wcum);
  free (sd->IM);
  free (sd->f);
  free (sd->cov);
  free (sd->cov_decomp);
  free (sd->sample_matrix);
  free (sd->invUsample);
  free (sd->cov_array);
  free (sd->lnnorms);
  free (sd->cv_trace);
  orc_nmsimplex2_free (sd->fmin);
  free (sd);
}

void orc_sd_set_over_smooth (orc_sd *sd, double os) { sd->over_smooth = os; }
void orc_sd_set_shrink (orc_sd *sd, double shrink) { sd->shrink = shrink; }
void orc_sd_set_split_frac (orc_sd *sd, double split_frac) { sd->split_frac = split_frac; }
void orc_sd_set_use_threads (orc_sd *sd, int use_threads) { sd->use_threads = use_threads; }
void orc_sd_set_cov_type (orc_sd *sd, int cov_type) { sd->cov_type = cov_type; }
void orc_sd_set_nearPD_maxiter (orc_sd *sd, int maxiter) { sd->nearPD_maxiter = maxiter; }
void orc_sd_set_local_frac (orc_sd *sd, double local_frac) { sd->local_frac = local_frac; }
void orc_sd_set_use_rot_href (orc_sd *sd, int use_rot_href) { sd->use_rot_href = use_rot_href; }

void
orc_sd_set_cov_fixed (orc_sd *sd, const double *cov, int ld)
{
  const int d = sd->d;
  int i, j;

  free (sd->cov_fixed);
  sd->cov_fixed = (double *) malloc (sizeof (double) * d * d);

  for (i = 0; i < d; i++)
    for (j = 0; j < d; j++)
      sd->cov_fixed[i * d + j] = cov[i * ld + j];
}

void
orc_sd_reset (orc_sd *sd)
{
  int i;

  for (i = 0; i < sd->n_sample; i++)
    free (sd->sample[i]);

  sd->n_sample = 0;
}

/* ncm_stats_dist.c:1681-1686: add_obs copies the vector */
void
orc_sd_add_obs (orc_sd *sd, const double *x)
{
  if (sd->n_sample == sd->cap_sample)
  {
    sd->cap_sample = sd->cap_sample ? 2 * sd->cap_sample : 64;
    sd->sample     = (double **) realloc (sd->sample, sizeof (double *) * sd->cap_sample);
  }

  sd->sample[sd->n_sample] = (double *) malloc (sizeof (double) * sd->d);
  memcpy (sd->sample[sd->n_sample], x, sizeof (double) * sd->d);
  sd->n_sample++;
}

/* ncm_matrix.c:1124-1130: ncm_lapack_dpotrf flips 'U' -> 'L' (ncm_lapack.c:58,216-226) */
int
orc_cholesky_decomp_U (double *a, int n, int ld)
{
  int info = 0;

  scipy_dpotrf_ ("L", &n, a, &ld, &info);

  return info;
}

/* ncm_matrix.c:1248-1343 (Higham nearPD, UL = 'U', cholesky_decomp = TRUE) */
static int
orc_matrix_nearPD (double *cm, int n, int maxiter)
{
  double *eva  = (double *) malloc (sizeof (double) * n);
  double *diag = (double *) malloc (sizeof (double) * n);
  double *eve  = (double *) malloc (sizeof (double) * n * n);
  double *D_S  = (double *) calloc ((size_t) n * n, sizeof (double));
  double *R    = (double *) malloc (sizeof (double) * n * n);
  int *isuppz  = (int *) malloc (sizeof (int) * 2 * n);
  int neva     = 0;
  int ret, i, iter = 0;

  for (i = 0; i < n; i++)
    diag[i] = cm[i * n + i];

  while (1)
  {
    double min_pos_ev = INFINITY;
    const double zero = 0.0;
    const int izero   = 0;
    int lwork = -1, liwork = -1, info = 0, liwq;
    double wq, *work;
    int *iwork;

    for (i = 0; i < n * n; i++)
      cm[i] -= D_S[i];

    memcpy (R, cm, sizeof (double) * n * n);

    /* ncm_lapack_dsyevr ('V', 'A', 'U' -> 'L', ...) */
    scipy_dsyevr_ ("V", "A", "L", &n, cm, &n, &zero, &zero, &izero, &izero, &zero, &neva, eva, eve, &n, isuppz, &wq, &lwork, &liwq, &liwork, &info);
    lwork  = (int) wq;
    liwork = liwq;
    work   = (double *) malloc (sizeof (double) * lwork);
    iwork  = (int *) malloc (sizeof (int) * liwork);
    scipy_dsyevr_ ("V", "A", "L", &n, cm, &n, &zero, &zero, &izero, &izero, &zero, &neva, eva, eve, &n, isuppz, work, &lwork, iwork, &liwork, &info);
    free (work);
    free (iwork);

    if (neva == 0)
    {
      ret = -1;
      break;
    }

    for (i = 0; i < neva; i++)
    {
      if (eva[i] > 0.0)
        min_pos_ev = (eva[i] < min_pos_ev) ? eva[i] : min_pos_ev;
    }

    for (i = 0; i < neva; i++)
    {
      if (eva[i] < 0.0)
        eva[i] = min_pos_ev * DBL_EPSILON;

      scipy_cblas_dscal (n, sqrt (eva[i]), &eve[i * n], 1);
    }

    scipy_cblas_dsyrk (OrcRowMajor, OrcUpper, OrcTrans, n, neva, 1.0, eve, n, 0.0, cm, n);

    for (i = 0; i < n * n; i++)
      D_S[i] = cm[i] - R[i];

    for (i = 0; i < n; i++)
      cm[i * n + i] = diag[i];

    memcpy (R, cm, sizeof (double) * n * n);

    if ((ret = orc_cholesky_decomp_U (R, n, n)) == 0)
      break;

    if (iter > maxiter)
      break;

    iter++;
  }

  memcpy (cm, R, sizeof (double) * n * n);

  free (eva);
  free (diag);
  free (eve);
  free (D_S);
  free (R);
  free (isuppz);

  return ret;
}

/* kde.c:344-367 == vkde.c:337-360 */
static void
_cholesky_decomp (double *cov_decomp, const double *cov, const int d, const int maxiter)
{
  memcpy (cov_decomp, cov, sizeof (double) * d * d);

  if (orc_cholesky_decomp_U (cov_decomp, d, d) != 0)
  {
    memcpy (cov_decomp, cov, sizeof (double) * d * d);

    if (orc_matrix_nearPD (cov_decomp, d, maxiter) != 0)
    {
      int i;

      memset (cov_decomp, 0, sizeof (double) * d * d);

      for (i = 0; i < d; i++)
        cov_decomp[i * d + i] = cov[i * d + i];

      orc_cholesky_decomp_U (cov_decomp, d, d);
    }
  }
}

/* NcmStatsVec with NCM_STATS_VEC_COV: ncm_stats_vec.c:510-551 update, :2375-2399 get_cov_matrix */
typedef struct orc_stats_vec
{
  int len;
  double weight, weight2, bias_wt;
  double *mean, *var, *cov;
} orc_stats_vec;

static void
svec_init (orc_stats_vec *s, int len)
{
  s->len  = len;
  s->mean = (double *) malloc (sizeof (double) * len);
  s->var  = (double *) malloc (sizeof (double) * len);
  s->cov  = (double *) malloc (sizeof (double) * len * len);
}

static void
svec_reset (orc_stats_vec *s)
{
  s->weight  = 0.0;
  s->weight2 = 0.0;
  s->bias_wt = 0.0;
  memset (s->mean, 0, sizeof (double) * s->len);
  memset (s->var, 0, sizeof (double) * s->len);
  memset (s->cov, 0, sizeof (double) * s->len * s->len);
}

static void
svec_free (orc_stats_vec *s)
{
  free (s->mean);
  free (s->var);
  free (s->cov);
}

static void
svec_append (orc_stats_vec *s, const double *x)
{
  const double w         = 1.0;
  const double curweight = s->weight + w;
  const int sveclen      = s->len;
  int i;

  for (i = 0; i < sveclen; i++)
  {
    int j;
    double mean_i        = s->mean[i];
    const double x_i     = x[i];
    const double delta_i = x_i - mean_i;
    const double R_i     = delta_i * w / curweight;
    const double var     = s->var[i];
    const double dvar    = s->weight * delta_i * R_i;

    mean_i    += R_i;
    s->mean[i] = mean_i;
    s->var[i]  = var + dvar;

    for (j = i + 1; j < sveclen; j++)
    {
      const double x_j    = x[j];
      const double mean_j = s->mean[j];
      const double dC_ij  = w * (x_i - mean_i) * (x_j - mean_j);
      const double oC_ij  = s->cov[i * sveclen + j];
      const double C_ij   = oC_ij + dC_ij;

      s->cov[i * sveclen + j] = C_ij;
      s->cov[j * sveclen + i] = C_ij;
    }
  }

  s->weight   = curweight;
  s->weight2 += w * w;
  s->bias_wt  = 1.0 / (s->weight - s->weight2 / s->weight);
}

static void
svec_get_cov_matrix (const orc_stats_vec *s, double *m)
{
  const int n = s->len;
  int i;

  memcpy (m, s->cov, sizeof (double) * n * n);

  for (i = 0; i < n; i++)
    m[i * n + i] = s->var[i];

  for (i = 0; i < n * n; i++)
    m[i] *= s->bias_wt;
}

/* ------------------------------------------------------------------------------------------------ */
/* Robust covariance types: ncm_stats_vec.c:1821-2090                       */
/* ------------------------------------------------------------------------------------------------ */

static int
dbl_cmp (const void *a, const void *b)
{
  const double x = *(const double *) a, y = *(const double *) b;

  return (x > y) - (x < y);
}

/* gsl_stats_Qn_from_sorted_data (GSL >= 2.5, statistics/Qn.c; GSL is absent from /root/reference and from this image:
 * PARITY UNPINNED for the finite-sample factors d_n below).  Rousseeuw & Croux's Q_n: the k-th smallest of the
 * n (n - 1) / 2 differences x_i - x_j (i > j), k = h (h - 1) / 2, h = n / 2 + 1, times 2.21914 times d_n.  GSL finds the
 * order statistic with the O(n log n) algorithm of Croux & Rousseeuw (1992); an order statistic does not depend on how
 * it is found, so the restatement simply sorts all the differences. */
double
orc_stats_Qn_from_sorted_data (const double *sorted, int n)
{
  const double scale = 2.21914;
  const long h = n / 2 + 1, k = h * (h - 1) / 2;
  const size_t npairs = (size_t) n * (n - 1) / 2;
  double *diff = (double *) malloc (sizeof (double) * (npairs > 0 ? npairs : 1));
  double Qn0, dn = 1.0;
  size_t t = 0;
  int i, j;

  for (i = 1; i < n; i++)
    for (j = 0; j < i; j++)
      diff[t++] = sorted[i] - sorted[j];

  qsort (diff, npairs, sizeof (double), dbl_cmp);
  Qn0 = (npairs > 0) ? diff[k - 1] : 0.0;
  free (diff);

  if (n <= 12)
  {
    static const double dn_small[13] = {1.0, 1.0, 0.399356, 0.99365, 0.51321, 0.84401, 0.61220, 0.85877, 0.66993, 0.87344, 0.72014, 0.88906, 0.75743};

    dn = dn_small[n];
  }
  else
  {
    if (n % 2 == 1)
      dn = 1.60188 + (-2.1284 - 5.172 / n) / n;
    else
      dn = 3.67561 + (1.9654 + (6.987 - 77.0 / n) / n) / n;

    dn = 1.0 / (dn / (double) n + 1.0);
  }

  return scale * dn * Qn0;
}

static double
column_Qn (const double *col, int stride, int n, double *data)
{
  int a;

  for (a = 0; a < n; a++)
    data[a] = col[(size_t) a * stride];

  qsort (data, n, sizeof (double), dbl_cmp); /* gsl_sort */

  return orc_stats_Qn_from_sorted_data (data, n);
}

/* ncm_stats_vec_compute_cov_robust_diag, ncm_stats_vec.c:1832-1882: rows[n] points to the saved observations */
int
orc_cov_robust_diag (const double *const *rows, int n, int d, double *cov)
{
  double *data = (double *) malloc (sizeof (double) * n);
  int i, a;

  if (n < 4)
  {
    free (data);

    return -7; /* g_error ("... too few points to estimate the covariance") */
  }

  memset (cov, 0, sizeof (double) * d * d);

  for (i = 0; i < d; i++)
  {
    double s;

    for (a = 0; a < n; a++)
      data[a] = rows[a][i];

    qsort (data, n, sizeof (double), dbl_cmp);
    s              = orc_stats_Qn_from_sorted_data (data, n);
    cov[i * d + i] = s * s;
  }

  free (data);

  return 0;
}

/* ncm_stats_vec_compute_cov_robust_ogk, ncm_stats_vec.c:1896-2072 (orthogonalised Gnanadesikan-Kettenring).  Only the
 * upper triangle of the result is defined by the reference (dsyrk 'U' over a matrix dsyevr destroyed); the restatement
 * mirrors it into the lower one. */
int
orc_cov_robust_ogk (const double *const *rows, int n, int d, double *cov)
{
  double *E       = (double *) malloc (sizeof (double) * d * d);
  double *y       = (double *) malloc (sizeof (double) * (size_t) n * d);
  double *z       = (double *) malloc (sizeof (double) * (size_t) n * d);
  double *sigma_x = (double *) malloc (sizeof (double) * d);
  double *sigma_z = (double *) malloc (sizeof (double) * d);
  double *data    = (double *) malloc (sizeof (double) * (n > d ? n : d));
  int a, i, j;

  if (n < 4)
  {
    free (E); free (y); free (z); free (sigma_x); free (sigma_z); free (data);

    return -7;
  }

  for (a = 0; a < n; a++)
    memcpy (&y[(size_t) a * d], rows[a], sizeof (double) * d);

  for (i = 0; i < d; i++)
  {
    const double sigma_i = column_Qn (&y[i], d, n, data);
    const double s       = 1.0 / sigma_i;

    sigma_x[i] = sigma_i;

    for (a = 0; a < n; a++)
      y[(size_t) a * d + i] *= s;
  }

  memset (cov, 0, sizeof (double) * d * d);

  for (i = 0; i < d; i++)
    cov[i * d + i] = 1.0;

  for (i = 0; i < d; i++)
  {
    for (j = i + 1; j < d; j++)
    {
      double s_ipj, s_imj;

      for (a = 0; a < n; a++)
        data[a] = y[(size_t) a * d + i] + y[(size_t) a * d + j];

      qsort (data, n, sizeof (double), dbl_cmp);
      s_ipj = orc_stats_Qn_from_sorted_data (data, n);

      for (a = 0; a < n; a++)
        data[a] = y[(size_t) a * d + i] - y[(size_t) a * d + j];

      qsort (data, n, sizeof (double), dbl_cmp);
      s_imj = orc_stats_Qn_from_sorted_data (data, n);

      cov[i * d + j] = 0.25 * (s_ipj * s_ipj - s_imj * s_imj);
    }
  }

  {
    /* ncm_lapack_dsyevr ('V', 'A', 'U' -> 'L', ...): eigenvector k lands in row k of the row-major E */
    const double zero = 0.0;
    const int izero   = 0;
    int lwork = -1, liwork = -1, info = 0, liwq, neval = 0;
    int *isuppz = (int *) malloc (sizeof (int) * 2 * d);
    double wq, *work;
    int *iwork;

    scipy_dsyevr_ ("V", "A", "L", &d, cov, &d, &zero, &zero, &izero, &izero, &zero, &neval, data, E, &d, isuppz, &wq, &lwork, &liwq, &liwork, &info);
    lwork  = (int) wq;
    liwork = liwq;
    work   = (double *) malloc (sizeof (double) * lwork);
    iwork  = (int *) malloc (sizeof (int) * liwork);
    scipy_dsyevr_ ("V", "A", "L", &d, cov, &d, &zero, &zero, &izero, &izero, &zero, &neval, data, E, &d, isuppz, work, &lwork, iwork, &liwork, &info);
    free (work);
    free (iwork);
    free (isuppz);

    if (info != 0)
    {
      free (E); free (y); free (z); free (sigma_x); free (sigma_z); free (data);

      return -8;
    }
  }

  /* z = E y^T  (d x n) */
  scipy_cblas_dgemm (OrcRowMajor, OrcNoTrans, OrcTrans, d, n, d, 1.0, E, d, y, d, 0.0, z, n);

  for (i = 0; i < d; i++)
    sigma_z[i] = column_Qn (&z[(size_t) i * n], 1, n, data);

  for (i = 0; i < d; i++)
    for (j = 0; j < d; j++)
      E[i * d + j] = sigma_z[i] * E[i * d + j] * sigma_x[j];

  memset (cov, 0, sizeof (double) * d * d);
  scipy_cblas_dsyrk (OrcRowMajor, OrcUpper, OrcTrans, d, d, 1.0, E, d, 0.0, cov, d);

  for (i = 0; i < d; i++)
    for (j = 0; j < i; j++)
      cov[i * d + j] = cov[j * d + i];

  free (E); free (y); free (z); free (sigma_x); free (sigma_z); free (data);

  return 0;
}

/* kde.c:378-490 */
static int
kde_prepare_kernel (orc_sd *sd)
{
  const int d = sd->d;
  orc_stats_vec sv;
  double *cov = (double *) malloc (sizeof (double) * d * d);
  int i;

  svec_init (&sv, d);
  svec_reset (&sv);

  for (i = 0; i < sd->n_kernels; i++)
    svec_append (&sv, sd->sample[i]);

  switch (sd->cov_type)
  {
    case ORC_COV_SAMPLE:
      svec_get_cov_matrix (&sv, cov);
      _cholesky_decomp (sd->cov_decomp, cov, d, sd->nearPD_maxiter);
      memcpy (sd->cov, cov, sizeof (double) * d * d);
      break;
    case ORC_COV_FIXED:

      if (sd->cov_fixed == NULL)
      {
        svec_free (&sv);
        free (cov);

        return -3;
      }

      /* kde.c:424-432: FIXED only saves cov; cov_decomp keeps its previous content.  The
       * reference relies on set_cov_fixed having decomposed it (kde.c:803-820). */
      memcpy (sd->cov, sd->cov_fixed, sizeof (double) * d * d);
      _cholesky_decomp (sd->cov_decomp, sd->cov_fixed, d, sd->nearPD_maxiter);
      break;
    case ORC_COV_ROBUST_DIAG:
    case ORC_COV_ROBUST:
    {
      /* kde.c:423-441 */
      const int rc = (sd->cov_type == ORC_COV_ROBUST_DIAG) ? orc_cov_robust_diag ((const double *const *) sd->sample, sd->n_kernels, d, cov)
                                                           : orc_cov_robust_ogk ((const double *const *) sd->sample, sd->n_kernels, d, cov);

      if (rc != 0)
      {
        svec_free (&sv);
        free (cov);

        return rc;
      }

      _cholesky_decomp (sd->cov_decomp, cov, d, sd->nearPD_maxiter);
      memcpy (sd->cov, cov, sizeof (double) * d * d);
      break;
    }
    default:
      svec_free (&sv);
      free (cov);

      return -2;
  }

  sd->kernel_lnnorm = orc_kernel_get_lnnorm (&sd->kernel, sd->cov_decomp, d);

  if ((sd->sample_matrix == NULL) || (sd->sm_rows != sd->n_obs))
  {
    free (sd->sample_matrix);
    free (sd->invUsample);
    sd->sample_matrix = (double *) malloc (sizeof (double) * sd->n_obs * d);
    sd->invUsample    = (double *) malloc (sizeof (double) * sd->n_obs * d);
    sd->sm_rows       = sd->n_obs;
  }

  for (i = 0; i < sd->n_obs; i++)
    memcpy (&sd->sample_matrix[(size_t) i * d], sd->sample[i], sizeof (double) * d);

  memcpy (sd->invUsample, sd->sample_matrix, sizeof (double) * sd->n_obs * d);

  /* gsl_blas_dtrsm (CblasRight, CblasUpper, CblasNoTrans, CblasNonUnit, 1.0, cov_decomp, invUsample) */
  scipy_cblas_dtrsm (OrcRowMajor, OrcRight, OrcUpper, OrcNoTrans, OrcNonUnit, sd->n_obs, d, 1.0, sd->cov_decomp, d, sd->invUsample, d);

  svec_free (&sv);
  free (cov);

  return 0;
}

typedef struct knn_item
{
  double dist;
  int idx;
} knn_item;

static int
knn_cmp (const void *a, const void *b)
{
  const knn_item *ka = (const knn_item *) a, *kb = (const knn_item *) b;

  if (ka->dist < kb->dist)
    return -1;

  if (ka->dist > kb->dist)
    return 1;

  return (ka->idx > kb->idx) - (ka->idx < kb->idx);
}

/* the neighbour search of vkde_build_cov_array on its own (test hook): the k nearest of points[n x d] to points[query],
 * ordered by (squared distance, index), as kdtree.c:229-321 + rb_knn_list.c:31-40 return them */
void
orc_knn_brute (const double *points, int n, int d, int query, int k, long *idx_out, double *dist_out)
{
  knn_item *items      = (knn_item *) malloc (sizeof (knn_item) * n);
  const double *target = &points[(size_t) query * d];
  int m, r;

  for (m = 0; m < n; m++)
  {
    const double *c1 = &points[(size_t) m * d];
    double dist      = 0;

    for (r = 0; r < d; r++)
    {
      const double df = c1[r] - target[r];

      dist += df * df;
    }

    items[m].dist = dist;
    items[m].idx  = m;
  }

  qsort (items, n, sizeof (knn_item), knn_cmp);

  for (m = 0; m < k; m++)
  {
    idx_out[m]  = items[m].idx;
    dist_out[m] = items[m].dist;
  }

  free (items);
}

/* vkde.c:362-496 */
static int
vkde_build_cov_array (orc_sd *sd)
{
  const int d = sd->d;
  const int n_obs = sd->n_obs, n_kernels = sd->n_kernels;
  /* vkde.c:426: const size_t k = GSL_MAX (local_frac * n_obs, 2) */
  const double kd = (sd->local_frac * n_obs > 2.0) ? sd->local_frac * n_obs : 2.0;
  const size_t k  = (size_t) kd;
  int i;

  int robust_rc = 0;

  if ((sd->cov_type < ORC_COV_SAMPLE) || (sd->cov_type > ORC_COV_ROBUST))
    return -2;

  if ((sd->cov_type >= ORC_COV_ROBUST_DIAG) && (k < 4))
    return -7; /* ncm_stats_vec.c:1842-1844: too few points to estimate the covariance */

  if (sd->cov_array_len != n_kernels)
  {
    free (sd->cov_array);
    free (sd->lnnorms);
    sd->cov_array     = (double *) malloc (sizeof (double) * (size_t) n_kernels * d * d);
    sd->lnnorms       = (double *) malloc (sizeof (double) * n_kernels);
    sd->cov_array_len = n_kernels;
  }

  const int blas_prev = blas_enter_omp ();

  #pragma omp parallel if (sd->use_threads)
  {
    orc_stats_vec sv;
    knn_item *items = (knn_item *) malloc (sizeof (knn_item) * n_obs);
    double *cov     = (double *) malloc (sizeof (double) * d * d);

    svec_init (&sv, d);

    #pragma omp for schedule(dynamic, 1)

    for (i = 0; i < n_kernels; i++)
    {
      const double *target = &sd->invUsample[(size_t) i * d];
      size_t j;
      int m;

      /* kdtree.c:27-38 distance(): sum of squared differences in index order (node - target) */
      for (m = 0; m < n_obs; m++)
      {
        const double *c1 = &sd->invUsample[(size_t) m * d];
        double dist      = 0;
        int r;

        for (r = 0; r < d; r++)
        {
          const double df = c1[r] - target[r];

          dist += df * df;
        }

        items[m].dist = dist;
        items[m].idx  = m;
      }

      qsort (items, n_obs, sizeof (knn_item), knn_cmp);

      svec_reset (&sv);

      for (j = 0; j < k; j++)
        svec_append (&sv, sd->sample[items[j].idx]);

      if (sd->cov_type >= ORC_COV_ROBUST_DIAG)
      {
        /* vkde.c:467-472: the neighbours in ascending-distance order are the saved rows of the NcmStatsVec */
        const double **rows = (const double **) malloc (sizeof (double *) * k);
        int rc;

        for (j = 0; j < k; j++)
          rows[j] = sd->sample[items[j].idx];

        rc = (sd->cov_type == ORC_COV_ROBUST_DIAG) ? orc_cov_robust_diag (rows, (int) k, d, cov) : orc_cov_robust_ogk (rows, (int) k, d, cov);
        free (rows);

        if (rc != 0)
          robust_rc = rc;
      }
      else
      {
        svec_get_cov_matrix (&sv, cov);
      }

      _cholesky_decomp (&sd->cov_array[(size_t) i * d * d], cov, d, sd->nearPD_maxiter);
      sd->lnnorms[i] = orc_kernel_get_lnnorm (&sd->kernel, &sd->cov_array[(size_t) i * d * d], d);
    }

    svec_free (&sv);
    free (items);
    free (cov);
  }

  blas_leave_omp (blas_prev);

  return robust_rc;
}

/* ncm_stats_dist.c:476-482 ; vkde.c:316-335 */
double
orc_sd_get_href (orc_sd *sd)
{
  const double base = sd->over_smooth * orc_kernel_get_rot_bandwidth (&sd->kernel, sd->n_kernels);

  if (sd->type == ORC_SD_KDE)
    return base;

  if (sd->use_rot_href)
    return base / sd->local_frac;
  else
    return sd->over_smooth;
}

static void
cv_trace_add (orc_sd *sd, double lnos, double val)
{
  if (sd->cv_trace_len == sd->cv_trace_cap)
  {
    sd->cv_trace_cap = sd->cv_trace_cap ? 2 * sd->cv_trace_cap : 64;
    sd->cv_trace     = (double *) realloc (sd->cv_trace, sizeof (double) * 2 * sd->cv_trace_cap);
  }

  sd->cv_trace[2 * sd->cv_trace_len + 0] = lnos;
  sd->cv_trace[2 * sd->cv_trace_len + 1] = val;
  sd->cv_trace_len++;
}

/* _ncm_stats_dist_m2lnp, ncm_stats_dist.c:484-511: the CV_SPLIT_NOFIT objective, minus twice the log-likelihood
 * of the held-out observations.  The reference accumulates inside an OpenMP loop without a reduction clause; the
 * defined (use_threads = FALSE) behaviour is the index-ordered sum restated here. */
static double
cv_obj_m2lnp (const double *v, int n, void *params)
{
  orc_sd *sd        = (orc_sd *) params;
  const double lnos = v[0];
  double m2lnp      = 0.0;
  int i;

  sd->over_smooth = exp (lnos);
  sd->href        = orc_sd_get_href (sd);

  for (i = sd->n_kernels; i < sd->n_obs; i++)
    m2lnp += orc_sd_eval_m2lnp (sd, sd->sample[i]);

  cv_trace_add (sd, lnos, m2lnp);

  return m2lnp;
}

static void
cv_alloc_IM (orc_sd *sd)
{
  if ((sd->n_obs != sd->alloc_n_obs) || (sd->n_kernels != sd->alloc_n_kernels))
  {
    free (sd->IM);
    free (sd->f);
    sd->IM = (double *) malloc (sizeof (double) * (size_t) sd->n_obs * sd->n_kernels);
    sd->f  = (double *) malloc (sizeof (double) * sd->n_obs);

    sd->alloc_n_obs     = sd->n_obs;
    sd->alloc_n_kernels = sd->n_kernels;
  }
}

/* _ncm_stats_dist_amise_kde_gauss, ncm_stats_dist.c:513-558 (CV_LOO, KDE with the Gaussian kernel) */
static double
cv_obj_amise_kde_gauss (const double *v, int n, void *params)
{
  orc_sd *sd        = (orc_sd *) params;
  const double lnos = v[0];
  const int nk      = sd->n_kernels;
  double amise      = 0.0;
  int i, j;

  sd->over_smooth = exp (lnos);
  sd->href        = sqrt (2.0) * orc_sd_get_href (sd);

  orc_sd_compute_IM (sd, sd->IM);

  for (i = 0; i < nk; i++)
    for (j = 0; j < nk; j++)
      amise += sd->IM[(size_t) i * nk + j] / ((double) nk * (double) nk);

  sd->over_smooth = exp (lnos);
  sd->href        = orc_sd_get_href (sd);

  orc_sd_compute_IM (sd, sd->IM);

  for (i = 0; i < nk; i++)
  {
    for (j = 0; j < i; j++)
      amise -= 2.0 * sd->IM[(size_t) i * nk + j] / (double) ((unsigned int) nk * (unsigned int) (nk - 1));

    for (j = i + 1; j < nk; j++)
      amise -= 2.0 * sd->IM[(size_t) i * nk + j] / (double) ((unsigned int) nk * (unsigned int) (nk - 1));
  }

  cv_trace_add (sd, lnos, amise);

  return amise;
}

static int idx_cmp_ctx (const void *a, const void *b, void *ctx);

/* ncm_stats_dist_sample2, ncm_stats_dist.c:1629-1651: antithetic pair of kernels, ranks i and len - 1 - i of the
 * sorted leave-one-out densities */
static void
cv_sample2 (orc_sd *sd, const size_t *sort, double *x1, double *x2, orc_rng *rng)
{
  const int i   = orc_sd_kernel_choose (sd, rng);
  const int o_i = (int) sort[i];

  orc_kernel_sample (&sd->kernel, orc_sd_peek_cov_decomp (sd, o_i), sd->d, sd->href, sd->sample[o_i], x1, rng);

  {
    const int j   = sd->n_sample - 1 - i;
    const int o_j = (int) sort[j];

    orc_kernel_sample (&sd->kernel, orc_sd_peek_cov_decomp (sd, o_j), sd->d, sd->href, sd->sample[o_j], x2, rng);
  }
}

/* _ncm_stats_dist_amise, ncm_stats_dist.c:562-658 (CV_LOO, every other class / kernel pair): leave-one-out term
 * from the interpolation matrix plus a Monte-Carlo estimate of the integral of p^2 */
static double
cv_obj_amise (const double *v, int n, void *params)
{
  orc_sd *sd        = (orc_sd *) params;
  const double lnos = v[0];
  const int nk      = sd->n_kernels;
  double amise      = 0.0;
  double *dens      = (double *) malloc (sizeof (double) * nk);
  size_t *sort      = (size_t *) malloc (sizeof (size_t) * nk);
  int i, j;

  sd->over_smooth = exp (lnos);
  sd->href        = orc_sd_get_href (sd);

  orc_sd_compute_IM (sd, sd->IM);

  for (i = 0; i < nk; i++)
  {
    double row_sum = 0.0;

    for (j = 0; j < i; j++)
      row_sum += sd->IM[(size_t) i * nk + j];

    for (j = i + 1; j < nk; j++)
      row_sum += sd->IM[(size_t) i * nk + j];

    amise -= 2.0 * row_sum / (double) ((unsigned int) nk * (unsigned int) (nk - 1));

    dens[i] = (row_sum + sd->IM[(size_t) i * nk + i]) / nk;
    sort[i] = i;
  }

  /* gsl_sort_index (heapsort, not stable): ties broken by index here */
  qsort_r (sort, nk, sizeof (size_t), idx_cmp_ctx, (void *) dens);

  {
    orc_rng rng;
    orc_stats_vec stats;
    double *x1 = (double *) malloc (sizeof (double) * sd->d);
    double *x2 = (double *) malloc (sizeof (double) * sd->d);
    const unsigned int max_iter = 100000000;
    double mean = 0.0, p12[2];
    unsigned int it;

    orc_rng_set (&rng, 0);
    svec_init (&stats, 2);
    svec_reset (&stats);

    for (it = 0; it < 100; it++)
    {
      cv_sample2 (sd, sort, x1, x2, &rng);
      p12[0] = orc_sd_eval (sd, x1);
      p12[1] = orc_sd_eval (sd, x2);
      svec_append (&stats, p12);
    }

    for (it = 0; it < max_iter; it++)
    {
      cv_sample2 (sd, sort, x1, x2, &rng);
      p12[0] = orc_sd_eval (sd, x1);
      p12[1] = orc_sd_eval (sd, x2);
      svec_append (&stats, p12);

      mean = 0.5 * (stats.mean[0] + stats.mean[1]);

      {
        const double var = 0.25 * (stats.var[0] * stats.bias_wt + stats.var[1] * stats.bias_wt + 2.0 * (stats.cov[0 * 2 + 1] * stats.bias_wt));
        const double msd = sqrt (var / (it + 101.0)) / mean;

        if (msd < 1.0e-2)
          break;
      }
    }

    amise += mean;

    svec_free (&stats);
    free (x1);
    free (x2);
  }

  free (dens);
  free (sort);

  cv_trace_add (sd, lnos, amise);

  return amise;
}

/* _ncm_stats_dist_minimize_obj, ncm_stats_dist.c:660-701.  The objective's side effects stay: over_smooth and
 * href are those of the LAST point the simplex evaluated, not of the best corner. */
static void
cv_minimize_obj (orc_sd *sd, orc_fmin_fn objective)
{
  const double s    = 0.1;
  const double lnos = log (sd->over_smooth);

  orc_nmsimplex2_minimize (sd->fmin, objective, sd, &lnos, &s, 1.0e-3, 1000);
}

/* ncm_stats_dist.c:703-789 */
int
orc_sd_prepare (orc_sd *sd)
{
  double t0 = now_s ();
  int ret, i;

  sd->cv_trace_len = 0;

  switch (sd->cv_type)
  {
    case ORC_CV_LOO:
      sd->n_obs     = sd->n_sample;
      sd->n_kernels = sd->n_sample;
      cv_alloc_IM (sd);
      break;
    case ORC_CV_NONE:
      sd->n_obs     = sd->n_sample;
      sd->n_kernels = sd->n_sample;
      break;
    case ORC_CV_SPLIT:
    case ORC_CV_SPLIT_NOFIT:
      sd->n_obs     = sd->n_sample;
      sd->n_kernels = (int) ceil (sd->n_sample * sd->split_frac);
      break;
    default:
      return -2;
  }

  if (sd->n_obs <= sd->d)
    return -1; /* g_error ("_ncm_stats_dist_prepare: the sample is too small.") */

  if (sd->type == ORC_SD_VKDE)
  {
    /* vkde.c:505-510 */
    if (sd->local_frac * sd->n_obs < 2)
      return -4; /* g_error ("Too few observations...") */
  }

  if ((ret = kde_prepare_kernel (sd)) != 0)
    return ret;

  if (sd->type == ORC_SD_VKDE)
    if ((ret = vkde_build_cov_array (sd)) != 0)
      return ret;

  if ((sd->weights == NULL) || (sd->n_kernels != sd->weights_len))
  {
    free (sd->weights);
    free (sd->wcum);
    sd->weights     = (double *) malloc (sizeof (double) * sd->n_kernels);
    sd->wcum        = (double *) malloc (sizeof (double) * (sd->n_kernels + 1));
    sd->weights_len = sd->n_kernels;
  }

  sd->href = orc_sd_get_href (sd);

  for (i = 0; i < sd->n_kernels; i++)
    sd->weights[i] = 1.0 / (1.0 * sd->n_kernels);

  sd->wcum_ready = 0;

  switch (sd->cv_type)
  {
    case ORC_CV_NONE:
    case ORC_CV_SPLIT:
      break;
    case ORC_CV_SPLIT_NOFIT:
      cv_minimize_obj (sd, &cv_obj_m2lnp);
      break;
    case ORC_CV_LOO:

      if ((sd->type == ORC_SD_KDE) && (sd->kernel.kind == ORC_KERNEL_GAUSS))
        cv_minimize_obj (sd, &cv_obj_amise_kde_gauss);
      else
        cv_minimize_obj (sd, &cv_obj_amise);

      break;
  }

  sd->timers[0] = now_s () - t0;

  return 0;
}

/* kde.c:492-557 */
static void
kde_compute_IM (orc_sd *sd, double *IM)
{
  const int d        = sd->d;
  const int nk       = sd->n_kernels;
  const double href2 = sd->href * sd->href;
  int i;

  for (i = 0; i < nk; i++)
  {
    const double *row_i = &sd->invUsample[(size_t) i * d];
    int j;

    IM[(size_t) i * nk + i] = 0.0;

    for (j = i + 1; j < nk; j++)
    {
      const double *row_j = &sd->invUsample[(size_t) j * d];
      double chi2_ij      = 0.0;
      int k;

      for (k = 0; k < d; k++)
      {
        const double df = row_i[k] - row_j[k];

        chi2_ij += df * df;
      }

      chi2_ij = chi2_ij / href2;

      IM[(size_t) i * nk + j] = chi2_ij;
      IM[(size_t) j * nk + i] = chi2_ij;
    }
  }

  for (i = nk; i < sd->n_obs; i++)
  {
    const double *row_i = &sd->invUsample[(size_t) i * d];
    int j;

    for (j = 0; j < nk; j++)
    {
      const double *row_j = &sd->invUsample[(size_t) j * d];
      double chi2_ij      = 0.0;
      int k;

      for (k = 0; k < d; k++)
      {
        const double df = row_i[k] - row_j[k];

        chi2_ij += df * df;
      }

      chi2_ij = chi2_ij / href2;

      IM[(size_t) i * nk + j] = chi2_ij;
    }
  }

  for (i = 0; i < sd->n_obs; i++)
    orc_kernel_eval_unnorm_vec (&sd->kernel, &IM[(size_t) i * nk], 1, &IM[(size_t) i * nk], 1, nk);

  {
    const double scale = exp (-(sd->kernel_lnnorm + d * log (sd->href)));
    size_t t;

    for (t = 0; t < (size_t) sd->n_obs * nk; t++)
      IM[t] *= scale;
  }
}

/* vkde.c:517-606 */
static void
vkde_compute_IM (orc_sd *sd, double *IM)
{
  const int d            = sd->d;
  const int nk           = sd->n_kernels;
  const int n_obs        = sd->n_obs;
  const double href2     = sd->href * sd->href;
  const double one_href2 = 1.0 / href2;
  const int blas_prev = blas_enter_omp ();
  int i;

  #pragma omp parallel if (sd->use_threads)
  {
    double *invUsample_matrix = (double *) malloc (sizeof (double) * (size_t) n_obs * d);

    #pragma omp for schedule(dynamic, 1)

    for (i = 0; i < nk; i++)
    {
      const double *cov_decomp_i = &sd->cov_array[(size_t) i * d * d];
      const double *theta_i      = sd->sample[i];
      int j;

      memcpy (invUsample_matrix, sd->sample_matrix, sizeof (double) * (size_t) n_obs * d);

      for (j = 0; j < n_obs; j++)
      {
        double *theta_j = &invUsample_matrix[(size_t) j * d];

        scipy_cblas_daxpy (d, -1.0, theta_i, 1, theta_j, 1);
      }

      scipy_cblas_dtrsm (OrcRowMajor, OrcRight, OrcUpper, OrcNoTrans, OrcNonUnit, n_obs, d, 1.0, cov_decomp_i, d, invUsample_matrix, d);

      for (j = 0; j < n_obs; j++)
      {
        double *theta_j = &invUsample_matrix[(size_t) j * d];
        double chi2_ij;

        chi2_ij = scipy_cblas_ddot (d, theta_j, 1, theta_j, 1) * one_href2;

        IM[(size_t) j * nk + i] = chi2_ij;
      }
    }

    free (invUsample_matrix);
  }

  blas_leave_omp (blas_prev);

  {
    const double lnnorm_href = d * log (sd->href);

    for (i = 0; i < n_obs; i++)
      orc_kernel_eval_unnorm_vec (&sd->kernel, &IM[(size_t) i * nk], 1, &IM[(size_t) i * nk], 1, nk);

    for (i = 0; i < nk; i++)
    {
      const double norm_i = exp (sd->lnnorms[i] + lnnorm_href);
      const double s      = 1.0 / norm_i;
      int j;

      for (j = 0; j < n_obs; j++)
        IM[(size_t) j * nk + i] *= s;
    }
  }
}

void
orc_sd_compute_IM (orc_sd *sd, double *IM)
{
  if (sd->type == ORC_SD_KDE)
    kde_compute_IM (sd, IM);
  else
    vkde_compute_IM (sd, IM);
}

static int
idx_cmp_ctx (const void *a, const void *b, void *ctx)
{
  const double *v = (const double *) ctx;
  const size_t ia = *(const size_t *) a, ib = *(const size_t *) b;

  if (v[ia] < v[ib])
    return -1;

  if (v[ia] > v[ib])
    return 1;

  return (ia > ib) - (ia < ib);
}

/* _ncm_stats_dist_compute_IM_full, ncm_stats_dist.c:791-804 */
static void
sd_compute_IM_full (orc_sd *sd)
{
  const int nk = sd->n_kernels;
  int i;

  orc_sd_compute_IM (sd, sd->IM);

  #pragma omp parallel for if (sd->use_threads)

  for (i = 0; i < sd->n_obs; i++)
  {
    const double s = 1.0 / sd->f[i];
    int j;

    for (j = 0; j < nk; j++)
      sd->IM[(size_t) i * nk + j] *= s;
  }
}

/* NCM_NNLS_SOLVE (self->nnls, self->sub_IM, self->sub_x, self->f1); reltol default GSL_DBL_EPSILON, ncm_nnls.c:271-275 */
static double
sd_nnls (orc_sd *sd)
{
  double *f1 = (double *) malloc (sizeof (double) * sd->n_obs);
  double rnorm;
  int i;

  for (i = 0; i < sd->n_obs; i++)
    f1[i] = 1.0;

  rnorm = orc_nnls_solve (sd->IM, sd->n_obs, sd->n_kernels, sd->n_kernels, sd->weights, f1, DBL_EPSILON, &sd->nnls_stats);
  free (f1);

  return rnorm;
}

double orc_sd_eval_m2lnp (orc_sd *sd, const double *x);

/* _ncm_stats_dist_prepare_interp_fit_nnls_f, ncm_stats_dist.c:815-851: residuals of the CV_SPLIT fit at ln over_smooth = p[0].
 * eval_m2lnp reads the raw NNLS solution left in self->weights (no normalisation, no shrink). */
static void
cv_fit_nnls_f (double *p, double *hx, int m, int n, void *adata)
{
  orc_sd *sd = (orc_sd *) adata;
  double rnorm;
  int i;

  sd->over_smooth = exp (p[0]);
  sd->href        = orc_sd_get_href (sd);

  sd_compute_IM_full (sd);
  rnorm = sd_nnls (sd);

  #pragma omp parallel for if (sd->use_threads)

  for (i = 0; i < sd->n_obs; i++)
  {
    const double m2lnpt_i = sd->cv_m2lnp[i] - sd->min_m2lnp;
    const double m2lnpi_i = orc_sd_eval_m2lnp (sd, sd->sample[i]);

    hx[i] = expm1 (-0.5 * (m2lnpi_i - m2lnpt_i));
  }

  cv_trace_add (sd, p[0], rnorm);
}

/* dlevmar_dif: the reference's own levmar when oracle/_ref/liblevmar_ref.so was registered (orc_set_levmar_dif),
 * the restatement of orc_optim.c otherwise */
typedef int (*dlevmar_dif_fn) (void (*func) (double *, double *, int, int, void *), double *p, double *x, int m, int n, int itmax, double *opts, double *info, double *work, double *covar, void *adata);
static dlevmar_dif_fn ref_dlevmar_dif = NULL;

void
orc_set_levmar_dif (void *fn)
{
  ref_dlevmar_dif = (dlevmar_dif_fn) fn;
}

static int
cv_lm_dif (orc_lm_fn func, double *p, const double *x, int m, int n, int itmax, const double *opts, double *info, void *adata)
{
  if (ref_dlevmar_dif != NULL)
  {
    double o[5] = {opts[0], opts[1], opts[2], opts[3], opts[4]};

    return ref_dlevmar_dif (func, p, (double *) x, m, n, itmax, o, info, NULL, NULL, adata);
  }

  return orc_lm_dif (func, p, x, m, n, itmax, opts, info, adata);
}

/* ncm_stats_dist.c:878-1094 */
int
orc_sd_prepare_interp (orc_sd *sd, const double *m2lnp, int n)
{
  const double dbl_limit = 2.0;
  const double range_max = -2.0 * dbl_limit * log (DBL_EPSILON); /* GSL_LOG_DBL_EPSILON */
  int ret, i;

  if ((ret = orc_sd_prepare (sd)) != 0)
    return ret;

  if (n != sd->n_obs)
    return -5;

  sd->min_m2lnp = INFINITY;
  sd->max_m2lnp = -INFINITY;

  for (i = 0; i < sd->n_kernels; i++)
  {
    const double m2lnp_i = m2lnp[i];

    sd->min_m2lnp = (sd->min_m2lnp < m2lnp_i) ? sd->min_m2lnp : m2lnp_i;
    sd->max_m2lnp = (sd->max_m2lnp > m2lnp_i) ? sd->max_m2lnp : m2lnp_i;
  }

  if (sd->max_m2lnp - sd->min_m2lnp > range_max)
  {
    /* ncm_stats_dist.c:906-982.  gsl_sort_index is a heapsort (not stable); ties are
     * broken here by index, which only matters for exactly equal m2lnp values. */
    /* The reference sorts the whole vector, ncm_vector_len (m2lnp) = n_obs entries (:912-915; its index array is sized
     * n_kernels, which under CV_SPLIT, n_obs > n_kernels, it overruns); the scan below reads the first n_kernels. */
    size_t *sort = (size_t *) malloc (sizeof (size_t) * sd->n_obs);
    int n_cut    = 0;

    for (i = 0; i < sd->n_obs; i++)
      sort[i] = i;

    qsort_r (sort, sd->n_obs, sizeof (size_t), idx_cmp_ctx, (void *) m2lnp);

    for (i = 0; i < sd->n_kernels; i++)
    {
      const size_t p       = sort[i];
      const double m2lnp_p = m2lnp[p];

      if (m2lnp_p - sd->min_m2lnp > range_max)
      {
        n_cut = i;
        break;
      }
    }

    if (n_cut < (int) (0.5 * sd->n_obs))
    {
      for (i = 0; i < sd->n_kernels; i++)
        sd->weights[i] = 0.1 / (sd->n_kernels - n_cut);

      for (i = 0; i < n_cut; i++)
        if ((int) sort[i] < sd->n_kernels)   /* held-out observations (CV_SPLIT) carry no weight */
          sd->weights[sort[i]] = 0.9 / n_cut;

      free (sort);

      return 0;
    }

    {
      double *m2lnp_cut, **cut;
      int j = 0;

      for (i = 0; i < sd->n_obs; i++)
        if (m2lnp[i] - sd->min_m2lnp <= range_max)
          j++;

      if (j != n_cut)   /* g_assert (j == n_cut), :965 */
      {
        free (sort);

        return -5;
      }

      m2lnp_cut = (double *) malloc (sizeof (double) * n_cut);
      cut       = (double **) malloc (sizeof (double *) * n_cut);
      j         = 0;

      for (i = 0; i < sd->n_obs; i++)
      {
        const double m2lnp_i = m2lnp[i];

        if (m2lnp_i - sd->min_m2lnp <= range_max)
        {
          m2lnp_cut[j] = m2lnp_i;
          cut[j]       = sd->sample[i];
          j++;
        }
        else
        {
          free (sd->sample[i]);
        }
      }

      for (i = 0; i < n_cut; i++)
        sd->sample[i] = cut[i];

      sd->n_sample = n_cut;

      ret = orc_sd_prepare_interp (sd, m2lnp_cut, n_cut);

      free (m2lnp_cut);
      free (cut);
      free (sort);

      return ret;
    }
  }

  if ((sd->n_obs != sd->alloc_n_obs) || (sd->n_kernels != sd->alloc_n_kernels))
  {
    free (sd->IM);
    free (sd->f);
    sd->IM = (double *) malloc (sizeof (double) * (size_t) sd->n_obs * sd->n_kernels);
    sd->f  = (double *) malloc (sizeof (double) * sd->n_obs);

    sd->alloc_n_obs     = sd->n_obs;
    sd->alloc_n_kernels = sd->n_kernels;
  }

  memset (sd->weights, 0, sizeof (double) * sd->n_kernels);

  for (i = 0; i < sd->n_obs; i++)
    sd->f[i] = exp (-0.5 * (m2lnp[i] - sd->min_m2lnp));

  switch (sd->cv_type)
  {
    case ORC_CV_SPLIT:
    {
      /* ncm_stats_dist.c:1018-1072 */
      const double opts[5] = {1e-3 /* LM_INIT_MU */, 1.0e-7, 1.0e-7, 1.0e-10, 1e-6 /* LM_DIFF_DELTA */};
      double info[10];
      double ln_os, rnorm0;
      double t0 = now_s ();

      ln_os = log (sd->over_smooth);

      sd_compute_IM_full (sd);
      rnorm0 = sd_nnls (sd);
      cv_trace_add (sd, ln_os, rnorm0);

      for (i = 0; i < 10; i++)
      {
        const double ln_os_try = orc_ran_gaussian (&sd->cv_rng, 0.5) + ln_os;
        double rnorm_try;

        sd->over_smooth = exp (ln_os_try);
        sd->href        = orc_sd_get_href (sd);

        sd_compute_IM_full (sd);
        rnorm_try = sd_nnls (sd);
        cv_trace_add (sd, ln_os_try, rnorm_try);

        if (rnorm_try < rnorm0)
        {
          ln_os  = ln_os_try;
          rnorm0 = rnorm_try;
        }
      }

      sd->cv_m2lnp = m2lnp;
      cv_lm_dif (&cv_fit_nnls_f, &ln_os, NULL, 1, sd->n_obs, 10000, opts, info, sd);
      sd->cv_m2lnp = NULL;

      sd->over_smooth = exp (ln_os);
      sd->href        = orc_sd_get_href (sd);

      sd_compute_IM_full (sd);
      sd->rnorm     = sd_nnls (sd);
      sd->timers[2] = now_s () - t0;
      break;
    }
    default:
    {
      double t0 = now_s (), t1;

      sd_compute_IM_full (sd);
      t1            = now_s ();
      sd->timers[1] = t1 - t0;
      sd->rnorm     = sd_nnls (sd);
      sd->timers[2] = now_s () - t1;
      break;
    }
  }

  {
    double total_weight = 0.0;

    for (i = 0; i < sd->n_kernels; i++)
      total_weight += sd->weights[i];

    if (!(total_weight > 0.0))
      return -6;

    {
      const double s = (1.0 - sd->shrink) / total_weight;
      const double c = sd->shrink / sd->n_kernels;

      for (i = 0; i < sd->n_kernels; i++)
        sd->weights[i] *= s;

      for (i = 0; i < sd->n_kernels; i++)
        sd->weights[i] += c;
    }
  }

  return 0;
}

/* kde.c:639-681 */
static double
kde_eval_m2lnp (orc_sd *sd, const double *x, double *v, double *chi2, double *lnK)
{
  const int d        = sd->d;
  const double href2 = sd->href * sd->href;
  double gamma, lambda;
  int i;

  memcpy (v, x, sizeof (double) * d);
  trsv_upper_trans (d, sd->cov_decomp, d, v);

  for (i = 0; i < sd->n_kernels; i++)
  {
    const double *row_i = &sd->invUsample[(size_t) i * d];
    double chi2_i       = 0.0;
    int k;

    for (k = 0; k < d; k++)
    {
      const double df = row_i[k] - v[k];

      chi2_i += df * df;
    }

    chi2_i = chi2_i / href2;

    chi2[i] = chi2_i;
  }

  orc_kernel_eval_sum1_gamma_lambda (&sd->kernel, chi2, sd->weights, sd->kernel_lnnorm, lnK, sd->n_kernels, &gamma, &lambda);

  return -2.0 * (gamma + log1p (lambda) - d * log (sd->href));
}

/* kde.c:596-637 */
static double
kde_eval (orc_sd *sd, const double *x, double *v, double *chi2)
{
  const int d        = sd->d;
  const double href2 = sd->href * sd->href;
  int i;

  memcpy (v, x, sizeof (double) * d);
  trsv_upper_trans (d, sd->cov_decomp, d, v);

  for (i = 0; i < sd->n_kernels; i++)
  {
    const double *row_i = &sd->invUsample[(size_t) i * d];
    double chi2_i       = 0.0;
    int k;

    for (k = 0; k < d; k++)
    {
      const double df = row_i[k] - v[k];

      chi2_i += df * df;
    }

    chi2[i] = chi2_i / href2;
  }

  orc_kernel_eval_unnorm_vec (&sd->kernel, chi2, 1, chi2, 1, sd->n_kernels);

  return scipy_cblas_ddot (sd->n_kernels, chi2, 1, sd->weights, 1) * exp (-(sd->kernel_lnnorm + d * log (sd->href)));
}

/* vkde.c:681-723 */
static double
vkde_eval_m2lnp (orc_sd *sd, const double *x, double *delta_x, double *chi2, double *lnK)
{
  const int d            = sd->d;
  const double href2     = sd->href * sd->href;
  const double one_href2 = 1.0 / href2;
  double gamma, lambda;
  int i;

  for (i = 0; i < sd->n_kernels; i++)
  {
    const double *cov_decomp_i = &sd->cov_array[(size_t) i * d * d];
    const double *theta_i      = sd->sample[i];

    {
      double dot = 0.0;
      int k;

      for (k = 0; k < d; k++)
        delta_x[k] = x[k] - theta_i[k];

      trsv_upper_trans (d, cov_decomp_i, d, delta_x);

      for (k = 0; k < d; k++)
        dot += delta_x[k] * delta_x[k];

      chi2[i] = dot * one_href2;
    }
  }

  orc_kernel_eval_sum0_gamma_lambda (&sd->kernel, chi2, sd->weights, sd->lnnorms, lnK, sd->n_kernels, &gamma, &lambda);

  return -2.0 * (gamma + log1p (lambda) - d * log (sd->href));
}

/* vkde.c:631-679 */
static double
vkde_eval (orc_sd *sd, const double *x, double *delta_x, double *chi2)
{
  const int d            = sd->d;
  const double href2     = sd->href * sd->href;
  const double one_href2 = 1.0 / href2;
  double s               = 0.0;
  int i;

  for (i = 0; i < sd->n_kernels; i++)
  {
    const double *cov_decomp_i = &sd->cov_array[(size_t) i * d * d];
    const double *theta_i      = sd->sample[i];

    {
      double dot = 0.0;
      int k;

      for (k = 0; k < d; k++)
        delta_x[k] = x[k] - theta_i[k];

      trsv_upper_trans (d, cov_decomp_i, d, delta_x);

      for (k = 0; k < d; k++)
        dot += delta_x[k] * delta_x[k];

      chi2[i] = dot * one_href2;
    }
  }

  orc_kernel_eval_unnorm_vec (&sd->kernel, chi2, 1, chi2, 1, sd->n_kernels);

  for (i = 0; i < sd->n_kernels; i++)
  {
    const double Ku_i = chi2[i];
    const double u_i  = exp (sd->lnnorms[i]);
    const double w_i  = sd->weights[i];

    s += w_i * (Ku_i / u_i);
  }

  return s / pow (sd->href, d);
}

double
orc_sd_eval_m2lnp (orc_sd *sd, const double *x)
{
  double *v    = (double *) malloc (sizeof (double) * sd->d);
  double *chi2 = (double *) malloc (sizeof (double) * sd->n_kernels);
  double *lnK  = (double *) malloc (sizeof (double) * sd->n_kernels);
  double res;

  if (sd->type == ORC_SD_KDE)
    res = kde_eval_m2lnp (sd, x, v, chi2, lnK);
  else
    res = vkde_eval_m2lnp (sd, x, v, chi2, lnK);

  free (v);
  free (chi2);
  free (lnK);

  return res;
}

double
orc_sd_eval (orc_sd *sd, const double *x)
{
  double *v    = (double *) malloc (sizeof (double) * sd->d);
  double *chi2 = (double *) malloc (sizeof (double) * sd->n_kernels);
  double res;

  if (sd->type == ORC_SD_KDE)
    res = kde_eval (sd, x, v, chi2);
  else
    res = vkde_eval (sd, x, v, chi2);

  free (v);
  free (chi2);

  return res;
}

/* One point per OpenMP thread, schedule(dynamic,1): the only parallelism eval_m2lnp
 * gets in the reference (ncm_fit_esmcmc.c:2158); per-thread scratch mirrors the
 * NcmMemoryPool of eval vars (kde.c:128-147, vkde.c:121-141). */
void
orc_sd_eval_m2lnp_batch (orc_sd *sd, const double *X, int ldx, int q, double *out, int nthreads)
{
  const int blas_prev = blas_enter_omp ();
  int i;

  #pragma omp parallel num_threads (nthreads > 0 ? nthreads : 1)
  {
    double *v    = (double *) malloc (sizeof (double) * sd->d);
    double *chi2 = (double *) malloc (sizeof (double) * sd->n_kernels);
    double *lnK  = (double *) malloc (sizeof (double) * sd->n_kernels);

    #pragma omp for schedule(dynamic, 1)

    for (i = 0; i < q; i++)
    {
      if (sd->type == ORC_SD_KDE)
        out[i] = kde_eval_m2lnp (sd, &X[(size_t) i * ldx], v, chi2, lnK);
      else
        out[i] = vkde_eval_m2lnp (sd, &X[(size_t) i * ldx], v, chi2, lnK);
    }

    free (v);
    free (chi2);
    free (lnK);
  }

  blas_leave_omp (blas_prev);
}

void
orc_sd_eval_batch (orc_sd *sd, const double *X, int ldx, int q, double *out, int nthreads)
{
  const int blas_prev = blas_enter_omp ();
  int i;

  #pragma omp parallel num_threads (nthreads > 0 ? nthreads : 1)
  {
    double *v    = (double *) malloc (sizeof (double) * sd->d);
    double *chi2 = (double *) malloc (sizeof (double) * sd->n_kernels);

    #pragma omp for schedule(dynamic, 1)

    for (i = 0; i < q; i++)
    {
      if (sd->type == ORC_SD_KDE)
        out[i] = kde_eval (sd, &X[(size_t) i * ldx], v, chi2);
      else
        out[i] = vkde_eval (sd, &X[(size_t) i * ldx], v, chi2);
    }

    free (v);
    free (chi2);
  }

  blas_leave_omp (blas_prev);
}

/* ncm_stats_dist.c:1565-1606 */
int
orc_sd_kernel_choose (orc_sd *sd, orc_rng *rng)
{
  int i;

  if (!sd->wcum_ready)
  {
    double cum = 0.0;

    sd->wcum[0] = cum;

    for (i = 0; i < sd->n_kernels; i++)
    {
      cum           += sd->weights[i];
      sd->wcum[i + 1] = cum;
    }

    {
      const double s = 1.0 / cum;

      for (i = 0; i < sd->n_kernels + 1; i++)
        sd->wcum[i] *= s;
    }

    sd->wcum_ready = 1;
  }

  {
    const double p = orc_ran_flat (rng, 0.0, 1.0);
    int ilo        = 0;
    int ihi        = sd->n_kernels;

    while (ihi > ilo + 1)
    {
      int mi = (ihi + ilo) / 2;

      if (sd->wcum[mi] > p)
        ihi = mi;
      else
        ilo = mi;
    }

    i = ilo;
  }

  return i;
}

/* ncm_stats_dist.c:1618-1627 */
void
orc_sd_sample (orc_sd *sd, double *x, orc_rng *rng)
{
  const int i         = orc_sd_kernel_choose (sd, rng);
  const double *x_i   = sd->sample[i];
  const double *cov_U = orc_sd_peek_cov_decomp (sd, i);

  orc_kernel_sample (&sd->kernel, cov_U, sd->d, sd->href, x_i, x, rng);
}

/* test hook: overwrite self->weights (the reference exposes them through peek_weights, ncm_stats_dist.c:1775-1780) */
void
orc_sd_set_weights (orc_sd *sd, const double *w)
{
  memcpy (sd->weights, w, sizeof (double) * sd->n_kernels);
  sd->wcum_ready = 0;
}

int orc_sd_get_dim (const orc_sd *sd) { return sd->d; }
double orc_sd_get_over_smooth (const orc_sd *sd) { return sd->over_smooth; }

int
orc_sd_get_cv_trace (const orc_sd *sd, double *lnos, double *val, int cap)
{
  int i;

  for (i = 0; i < sd->cv_trace_len && i < cap; i++)
  {
    lnos[i] = sd->cv_trace[2 * i + 0];
    val[i]  = sd->cv_trace[2 * i + 1];
  }

  return sd->cv_trace_len;
}

int orc_sd_get_sample_size (const orc_sd *sd) { return sd->n_sample; }
int orc_sd_get_n_obs (const orc_sd *sd) { return sd->n_obs; }
int orc_sd_get_n_kernels (const orc_sd *sd) { return sd->n_kernels; }
double orc_sd_get_rnorm (const orc_sd *sd) { return sd->rnorm * sd->rnorm; }
const double *orc_sd_peek_weights (const orc_sd *sd) { return sd->weights; }
const double *orc_sd_peek_full_cov (const orc_sd *sd) { return sd->cov; }
const double *orc_sd_peek_full_cov_decomp (const orc_sd *sd) { return sd->cov_decomp; }
const double *orc_sd_peek_sample (const orc_sd *sd, int i) { return sd->sample[i]; }
const double *orc_sd_peek_IM (const orc_sd *sd) { return sd->IM; }
const double *orc_sd_peek_lnnorms (const orc_sd *sd) { return sd->lnnorms; }
const double *orc_sd_peek_invUsample (const orc_sd *sd) { return sd->invUsample; }

/* kde.c:559-566 ; vkde.c:608-617 */
const double *
orc_sd_peek_cov_decomp (const orc_sd *sd, int i)
{
  if (sd->type == ORC_SD_KDE)
    return sd->cov_decomp;
  else
    return &sd->cov_array[(size_t) i * sd->d * sd->d];
}

/* kde.c:586-594 ; vkde.c:619-629 */
double
orc_sd_get_lnnorm (orc_sd *sd, int i)
{
  if (sd->type == ORC_SD_KDE)
    return sd->kernel_lnnorm + sd->d * log (sd->href);
  else
    return sd->lnnorms[i] + sd->d * log (sd->href);
}

void
orc_sd_get_nnls_stats (const orc_sd *sd, orc_nnls_stats *st)
{
  *st = sd->nnls_stats;
}

void
orc_sd_get_timers (const orc_sd *sd, double *t3)
{
  t3[0] = sd->timers[0];
  t3[1] = sd->timers[1];
  t3[2] = sd->timers[2];
}

void
orc_set_blas_threads (int n)
{
  scipy_openblas_set_num_threads (n);
}

int
orc_get_max_threads (void)
{
#ifdef _OPENMP

  return omp_get_max_threads ();

#else

  return 1;

#endif
}
