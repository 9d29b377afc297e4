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
 * Cross-validation modes other than CV_NONE and the ROBUST covariance types are
 * outside the APES path (last paragraph) and return -2 here.
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

  if ((sd->cov_type != ORC_COV_SAMPLE) && (sd->cov_type != ORC_COV_FIXED))
    return -2;

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

      svec_get_cov_matrix (&sv, cov);
      _cholesky_decomp (&sd->cov_array[(size_t) i * d * d], cov, d, sd->nearPD_maxiter);
      sd->lnnorms[i] = orc_kernel_get_lnnorm (&sd->kernel, &sd->cov_array[(size_t) i * d * d], d);
    }

    svec_free (&sv);
    free (items);
    free (cov);
  }

  blas_leave_omp (blas_prev);

  return 0;
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

/* ncm_stats_dist.c:703-789 (CV_NONE only) */
int
orc_sd_prepare (orc_sd *sd)
{
  double t0 = now_s ();
  int ret, i;

  if (sd->cv_type != ORC_CV_NONE)
    return -2;

  sd->n_obs     = sd->n_sample;
  sd->n_kernels = sd->n_sample;

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
  sd->timers[0]  = now_s () - t0;

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

/* ncm_stats_dist.c:878-1094 (CV_NONE branch) */
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
    size_t *sort = (size_t *) malloc (sizeof (size_t) * sd->n_kernels);
    int n_cut    = 0;

    for (i = 0; i < sd->n_kernels; i++)
      sort[i] = i;

    qsort_r (sort, sd->n_kernels, sizeof (size_t), idx_cmp_ctx, (void *) m2lnp);

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
        sd->weights[sort[i]] = 0.9 / n_cut;

      free (sort);

      return 0;
    }

    {
      double *m2lnp_cut = (double *) malloc (sizeof (double) * n_cut);
      double **cut      = (double **) malloc (sizeof (double *) * n_cut);
      int j             = 0;

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

  {
    double *f1 = (double *) malloc (sizeof (double) * sd->n_obs);
    double t0  = now_s (), t1;
    const int nk = sd->n_kernels;

    for (i = 0; i < sd->n_obs; i++)
      f1[i] = 1.0;

    /* _ncm_stats_dist_compute_IM_full: ncm_stats_dist.c:791-804 */
    orc_sd_compute_IM (sd, sd->IM);

    #pragma omp parallel for if (sd->use_threads)

    for (i = 0; i < sd->n_obs; i++)
    {
      const double s = 1.0 / sd->f[i];
      int j;

      for (j = 0; j < nk; j++)
        sd->IM[(size_t) i * nk + j] *= s;
    }

    t1            = now_s ();
    sd->timers[1] = t1 - t0;

    /* reltol default GSL_DBL_EPSILON, ncm_nnls.c:271-275 */
    sd->rnorm     = orc_nnls_solve (sd->IM, sd->n_obs, sd->n_kernels, sd->n_kernels, sd->weights, f1, DBL_EPSILON, &sd->nnls_stats);
    sd->timers[2] = now_s () - t1;

    free (f1);
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
