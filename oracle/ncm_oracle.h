/*
 * oracle/ncm_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, SciPy's OpenBLAS for the BLAS/LAPACK calls) of the
 * NumCosmo APES density-estimation hot path: NcmStatsDist{,KDE,VKDE},
 * NcmStatsDistKernel{Gauss,ST}, NcmNNLS (+NcmISet), the GSL RNG behind NcmRNG
 * and the APES walker / ESMCMC accept loop.  Every function cites the
 * reference file:line it follows (paths relative to /root/reference).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product
 * (numcosmo_b200/) never links, imports or executes it.
 *
 * PARITY STATUS.  The reference cannot be built here (no GLib/GSL/meson) and
 * holds no golden vectors for this path, so the oracle is
 * pinned against (i) the closed-form known answers of
 * tests/c/ncm/stats/test_ncm_stats_dist_kernel.c:179-428 and the identities of
 * tests/c/ncm/stats/test_ncm_stats_dist.c:922-1117, restated in
 * tests/test_oracle_*.py, (ii) an independent numpy/scipy restatement of the
 * formulas of SURVEY.md Appendix C, (iii) MT19937's published known answers.
 * The GSL gamma/chisq/beta streams are "parity unpinned" (see orc_rng.c).
 */
#ifndef NCM_ORACLE_H
#define NCM_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------- RNG (orc_rng.c) ---------------- */
typedef struct orc_rng
{
  unsigned long mt[624];
  int mti;
} orc_rng;

void orc_rng_set (orc_rng *r, unsigned long seed);
unsigned long orc_rng_get (orc_rng *r);
double orc_rng_uniform (orc_rng *r);
double orc_rng_uniform_pos (orc_rng *r);
double orc_ran_flat (orc_rng *r, const double a, const double b);
double orc_ran_gaussian (orc_rng *r, const double sigma);
double orc_ran_ugaussian (orc_rng *r);
double orc_ran_gaussian_ziggurat (orc_rng *r, const double sigma);
double orc_ran_gamma (orc_rng *r, const double a, const double b);
double orc_ran_chisq (orc_rng *r, const double nu);
double orc_ran_beta (orc_rng *r, const double a, const double b);

/* ---------------- kernels (orc_kernel.c) ---------------- */
enum { ORC_KERNEL_GAUSS = 0, ORC_KERNEL_ST = 1 };

typedef struct orc_kernel
{
  int kind;
  int d;
  double nu;
} orc_kernel;

double orc_kernel_get_rot_bandwidth (const orc_kernel *k, const double n);
double orc_cholesky_lndet (const double *U, int n, int ld);
double orc_kernel_get_lnnorm (const orc_kernel *k, const double *cov_decomp, int ld);
double orc_kernel_eval_unnorm (const orc_kernel *k, const double chi2);
void orc_kernel_eval_unnorm_vec (const orc_kernel *k, const double *chi2, int chi2_stride, double *Ku, int Ku_stride, int n);
void orc_kernel_eval_sum0_gamma_lambda (const orc_kernel *k, const double *chi2, const double *weights, const double *lnnorms, double *lnK, int n, double *gamma, double *lambda);
void orc_kernel_eval_sum1_gamma_lambda (const orc_kernel *k, const double *chi2, const double *weights, double lnnorm, double *lnK, int n, double *gamma, double *lambda);
void orc_kernel_sample (const orc_kernel *k, const double *cov_decomp, int ld, const double href, const double *mu, double *x, orc_rng *rng);

/* ---------------- NNLS (orc_nnls.c) ---------------- */
typedef struct orc_nnls_stats
{
  int n_chol;       /* dposv calls */
  int n_lu;         /* dsysv fallbacks */
  int n_qr;         /* dgels fallbacks */
  int n_outer;      /* accepted outer iterations */
  int n_passive;    /* final |P| */
} orc_nnls_stats;

double orc_nnls_solve (const double *A, int nrows, int ncols, int lda, double *x, const double *f, double reltol, orc_nnls_stats *stats);
/* GSL subset selection used by NcmISet (restated, see orc_nnls.c) */
void orc_sort_smallest_index (int *p, int k, const double *src, int stride, int n);
void orc_sort_largest_index (int *p, int k, const double *src, int stride, int n);

/* ---------------- optimisers behind the cross-validation modes (orc_optim.c) ---------------- */
typedef double (*orc_fmin_fn) (const double *x, int n, void *params);
typedef struct orc_nmsimplex2 orc_nmsimplex2;   /* gsl_multimin_fminimizer_nmsimplex2 restated (parity unpinned: GSL absent) */
orc_nmsimplex2 *orc_nmsimplex2_new (int n);
void orc_nmsimplex2_free (orc_nmsimplex2 *s);
int orc_nmsimplex2_set (orc_nmsimplex2 *s, orc_fmin_fn f, void *params, const double *x, const double *step_size);
int orc_nmsimplex2_iterate (orc_nmsimplex2 *s);
const double *orc_nmsimplex2_x (const orc_nmsimplex2 *s);
double orc_nmsimplex2_fval (const orc_nmsimplex2 *s);
double orc_nmsimplex2_size (const orc_nmsimplex2 *s);
/* the iterate / gsl_multimin_test_size loop of ncm_stats_dist.c:681-692; returns the iteration count */
int orc_nmsimplex2_minimize (orc_nmsimplex2 *s, orc_fmin_fn f, void *params, const double *x0, const double *step, double size_tol, int max_iter);
/* levmar's dlevmar_dif (numcosmo/external/levmar/lm_core.c:436-851) restated; pinned against oracle/_ref/liblevmar_ref.so */
typedef void (*orc_lm_fn) (double *p, double *hx, int m, int n, void *adata);
/* register the reference's own dlevmar_dif (from oracle/_ref/liblevmar_ref.so) for the CV_SPLIT fit; NULL = the restatement */
void orc_set_levmar_dif (void *dlevmar_dif_fn);
int orc_lm_dif (orc_lm_fn func, double *p, const double *x, int m, int n, int itmax, const double *opts, double *info, void *adata);

/* ---------------- NcmStatsDist (orc_stats_dist.c) ---------------- */
enum { ORC_SD_KDE = 0, ORC_SD_VKDE = 1 };
enum { ORC_CV_NONE = 0, ORC_CV_SPLIT, ORC_CV_SPLIT_NOFIT, ORC_CV_LOO };
enum { ORC_COV_SAMPLE = 0, ORC_COV_FIXED, ORC_COV_ROBUST_DIAG, ORC_COV_ROBUST };

typedef struct orc_sd orc_sd;

/* the exact kNN ordering used by the VKDE prepare_kernel restatement (pinned against the reference's kd-tree, oracle/_ref/libkdtree_ref.so) */
void orc_knn_brute (const double *points, int n, int d, int query, int k, long *idx_out, double *dist_out);
/* robust covariance estimators of NcmStatsVec (ncm_stats_vec.c:1821-2072); rows[n] point to d-vectors; 0 or <0 (too few points) */
double orc_stats_Qn_from_sorted_data (const double *sorted, int n);   /* gsl_stats_Qn_from_sorted_data restated: parity unpinned (GSL absent) */
int orc_cov_robust_diag (const double *const *rows, int n, int d, double *cov);
int orc_cov_robust_ogk (const double *const *rows, int n, int d, double *cov);

orc_sd *orc_sd_new (int type, int kernel_kind, double nu, int d, int cv_type);
void orc_sd_free (orc_sd *sd);
void orc_sd_set_over_smooth (orc_sd *sd, double os);
void orc_sd_set_shrink (orc_sd *sd, double shrink);
void orc_sd_set_split_frac (orc_sd *sd, double split_frac);
void orc_sd_set_use_threads (orc_sd *sd, int use_threads);
void orc_sd_set_cov_type (orc_sd *sd, int cov_type);
void orc_sd_set_cov_fixed (orc_sd *sd, const double *cov, int ld);
void orc_sd_set_nearPD_maxiter (orc_sd *sd, int maxiter);
void orc_sd_set_local_frac (orc_sd *sd, double local_frac);
void orc_sd_set_use_rot_href (orc_sd *sd, int use_rot_href);

void orc_sd_reset (orc_sd *sd);
void orc_sd_add_obs (orc_sd *sd, const double *x);
int orc_sd_prepare (orc_sd *sd);                                  /* returns 0, or <0 on the reference's g_error paths */
int orc_sd_prepare_interp (orc_sd *sd, const double *m2lnp, int n);
double orc_sd_eval (orc_sd *sd, const double *x);
double orc_sd_eval_m2lnp (orc_sd *sd, const double *x);
/* q points, one per OpenMP thread, schedule(dynamic,1): ncm_fit_esmcmc.c:2158 */
void orc_sd_eval_m2lnp_batch (orc_sd *sd, const double *X, int ldx, int q, double *out, int nthreads);
void orc_sd_eval_batch (orc_sd *sd, const double *X, int ldx, int q, double *out, int nthreads);
int orc_sd_kernel_choose (orc_sd *sd, orc_rng *rng);
void orc_sd_sample (orc_sd *sd, double *x, orc_rng *rng);

void orc_sd_set_weights (orc_sd *sd, const double *w);
int orc_sd_get_dim (const orc_sd *sd);
int orc_sd_get_sample_size (const orc_sd *sd);
int orc_sd_get_n_obs (const orc_sd *sd);
int orc_sd_get_n_kernels (const orc_sd *sd);
double orc_sd_get_over_smooth (const orc_sd *sd);       /* the cross-validation modes leave their fitted value here */
/* optimiser trace of the last prepare / prepare_interp with a CV mode: number of objective evaluations and, per
 * evaluation, (ln over_smooth, objective or rnorm); at most cap entries are copied */
int orc_sd_get_cv_trace (const orc_sd *sd, double *lnos, double *val, int cap);
double orc_sd_get_href (orc_sd *sd);
double orc_sd_get_rnorm (const orc_sd *sd);             /* returns rnorm^2 as ncm_stats_dist.c:1664-1669 */
double orc_sd_get_lnnorm (orc_sd *sd, int i);
const double *orc_sd_peek_weights (const orc_sd *sd);
const double *orc_sd_peek_cov_decomp (const orc_sd *sd, int i);  /* d x d, ld = d */
const double *orc_sd_peek_full_cov (const orc_sd *sd);
const double *orc_sd_peek_full_cov_decomp (const orc_sd *sd);
const double *orc_sd_peek_sample (const orc_sd *sd, int i);
const double *orc_sd_peek_IM (const orc_sd *sd);                 /* n_obs x n_kernels after prepare_interp */
const double *orc_sd_peek_lnnorms (const orc_sd *sd);            /* VKDE: per-kernel lnnorm without d ln href */
const double *orc_sd_peek_invUsample (const orc_sd *sd);
void orc_sd_get_nnls_stats (const orc_sd *sd, orc_nnls_stats *st);
/* compute the interpolation matrix only (klass->compute_IM, no 1/f scaling) into IM[n_obs x n_kernels] */
void orc_sd_compute_IM (orc_sd *sd, double *IM);
/* stage timers of the last prepare_interp, seconds: [0] prepare_kernel [1] IM [2] NNLS */
void orc_sd_get_timers (const orc_sd *sd, double *t3);

/* ---------------- APES walker + ESMCMC accept loop (orc_apes.c) ---------------- */
enum { ORC_TARGET_MVND = 0, ORC_TARGET_ROSENBROCK = 1, ORC_TARGET_FUNNEL = 2 };

typedef struct orc_target
{
  int kind;
  int d;
  const double *mu;       /* MVND mean [d] */
  const double *cov_inv_U;/* MVND: upper Cholesky factor U of the covariance, d x d (chi2 = |U^-T (x-mu)|^2) */
  const double *lb;       /* bounds [d] */
  const double *ub;
} orc_target;

double orc_target_m2lnL (const orc_target *t, const double *x);

typedef struct orc_apes orc_apes;

orc_apes *orc_apes_new (int nwalkers, int d, int method_vkde_obj, int kernel_kind, double nu, double over_smooth, int use_interp, double shrink, double random_walk_prob, double local_frac, int use_threads);
void orc_apes_free (orc_apes *a);
void orc_apes_set_cov_type (orc_apes *a, int cov_type, const double *cov_fixed, int ld);
void orc_apes_set_exploration (orc_apes *a, unsigned int exploration);
/* theta [nwalkers x d] and m2lnL [nwalkers] are updated in place; accepted[nwalkers*iters] receives the accept flags
 * in walker order per iteration.  Follows ncm_fit_esmcmc.c:2235-2288 + walker_apes.c:742-919. */
void orc_apes_run (orc_apes *a, const orc_target *t, double *theta, double *m2lnL, int iters, orc_rng *rng, unsigned char *accepted, int nthreads);
/* stage timers accumulated over run: [0] prepare_kernel [1] IM [2] NNLS [3] sample [4] eval [5] likelihood+accept */
void orc_apes_get_timers (const orc_apes *a, double *t6);
void orc_apes_get_fallback_counts (const orc_apes *a, long *n3);
const double *orc_apes_peek_thetastar (const orc_apes *a);
const double *orc_apes_peek_m2lnp_star (const orc_apes *a);
const double *orc_apes_peek_m2lnp_cur (const orc_apes *a);

/* misc helpers */
double orc_log_gaussian_integral (double xl, double xu, double mu, double sigma, double *sign);
void orc_fill_rand_cov (double *cm, int n, double sigma_min, double sigma_max, double cor_level, orc_rng *rng);
int orc_cholesky_decomp_U (double *a, int n, int ld);  /* ncm_matrix_cholesky_decomp (cm, 'U') */
void orc_set_blas_threads (int n);
int orc_get_max_threads (void);

#ifdef __cplusplus
}
#endif

#endif
