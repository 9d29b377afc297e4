/*
 * include/ncm_stats_dist_b200.h -- host-side mirror of the reference's C interface for the
 * APES density-estimation path, served by the B200 kernels of libncm_sd_gpu.so
 * (libncm_stats_dist_b200.so).
 *
 * Function names, argument meaning and error behaviour follow
 *   numcosmo/ncm/stats/ncm_stats_dist.h:87-141
 *   numcosmo/ncm/stats/ncm_stats_dist_kde.h:74-86
 *   numcosmo/ncm/stats/ncm_stats_dist_vkde.h:53-62
 *   numcosmo/ncm/stats/ncm_stats_dist_kernel.h:76-90, _kernel_gauss.h:41-45, _kernel_st.h:41-48
 *   numcosmo/ncm/fit/ncm_fit_esmcmc_walker_apes.h:77-110
 * GLib/GObject and GSL are absent from this image, so the few GLib typedefs and the
 * NcmVector / NcmMatrix / NcmRNG / GPtrArray surface those signatures need are declared here
 * with the same names (row-major doubles, stride/tda as in ncm_vector.h / ncm_matrix.h).  In a
 * NumCosmo build tree this header is NOT used: the vtable bodies of the reference classes call
 * include/ncm_sd_gpu.h directly.
 *
 * Errors: the reference reports precondition failures with g_error (abort).  Here ncm_b200_error
 * prints the same message to stderr and aborts, unless a handler was installed with
 * ncm_b200_set_error_handler (the call then returns after the handler ran).
 */
#ifndef NCM_STATS_DIST_B200_H
#define NCM_STATS_DIST_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef double gdouble;
typedef unsigned int guint;
typedef int gint;
typedef int gboolean;
typedef char gchar;
typedef unsigned long gulong;
#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif

typedef void (*NcmB200ErrorHandler) (const char *msg, void *user_data);
void ncm_b200_set_error_handler (NcmB200ErrorHandler handler, void *user_data);
/* device used by objects created afterwards (default: $NCM_SD_GPU_DEVICE or 0) */
void ncm_b200_set_device (gint device);
/* OpenMP threads of the host-side prepare_kernel (use cores / ranks when several ranks share a host) */
void ncm_b200_set_num_threads (gint n);
/* parity switch: TRUE runs the VKDE prepare_kernel loop (kNN + local covariance + Cholesky) on the host instead of
 * the device (default: device; also set by $NCM_B200_HOST_PREPARE_KERNEL).  Both produce bit-identical factors. */
void ncm_b200_set_host_prepare_kernel (gboolean on);

/* ---- NcmVector / NcmMatrix (ncm_vector.h, ncm_matrix.h) ---- */
typedef struct _NcmVector NcmVector;
typedef struct _NcmMatrix NcmMatrix;

NcmVector *ncm_vector_new (const guint n);
NcmVector *ncm_vector_new_data_static (gdouble *d, const guint size, const guint stride);
NcmVector *ncm_vector_ref (NcmVector *cv);
NcmVector *ncm_vector_dup (const NcmVector *cv);
void ncm_vector_free (NcmVector *cv);
void ncm_vector_clear (NcmVector **cv);
guint ncm_vector_len (const NcmVector *cv);
guint ncm_vector_stride (const NcmVector *cv);
gdouble *ncm_vector_data (NcmVector *cv);
gdouble ncm_vector_get (const NcmVector *cv, const guint i);
void ncm_vector_set (NcmVector *cv, const guint i, const gdouble val);
void ncm_vector_set_all (NcmVector *cv, const gdouble val);
void ncm_vector_memcpy (NcmVector *cv1, const NcmVector *cv2);

NcmMatrix *ncm_matrix_new (const guint nrows, const guint ncols);
NcmMatrix *ncm_matrix_ref (NcmMatrix *cm);
NcmMatrix *ncm_matrix_dup (const NcmMatrix *cm);
void ncm_matrix_free (NcmMatrix *cm);
void ncm_matrix_clear (NcmMatrix **cm);
guint ncm_matrix_nrows (const NcmMatrix *cm);
guint ncm_matrix_ncols (const NcmMatrix *cm);
guint ncm_matrix_tda (const NcmMatrix *cm);
gdouble *ncm_matrix_data (NcmMatrix *cm);
gdouble ncm_matrix_get (const NcmMatrix *cm, const guint i, const guint j);
void ncm_matrix_set (NcmMatrix *cm, const guint i, const guint j, const gdouble val);

/* ---- GPtrArray of NcmVector (only what peek_sample_array needs) ---- */
typedef struct _GPtrArray
{
  void **pdata;
  guint len;
} GPtrArray;
#define g_ptr_array_index(array, index_) ((array)->pdata)[index_]

/* ---- NcmRNG (ncm_rng.h): gsl_rng_mt19937 and the GSL distributions the path draws from ---- */
typedef struct _NcmRNG NcmRNG;

NcmRNG *ncm_rng_new (const gchar *algo);
NcmRNG *ncm_rng_seeded_new (const gchar *algo, gulong seed);
void ncm_rng_free (NcmRNG *rng);
void ncm_rng_clear (NcmRNG **rng);
void ncm_rng_set_seed (NcmRNG *rng, gulong seed);
gulong ncm_rng_get_seed (NcmRNG *rng);
gulong ncm_rng_gen_ulong (NcmRNG *rng);
gdouble ncm_rng_uniform01_gen (NcmRNG *rng);
gdouble ncm_rng_uniform01_pos_gen (NcmRNG *rng);
gdouble ncm_rng_uniform_gen (NcmRNG *rng, const gdouble xl, const gdouble xu);
gdouble ncm_rng_gaussian_gen (NcmRNG *rng, const gdouble mu, const gdouble sigma);
gdouble ncm_rng_ugaussian_gen (NcmRNG *rng);
gdouble ncm_rng_chisq_gen (NcmRNG *rng, const gdouble nu);
gdouble ncm_rng_beta_gen (NcmRNG *rng, const gdouble a, const gdouble b);

/* ---- NcmStatsDistKernel ---- */
typedef struct _NcmStatsDistKernel NcmStatsDistKernel;
typedef struct _NcmStatsDistKernel NcmStatsDistKernelGauss;
typedef struct _NcmStatsDistKernel NcmStatsDistKernelST;
#define NCM_STATS_DIST_KERNEL(obj) ((NcmStatsDistKernel *) (obj))

NcmStatsDistKernelGauss *ncm_stats_dist_kernel_gauss_new (const guint dim);
NcmStatsDistKernelGauss *ncm_stats_dist_kernel_gauss_ref (NcmStatsDistKernelGauss *sdkg);
void ncm_stats_dist_kernel_gauss_free (NcmStatsDistKernelGauss *sdkg);
void ncm_stats_dist_kernel_gauss_clear (NcmStatsDistKernelGauss **sdkg);

NcmStatsDistKernelST *ncm_stats_dist_kernel_st_new (const guint dim, const gdouble nu);
NcmStatsDistKernelST *ncm_stats_dist_kernel_st_ref (NcmStatsDistKernelST *sdkst);
void ncm_stats_dist_kernel_st_free (NcmStatsDistKernelST *sdkst);
void ncm_stats_dist_kernel_st_clear (NcmStatsDistKernelST **sdkst);
void ncm_stats_dist_kernel_st_set_nu (NcmStatsDistKernelST *sdkst, const gdouble nu);
gdouble ncm_stats_dist_kernel_st_get_nu (NcmStatsDistKernelST *sdkst);

NcmStatsDistKernel *ncm_stats_dist_kernel_ref (NcmStatsDistKernel *sdk);
void ncm_stats_dist_kernel_free (NcmStatsDistKernel *sdk);
void ncm_stats_dist_kernel_clear (NcmStatsDistKernel **sdk);
guint ncm_stats_dist_kernel_get_dim (NcmStatsDistKernel *sdk);
gdouble ncm_stats_dist_kernel_get_rot_bandwidth (NcmStatsDistKernel *sdk, const gdouble n);
gdouble ncm_stats_dist_kernel_get_lnnorm (NcmStatsDistKernel *sdk, NcmMatrix *cov_decomp);
gdouble ncm_stats_dist_kernel_eval_unnorm (NcmStatsDistKernel *sdk, const gdouble chi2);
void ncm_stats_dist_kernel_eval_unnorm_vec (NcmStatsDistKernel *sdk, NcmVector *chi2, NcmVector *Ku);
void ncm_stats_dist_kernel_eval_sum0_gamma_lambda (NcmStatsDistKernel *sdk, NcmVector *chi2, NcmVector *weights, NcmVector *lnnorms, NcmVector *lnK, gdouble *gamma, gdouble *lambda);
void ncm_stats_dist_kernel_eval_sum1_gamma_lambda (NcmStatsDistKernel *sdk, NcmVector *chi2, NcmVector *weights, gdouble lnnorm, NcmVector *lnK, gdouble *gamma, gdouble *lambda);
void ncm_stats_dist_kernel_sample (NcmStatsDistKernel *sdk, NcmMatrix *cov_decomp, const gdouble href, NcmVector *mu, NcmVector *y, NcmRNG *rng);

/* ---- NcmStatsDist / KDE / VKDE ---- */
typedef struct _NcmStatsDist NcmStatsDist;
typedef struct _NcmStatsDist NcmStatsDistKDE;
typedef struct _NcmStatsDist NcmStatsDistVKDE;
#define NCM_STATS_DIST(obj) ((NcmStatsDist *) (obj))
#define NCM_STATS_DIST_KDE(obj) ((NcmStatsDistKDE *) (obj))
#define NCM_STATS_DIST_VKDE(obj) ((NcmStatsDistVKDE *) (obj))

typedef enum _NcmStatsDistCV
{
  NCM_STATS_DIST_CV_NONE,
  NCM_STATS_DIST_CV_SPLIT,
  NCM_STATS_DIST_CV_SPLIT_NOFIT,
  NCM_STATS_DIST_CV_LOO,
  NCM_STATS_DIST_CV_LEN,
} NcmStatsDistCV;

typedef enum _NcmStatsDistKDECovType
{
  NCM_STATS_DIST_KDE_COV_TYPE_SAMPLE,
  NCM_STATS_DIST_KDE_COV_TYPE_FIXED,
  NCM_STATS_DIST_KDE_COV_TYPE_ROBUST_DIAG,
  NCM_STATS_DIST_KDE_COV_TYPE_ROBUST,
  NCM_STATS_DIST_KDE_COV_TYPE_LEN,
} NcmStatsDistKDECovType;

NcmStatsDist *ncm_stats_dist_ref (NcmStatsDist *sd);
void ncm_stats_dist_free (NcmStatsDist *sd);
void ncm_stats_dist_clear (NcmStatsDist **sd);

void ncm_stats_dist_set_kernel (NcmStatsDist *sd, NcmStatsDistKernel *sdk);
NcmStatsDistKernel *ncm_stats_dist_peek_kernel (NcmStatsDist *sd);
NcmStatsDistKernel *ncm_stats_dist_get_kernel (NcmStatsDist *sd);

guint ncm_stats_dist_get_dim (NcmStatsDist *sd);
guint ncm_stats_dist_get_sample_size (NcmStatsDist *sd);
guint ncm_stats_dist_get_n_kernels (NcmStatsDist *sd);
gdouble ncm_stats_dist_get_href (NcmStatsDist *sd);

void ncm_stats_dist_set_over_smooth (NcmStatsDist *sd, const gdouble over_smooth);
gdouble ncm_stats_dist_get_over_smooth (NcmStatsDist *sd);
void ncm_stats_dist_set_split_frac (NcmStatsDist *sd, const gdouble split_frac);
gdouble ncm_stats_dist_get_split_frac (NcmStatsDist *sd);
void ncm_stats_dist_set_shrink (NcmStatsDist *sd, const gdouble shrink);
gdouble ncm_stats_dist_get_shrink (NcmStatsDist *sd);
void ncm_stats_dist_set_print_fit (NcmStatsDist *sd, const gboolean print_fit);
gboolean ncm_stats_dist_get_print_fit (NcmStatsDist *sd);
void ncm_stats_dist_set_cv_type (NcmStatsDist *sd, const NcmStatsDistCV cv_type);
NcmStatsDistCV ncm_stats_dist_get_cv_type (NcmStatsDist *sd);
void ncm_stats_dist_set_use_threads (NcmStatsDist *sd, const gboolean use_threads);
gboolean ncm_stats_dist_get_use_threads (NcmStatsDist *sd);

void ncm_stats_dist_prepare_kernel (NcmStatsDist *sd, GPtrArray *sample_array);
void ncm_stats_dist_prepare (NcmStatsDist *sd);
void ncm_stats_dist_prepare_interp (NcmStatsDist *sd, NcmVector *m2lnp);

gdouble ncm_stats_dist_eval (NcmStatsDist *sd, NcmVector *x);
gdouble ncm_stats_dist_eval_m2lnp (NcmStatsDist *sd, NcmVector *x);
/* NEW (one padding[] slot of NcmStatsDistClass, ncm_stats_dist.h:63-64): all rows of X at once */
void ncm_stats_dist_eval_m2lnp_array (NcmStatsDist *sd, NcmMatrix *X, NcmVector *m2lnp_out);
void ncm_stats_dist_eval_array (NcmStatsDist *sd, NcmMatrix *X, NcmVector *p_out);

guint ncm_stats_dist_kernel_choose (NcmStatsDist *sd, NcmRNG *rng);
void ncm_stats_dist_sample (NcmStatsDist *sd, NcmVector *x, NcmRNG *rng);

gdouble ncm_stats_dist_get_rnorm (NcmStatsDist *sd);

void ncm_stats_dist_add_obs (NcmStatsDist *sd, NcmVector *y);

GPtrArray *ncm_stats_dist_peek_sample_array (NcmStatsDist *sd);
NcmMatrix *ncm_stats_dist_peek_cov_decomp (NcmStatsDist *sd, guint i);
NcmMatrix *ncm_stats_dist_peek_full_cov_decomp (NcmStatsDist *sd);
NcmMatrix *ncm_stats_dist_peek_full_cov (NcmStatsDist *sd);
gdouble ncm_stats_dist_get_lnnorm (NcmStatsDist *sd, guint i);
NcmVector *ncm_stats_dist_peek_weights (NcmStatsDist *sd);
void ncm_stats_dist_get_Ki (NcmStatsDist *sd, const guint i, NcmVector **y_i, NcmMatrix **cov_i, gdouble *n_i, gdouble *w_i);
void ncm_stats_dist_reset (NcmStatsDist *sd);

NcmStatsDistKDE *ncm_stats_dist_kde_new (NcmStatsDistKernel *sdk, NcmStatsDistCV CV_type);
NcmStatsDistKDE *ncm_stats_dist_kde_ref (NcmStatsDistKDE *sdkde);
void ncm_stats_dist_kde_free (NcmStatsDistKDE *sdkde);
void ncm_stats_dist_kde_clear (NcmStatsDistKDE **sdkde);
void ncm_stats_dist_kde_set_nearPD_maxiter (NcmStatsDistKDE *sdkde, const guint maxiter);
guint ncm_stats_dist_kde_get_nearPD_maxiter (NcmStatsDistKDE *sdkde);
void ncm_stats_dist_kde_set_cov_type (NcmStatsDistKDE *sdkde, NcmStatsDistKDECovType cov_type);
NcmStatsDistKDECovType ncm_stats_dist_kde_get_cov_type (NcmStatsDistKDE *sdkde);
void ncm_stats_dist_kde_set_cov_fixed (NcmStatsDistKDE *sdkde, NcmMatrix *cov_fixed);
NcmMatrix *ncm_stats_dist_kde_peek_cov_fixed (NcmStatsDistKDE *sdkde);

NcmStatsDistVKDE *ncm_stats_dist_vkde_new (NcmStatsDistKernel *sdk, NcmStatsDistCV CV_type);
NcmStatsDistVKDE *ncm_stats_dist_vkde_ref (NcmStatsDistVKDE *sdvkde);
void ncm_stats_dist_vkde_free (NcmStatsDistVKDE *sdvkde);
void ncm_stats_dist_vkde_clear (NcmStatsDistVKDE **sdvkde);
void ncm_stats_dist_vkde_set_local_frac (NcmStatsDistVKDE *sdvkde, const gdouble local_frac);
gdouble ncm_stats_dist_vkde_get_local_frac (NcmStatsDistVKDE *sdvkde);
void ncm_stats_dist_vkde_set_use_rot_href (NcmStatsDistVKDE *sdvkde, const gboolean use_rot_href);
gboolean ncm_stats_dist_vkde_get_use_rot_href (NcmStatsDistVKDE *sdvkde);

/* instrumentation of the GPU path behind an object (not in the reference) */
/* optimiser trace of the last prepare / prepare_interp under a cross-validation mode: the number of objective evaluations and,
 * for the first cap of them, (ln over_smooth, objective value or NNLS rnorm) in evaluation order */
gint ncm_stats_dist_b200_get_cv_trace (NcmStatsDist *sd, gdouble *lnos, gdouble *val, gint cap);
/* Multi-rank (SPMD) mode, one process per GPU: every rank builds the same object, feeds it the same observations and makes the same
 * calls; after comm_init the rows of the interpolation matrix (ncm_stats_dist.c:878-1094) and the query rows of the batched evaluation
 * are sharded over the ranks, the NNLS normal equations are all-reduced and the densities all-gathered with NCCL, and every rank
 * receives the complete, identical results.  id comes from ncm_stats_dist_b200_comm_unique_id on one rank, distributed by the caller. */
gint ncm_stats_dist_b200_comm_unique_id (gchar id_out[128]);
gboolean ncm_stats_dist_b200_comm_init (NcmStatsDist *sd, gint nranks, gint rank, const gchar id[128]);
void ncm_stats_dist_b200_get_nnls_stats (NcmStatsDist *sd, gint *n_chol, gint *n_lu, gint *n_outer, gint *n_passive);
/* of the last NNLS: systems that took the reference's fallbacks, dsysv (ncm_nnls.c:573-606) and dgels (:608-638) */
void ncm_stats_dist_b200_get_nnls_fallback_stats (NcmStatsDist *sd, gint *n_lu, gint *n_qr);
/* of the last NNLS: systems solved by low-rank modification of an earlier factor instead of a fresh dposv (csrc/lowrank.cu), how many of
 * those fell back to a fresh factorisation, triangular inverses formed, largest |D| + |A| */
void ncm_stats_dist_b200_get_nnls_lowrank_stats (NcmStatsDist *sd, gint *n_lowrank, gint *n_fallback, gint *n_trinv, gint *max_k);
void ncm_stats_dist_b200_get_timers (NcmStatsDist *sd, gdouble *ms7, long long *n_launches, gdouble *host_prepare_kernel_ms);
void ncm_stats_dist_b200_enable_timers (NcmStatsDist *sd, gboolean on);
void *ncm_stats_dist_b200_peek_ctx (NcmStatsDist *sd);

/* ---- APES walker (ncm_fit_esmcmc_walker_apes.h) + the ESMCMC accept loop around it ---- */
typedef struct _NcmFitESMCMCWalkerAPES NcmFitESMCMCWalkerAPES;

typedef enum _NcmFitESMCMCWalkerAPESMethod
{
  NCM_FIT_ESMCMC_WALKER_APES_METHOD_KDE = 0,
  NCM_FIT_ESMCMC_WALKER_APES_METHOD_VKDE,
  NCM_FIT_ESMCMC_WALKER_APES_METHOD_LEN,
} NcmFitESMCMCWalkerAPESMethod;

typedef enum _NcmFitESMCMCWalkerAPESKType
{
  NCM_FIT_ESMCMC_WALKER_APES_KTYPE_CAUCHY = 0,
  NCM_FIT_ESMCMC_WALKER_APES_KTYPE_ST3,
  NCM_FIT_ESMCMC_WALKER_APES_KTYPE_GAUSS,
  NCM_FIT_ESMCMC_WALKER_APES_KTYPE_LEN,
} NcmFitESMCMCWalkerAPESKType;

NcmFitESMCMCWalkerAPES *ncm_fit_esmcmc_walker_apes_new (guint nwalkers, guint nparams);
NcmFitESMCMCWalkerAPES *ncm_fit_esmcmc_walker_apes_new_full (guint nwalkers, guint nparams, NcmFitESMCMCWalkerAPESMethod method, NcmFitESMCMCWalkerAPESKType k_type, gdouble over_smooth, gboolean use_interp);
NcmFitESMCMCWalkerAPES *ncm_fit_esmcmc_walker_apes_ref (NcmFitESMCMCWalkerAPES *apes);
void ncm_fit_esmcmc_walker_apes_free (NcmFitESMCMCWalkerAPES *apes);
void ncm_fit_esmcmc_walker_apes_clear (NcmFitESMCMCWalkerAPES **apes);
void ncm_fit_esmcmc_walker_apes_set_method (NcmFitESMCMCWalkerAPES *apes, NcmFitESMCMCWalkerAPESMethod method);
void ncm_fit_esmcmc_walker_apes_set_k_type (NcmFitESMCMCWalkerAPES *apes, NcmFitESMCMCWalkerAPESKType k_type);
void ncm_fit_esmcmc_walker_apes_set_over_smooth (NcmFitESMCMCWalkerAPES *apes, const gdouble os);
void ncm_fit_esmcmc_walker_apes_set_shrink (NcmFitESMCMCWalkerAPES *apes, const gdouble shrink);
void ncm_fit_esmcmc_walker_apes_set_random_walk_prob (NcmFitESMCMCWalkerAPES *apes, const gdouble prob);
void ncm_fit_esmcmc_walker_apes_set_random_walk_scale (NcmFitESMCMCWalkerAPES *apes, const gdouble scale);
NcmFitESMCMCWalkerAPESMethod ncm_fit_esmcmc_walker_apes_get_method (NcmFitESMCMCWalkerAPES *apes);
NcmFitESMCMCWalkerAPESKType ncm_fit_esmcmc_walker_apes_get_k_type (NcmFitESMCMCWalkerAPES *apes);
gdouble ncm_fit_esmcmc_walker_apes_get_over_smooth (NcmFitESMCMCWalkerAPES *apes);
gdouble ncm_fit_esmcmc_walker_apes_get_shrink (NcmFitESMCMCWalkerAPES *apes);
gdouble ncm_fit_esmcmc_walker_apes_get_random_walk_prob (NcmFitESMCMCWalkerAPES *apes);
gdouble ncm_fit_esmcmc_walker_apes_get_random_walk_scale (NcmFitESMCMCWalkerAPES *apes);
void ncm_fit_esmcmc_walker_apes_use_interp (NcmFitESMCMCWalkerAPES *apes, gboolean use_interp);
gboolean ncm_fit_esmcmc_walker_apes_interp (NcmFitESMCMCWalkerAPES *apes);
void ncm_fit_esmcmc_walker_apes_set_use_threads (NcmFitESMCMCWalkerAPES *apes, gboolean use_threads);
gboolean ncm_fit_esmcmc_walker_apes_get_use_threads (NcmFitESMCMCWalkerAPES *apes);
void ncm_fit_esmcmc_walker_apes_peek_sds (NcmFitESMCMCWalkerAPES *apes, NcmStatsDist **sd0, NcmStatsDist **sd1);
void ncm_fit_esmcmc_walker_apes_set_local_frac (NcmFitESMCMCWalkerAPES *apes, gdouble local_frac);
void ncm_fit_esmcmc_walker_apes_set_exploration (NcmFitESMCMCWalkerAPES *apes, guint exploration);
/* ncm_fit_esmcmc_walker_apes.h:106-108.  The NcmMSet argument of the reference is replaced by what it reads from it: the scales of the
 * free parameters (ncm_mset_fparam_get_scale (mset, i), i < nparams). */
void ncm_fit_esmcmc_walker_apes_set_cov_fixed_from_mset (NcmFitESMCMCWalkerAPES *apes, const gdouble *fparam_scales);
void ncm_fit_esmcmc_walker_apes_set_cov_robust_diag (NcmFitESMCMCWalkerAPES *apes);
void ncm_fit_esmcmc_walker_apes_set_cov_robust (NcmFitESMCMCWalkerAPES *apes);
/* instrumentation of the proposal draws generated ahead of the weights (host/apes.cc): blocks pre-generated, blocks replayed serially */
void ncm_fit_esmcmc_walker_apes_b200_get_pregen_stats (NcmFitESMCMCWalkerAPES *a, long long *n_blocks, long long *n_fallbacks);

/* The walker vtable entries (ncm_fit_esmcmc_walker.h: setup / step / prob_norm) in array form:
 * theta [nwalkers x nparams] row-major, m2lnL [nwalkers], bounds lb/ub [nparams].
 * setup builds the density from the OTHER half, draws the proposals for walkers [ki, kf) in
 * reference RNG order and -- new -- evaluates both transition densities of every walker of the
 * block in one batched GPU call, so that step() only reads the cache. */
void ncm_fit_esmcmc_walker_apes_setup (NcmFitESMCMCWalkerAPES *apes, const gdouble *lb, const gdouble *ub, const gdouble *theta, const gdouble *m2lnL, guint ki, guint kf, NcmRNG *rng);
void ncm_fit_esmcmc_walker_apes_step (NcmFitESMCMCWalkerAPES *apes, const gdouble *theta, gdouble *thetastar, guint k);
gdouble ncm_fit_esmcmc_walker_apes_prob_norm (NcmFitESMCMCWalkerAPES *apes, guint k);
const gdouble *ncm_fit_esmcmc_walker_apes_peek_thetastar (NcmFitESMCMCWalkerAPES *apes);
const gdouble *ncm_fit_esmcmc_walker_apes_peek_m2lnp_star (NcmFitESMCMCWalkerAPES *apes);
const gdouble *ncm_fit_esmcmc_walker_apes_peek_m2lnp_cur (NcmFitESMCMCWalkerAPES *apes);

/* likelihood callback: fills m2lnL[0..n) for the n points X [n x nparams] */
typedef void (*NcmB200M2lnLFunc) (const gdouble *X, guint n, guint nparams, gdouble *m2lnL, void *user_data);
/* _ncm_fit_esmcmc_run for `iters` whole-ensemble iterations (ncm_fit_esmcmc.c:2235-2288, ki = 0):
 * jumps, setup + run of both blocks, accept/reject in place.  accepted may be NULL, else
 * [iters x nwalkers].  timers_ms[8]: prepare_kernel, IM, NNLS, sample, eval, likelihood+accept, upload, total. */
void ncm_b200_esmcmc_run (NcmFitESMCMCWalkerAPES *apes, NcmB200M2lnLFunc m2lnL_func, void *user_data, const gdouble *lb, const gdouble *ub,
                          gdouble *theta, gdouble *m2lnL, guint iters, NcmRNG *rng, unsigned char *accepted, gdouble *timers_ms);
/* built-in synthetic targets (formulas of ncm_data_rosenbrock.c:106-113, ncm_data_funnel.c:113-132, MVND) */
void ncm_b200_target_rosenbrock (const gdouble *X, guint n, guint nparams, gdouble *m2lnL, void *user_data);
void ncm_b200_target_funnel (const gdouble *X, guint n, guint nparams, gdouble *m2lnL, void *user_data);
typedef struct _NcmB200MVND { guint d; const gdouble *mu; const gdouble *U; } NcmB200MVND; /* cov = U^T U, U upper [d x d] */
void ncm_b200_target_mvnd (const gdouble *X, guint n, guint nparams, gdouble *m2lnL, void *user_data);

#ifdef __cplusplus
}
#endif

#endif /* NCM_STATS_DIST_B200_H */
