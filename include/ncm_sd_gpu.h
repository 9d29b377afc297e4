/*
 * include/ncm_sd_gpu.h -- C ABI of the B200 (sm_100a) density-estimation path that sits
 * under NumCosmo's NcmStatsDist vtable (libncm_sd_gpu.so).
 *
 * Plain C: opaque context, doubles, row-major matrices with explicit leading dimensions,
 * int status returns (0 = ok; the host glue turns non-zero into g_error, the reference's
 * only error channel).  No GLib and no torch types.
 *
 * Each entry point names the reference interface it replaces (paths relative to the
 * NumCosmo source tree, v0.27.0).  There is NO CPU fallback behind any of them: without a
 * CUDA device ncm_sd_gpu_ctx_new fails with NCM_SD_GPU_ENODEV.
 */
#ifndef NCM_SD_GPU_H
#define NCM_SD_GPU_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ncm_sd_gpu_ctx ncm_sd_gpu_ctx;

enum
{
  NCM_SD_GPU_OK      = 0,
  NCM_SD_GPU_EINVAL  = 1, /* bad argument / call order */
  NCM_SD_GPU_ENODEV  = 2, /* no usable CUDA device */
  NCM_SD_GPU_ECUDA   = 3, /* CUDA runtime error, see ncm_sd_gpu_last_error */
  NCM_SD_GPU_ENOTPD  = 4, /* NNLS: normal matrix not positive definite even after the diagonal retry */
  NCM_SD_GPU_ENCCL   = 5, /* NCCL error */
  NCM_SD_GPU_ENOMEM  = 6
};

enum { NCM_SD_GPU_KERNEL_GAUSS = 0, NCM_SD_GPU_KERNEL_ST = 1 };
enum { NCM_SD_GPU_KDE = 0, NCM_SD_GPU_VKDE = 1 };

#define NCM_SD_GPU_MAX_DIM 32

/* ---- context ---------------------------------------------------------------------- */

/* One context per (process, device): owns the stream, the device buffers and (optionally)
 * an NCCL communicator.  Plays the role of the per-object scratch the reference keeps in
 * NcmMemoryPool (ncm_stats_dist_kde.c:128-147, ncm_stats_dist_vkde.c:121-141). */
int ncm_sd_gpu_ctx_new (ncm_sd_gpu_ctx **ctx, int device);
int ncm_sd_gpu_ctx_free (ncm_sd_gpu_ctx *ctx);
const char *ncm_sd_gpu_last_error (const ncm_sd_gpu_ctx *ctx);
int ncm_sd_gpu_device_count (void);
/* cudaStream_t the context launches on (as void*), for callers that own device buffers. */
void *ncm_sd_gpu_stream (ncm_sd_gpu_ctx *ctx);
int ncm_sd_gpu_synchronize (ncm_sd_gpu_ctx *ctx);

/* ---- kernel + bandwidth -------------------------------------------------------------
 * ncm_stats_dist_kernel_gauss_new / ncm_stats_dist_kernel_st_new
 * (ncm_stats_dist_kernel_gauss.c:357-373, ncm_stats_dist_kernel_st.c:416-435). */
int ncm_sd_gpu_set_kernel (ncm_sd_gpu_ctx *ctx, int kind, double nu, int d);

/* ---- centres ------------------------------------------------------------------------
 * KDE: result of _ncm_stats_dist_kde_prepare_kernel (ncm_stats_dist_kde.c:378-490):
 *   invUsample [n_obs x d] = sample . U^-1 (rows 0..n_kernels-1 are the kernel centres),
 *   U [d x d] upper Cholesky factor of the covariance, lnnorm = kernel_lnnorm (no d ln h). */
int ncm_sd_gpu_upload_kde (ncm_sd_gpu_ctx *ctx, int n_obs, int n_kernels, const double *invUsample, int ld,
                           const double *U, int ldu, double lnnorm);
/* VKDE: result of _ncm_stats_dist_vkde_build_cov_array_kdtree (ncm_stats_dist_vkde.c:362-496):
 *   sample [n_obs x d] raw points, U_all [n_kernels x d x d] upper factors (dense, row-major),
 *   lnnorms [n_kernels] (no d ln h). */
int ncm_sd_gpu_upload_vkde (ncm_sd_gpu_ctx *ctx, int n_obs, int n_kernels, const double *sample, int ld,
                            const double *U_all, const double *lnnorms);
/* VKDE prepare_kernel ON THE DEVICE: the OpenMP loop of _ncm_stats_dist_vkde_build_cov_array_kdtree
 * (ncm_stats_dist_vkde.c:428-493).  sample [n_obs x d] raw, invUsample [n_obs x d] whitened (both host),
 * k = max (local_frac * n_obs, 2) neighbours.  Returns the upper factors U_all_out [n_kernels x d x d] and
 * fail_out [n_kernels] (1 = covariance not positive definite: the caller applies the reference's
 * nearPD / diagonal fallback, kde.c:344-367, and passes the repaired factors to vkde_finish).  The points
 * and factors stay on the device; vkde_finish uploads the per-kernel lnnorms (which the host computes from
 * the factors, _kernel_gauss.c:201-207 / _kernel_st.c:240-252) and packs the records.  Beyond n_obs ~ 25000 one row of
 * distances no longer fits in shared memory and the n_kernels x n_obs distance matrix must fit in device memory. */
int ncm_sd_gpu_vkde_prepare (ncm_sd_gpu_ctx *ctx, int n_obs, int n_kernels, const double *sample, int ld, const double *invUsample, int ldz,
                             int k, double *U_all_out, int *fail_out);
int ncm_sd_gpu_vkde_finish (ncm_sd_gpu_ctx *ctx, const double *lnnorms, int n_fixed, const int *fixed_idx, const double *fixed_U);
/* weights + bandwidth: self->weights / self->href of NcmStatsDistPrivate
 * (ncm_stats_dist_private.h:39-76), set by _ncm_stats_dist_prepare (ncm_stats_dist.c:753-767). */
int ncm_sd_gpu_set_weights (ncm_sd_gpu_ctx *ctx, int n_kernels, const double *weights, double href);
int ncm_sd_gpu_set_href (ncm_sd_gpu_ctx *ctx, double href);
int ncm_sd_gpu_get_weights (ncm_sd_gpu_ctx *ctx, int n_kernels, double *weights);

/* ---- batched evaluation (the new "eval_vec") -----------------------------------------
 * q calls of ncm_stats_dist_eval_m2lnp / ncm_stats_dist_eval (ncm_stats_dist.c:1527-1554 ->
 * ncm_stats_dist_kde.c:596-681, ncm_stats_dist_vkde.c:631-723, kernels'
 * eval_sum0/sum1_gamma_lambda).  X [q x d] host, out [q] host. */
int ncm_sd_gpu_eval_m2lnp (ncm_sd_gpu_ctx *ctx, int q, const double *X, int ldx, double *m2lnp_out);
int ncm_sd_gpu_eval (ncm_sd_gpu_ctx *ctx, int q, const double *X, int ldx, double *p_out);
/* same with device-resident buffers (dX [q x d] ld = ldx, dOut [q]); asynchronous on the ctx stream */
int ncm_sd_gpu_eval_m2lnp_dev (ncm_sd_gpu_ctx *ctx, int q, const double *dX, int ldx, double *dOut);

/* ---- interpolation matrix -------------------------------------------------------------
 * klass->compute_IM followed by the 1/f_i row scaling of _ncm_stats_dist_compute_IM_full
 * (ncm_stats_dist.c:791-804; ncm_stats_dist_kde.c:492-557; ncm_stats_dist_vkde.c:517-606).
 * Rows = the n_obs uploaded points, columns = the n_kernels centres; row_scale[n_obs]
 * (= 1/f_i) may be NULL.  The matrix stays resident on the device for ncm_sd_gpu_nnls_solve;
 * it is copied to IM_host [n_obs x n_kernels, ld = n_kernels] only if IM_host != NULL. */
int ncm_sd_gpu_compute_IM (ncm_sd_gpu_ctx *ctx, const double *row_scale, double *IM_host);

/* ---- NNLS ------------------------------------------------------------------------------
 * ncm_nnls_solve with NCM_NNLS_UMETHOD_NORMAL on the resident IM and f = 1
 * (ncm_nnls.c:767-871; NcmISet logic ncm_iset.c:853-1077).  x_out [n_kernels].
 * The un-normalised solution also replaces the resident weights (call set_weights after the
 * host applied the shrink normalisation of ncm_stats_dist.c:1087-1093). */
typedef struct ncm_sd_gpu_nnls_stats
{
  int n_chol;    /* Cholesky factorisations */
  int n_lu;      /* systems dposv found not positive definite, solved by the symmetric-indefinite L D L^T (dsysv, ncm_nnls.c:573-606) */
  int n_outer;   /* accepted outer iterations */
  int n_passive; /* final passive-set size */
  double chol_flops; /* sum over the factorisations of |P|^3 / 3 (algorithmic flops of the Cholesky solves) */
  double syrk_flops; /* nrows * ncols^2 (normal equations) */
  int n_lowrank;     /* passive-set systems solved by low-rank modification of an earlier factor (lowrank.cu) instead of dposv */
  int n_lowrank_fallback; /* ... whose refinement correction was too large: solved again by a fresh factorisation */
  int n_trinv;       /* triangular inverses formed for those solves */
  int max_lowrank_k; /* largest |D| + |A| served that way */
  double lowrank_flops;   /* 2 |B|^2 (k + 1) per such solve + |B|^3 / 3 per triangular inverse */
  int n_dist_chol;   /* factorisations whose trailing updates were distributed over the ranks (dist_chol.cu) */
  int n_qr;      /* ... whose L D L^T met an exactly singular pivot: least squares by Householder QR (dgels, ncm_nnls.c:608-638) */
  int n_lowrank_nested; /* low-rank solves whose removed set contained the previous one: the k x k factor was extended, not rebuilt */
  int reserved_;
} ncm_sd_gpu_nnls_stats;

int ncm_sd_gpu_nnls_solve (ncm_sd_gpu_ctx *ctx, double reltol, double *x_out, double *rnorm_out, ncm_sd_gpu_nnls_stats *stats);
/* generic entry (used by the NNLS parity tests): A [nrows x ncols] host row-major, f [nrows] host */
int ncm_sd_gpu_nnls_solve_host (ncm_sd_gpu_ctx *ctx, int nrows, int ncols, const double *A, int lda, const double *f,
                                double reltol, double *x_out, double *rnorm_out, ncm_sd_gpu_nnls_stats *stats);

/* ---- proposal sampling -------------------------------------------------------------------
 * The affine map of kernel->sample (ncm_stats_dist_kernel_gauss.c:335-355,
 * ncm_stats_dist_kernel_st.c:388-414): X_out[r] = centre[kidx[r]] + scale[r] * U_{kidx[r]}^T (href * Z[r]),
 * with the standard normals Z [q x d] (and scale = sqrt(nu / chi2_nu) for ST, NULL for Gauss)
 * drawn by the caller's RNG in reference order. */
int ncm_sd_gpu_sample_apply (ncm_sd_gpu_ctx *ctx, int q, const int *kidx, const double *Z, int ldz, const double *scale,
                             double *X_out, int ldx);
/* counter-based (Philox4x32-10) throughput mode, one stream per proposal row; NOT stream-
 * compatible with the reference RNG (hard part a). */
int ncm_sd_gpu_sample_philox (ncm_sd_gpu_ctx *ctx, int q, unsigned long long seed, unsigned long long offset,
                              double *X_out, int ldx, int *kidx_out);

/* ---- multi-GPU ------------------------------------------------------------------------------
 * One process per GPU, centres and factors replicated.  Either the caller shards by hand
 * (ncm_sd_gpu_set_row_shard: this rank's IM row block; its own query rows) or, in SPMD callers,
 * ncm_sd_gpu_set_auto_shard does it behind unchanged calls.  The exchanges, all NCCL on the context's
 * streams: all-reduce of the NNLS normal equations (M = IM^T IM, b = IM^T 1) and of A^T r / |r|^2 per
 * outer iteration; all-gather of the densities of a sharded evaluation; all-gather of the factors of a
 * sharded ncm_sd_gpu_vkde_prepare; broadcast / all-gather of the panels of the distributed Cholesky that
 * ncm_sd_gpu_nnls_solve uses for passive sets of NCM_SD_GPU_DIST_CHOL_MIN_N (8192) indices and more
 * (replaces ncm_matrix_cholesky_solve, ncm_matrix.c:1199-1210, as called from ncm_nnls.c:655-666). */
int ncm_sd_gpu_comm_unique_id (char id_out[128]);
int ncm_sd_gpu_comm_init (ncm_sd_gpu_ctx *ctx, int nranks, int rank, const char id[128]);
/* restrict compute_IM to observation rows [row0, row0 + nrows) on this rank */
int ncm_sd_gpu_set_row_shard (ncm_sd_gpu_ctx *ctx, int row0, int nrows);
/* Automatic sharding for SPMD callers (every rank makes the same calls with the same host arrays, as the ranks of a multi-rank
 * APES run do): with on != 0 and a communicator, compute_IM (IM_host == NULL) takes the row block [n_obs rank / G, n_obs (rank + 1) / G)
 * of this rank and the NNLS all-reduces the normal equations; eval / eval_m2lnp upload and evaluate this rank's block of the query
 * rows and ncclAllGather the results on the device, so that every rank receives the complete output; ncm_sd_gpu_vkde_prepare builds the
 * factors of a contiguous block of centres per rank and all-gathers them.  Centres and factors stay replicated.  Shards ncm_stats_dist.c:878-1094 (rows of the interpolation matrix) and the per-walker density calls of
 * walker_apes.c:742-812. */
int ncm_sd_gpu_set_auto_shard (ncm_sd_gpu_ctx *ctx, int on);
/* ncclAllGather of `count` doubles per rank on the context stream (device pointers; drecv holds nranks * count): the exchange that
 * follows a query-sharded ncm_sd_gpu_eval_m2lnp_dev */
int ncm_sd_gpu_allgather_dev (ncm_sd_gpu_ctx *ctx, const double *dsend, double *drecv, int count);

/* ---- instrumentation ------------------------------------------------------------------------- */
enum
{
  NCM_SD_GPU_T_EVAL = 0,  /* eval kernels */
  NCM_SD_GPU_T_IM,        /* IM kernel */
  NCM_SD_GPU_T_SYRK,      /* normal equations */
  NCM_SD_GPU_T_CHOL,      /* Cholesky + triangular solves */
  NCM_SD_GPU_T_NNLS_MISC, /* gathers, residuals, gradients */
  NCM_SD_GPU_T_H2D,
  NCM_SD_GPU_T_D2H,
  NCM_SD_GPU_T_PREP,      /* VKDE prepare_kernel: kNN + local covariance + Cholesky */
  NCM_SD_GPU_T_LOWRANK,   /* passive-set solves by low-rank modification: triangular inverse, bordered solve, refinement */
  NCM_SD_GPU_T_COMM,      /* NCCL collectives on the data path: all-reduce of the normal equations, all-gather of the densities */
  NCM_SD_GPU_T_LEN
};
/* device time (CUDA events on the ctx stream) accumulated per stage, milliseconds, and the
 * number of kernels launched since the last reset */
int ncm_sd_gpu_get_timers (ncm_sd_gpu_ctx *ctx, double ms[NCM_SD_GPU_T_LEN], long long *n_launches);
int ncm_sd_gpu_reset_timers (ncm_sd_gpu_ctx *ctx);
/* bytes this context copied host->device / device->host since the last reset_timers */
int ncm_sd_gpu_get_traffic (ncm_sd_gpu_ctx *ctx, long long *h2d_bytes, long long *d2h_bytes);
int ncm_sd_gpu_enable_timers (ncm_sd_gpu_ctx *ctx, int enable);

/* page-locked host memory for the large arrays that cross PCIe every prepare (the factor slab U_all_out): cudaHostAlloc / cudaFreeHost */
int ncm_sd_gpu_host_alloc (void **ptr, size_t bytes);
int ncm_sd_gpu_host_free (void *ptr);

/* which VKDE evaluation kernel serves the uploaded factors: 1 = DMMA path with explicit inverses (d >= 13 and
 * max_i cond_1 (U_i) <= 1e5, reported in cond_max), 0 = forward substitution (vkde.cu).  Both replace
 * ncm_stats_dist_vkde.c:631-723 / :517-606. */
int ncm_sd_gpu_vkde_path (ncm_sd_gpu_ctx *ctx, int *uses_mma, double *cond_max);

/* plain FP64 building blocks exported for tests and microbenchmarks (device pointers) */
int ncm_sd_gpu_dsyrk_ata_dev (ncm_sd_gpu_ctx *ctx, int nrows, int ncols, const double *dA, int lda, double *dM, int ldm);
int ncm_sd_gpu_dpotrf_upper_dev (ncm_sd_gpu_ctx *ctx, int n, double *dM, int ldm, int *info_host);
/* dposv 'U' as ncm_matrix_cholesky_solve calls it (ncm_matrix.c:1199-1210): factor the upper triangle of dM in place and
 * overwrite dRhs [n] with the solution of M x = rhs */
int ncm_sd_gpu_dposv_upper_dev (ncm_sd_gpu_ctx *ctx, int n, double *dM, int ldm, double *dRhs, int *info_host);
/* dtrtri 'U' 'N': dW = dU^-1 for an upper-triangular row-major factor (the lower triangle of dU is not read, that of dW is
 * zeroed); dScratch is n x ld doubles.  Recursive doubling over DMMA GEMMs (csrc/lowrank.cu): the building block that lets
 * the NNLS solve the passive-set systems after the first (ncm_nnls.c:728-751) without a new dposv each. */
int ncm_sd_gpu_dtrtri_upper_dev (ncm_sd_gpu_ctx *ctx, int n, const double *dU, int ld, double *dW, double *dScratch);
/* dsysv 'U' as _ncm_nnls_solve_normal_LU calls it (ncm_nnls.c:573-606 over ncm_lapack.c:798): Bunch-Kaufman L D L^T of the symmetric
 * (possibly indefinite) matrix in the upper triangle of the row-major dM, then the solve; dM is destroyed, dRhs [n] becomes x.
 * info_host: 0, or the 1-based index of an exactly singular pivot (nothing solved then, as dsysv).  csrc/ldl_bk.cu */
int ncm_sd_gpu_dsysv_upper_dev (ncm_sd_gpu_ctx *ctx, int n, double *dM, int ldm, double *dRhs, int *info_host);
/* dgels 'N' on the columns dIdx [n] (ascending, device) of the row-major dA [m x lda] with right-hand side dF [m], as
 * _ncm_nnls_solve_normal_QR does (ncm_nnls.c:608-638): Householder QR, dX [n] = least-squares solution.  info_host: 0 or the
 * 1-based index of an exactly zero diagonal entry of R.  csrc/qr_ls.cu */
int ncm_sd_gpu_dgels_cols_dev (ncm_sd_gpu_ctx *ctx, int m, int n, const double *dA, int lda, const int *dIdx, const double *dF, double *dX, int *info_host);

#ifdef __cplusplus
}
#endif

#endif /* NCM_SD_GPU_H */
