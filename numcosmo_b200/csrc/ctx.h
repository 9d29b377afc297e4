// Internal context of libncm_sd_gpu: stream, device buffers, timers.
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <vector>
#include "../../include/ncm_sd_gpu.h"
#include "common.cuh"

struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  // grow-only device allocation; returns false on failure
  bool reserve(size_t bytes);
  void release();
  template <typename T>
  T *as() const { return static_cast<T *>(p); }
};

struct PinBuf {
  void *p = nullptr;
  size_t cap = 0;
  bool reserve(size_t bytes);
  void release();
  template <typename T>
  T *as() const { return static_cast<T *>(p); }
};

struct ncm_sd_gpu_ctx {
  int device = 0;
  int n_sm = 148;
  cudaStream_t stream = nullptr;
  std::string err;

  // kernel
  int kind = 0;
  double nu = 3.0;
  int d = 0;
  int dp = 0;   // VKDE template dimension (d padded to the next instantiated size)

  // model
  int type = -1;   // NCM_SD_GPU_KDE / VKDE
  int n_obs = 0, n_kernels = 0;
  double href = 1.0;
  double lnnorm = 0.0;       // KDE common lnnorm
  int row0 = 0, nrows = 0;   // IM row shard on this rank
  bool have_weights = false;
  bool prep_pending = false;   // vkde_prepare done, vkde_finish (lnnorms + record packing) still to come

  // VKDE buffers
  DevBuf sample;     // [n_obs x d] raw points (row-major, ld = d)
  DevBuf vrec;       // [n_kernels x vrec_len] packed records: theta[dp], Lp[dp(dp+1)/2] (diag = 1/U_kk)
  int vrec_len = 0;
  DevBuf vrec_mma;   // tensor-core records (vkde_mma.cu): theta[dp], W = L^-1 in DMMA fragment order
  int vrec_mma_len = 0;
  bool vkde_mma = false;      // the tensor-core path serves eval / IM (d >= 13 and max cond below VKDE_MMA_MAX_COND)
  double vkde_cond = 0.0;     // max_i |L_i|_1 |L_i^-1|_1 of the last pack
  DevBuf lnu;        // [n_kernels] per-kernel lnnorm (VKDE) -- without d ln h
  DevBuf cterm;      // [n_kernels] ln w_i - lnu_i
  DevBuf clin;       // Student-t linear-domain eval: [n_alloc] exp(cterm_i - cmax), then {cmax} and an int flag (see common.cuh)
  int clin_n = 0;    // entries of clin (cmax sits at clin[clin_n], the underflow flag at clin[clin_n + 1])
  DevBuf weights;    // [n_kernels]
  DevBuf Ufull;      // VKDE: [n_kernels x d x d] dense factors (sample_apply) ; KDE: [d x d]

  // KDE buffers
  int kp = 0;        // padded K of the augmented GEMM
  DevBuf zc;         // [n_obs x d] whitened centred points
  DevBuf zmean;      // [d]
  DevBuf bfrag;      // fragment-major B operand of the kernels [n_kernels/8][kp/4][32]
  DevBuf kde_U;      // [d x d]

  // queries / outputs / partials
  DevBuf qX, qOut, qA, part;
  PinBuf pinX, pinOut;

  // IM + NNLS
  DevBuf IM;         // [nrows_local x n_kernels]
  DevBuf rowscale;   // [n_obs]
  DevBuf M, MU, nn_b, nn_x, nn_r, nn_g, nn_tmp, nn_idx, nn_f;
  DevBuf dist;   // VKDE prepare_kernel: squared distances centre x point [n_kernels x n_obs]
  DevBuf lrW, lrWt, lrS, lrV, lrT, lrPart, lrSmall, lrVec, lrIdx;   // low-rank passive-set solves (lowrank.cu): base inverse, scratch, bordered blocks
  DevBuf bkWork, qrWork;   // fallback solvers of the NNLS (ldl_bk.cu: pivots, info, barrier; qr_ls.cu: gathered columns)
  DevBuf chol_flags, chol_part;   // single-launch Cholesky (chol_fused.cu): dependency flags, back-substitution contributions
  int chol_epoch = 0;
  long long *chol_trace = nullptr;   // device buffer [n_sm][cap][2] set by ncm_sd_gpu_chol_trace (debugging aid)
  int chol_trace_cap = 0;
  PinBuf pin_nn;

  // NCCL
  void *nccl_comm = nullptr;
  int nranks = 1, rank = 0;
  bool auto_shard = false;   // ncm_sd_gpu_set_auto_shard: compute_IM takes this rank's row block, eval its query block + all-gather
  bool im_sharded = false;   // the resident IM holds this rank's row block only
  DevBuf gath;               // all-gather staging [nranks x cap]
  DevBuf dcPack, dcW, dcStage, dcVec, dcTiles;   // distributed Cholesky (dist_chol.cu): broadcast block, kept inverses, panel staging, tile lists

  // timers
  bool timers_on = false;
  double t_ms[NCM_SD_GPU_T_LEN] = {0};
  long long n_launches = 0;
  long long h2d_bytes = 0, d2h_bytes = 0;   // bytes moved over PCIe by this context (ncm_sd_gpu_get_traffic)
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
  // look-ahead Cholesky (chol.cu): a highest-priority stream for the latency-bound panel kernels, run concurrently with the
  // trailing updates on `stream`; created on first use
  cudaStream_t stream_hi = nullptr;
  // distributed Cholesky (dist_chol.cu): stream of the diagonal-block inversions, events between its three streams
  cudaStream_t dc_sP = nullptr, dc_sW = nullptr;
  cudaEvent_t dc_evP = nullptr, dc_evA = nullptr, dc_evD = nullptr, dc_evW = nullptr, dc_evU = nullptr;
  cudaEvent_t ev_panel = nullptr, ev_tail = nullptr;

  int fail(int code, const std::string &msg) {
    err = msg;
    return code;
  }
};

#define NCM_CUDA_OK(ctx, call)                                                                                     \
  do {                                                                                                             \
    cudaError_t e__ = (call);                                                                                      \
    if (e__ != cudaSuccess)                                                                                        \
      return (ctx)->fail(NCM_SD_GPU_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e__) + " (" + __FILE__ + ":" + \
                                               std::to_string(__LINE__) + ")");                                   \
  } while (0)

// host <-> device copies go through these so that the context can report its PCIe traffic
inline cudaError_t ncm_memcpy_async(ncm_sd_gpu_ctx *c, void *dst, const void *src, size_t bytes, cudaMemcpyKind kind, cudaStream_t s) {
  if (kind == cudaMemcpyHostToDevice) c->h2d_bytes += (long long) bytes;
  if (kind == cudaMemcpyDeviceToHost) c->d2h_bytes += (long long) bytes;
  return cudaMemcpyAsync(dst, src, bytes, kind, s);
}
inline cudaError_t ncm_memcpy2d_async(ncm_sd_gpu_ctx *c, void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height,
                                      cudaMemcpyKind kind, cudaStream_t s) {
  if (kind == cudaMemcpyHostToDevice) c->h2d_bytes += (long long) (width * height);
  if (kind == cudaMemcpyDeviceToHost) c->d2h_bytes += (long long) (width * height);
  // densely packed on both sides (the usual case: ld == d): one linear copy instead of `height` row descriptors
  if (dpitch == width && spitch == width) return cudaMemcpyAsync(dst, src, width * height, kind, s);
  return cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, kind, s);
}

// The collectives are timed with their own event pair so that a COMM timer may sit inside another stage's scope (the time is then
// counted in both; the all-reduce of the normal matrix, the only large one, is kept outside the SYRK scope in nnls.cu).
struct StageTimer {
  ncm_sd_gpu_ctx *c;
  int stage;
  cudaEvent_t e0, e1;
  StageTimer(ncm_sd_gpu_ctx *ctx, int st) : c(ctx), stage(st) {
    const bool comm = st == NCM_SD_GPU_T_COMM;
    e0 = comm ? c->ev2 : c->ev0;
    e1 = comm ? c->ev3 : c->ev1;
    if (c->timers_on) cudaEventRecord(e0, c->stream);
  }
  ~StageTimer() {
    if (c->timers_on) {
      cudaEventRecord(e1, c->stream);
      cudaEventSynchronize(e1);
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e0, e1);
      c->t_ms[stage] += ms;
    }
  }
};

// Number of centre splits for the eval / IM grids: q_tiles x splits CTAs on `slots` resident CTA slots.  The smallest
// split count that fills the device and leaves a last wave at least 92 % full (a 512-tile grid on 296 slots is 1.73 waves:
// 13 % of the launch is tail, ncu r01c); more splits only cost 16 bytes of partials per query and split.
inline int ncm_pick_splits(int slots, int q_tiles, int max_splits) {
  if (max_splits < 1) max_splits = 1;
  int first = (slots + q_tiles - 1) / q_tiles;
  if (first < 1) first = 1;
  if (first > max_splits) return max_splits;
  int best = first;
  double best_eff = 0.0;
  const int hi = first + 16 < max_splits ? first + 16 : max_splits;
  for (int s = first; s <= hi; ++s) {
    const long ctas  = (long) q_tiles * s;
    const long waves = (ctas + slots - 1) / slots;
    const double eff = (double) ctas / (double) (waves * slots);
    if (eff > best_eff + 1e-9) {
      best_eff = eff;
      best     = s;
    }
    if (eff >= 0.92) return s;
  }
  return best;
}

// ---- kernels' host launchers (defined in the .cu files) --------------------------------------------
int vkde_pad_dim(int d);
int vkde_pack(ncm_sd_gpu_ctx *c, const double *dU_all /* n x d x d */);
int vkde_prepare_dev(ncm_sd_gpu_ctx *c, int n_obs, int n_kernels, int k, const double *dZ, const double *dX, int *dNbr, double *dU_all, int *dFail, int cbeg = 0);
int vkde_mma_pad_dim(int d);
int vkde_mma_pack(ncm_sd_gpu_ctx *c, const double *dU_all, double *cond_max_host);
int vkde_mma_eval_launch(ncm_sd_gpu_ctx *c, int q, const double *dX, int ldx, double *dOut, bool as_density);
int vkde_mma_im_launch(ncm_sd_gpu_ctx *c, const double *dInvNorm, const double *dRowScale);
int vkde_eval_launch(ncm_sd_gpu_ctx *c, int q, const double *dX, int ldx, double *dOut, bool as_density);
int vkde_im_launch(ncm_sd_gpu_ctx *c, const double *dRowScale);

int kde_prepare(ncm_sd_gpu_ctx *c, const double *dInvU /* n_obs x d */);
int kde_set_weights(ncm_sd_gpu_ctx *c);
int kde_eval_launch(ncm_sd_gpu_ctx *c, int q, const double *dX, int ldx, double *dOut, bool as_density);
int kde_im_launch(ncm_sd_gpu_ctx *c, const double *dRowScale);

int update_cterm(ncm_sd_gpu_ctx *c);
int lse_finalize_launch(ncm_sd_gpu_ctx *c, const double *pm, const double *ps, const double *row_add, int q, int n_splits, double shift,
                        bool as_density, double *dOut, int *dFlagOut = nullptr, const int *dOnlyIf = nullptr);
// Student-t with integer nu: fills kp.m2 / lin / cmax (common.cuh) for the current kernel; clin must have been built by build_clin
void ncm_fill_kp(const ncm_sd_gpu_ctx *c, KernParams &kp, bool eval);
int build_clin(ncm_sd_gpu_ctx *c, const double *dCterm, int n_alloc);
bool st_linear_enabled(const ncm_sd_gpu_ctx *c);

int chol_fused_max_n();
int dpotrf_upper_solve_fused(ncm_sd_gpu_ctx *c, int n, double *dM, int ldm, double *dRhs, int *info_host);
int dpotrf_upper_solve_fused_on(ncm_sd_gpu_ctx *c, cudaStream_t st, int max_ctas, int n, double *dM, int ldm, double *dRhs, int *info_host);
int dpotrf_upper_solve_any(ncm_sd_gpu_ctx *c, int n, double *dM, int ldm, double *dRhs, double *dDinv, int *dInfo, int *info_host);
int dsyrk_ata(ncm_sd_gpu_ctx *c, int nrows, int ncols, const double *dA, int lda, double *dM, int ldm);
int dpotrf_upper(ncm_sd_gpu_ctx *c, int n, double *dM, int ldm, double *dRhs /* optional, solved in place */, int *info_host);
int nnls_solve_dev(ncm_sd_gpu_ctx *c, int nrows, int ncols, const double *dA, int lda, const double *dF, double reltol, double *x_host,
                   double *rnorm_host, ncm_sd_gpu_nnls_stats *stats);
// lowrank.cu: passive-set solves by low-rank modification of a base factor
struct LowrankBufs {
  double *W, *Wt, *S;     // n x ldm each: W = U^-1 of the base factor, its transpose, scratch of the recursive doubling
  double *V, *T;          // nB x ldv each: [M_BA | E_D | b_B] and W^T times it
  double *part;           // split-K partial sums (lowrank_part_doubles)
  double *Lg;             // packed L J L^T factor of the k x k system + its inverse diagonal
  double *z, *z2, *y, *xB, *dxB, *xfull, *rfull, *rB, *tr, *out;
  double *tb;             // W^T b_B of the current base (set by its own k = 0 solve)
  bool tb_valid = false;
  int *idxB, *idxA, *posD, *bsel, *psrc, *info;
};
int dsysv_upper_solve(ncm_sd_gpu_ctx *c, int n, double *dS, int lds, double *dRhs, int *info_host);
int dgels_cols_solve(ncm_sd_gpu_ctx *c, int m, int n, const double *dA, int lda, const int *dIdx, const double *dF, double *dX, int *info_host);
int dist_chol_min_n();
int dpotrf_upper_solve_dist(ncm_sd_gpu_ctx *c, int n, double *dM, int ldm, double *dRhs, int *info_host);
int lowrank_kmax();
size_t lowrank_part_doubles(int n, int ldv);
int symmetrize_upper(ncm_sd_gpu_ctx *c, int n, double *dM, int ld);
int trinv_upper(ncm_sd_gpu_ctx *c, int n, const double *dU, double *dW, double *dS, int ld, double *dWt);
int trinv_upper_on(ncm_sd_gpu_ctx *c, cudaStream_t st, int n, const double *dU, double *dW, double *dS, int ld, double *dWt);
int dsyrk_ata_tiles_on(ncm_sd_gpu_ctx *c, cudaStream_t st, int K, int n, const double *dP, int ldp, double *dC, int ldc, const int *dTiles, int ntiles, int max_ctas = 0);
int lowrank_solve(ncm_sd_gpu_ctx *c, const double *dM, int ldm, int n, const double *db, int nB, int na, int nd, int np, const LowrankBufs &w, int ldv,
                  bool refine, int kold = 0);
int sample_apply_launch(ncm_sd_gpu_ctx *c, int q, const int *dIdx, const double *dZ, int ldz, const double *dScale, double *dX, int ldx);
