// Cholesky solve of the NNLS passive-set systems distributed over the ranks of one NVLink / NVSwitch box.
//
// Replaces ncm_matrix_cholesky_solve (dposv 'U', ncm_matrix.c:1199-1210) as called from _ncm_nnls_solve_normal_cholesky
// (ncm_nnls.c:655-666) when |P| >= 8192 and the context carries a communicator.  In the row-sharded NNLS every rank holds the
// same all-reduced normal matrix, and until now every rank factorised it by itself (at N = 16384 the replicated
// factorisations were all that was left of the step on 8 GPUs).  Here the O(n^3) trailing updates are split:
//
//   block columns of DB = 512, column j owned by rank j mod G (1-D block-cyclic); for step k = 0 .. nb-1
//     owner(k):  factor the diagonal block (chol_fused.cu), W_kk = U_kk^-1 (lowrank.cu), ncclBroadcast {U_kk, W_kk, info}
//     every rank, for its own columns j > k:   U_kj = W_kk^T A_kj          (DMMA GEMM, K = 512)
//     ncclAllGather of the panel row (each rank's tiles, 2 MB each), written back into rows k of every rank's matrix
//     every rank, for its own columns j > k:   A_ij -= U_ki^T U_kj, k < i <= j   (ata_kernel, 128 x 128 DMMA tiles, K = 512)
//
// Every rank ends with the complete factor (bit-identical: each tile is computed once and copied), so the two triangular solves
// run replicated: y_k = W_kk^T (b_k - U[0:k, k]^T y), x_k = W_kk (y_k - U[k, k+1:] x) by block rows with the kept inverses.
// The decisions of the NNLS stay replicated and identical.
#include <algorithm>
#include <cstring>
#include <vector>
#include "ctx.h"
#include "gemm_tile.cuh"
#include "nccl_shim.h"

int dpotrf_upper_solve_any(ncm_sd_gpu_ctx *c, int n, double *dM, int ldm, double *dRhs, double *dDinv, int *dInfo, int *info_host);
int dsyrk_ata_tiles(ncm_sd_gpu_ctx *c, int K, int n, const double *dP, int ldp, double *dC, int ldc, const int *dTiles, int ntiles);
int gemv_t(ncm_sd_gpu_ctx *c, const double *dA, int lda, int nrows, int ncols, const double *dv, double *dOut, DevBuf &tmp);
int row_dot_upper(ncm_sd_gpu_ctx *c, const double *dA, int lda, int n, const double *dv, double *dOut);

namespace {

using namespace ncm_gemm;
constexpr int DB = 512;

// dst (ld = DB, bs x bs) <- upper triangle of src (ld = lds), zeros below; or the reverse (upper triangle only)
__global__ void pack_upper_kernel(const double *__restrict__ src, int lds, double *__restrict__ dst, int ldd, int bs, int zero_lower) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (i >= bs || j >= bs) return;
  if (j >= i)
    dst[(size_t) i * ldd + j] = src[(size_t) i * lds + j];
  else if (zero_lower)
    dst[(size_t) i * ldd + j] = 0.0;
}

// stage[(g * cap + m)][DB x DB] tiles  ->  rows k0 .. k0 + bs of M, block column j = first(g) + m G
__global__ void unpack_panel_kernel(const double *__restrict__ stage, int cap, int G, int k, int nb, double *__restrict__ M, int ldm, int n, int bs) {
  const int g = blockIdx.z / cap, m = blockIdx.z % cap;
  int first = k + 1 + (((g - (k + 1)) % G) + G) % G;   // smallest j > k with j mod G == g
  const int j = first + m * G;
  if (j >= nb) return;
  const int j0 = j * DB, wj = min(DB, n - j0);
  const int cc = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (r >= bs || cc >= wj) return;
  M[(size_t) (k * DB + r) * ldm + j0 + cc] = stage[((size_t) blockIdx.z * DB + r) * DB + cc];
}

// U_kj = W_kk^T A_kj for this rank's block columns j = first + m G (blockIdx.z = m): A = W (bs x bs upper, ld DB), B = M rows k, C = stage tile
__global__ void __launch_bounds__(GTHREADS) panel_solve_kernel(const double *__restrict__ W, const double *__restrict__ M, int ldm, int n, int k0, int bs, int first, int G,
                                                               int nb, double *__restrict__ stage_me) {
  extern __shared__ __align__(16) double smem[];
  const int j = first + blockIdx.z * G;
  if (j >= nb) return;
  const int j0 = j * DB, wj = min(DB, n - j0);
  const int i0 = blockIdx.y * GT, c0 = blockIdx.x * GT;
  if (i0 >= bs || c0 >= wj) return;
  gemm_tile<true>(W, DB, M + (size_t) k0 * ldm + j0, ldm, stage_me + (size_t) blockIdx.z * DB * DB, DB, bs, wj, i0, c0, 0, min(bs, i0 + GT), 1.0, smem);
}

// owner: info slot of the broadcast block <- status of the diagonal factorisation left on the device by the fused kernel
__global__ void dc_set_info_kernel(const int *__restrict__ fused_info, int k0, double *__restrict__ pk_info) {
  const int i = fused_info[0];
  pk_info[0]  = i == 0 ? 0.0 : (double) (k0 + i);
}
// every rank: remember the first failing pivot (checked once, after the last step: no host round trip per block column)
__global__ void dc_acc_info_kernel(const double *__restrict__ pk_info, double *__restrict__ acc) {
  if (acc[0] == 0.0 && pk_info[0] != 0.0) acc[0] = pk_info[0];
}

__global__ void axpby_kernel(const double *__restrict__ a, const double *__restrict__ b, double sb, int n, double *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + sb * b[i];
}
// r[i] = f[i] - sum_j A[i][j] x[j]   (warp per row)
__global__ void row_resid_kernel(const double *__restrict__ A, int lda, int nrows, int ncols, const double *__restrict__ x, const double *__restrict__ f,
                                 double *__restrict__ r) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row  = blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= nrows) return;
  const double *a = A + (size_t) row * lda;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int j = lane;
  for (; j + 96 < ncols; j += 128) {
    s0 = fma(a[j], x[j], s0);
    s1 = fma(a[j + 32], x[j + 32], s1);
    s2 = fma(a[j + 64], x[j + 64], s2);
    s3 = fma(a[j + 96], x[j + 96], s3);
  }
  for (; j < ncols; j += 32) s0 = fma(a[j], x[j], s0);
  double s = (s0 + s1) + (s2 + s3);
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) r[row] = f[row] - s;
}

}   // namespace

int trinv_upper(ncm_sd_gpu_ctx *c, int n, const double *dU, double *dW, double *dS, int ld, double *dWt);

int dist_chol_min_n() {
  static const int v = [] {
    const char *e = getenv("NCM_SD_GPU_DIST_CHOL_MIN_N");
    return e != nullptr ? atoi(e) : 8192;
  }();
  return v;
}

// In-place factorisation of the upper triangle of dM (identical on every rank) and, when dRhs != nullptr, solution of M x = rhs in place,
// with the trailing updates distributed over the ranks of the context's communicator.  info_host: 0 or the 1-based failing pivot.
int dpotrf_upper_solve_dist(ncm_sd_gpu_ctx *c, int n, double *dM, int ldm, double *dRhs, int *info_host) {
  NcclApi &api = nccl_api();
  if (c->nccl_comm == nullptr || !api.ok) return c->fail(NCM_SD_GPU_EINVAL, "dist_chol: communicator required");
  const int G = c->nranks, me = c->rank;
  const int nb = (n + DB - 1) / DB;
  const int cap0 = (nb - 1 + G - 1) / G > 0 ? (nb - 1 + G - 1) / G : 1;
  const size_t tile = (size_t) DB * DB;
  // buffers: pk = {U_kk, W_kk, scratch, info}; Wall = nb kept inverses; stage = G x cap0 tiles; vectors
  if (!c->dcPack.reserve((3 * tile + 16) * sizeof(double)) || !c->dcW.reserve((size_t) nb * tile * sizeof(double)) ||
      !c->dcStage.reserve((size_t) G * cap0 * tile * sizeof(double)) || !c->dcVec.reserve((size_t) (4 * DB + 128) * sizeof(double)))
    return c->fail(NCM_SD_GPU_ENOMEM, "dist_chol: out of device memory");
  double *pkU = c->dcPack.as<double>(), *pkW = pkU + tile, *pkInfo = pkW + tile, *pkS = pkInfo + 16;
  double *Wall = c->dcW.as<double>(), *stage = c->dcStage.as<double>();
  cudaStream_t st = c->stream;
  static bool attr_set[NCM_MAX_DEVICES] = {};
  {
    int dev = c->device < 0 || c->device >= NCM_MAX_DEVICES ? 0 : c->device;
    if (!attr_set[dev]) {
      NCM_CUDA_OK(c, cudaFuncSetAttribute(panel_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) GEMM_SMEM));
      attr_set[dev] = true;
    }
  }
  // tile lists of the trailing updates of every step, built once and uploaded in one copy: step k updates the 128-tiles (ti, tj) with
  // ti >= 4 (k + 1), ti <= tj, block column of tj owned by this rank
  const int T128 = DB / 128, nt = (n + 127) / 128;
  std::vector<int> tiles;
  std::vector<size_t> off(nb + 1, 0);
  for (int k = 0; k < nb; ++k) {
    off[k] = tiles.size() / 2;
    for (int j = k + 1; j < nb; ++j) {
      if (j % G != me) continue;
      for (int tj = j * T128; tj < std::min((j + 1) * T128, nt); ++tj)
        for (int ti = (k + 1) * T128; ti <= tj; ++ti) {
          tiles.push_back(ti);
          tiles.push_back(tj);
        }
    }
  }
  off[nb] = tiles.size() / 2;
  if (!c->dcTiles.reserve((tiles.size() + 2) * sizeof(int))) return c->fail(NCM_SD_GPU_ENOMEM, "dist_chol: out of device memory");
  if (!tiles.empty()) NCM_CUDA_OK(c, cudaMemcpyAsync(c->dcTiles.p, tiles.data(), tiles.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  NCM_CUDA_OK(c, cudaStreamSynchronize(st));   // `tiles` is pageable host memory: the copy has to be done before it goes out of scope
  const int *dTiles = c->dcTiles.as<int>();

  int info_all = 0;
  double *dInfoAcc = pkInfo + 8;   // inside the 16-double info slot; slot 0 travels with the broadcast, slot 8 accumulates locally
  NCM_CUDA_OK(c, cudaMemsetAsync(pkInfo, 0, 16 * sizeof(double), st));
  if (DB > chol_fused_max_n()) return c->fail(NCM_SD_GPU_EINVAL, "dist_chol: block size exceeds the fused factorisation");
  for (int k = 0; k < nb; ++k) {
    const int k0 = k * DB, bs = std::min(DB, n - k0), owner = k % G;
    double *Akk = dM + (size_t) k0 * ldm + k0;
    if (me == owner) {
      // the fused kernel leaves its status on the device (chol_flags[0]); nothing here waits for the host
      int rc = dpotrf_upper_solve_fused(c, bs, Akk, ldm, nullptr, nullptr);
      if (rc != NCM_SD_GPU_OK) return rc;
      dc_set_info_kernel<<<1, 1, 0, st>>>(c->chol_flags.as<int>(), k0, pkInfo);
      pack_upper_kernel<<<dim3((bs + 127) / 128, bs), 128, 0, st>>>(Akk, ldm, pkU, DB, bs, 1);
      rc = trinv_upper(c, bs, pkU, pkW, pkS, DB, nullptr);
      if (rc != NCM_SD_GPU_OK) return rc;
      c->n_launches += 2;
    }
    {
      // {U_kk, W_kk, info}: the first half of the info slot travels, the accumulator in its second half stays local
      ncclResult_t r = api.Broadcast(pkU, pkU, 2 * tile + 8, ncclDouble, owner, (ncclComm_t) c->nccl_comm, st);
      if (r != ncclSuccess) return c->fail(NCM_SD_GPU_ENCCL, std::string("ncclBroadcast: ") + api.GetErrorString(r));
    }
    dc_acc_info_kernel<<<1, 1, 0, st>>>(pkInfo, dInfoAcc);
    c->n_launches++;
    if (me != owner) pack_upper_kernel<<<dim3((bs + 127) / 128, bs), 128, 0, st>>>(pkU, DB, Akk, ldm, bs, 0);
    NCM_CUDA_OK(c, cudaMemcpyAsync(Wall + (size_t) k * tile, pkW, tile * sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (k + 1 >= nb) break;
    const int cap   = (nb - k - 1 + G - 1) / G;
    const int first = k + 1 + (((me - (k + 1)) % G) + G) % G;
    const int nown  = first < nb ? (nb - 1 - first) / G + 1 : 0;
    double *stage_me = stage + (size_t) me * cap * tile;
    if (nown > 0) {
      panel_solve_kernel<<<dim3(DB / GT, (bs + GT - 1) / GT, nown), GTHREADS, GEMM_SMEM, st>>>(pkW, dM, ldm, n, k0, bs, first, G, nb, stage_me);
      c->n_launches++;
    }
    {
      ncclResult_t r = api.AllGather(stage_me, stage, (size_t) cap * tile, ncclDouble, (ncclComm_t) c->nccl_comm, st);
      if (r != ncclSuccess) return c->fail(NCM_SD_GPU_ENCCL, std::string("ncclAllGather: ") + api.GetErrorString(r));
    }
    unpack_panel_kernel<<<dim3(DB / 128, bs, G * cap), 128, 0, st>>>(stage, cap, G, k, nb, dM, ldm, n, bs);
    c->n_launches++;
    const int ntl = (int) (off[k + 1] - off[k]);
    if (ntl > 0) {
      int rc = dsyrk_ata_tiles(c, bs, n, dM + (size_t) k0 * ldm, ldm, dM, ldm, dTiles + 2 * off[k], ntl);
      if (rc != NCM_SD_GPU_OK) return rc;
    }
  }
  NCM_CUDA_OK(c, cudaGetLastError());
  {
    double finfo = 0.0;
    NCM_CUDA_OK(c, cudaMemcpyAsync(&finfo, dInfoAcc, sizeof(double), cudaMemcpyDeviceToHost, st));
    NCM_CUDA_OK(c, cudaStreamSynchronize(st));
    info_all = (int) finfo;
  }
  if (info_host != nullptr) *info_host = info_all;
  if (info_all != 0 || dRhs == nullptr) {
    NCM_CUDA_OK(c, cudaStreamSynchronize(st));
    return NCM_SD_GPU_OK;
  }
  // triangular solves, replicated: y = U^-T b (block rows ascending), x = U^-1 y (descending); the right-hand side is overwritten
  double *v1 = c->dcVec.as<double>(), *v2 = v1 + DB + 16;
  for (int k = 0; k < nb; ++k) {
    const int k0 = k * DB, bs = std::min(DB, n - k0);
    const double *rhs_k = dRhs + k0;
    if (k > 0) {
      int rc = gemv_t(c, dM + k0, ldm, k0, bs, dRhs, v1, c->nn_tmp);          // U[0:k0, k-block]^T y[0:k0]
      if (rc != NCM_SD_GPU_OK) return rc;
      axpby_kernel<<<(bs + 255) / 256, 256, 0, st>>>(rhs_k, v1, -1.0, bs, v2);
      rhs_k = v2;
      c->n_launches++;
    }
    int rc = gemv_t(c, Wall + (size_t) k * tile, DB, bs, bs, rhs_k, dRhs + k0, c->nn_tmp);   // y_k = W_kk^T ( . )
    if (rc != NCM_SD_GPU_OK) return rc;
  }
  for (int k = nb - 1; k >= 0; --k) {
    const int k0 = k * DB, bs = std::min(DB, n - k0), k1 = k0 + bs;
    const double *yk = dRhs + k0;
    if (k1 < n) {
      row_resid_kernel<<<(bs + 7) / 8, 256, 0, st>>>(dM + (size_t) k0 * ldm + k1, ldm, bs, n - k1, dRhs + k1, dRhs + k0, v1);
      yk = v1;
      c->n_launches++;
    }
    int rc = row_dot_upper(c, Wall + (size_t) k * tile, DB, bs, yk, v2);      // x_k = W_kk ( . )
    if (rc != NCM_SD_GPU_OK) return rc;
    NCM_CUDA_OK(c, cudaMemcpyAsync(dRhs + k0, v2, bs * sizeof(double), cudaMemcpyDeviceToDevice, st));
  }
  NCM_CUDA_OK(c, cudaGetLastError());
  NCM_CUDA_OK(c, cudaStreamSynchronize(st));
  return NCM_SD_GPU_OK;
}
