// Cholesky solve of the NNLS passive-set systems distributed over the ranks of one NVLink / NVSwitch box.
//
// Replaces ncm_matrix_cholesky_solve (dposv 'U', ncm_matrix.c:1199-1210) as called from _ncm_nnls_solve_normal_cholesky
// (ncm_nnls.c:655-666) when |P| >= 8192 and the context carries a communicator.  In the row-sharded NNLS every rank holds the
// same all-reduced normal matrix; here the O(n^3) trailing updates of its factorisation are split over the ranks:
//
//   block columns of DB = 512, column j owned by rank j mod G (1-D block-cyclic); for step k = 0 .. nb-1
//     owner(k):  factor the diagonal block (chol_fused.cu, on a few SMs), ncclBroadcast {U_kk, info}
//     every rank, for its own columns j > k:   U_kj = U_kk^-T A_kj     (panel_trsm_kernel: blocked forward substitution, the strip of
//                                                                        the solution in shared memory, U_kk read through L1)
//     ncclAllGather of the panel row (each rank's tiles, 2 MB each), written back into rows k of every rank's matrix
//     every rank, for its own columns j > k:   A_ij -= U_ki^T U_kj, k < i <= j   (ata_kernel, 128 x 128 DMMA tiles, K = 512)
//
// Look-ahead.  The chain factor -> broadcast -> panel solve -> all-gather is what bounds a step once the updates are split eight
// ways, so it runs on its own highest-priority stream and only waits for the part of the previous update it reads: the update of
// step k is issued in two launches, block row k+1 first (event), the rest after; the diagonal factorisation of step k+1 takes 40
// CTAs, not the device, so it starts while the rest of update k still runs on the context stream.  The inverses of the diagonal
// blocks, which only the final triangular solves need, are formed by their owners on a third stream and all-gathered once.
//
// Every rank ends with the complete factor (bit-identical: each tile is computed once and copied), so the two triangular solves
// run replicated: y_k = W_kk^T (b_k - U[0:k, k]^T y), x_k = W_kk (y_k - U[k, k+1:] x) by block rows with the gathered inverses.
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

// U_kj = U_kk^-T A_kj for this rank's block columns j = first + m G (blockIdx.z = m), one CTA per strip of 32 columns.
// Forward substitution by 64-row blocks: thread (tr, tc) holds rows b*64 + tr*8 .. +8 of column tc in registers and subtracts the
// contributions of the solved rows (read back from the strip in shared memory); the 64 x 64 blocks of U_kk it needs, (a, b) for
// a <= b, stream through a double-buffered shared-memory tile (cp.async, the next block in flight while this one is used); in the
// diagonal block the 8 row groups are solved one after the other, each from registers.  Ukk: bs x bs upper, ld = DB, zeros below.
constexpr int TS_W = 32, TS_T = 256, TS_UP = 66;   // strip width, threads, pitch of the staged U block (16-byte aligned rows)
constexpr size_t TS_SMEM = ((size_t) DB * TS_W + 2 * 64 * TS_UP) * sizeof(double);
__global__ void __launch_bounds__(TS_T, 1) panel_trsm_kernel(const double *__restrict__ Ukk, const double *__restrict__ M, int ldm, int n, int k0, int bs, int first,
                                                               int G, int nb, double *__restrict__ stage_me) {
  extern __shared__ __align__(16) double ts_sm[];
  double *Xs = ts_sm;                     // [DB][TS_W]
  double *Ub = ts_sm + (size_t) DB * TS_W;   // [2][64][TS_UP]
  const int j = first + blockIdx.z * G;
  if (j >= nb) return;
  const int j0 = j * DB, wj = min(DB, n - j0);
  const int c0 = blockIdx.x * TS_W;
  if (c0 >= wj) return;
  const int tid = threadIdx.x, tc = tid & 31, tr = tid >> 5;
  const int col = c0 + tc;
  const bool valid = col < wj;
  const int nblk = (bs + 63) / 64;
  // block (a, b) of U_kk -> buffer: thread t copies 16 doubles of row t / 4 (Ukk is a full DB x DB tile: no bounds to check)
  auto stage_u = [&](int a, int b, int buf) {
    const int r = tid >> 2, cs = (tid & 3) * 16;
    const double *src = Ukk + (size_t) (a * 64 + r) * DB + b * 64 + cs;
    double *dst       = Ub + (size_t) buf * 64 * TS_UP + r * TS_UP + cs;
#pragma unroll
    for (int q = 0; q < 8; ++q) cp_async16(dst + 2 * q, src + 2 * q);
    cp_async_commit();
  };
  int buf = 0;
  stage_u(0, 0, 0);
  for (int b = 0; b < nblk; ++b) {
    const int r0 = b * 64 + tr * 8;
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = (valid && r0 + i < bs) ? M[(size_t) (k0 + r0 + i) * ldm + j0 + col] : 0.0;
    for (int a = 0; a <= b; ++a) {
      cp_async_wait<0>();
      __syncthreads();   // block (a, b) has landed; everybody is done with the other buffer
      if (a < b)
        stage_u(a + 1, b, buf ^ 1);
      else if (b + 1 < nblk)
        stage_u(0, b + 1, buf ^ 1);
      const double *U = Ub + (size_t) buf * 64 * TS_UP;
      if (a < b) {
#pragma unroll 4
        for (int p = 0; p < 64; ++p) {
          const double xv   = Xs[(a * 64 + p) * TS_W + tc];
          const double2 *u2 = reinterpret_cast<const double2 *>(U + p * TS_UP + tr * 8);
          const double2 u01 = u2[0], u23 = u2[1], u45 = u2[2], u67 = u2[3];
          acc[0] = fma(-u01.x, xv, acc[0]);
          acc[1] = fma(-u01.y, xv, acc[1]);
          acc[2] = fma(-u23.x, xv, acc[2]);
          acc[3] = fma(-u23.y, xv, acc[3]);
          acc[4] = fma(-u45.x, xv, acc[4]);
          acc[5] = fma(-u45.y, xv, acc[5]);
          acc[6] = fma(-u67.x, xv, acc[6]);
          acc[7] = fma(-u67.y, xv, acc[7]);
        }
      } else {
        for (int g = 0; g < 8; ++g) {
          if (tr == g) {
            double u[8][8], x[8];
#pragma unroll
            for (int q = 0; q < 8; ++q)
#pragma unroll
              for (int i = q; i < 8; ++i) u[q][i] = U[(g * 8 + q) * TS_UP + g * 8 + i];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              double sacc = acc[i];
#pragma unroll
              for (int q = 0; q < i; ++q) sacc = fma(-u[q][i], x[q], sacc);
              x[i] = (r0 + i < bs) ? sacc / u[i][i] : 0.0;
              Xs[(r0 + i) * TS_W + tc] = x[i];
            }
          }
          __syncthreads();
          if (tr > g) {
#pragma unroll
            for (int p = 0; p < 8; ++p) {
              const double xv   = Xs[(b * 64 + g * 8 + p) * TS_W + tc];
              const double2 *u2 = reinterpret_cast<const double2 *>(U + (g * 8 + p) * TS_UP + tr * 8);
              const double2 u01 = u2[0], u23 = u2[1], u45 = u2[2], u67 = u2[3];
              acc[0] = fma(-u01.x, xv, acc[0]);
              acc[1] = fma(-u01.y, xv, acc[1]);
              acc[2] = fma(-u23.x, xv, acc[2]);
              acc[3] = fma(-u23.y, xv, acc[3]);
              acc[4] = fma(-u45.x, xv, acc[4]);
              acc[5] = fma(-u45.y, xv, acc[5]);
              acc[6] = fma(-u67.x, xv, acc[6]);
              acc[7] = fma(-u67.y, xv, acc[7]);
            }
          }
        }
      }
      buf ^= 1;
    }
  }
  cp_async_wait<0>();
  __syncthreads();
  double *out = stage_me + (size_t) blockIdx.z * DB * DB;
  if (valid)
    for (int r = tr; r < bs; r += TS_T / 32) out[(size_t) r * DB + col] = Xs[r * TS_W + tc];
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
  const int capk = (nb + G - 1) / G;   // diagonal blocks per rank: block k lives in slot (k mod G) * capk + k / G of the gathered inverses
  const size_t tile = (size_t) DB * DB;
  // buffers: pk = {U_kk, info, scratch of the inversions}; Wall = G x capk inverses; stage = G x cap0 tiles; vectors
  if (!c->dcPack.reserve((3 * tile + 16) * sizeof(double)) || !c->dcW.reserve((size_t) G * capk * tile * sizeof(double)) ||
      !c->dcStage.reserve((size_t) G * cap0 * tile * sizeof(double)) || !c->dcVec.reserve((size_t) (4 * DB + 128) * sizeof(double)))
    return c->fail(NCM_SD_GPU_ENOMEM, "dist_chol: out of device memory");
  if (DB > chol_fused_max_n()) return c->fail(NCM_SD_GPU_EINVAL, "dist_chol: block size exceeds the fused factorisation");
  double *pkU = c->dcPack.as<double>(), *pkInfo = pkU + tile, *pkS = pkInfo + 16, *pkU2 = pkS + tile;
  double *Wall = c->dcW.as<double>(), *stage = c->dcStage.as<double>();
  double *dInfoAcc = pkInfo + 8;   // slot 0 travels with the broadcast, slot 8 accumulates locally
  auto wall = [&](int k) { return Wall + ((size_t) (k % G) * capk + k / G) * tile; };
  // streams: sU = the context stream (trailing updates, and whatever the caller queued before / queues after), sP = the chain of the
  // panels (highest priority, also carries the collectives), sW = inversions of the diagonal blocks
  cudaStream_t sU = c->stream;
  if (c->dc_sW == nullptr) {
    int least = 0, greatest = 0;
    NCM_CUDA_OK(c, cudaDeviceGetStreamPriorityRange(&least, &greatest));
    NCM_CUDA_OK(c, cudaStreamCreateWithPriority(&c->dc_sP, cudaStreamNonBlocking, greatest));
    NCM_CUDA_OK(c, cudaStreamCreateWithFlags(&c->dc_sW, cudaStreamNonBlocking));
    for (cudaEvent_t *e : {&c->dc_evP, &c->dc_evA, &c->dc_evD, &c->dc_evW, &c->dc_evU}) NCM_CUDA_OK(c, cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  }
  static const bool la_on = getenv("NCM_SD_GPU_DIST_CHOL_LOOKAHEAD") == nullptr || atoi(getenv("NCM_SD_GPU_DIST_CHOL_LOOKAHEAD")) != 0;
  static const int reserve_env = getenv("NCM_SD_GPU_DIST_CHOL_RESERVE") != nullptr ? atoi(getenv("NCM_SD_GPU_DIST_CHOL_RESERVE")) : -1;
  // SMs kept free of update CTAs for the panel chain (experiment knob, off by default: at 8 GPUs 64 reserved SMs changed nothing --
  // 52.4 against 52.1 ms for two factorisations of order 16384 -- and 96 cost 12 ms: the chain is not waiting for SMs)
  const int reserve = !la_on ? 0 : (reserve_env >= 0 ? std::min(reserve_env, c->n_sm - 16) : 0);
  cudaStream_t sP = la_on ? c->dc_sP : sU, sW = la_on ? c->dc_sW : sU;   // =0: everything in order on the context stream (A/B switch)
  static bool attr_set[NCM_MAX_DEVICES] = {};
  {
    int dev = c->device < 0 || c->device >= NCM_MAX_DEVICES ? 0 : c->device;
    if (!attr_set[dev]) {
      NCM_CUDA_OK(c, cudaFuncSetAttribute(panel_trsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) TS_SMEM));
      attr_set[dev] = true;
    }
  }
  // tile lists of the trailing updates of every step, built once and uploaded in one copy: step k updates the 128-tiles (ti, tj) with
  // ti >= 4 (k + 1), ti <= tj, block column of tj owned by this rank -- first those of block row k + 1 (what step k + 1's chain reads),
  // then the rest
  const int T128 = DB / 128, nt = (n + 127) / 128;
  std::vector<int> tiles;
  std::vector<size_t> offA(nb + 1, 0), offB(nb + 1, 0);
  for (int k = 0; k < nb; ++k) {
    for (int pass = 0; pass < 2; ++pass) {
      (pass == 0 ? offA : offB)[k] = tiles.size() / 2;
      for (int j = k + 1; j < nb; ++j) {
        if (j % G != me) continue;
        for (int tj = j * T128; tj < std::min((j + 1) * T128, nt); ++tj)
          for (int ti = (k + 1) * T128; ti <= tj; ++ti) {
            const bool rowA = ti < (k + 2) * T128;
            if (rowA != (pass == 0)) continue;
            tiles.push_back(ti);
            tiles.push_back(tj);
          }
      }
    }
  }
  offA[nb] = tiles.size() / 2;
  if (!c->dcTiles.reserve((tiles.size() + 2) * sizeof(int))) return c->fail(NCM_SD_GPU_ENOMEM, "dist_chol: out of device memory");
  if (!tiles.empty()) NCM_CUDA_OK(c, cudaMemcpyAsync(c->dcTiles.p, tiles.data(), tiles.size() * sizeof(int), cudaMemcpyHostToDevice, sU));
  NCM_CUDA_OK(c, cudaMemsetAsync(pkInfo, 0, 16 * sizeof(double), sU));
  NCM_CUDA_OK(c, cudaStreamSynchronize(sU));   // `tiles` is pageable host memory; this also orders everything queued before behind us
  const int *dTiles = c->dcTiles.as<int>();

  for (int k = 0; k < nb; ++k) {
    const int k0 = k * DB, bs = std::min(DB, n - k0), owner = k % G;
    double *Akk = dM + (size_t) k0 * ldm + k0;
    if (k > 0) NCM_CUDA_OK(c, cudaStreamWaitEvent(sP, c->dc_evA, 0));   // block row k carries the update of step k - 1
    if (me == owner) {
      // 40 CTAs: the 36 tiles of a 512 block keep them all busy, and the launch does not wait for the whole device
      int rc = dpotrf_upper_solve_fused_on(c, sP, 40, bs, Akk, ldm, nullptr, nullptr);
      if (rc != NCM_SD_GPU_OK) return rc;
      dc_set_info_kernel<<<1, 1, 0, sP>>>(c->chol_flags.as<int>(), k0, pkInfo);
      pack_upper_kernel<<<dim3((bs + 127) / 128, bs), 128, 0, sP>>>(Akk, ldm, pkU, DB, bs, 1);
      c->n_launches += 2;
      // the inverse of the block is only needed by the triangular solves at the very end: its owner forms it off the chain
      NCM_CUDA_OK(c, cudaEventRecord(c->dc_evD, sP));
      NCM_CUDA_OK(c, cudaStreamWaitEvent(sW, c->dc_evD, 0));
      pack_upper_kernel<<<dim3((bs + 127) / 128, bs), 128, 0, sW>>>(Akk, ldm, pkU2, DB, bs, 1);
      rc = trinv_upper_on(c, sW, bs, pkU2, wall(k), pkS, DB, nullptr);
      if (rc != NCM_SD_GPU_OK) return rc;
    }
    {
      ncclResult_t r = api.Broadcast(pkU, pkU, tile + 8, ncclDouble, owner, (ncclComm_t) c->nccl_comm, sP);
      if (r != ncclSuccess) return c->fail(NCM_SD_GPU_ENCCL, std::string("ncclBroadcast: ") + api.GetErrorString(r));
    }
    dc_acc_info_kernel<<<1, 1, 0, sP>>>(pkInfo, dInfoAcc);
    c->n_launches++;
    if (me != owner) pack_upper_kernel<<<dim3((bs + 127) / 128, bs), 128, 0, sP>>>(pkU, DB, Akk, ldm, bs, 0);
    if (k + 1 >= nb) break;
    const int cap   = (nb - k - 1 + G - 1) / G;
    const int first = k + 1 + (((me - (k + 1)) % G) + G) % G;
    const int nown  = first < nb ? (nb - 1 - first) / G + 1 : 0;
    double *stage_me = stage + (size_t) me * cap * tile;
    if (nown > 0) {
      panel_trsm_kernel<<<dim3(DB / TS_W, 1, nown), TS_T, TS_SMEM, sP>>>(pkU, dM, ldm, n, k0, bs, first, G, nb, stage_me);
      c->n_launches++;
    }
    {
      ncclResult_t r = api.AllGather(stage_me, stage, (size_t) cap * tile, ncclDouble, (ncclComm_t) c->nccl_comm, sP);
      if (r != ncclSuccess) return c->fail(NCM_SD_GPU_ENCCL, std::string("ncclAllGather: ") + api.GetErrorString(r));
    }
    // rows k of the trailing columns: nobody on sU touches them (update k - 1 works on rows >= k + 1 by now)
    unpack_panel_kernel<<<dim3(DB / 128, bs, G * cap), 128, 0, sP>>>(stage, cap, G, k, nb, dM, ldm, n, bs);
    c->n_launches++;
    NCM_CUDA_OK(c, cudaEventRecord(c->dc_evP, sP));
    NCM_CUDA_OK(c, cudaStreamWaitEvent(sU, c->dc_evP, 0));
    const int nA = (int) (offB[k] - offA[k]), nB = (int) (offA[k + 1] - offB[k]);
    if (nA > 0) {
      int rc = dsyrk_ata_tiles_on(c, sU, bs, n, dM + (size_t) k0 * ldm, ldm, dM, ldm, dTiles + 2 * offA[k], nA);
      if (rc != NCM_SD_GPU_OK) return rc;
    }
    NCM_CUDA_OK(c, cudaEventRecord(c->dc_evA, sU));
    if (nB > 0) {
      // the bulk of the update (optionally on fewer CTAs than SMs, see `reserve` above)
      int rc = dsyrk_ata_tiles_on(c, sU, bs, n, dM + (size_t) k0 * ldm, ldm, dM, ldm, dTiles + 2 * offB[k], nB, reserve > 0 ? c->n_sm - reserve : 0);
      if (rc != NCM_SD_GPU_OK) return rc;
    }
  }
  // join: the gathered inverses (one all-gather of every rank's blocks) and the status, back on the context stream
  NCM_CUDA_OK(c, cudaEventRecord(c->dc_evW, sW));
  NCM_CUDA_OK(c, cudaStreamWaitEvent(sP, c->dc_evW, 0));
  if (dRhs != nullptr) {
    ncclResult_t r = api.AllGather(Wall + (size_t) me * capk * tile, Wall, (size_t) capk * tile, ncclDouble, (ncclComm_t) c->nccl_comm, sP);
    if (r != ncclSuccess) return c->fail(NCM_SD_GPU_ENCCL, std::string("ncclAllGather: ") + api.GetErrorString(r));
  }
  NCM_CUDA_OK(c, cudaEventRecord(c->dc_evU, sP));
  NCM_CUDA_OK(c, cudaStreamWaitEvent(sU, c->dc_evU, 0));
  cudaStream_t st = sU;
  NCM_CUDA_OK(c, cudaGetLastError());
  int info_all = 0;
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
    int rc = gemv_t(c, wall(k), DB, bs, bs, rhs_k, dRhs + k0, c->nn_tmp);   // y_k = W_kk^T ( . )
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
    int rc = row_dot_upper(c, wall(k), DB, bs, yk, v2);      // x_k = W_kk ( . )
    if (rc != NCM_SD_GPU_OK) return rc;
    NCM_CUDA_OK(c, cudaMemcpyAsync(dRhs + k0, v2, bs * sizeof(double), cudaMemcpyDeviceToDevice, st));
  }
  NCM_CUDA_OK(c, cudaGetLastError());
  NCM_CUDA_OK(c, cudaStreamSynchronize(st));
  return NCM_SD_GPU_OK;
}
