// FP64 tensor-core (DMMA.8x8x4) building blocks of the NNLS solve:
//   ata_kernel      C(upper tiles) = beta C + alpha P^T P      P: K x n row-major
//                   - normal equations M = IM^T IM  (cblas_dsyrk Upper/Trans, ncm_matrix.c:1548-1579)
//                   - trailing update of the blocked Cholesky (alpha = -1, beta = 1, P = panel rows of U)
//   gemv kernels    b = A^T f, r = f - A x, g = A^T r           (cblas_dgemv, ncm_nnls.c:710-726, 791-792)
//
// ata_kernel: 128 x 128 (or 64 x 64) CTA tile, warp tiles of 8 x 4 (4 x 4) DMMA tiles,
// K consumed 16 rows per stage through a 4-stage cp.async ring.  P^T is never materialised: both
// DMMA operands are read from the same row-major K x 128 slabs (fragment element (i, r) = P[r][i]),
// stored with a row pitch of 132 doubles so that the 4 rows x 4 columns a half-warp touches fall
// in 16 distinct 8-byte bank pairs.
#include "ctx.h"

namespace {

constexpr int BK = 16;

// Tile configurations.  Big: 128 x 128 CTA tile, 8 warps as 2 x 4, warp tile 64 x 32 (8 x 4 DMMA tiles) -- the
// throughput shape (SYRK of the normal equations, large trailing updates).  Small: 64 x 64 CTA tile, 4 warps as
// 2 x 2, warp tile 32 x 32 -- four times more CTAs of a quarter of the latency each, for the K = 64 trailing
// updates of mid-size Cholesky factorisations where a 128-tile grid would be a single under-filled wave.
struct AtaBig {
  static constexpr int TB = 128, WN = 4, MI = 8, NI = 4, THREADS = 256, MINB = 1, STAGES = 4;
};
struct AtaSmall {
  static constexpr int TB = 64, WN = 2, MI = 4, NI = 4, THREADS = 128, MINB = 4, STAGES = 3;
};
template <typename Cfg>
struct AtaDerived {
  static constexpr int PITCH = Cfg::TB + 4;            // (TB + 4) * 8 B = k * 128 + 32: conflict-free fragment loads
  static constexpr int SLAB  = BK * PITCH;             // doubles per operand per stage
  static constexpr size_t SMEM = (size_t) Cfg::STAGES * 2 * SLAB * sizeof(double);
  static constexpr int CHUNKS = BK * (Cfg::TB / 2) / Cfg::THREADS;   // 16-byte chunks per thread per operand per stage
};

__device__ __forceinline__ void tile_from_linear(int t, int nt, int &ti, int &tj) {
  // upper-triangular tiles enumerated row by row: row ti has (nt - ti) tiles
  double disc = (2.0 * nt + 1.0) * (2.0 * nt + 1.0) - 8.0 * t;
  int i       = (int) ((2.0 * nt + 1.0 - sqrt(disc)) * 0.5);
  if (i < 0) i = 0;
  while (i > 0 && (i * (2 * nt - i + 1)) / 2 > t) --i;
  while (((i + 1) * (2 * nt - i)) / 2 <= t) ++i;
  ti = i;
  tj = i + (t - (i * (2 * nt - i + 1)) / 2);
}

// SUBC: C -= P^T P (alpha = -1, beta = 1): C is loaded into the accumulators up front (the loads overlap with the
// cp.async prologue), the A fragments are negated, and the epilogue is stores only.
template <typename Cfg, bool SUBC>
__device__ __forceinline__ void ata_tile(const double *__restrict__ P, int ldp, int K, int n, double *__restrict__ C, int ldc, double alpha, double beta,
                                         int ti, int tj, double *smem) {
  using D = AtaDerived<Cfg>;
  constexpr int TB = Cfg::TB, PITCH = D::PITCH, SLAB = D::SLAB, MI = Cfg::MI, NI = Cfg::NI, STAGES = Cfg::STAGES;
  const int i0 = ti * TB, j0 = tj * TB;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp / Cfg::WN, wn = warp % Cfg::WN;
  const int lr = lane & 3, lc = lane >> 2;           // fragment coordinates
  const int rm = wm * (MI * 8), cn = wn * (NI * 8);  // warp tile origin inside the CTA tile

  auto load_stage = [&](int kb, int st) {
    double *sA = smem + (size_t) st * 2 * SLAB;
    double *sB = sA + SLAB;
    const int r0 = kb * BK;
#pragma unroll
    for (int it = 0; it < D::CHUNKS; ++it) {
      const int chunk = tid + it * Cfg::THREADS;
      const int r     = chunk / (TB / 2);
      const int cc    = (chunk % (TB / 2)) * 2;       // column (doubles) within the slab
      const int gr    = r0 + r;
      const bool rv   = gr < K;
      {
        const int gc = i0 + cc;
        int bytes    = rv ? (n - gc) * 8 : 0;
        bytes        = bytes < 0 ? 0 : (bytes > 16 ? 16 : bytes);
        const double *src = (bytes > 0) ? (P + (size_t) gr * ldp + gc) : P;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(sA + r * PITCH + cc)), "l"(src), "r"(bytes) : "memory");
      }
      {
        const int gc = j0 + cc;
        int bytes    = rv ? (n - gc) * 8 : 0;
        bytes        = bytes < 0 ? 0 : (bytes > 16 ? 16 : bytes);
        const double *src = (bytes > 0) ? (P + (size_t) gr * ldp + gc) : P;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(sB + r * PITCH + cc)), "l"(src), "r"(bytes) : "memory");
      }
    }
  };

  const int nkb = (K + BK - 1) / BK;
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nkb) load_stage(s, s);
    cp_async_commit();
  }

  double acc[MI][NI][2];
#pragma unroll
  for (int a = 0; a < MI; ++a) {
    const int gi = i0 + rm + a * 8 + lc;
#pragma unroll
    for (int b = 0; b < NI; ++b) {
      acc[a][b][0] = acc[a][b][1] = 0.0;
      if (SUBC) {
        const int gj = j0 + cn + b * 8 + 2 * lr;
        if (gi < n && gj + 1 < n) {
          const double2 v = *reinterpret_cast<const double2 *>(C + (size_t) gi * ldc + gj);
          acc[a][b][0]    = v.x;
          acc[a][b][1]    = v.y;
        } else if (gi < n && gj < n) {
          acc[a][b][0] = C[(size_t) gi * ldc + gj];
        }
      }
    }
  }

  for (int kb = 0; kb < nkb; ++kb) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const int nk = kb + STAGES - 1;
      if (nk < nkb) load_stage(nk, nk % STAGES);
      cp_async_commit();
    }
    const double *sA = smem + (size_t) (kb % STAGES) * 2 * SLAB;
    const double *sB = sA + SLAB;
#pragma unroll
    for (int ks = 0; ks < BK / 4; ++ks) {
      double af[MI], bf[NI];
      const double *pa = sA + (ks * 4 + lr) * PITCH + rm + lc;
      const double *pb = sB + (ks * 4 + lr) * PITCH + cn + lc;
#pragma unroll
      for (int a = 0; a < MI; ++a) af[a] = SUBC ? -pa[a * 8] : pa[a * 8];
#pragma unroll
      for (int b = 0; b < NI; ++b) bf[b] = pb[b * 8];
#pragma unroll
      for (int a = 0; a < MI; ++a)
#pragma unroll
        for (int b = 0; b < NI; ++b) dmma884(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
    }
  }
  cp_async_wait<0>();

  // epilogue: lane holds C[row = lc][cols 2 lr, 2 lr + 1] of each 8 x 8 tile
#pragma unroll
  for (int a = 0; a < MI; ++a) {
    const int gi = i0 + rm + a * 8 + lc;
    if (gi >= n) continue;
#pragma unroll
    for (int b = 0; b < NI; ++b) {
      const int gj = j0 + cn + b * 8 + 2 * lr;
      double *pc   = C + (size_t) gi * ldc + gj;
      if (SUBC) {
        if (gj + 1 < n)
          *reinterpret_cast<double2 *>(pc) = make_double2(acc[a][b][0], acc[a][b][1]);
        else if (gj < n)
          pc[0] = acc[a][b][0];
        continue;
      }
      if (gj + 1 < n) {
        double2 v;
        if (beta != 0.0) {
          v   = *reinterpret_cast<const double2 *>(pc);
          v.x = fma(alpha, acc[a][b][0], beta * v.x);
          v.y = fma(alpha, acc[a][b][1], beta * v.y);
        } else {
          v.x = alpha * acc[a][b][0];
          v.y = alpha * acc[a][b][1];
        }
        *reinterpret_cast<double2 *>(pc) = v;
      } else if (gj < n) {
        pc[0] = (beta != 0.0) ? fma(alpha, acc[a][b][0], beta * pc[0]) : alpha * acc[a][b][0];
      }
    }
  }
}

template <typename Cfg, bool SUBC>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MINB)
ata_kernel(const double *__restrict__ P, int ldp, int K, int n, double *__restrict__ C, int ldc, double alpha, double beta, int nt,
           const int2 *__restrict__ tiles) {
  extern __shared__ __align__(16) double smem[];
  int ti, tj;
  if (tiles != nullptr) {   // explicit tile list: the column blocks one rank owns in the distributed factorisation (dist_chol.cu)
    ti = tiles[blockIdx.x].x;
    tj = tiles[blockIdx.x].y;
  } else {
    tile_from_linear(blockIdx.x, nt, ti, tj);
  }
  ata_tile<Cfg, SUBC>(P, ldp, K, n, C, ldc, alpha, beta, ti, tj, smem);
}

// the same over an explicit tile list with FEWER CTAs than tiles (each loops over its share): the trailing update of the distributed
// factorisation then leaves some SMs to the panel chain of the next step instead of making every one of its kernels wait for a CTA
// of this one to retire (dist_chol.cu)
__global__ void __launch_bounds__(AtaBig::THREADS, AtaBig::MINB)
ata_tiles_persistent_kernel(const double *__restrict__ P, int ldp, int K, int n, double *__restrict__ C, int ldc, const int2 *__restrict__ tiles, int ntiles) {
  extern __shared__ __align__(16) double smem[];
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    ata_tile<AtaBig, true>(P, ldp, K, n, C, ldc, -1.0, 1.0, tiles[t].x, tiles[t].y, smem);
    __syncthreads();   // the staging ring is reused by the next tile
  }
}

template <typename Cfg, bool SUBC>
cudaError_t ata_launch(cudaStream_t st, const double *dP, int ldp, int K, int n, double *dC, int ldc, double alpha, double beta, int head_tile_rows = 0) {
  using D = AtaDerived<Cfg>;
  static bool attr_set[NCM_MAX_DEVICES] = {};   // function attributes are per device
  int dev__ = 0;
  cudaGetDevice(&dev__);
  dev__ = dev__ < 0 || dev__ >= NCM_MAX_DEVICES ? 0 : dev__;
  if (!attr_set[dev__]) {
    cudaError_t e = cudaFuncSetAttribute(ata_kernel<Cfg, SUBC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) D::SMEM);
    if (e != cudaSuccess) return e;
    attr_set[dev__] = true;
  }
  const int nt = (n + Cfg::TB - 1) / Cfg::TB;
  // tiles are enumerated row by row over the upper triangle: the first nt of them are tile row 0, the next nt - 1 tile row 1, ...
  const int hr    = head_tile_rows > nt ? nt : head_tile_rows;
  const int tiles = hr > 0 ? (hr * (2 * nt - hr + 1)) / 2 : nt * (nt + 1) / 2;
  ata_kernel<Cfg, SUBC><<<tiles, Cfg::THREADS, D::SMEM, st>>>(dP, ldp, K, n, dC, ldc, alpha, beta, nt, nullptr);
  return cudaGetLastError();
}

// C[tile] -= P^T P for an explicit list of 128 x 128 tiles (ti <= tj, tile units)
cudaError_t ata_launch_tiles(cudaStream_t st, const double *dP, int ldp, int K, int n, double *dC, int ldc, const int2 *dTiles, int ntiles, int max_ctas) {
  using D = AtaDerived<AtaBig>;
  static bool attr_set[NCM_MAX_DEVICES] = {};
  int dev__ = 0;
  cudaGetDevice(&dev__);
  dev__ = dev__ < 0 || dev__ >= NCM_MAX_DEVICES ? 0 : dev__;
  if (!attr_set[dev__]) {
    cudaError_t e = cudaFuncSetAttribute(ata_kernel<AtaBig, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) D::SMEM);
    if (e != cudaSuccess) return e;
    attr_set[dev__] = true;
  }
  const int nt = (n + AtaBig::TB - 1) / AtaBig::TB;
  if (max_ctas > 0 && max_ctas < ntiles) {
    static bool attr2_set[NCM_MAX_DEVICES] = {};
    if (!attr2_set[dev__]) {
      cudaError_t e = cudaFuncSetAttribute(ata_tiles_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) D::SMEM);
      if (e != cudaSuccess) return e;
      attr2_set[dev__] = true;
    }
    ata_tiles_persistent_kernel<<<max_ctas, AtaBig::THREADS, D::SMEM, st>>>(dP, ldp, K, n, dC, ldc, dTiles, ntiles);
    return cudaGetLastError();
  }
  ata_kernel<AtaBig, true><<<ntiles, AtaBig::THREADS, D::SMEM, st>>>(dP, ldp, K, n, dC, ldc, -1.0, 1.0, nt, dTiles);
  return cudaGetLastError();
}

// out[j] (+)= sum_r A[r][j] v[r]   (v == nullptr means v = 1): two-pass, deterministic.
__global__ void gemv_t_partial_kernel(const double *__restrict__ A, int lda, int nrows, int ncols, const double *__restrict__ v,
                                      double *__restrict__ part, int rows_per_block) {
  const int j  = blockIdx.x * blockDim.x + threadIdx.x;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(nrows, r0 + rows_per_block);
  if (j >= ncols) return;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int r = r0;
  for (; r + 3 < r1; r += 4) {
    s0 = fma(A[(size_t) r * lda + j], v ? v[r] : 1.0, s0);
    s1 = fma(A[(size_t) (r + 1) * lda + j], v ? v[r + 1] : 1.0, s1);
    s2 = fma(A[(size_t) (r + 2) * lda + j], v ? v[r + 2] : 1.0, s2);
    s3 = fma(A[(size_t) (r + 3) * lda + j], v ? v[r + 3] : 1.0, s3);
  }
  for (; r < r1; ++r) s0 = fma(A[(size_t) r * lda + j], v ? v[r] : 1.0, s0);
  part[(size_t) blockIdx.y * ncols + j] = (s0 + s1) + (s2 + s3);
}

__global__ void reduce_parts_kernel(const double *__restrict__ part, int nparts, int n, double *__restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  double s = 0.0;
  for (int p = 0; p < nparts; ++p) s += part[(size_t) p * n + j];
  out[j] = s;
}

// r[i] = f[i] - sum_j A[i][j] x[j]   (f == nullptr means f = 1); one warp per row; also per-block sum of r^2
__global__ void residual_kernel(const double *__restrict__ A, int lda, int nrows, int ncols, const double *__restrict__ x,
                                const double *__restrict__ f, double *__restrict__ r, double *__restrict__ ss_part) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row  = blockIdx.x * (blockDim.x >> 5) + warp;
  double ri      = 0.0;
  if (row < nrows) {
    const double *a = A + (size_t) row * lda;
    double s0 = 0.0, s1 = 0.0;
    int j = lane;
    for (; j + 32 < ncols; j += 64) {
      s0 = fma(a[j], x[j], s0);
      s1 = fma(a[j + 32], x[j + 32], s1);
    }
    for (; j < ncols; j += 32) s0 = fma(a[j], x[j], s0);
    double s = s0 + s1;
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    ri = (f ? f[row] : 1.0) - s;
    if (lane == 0) r[row] = ri;
  }
  __shared__ double sh[32];
  if (lane == 0) sh[warp] = (row < nrows) ? ri * ri : 0.0;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += sh[w];
    ss_part[blockIdx.x] = t;
  }
}

}   // namespace

int dsyrk_ata_general_on(ncm_sd_gpu_ctx *c, cudaStream_t st, int K, int n, const double *dP, int ldp, double *dC, int ldc, double alpha, double beta);

int dsyrk_ata_general(ncm_sd_gpu_ctx *c, int K, int n, const double *dP, int ldp, double *dC, int ldc, double alpha, double beta) {
  return dsyrk_ata_general_on(c, c->stream, K, n, dP, ldp, dC, ldc, alpha, beta);
}

// C -= P^T P on the listed upper 128-tiles only (P: K x n row-major, C: n x n)
int dsyrk_ata_tiles(ncm_sd_gpu_ctx *c, int K, int n, const double *dP, int ldp, double *dC, int ldc, const int *dTiles /* pairs (ti, tj) */, int ntiles) {
  return dsyrk_ata_tiles_on(c, c->stream, K, n, dP, ldp, dC, ldc, dTiles, ntiles, 0);
}
// max_ctas > 0: at most that many CTAs, each looping over its share of the tiles
int dsyrk_ata_tiles_on(ncm_sd_gpu_ctx *c, cudaStream_t st, int K, int n, const double *dP, int ldp, double *dC, int ldc, const int *dTiles, int ntiles, int max_ctas) {
  if (ntiles <= 0 || K <= 0) return NCM_SD_GPU_OK;
  if ((ldp & 1) || (ldc & 1) || (((uintptr_t) dP) & 15) || (((uintptr_t) dC) & 15))
    return c->fail(NCM_SD_GPU_EINVAL, "ata: operands must be 16-byte aligned with even leading dimensions");
  NCM_CUDA_OK(c, ata_launch_tiles(st, dP, ldp, K, n, dC, ldc, reinterpret_cast<const int2 *>(dTiles), ntiles, max_ctas));
  c->n_launches++;
  return NCM_SD_GPU_OK;
}

int dsyrk_ata_general_on(ncm_sd_gpu_ctx *c, cudaStream_t st, int K, int n, const double *dP, int ldp, double *dC, int ldc, double alpha, double beta) {
  if (n <= 0 || K <= 0) return NCM_SD_GPU_OK;
  if ((ldp & 1) || (ldc & 1) || (((uintptr_t) dP) & 15) || (((uintptr_t) dC) & 15))
    return c->fail(NCM_SD_GPU_EINVAL, "ata: operands must be 16-byte aligned with even leading dimensions");
  // small tiles when K is short and the whole 64-tile grid is co-resident (latency-bound regime)
  const int nt64     = (n + 63) / 64;
  const bool small   = (K <= 128) && (nt64 * (nt64 + 1) / 2 <= 4 * c->n_sm);   // all 64-tiles resident at 4 CTAs per SM
  const bool subc    = (alpha == -1.0 && beta == 1.0);
  cudaError_t e;
  if (small)
    e = subc ? ata_launch<AtaSmall, true>(st, dP, ldp, K, n, dC, ldc, alpha, beta) : ata_launch<AtaSmall, false>(st, dP, ldp, K, n, dC, ldc, alpha, beta);
  else
    e = subc ? ata_launch<AtaBig, true>(st, dP, ldp, K, n, dC, ldc, alpha, beta) : ata_launch<AtaBig, false>(st, dP, ldp, K, n, dC, ldc, alpha, beta);
  NCM_CUDA_OK(c, e);
  c->n_launches++;
  NCM_CUDA_OK(c, cudaGetLastError());
  return NCM_SD_GPU_OK;
}

// C[0:64, 0:n] -= P[:, 0:64]^T P  (the first 64-row block of the trailing update only): lets the blocked Cholesky factor two
// 64-wide panels before it touches the rest of the matrix, so that the big update runs with K = 128
int dsyrk_ata_first_rows64_on(ncm_sd_gpu_ctx *c, cudaStream_t st, int K, int n, const double *dP, int ldp, double *dC, int ldc) {
  if (n <= 0 || K <= 0) return NCM_SD_GPU_OK;
  if ((ldp & 1) || (ldc & 1) || (((uintptr_t) dP) & 15) || (((uintptr_t) dC) & 15))
    return c->fail(NCM_SD_GPU_EINVAL, "ata: operands must be 16-byte aligned with even leading dimensions");
  NCM_CUDA_OK(c, (ata_launch<AtaSmall, true>(st, dP, ldp, K, n, dC, ldc, -1.0, 1.0, 1)));
  c->n_launches++;
  NCM_CUDA_OK(c, cudaGetLastError());
  return NCM_SD_GPU_OK;
}

int dsyrk_ata_first_rows64(ncm_sd_gpu_ctx *c, int K, int n, const double *dP, int ldp, double *dC, int ldc) {
  return dsyrk_ata_first_rows64_on(c, c->stream, K, n, dP, ldp, dC, ldc);
}

// C[0:rows, 0:n] -= P[:, 0:rows]^T P with 128 x 128 tiles (rows a multiple of 128): the head strip of a trailing update, i.e. exactly
// the block rows the next panel group of the look-ahead Cholesky factors
int dsyrk_ata_head_rows_on(ncm_sd_gpu_ctx *c, cudaStream_t st, int K, int n, int rows, const double *dP, int ldp, double *dC, int ldc) {
  if (n <= 0 || K <= 0 || rows <= 0) return NCM_SD_GPU_OK;
  if ((ldp & 1) || (ldc & 1) || (((uintptr_t) dP) & 15) || (((uintptr_t) dC) & 15) || (rows % AtaBig::TB) != 0)
    return c->fail(NCM_SD_GPU_EINVAL, "ata: operands must be 16-byte aligned with even leading dimensions, head rows a multiple of the tile");
  NCM_CUDA_OK(c, (ata_launch<AtaBig, true>(st, dP, ldp, K, n, dC, ldc, -1.0, 1.0, rows / AtaBig::TB)));
  c->n_launches++;
  NCM_CUDA_OK(c, cudaGetLastError());
  return NCM_SD_GPU_OK;
}

int dsyrk_ata(ncm_sd_gpu_ctx *c, int nrows, int ncols, const double *dA, int lda, double *dM, int ldm) {
  return dsyrk_ata_general(c, nrows, ncols, dA, lda, dM, ldm, 1.0, 0.0);
}

// out = A^T v (v may be null = ones); tmp must hold ceil(nrows / rows_per_block) * ncols doubles
int gemv_t(ncm_sd_gpu_ctx *c, const double *dA, int lda, int nrows, int ncols, const double *dv, double *dOut, DevBuf &tmp) {
  int nblk = (c->n_sm * 2 * 256 + ncols - 1) / ncols;   // enough row blocks for ~2 CTAs per SM
  if (nblk < 1) nblk = 1;
  if (nblk > (nrows + 63) / 64) nblk = (nrows + 63) / 64;
  if (nblk < 1) nblk = 1;
  const int rpb = (nrows + nblk - 1) / nblk;
  nblk          = (nrows + rpb - 1) / rpb;
  if (!tmp.reserve((size_t) nblk * ncols * sizeof(double))) return c->fail(NCM_SD_GPU_ENOMEM, "gemv_t: out of device memory");
  dim3 grid((ncols + 255) / 256, nblk);
  gemv_t_partial_kernel<<<grid, 256, 0, c->stream>>>(dA, lda, nrows, ncols, dv, tmp.as<double>(), rpb);
  reduce_parts_kernel<<<(ncols + 255) / 256, 256, 0, c->stream>>>(tmp.as<double>(), nblk, ncols, dOut);
  c->n_launches += 2;
  NCM_CUDA_OK(c, cudaGetLastError());
  return NCM_SD_GPU_OK;
}

// r = f - A x ; ss_part[nblocks] partial sums of squares (summed on the host in block order)
int residual(ncm_sd_gpu_ctx *c, const double *dA, int lda, int nrows, int ncols, const double *dx, const double *df, double *dr,
             double *d_ss_part, int *nblocks_out) {
  const int wpb = 8;
  const int nb  = (nrows + wpb - 1) / wpb;
  residual_kernel<<<nb, wpb * 32, 0, c->stream>>>(dA, lda, nrows, ncols, dx, df, dr, d_ss_part);
  c->n_launches++;
  *nblocks_out = nb;
  NCM_CUDA_OK(c, cudaGetLastError());
  return NCM_SD_GPU_OK;
}
