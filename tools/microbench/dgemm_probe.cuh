// FP64 tensor-core (DMMA.8x8x4) building blocks of the NNLS solve:
//   ata_kernel      C(upper tiles) = beta C + alpha P^T P      P: K x n row-major
//                   - normal equations M = IM^T IM  (cblas_dsyrk Upper/Trans, ncm_matrix.c:1548-1579)
//                   - trailing update of the blocked Cholesky (alpha = -1, beta = 1, P = panel rows of U)
//   gemv kernels    b = A^T f, r = f - A x, g = A^T r           (cblas_dgemv, ncm_nnls.c:710-726, 791-792)
//
// ata_kernel: 128 x 128 CTA tile, 8 warps as 2 (M) x 4 (N), warp tile 64 x 32 = 8 x 4 DMMA tiles,
// K consumed 16 rows per stage through a 4-stage cp.async ring.  P^T is never materialised: both
// DMMA operands are read from the same row-major K x 128 slabs (fragment element (i, r) = P[r][i]),
// stored with a row pitch of 132 doubles so that the 4 rows x 4 columns a half-warp touches fall
// in 16 distinct 8-byte bank pairs.
#include "ctx.h"

namespace {

__device__ long long g_aprobe[64];
__device__ int g_api;
constexpr int BM = 128, BN = 128, BK = 16, STAGES = 4;
constexpr int PITCH = 132;                       // doubles; 132 * 8 B = 1056 = 8 * 128 + 32
constexpr int SLAB  = BK * PITCH;                // doubles per operand per stage
constexpr int ATA_THREADS = 256;
constexpr size_t ATA_SMEM = (size_t) STAGES * 2 * SLAB * sizeof(double);   // 135168 B

__device__ __forceinline__ void tile_from_linear(int t, int nt, int &ti, int &tj) {
  // upper-triangular tiles enumerated row by row: row ti has (nt - ti) tiles
  int i = 0, rem = t;
  // closed form with a correction loop (nt is at most a few hundred)
  double disc = (2.0 * nt + 1.0) * (2.0 * nt + 1.0) - 8.0 * t;
  i = (int) ((2.0 * nt + 1.0 - sqrt(disc)) * 0.5);
  if (i < 0) i = 0;
  while (i > 0 && (i * (2 * nt - i + 1)) / 2 > t) --i;
  while (((i + 1) * (2 * nt - i)) / 2 <= t) ++i;
  rem = t - (i * (2 * nt - i + 1)) / 2;
  ti  = i;
  tj  = i + rem;
}

__global__ void __launch_bounds__(ATA_THREADS, 1)
ata_kernel(const double *__restrict__ P, int ldp, int K, int n, double *__restrict__ C, int ldc, double alpha, double beta, int nt) {
  extern __shared__ __align__(16) double smem[];
  int ti, tj;
  tile_from_linear(blockIdx.x, nt, ti, tj);
  if (threadIdx.x == 0 && blockIdx.x == 0) g_aprobe[g_api++] = clock64();  // start
  const int i0 = ti * BM, j0 = tj * BN;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;           // 2 x 4 warps
  const int lr = lane & 3, lc = lane >> 2;           // fragment coordinates

  auto load_stage = [&](int kb, int st) {
    double *sA = smem + (size_t) st * 2 * SLAB;
    double *sB = sA + SLAB;
    const int r0 = kb * BK;
    // each operand slab: BK rows x 128 doubles = 16 x 64 chunks of 16 B; 256 threads -> 4 chunks each per operand
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int chunk = tid + it * ATA_THREADS;      // 0 .. 1023
      const int r     = chunk >> 6;                  // 0 .. 15
      const int cc    = (chunk & 63) * 2;            // column (doubles) within the slab
      const int gr    = r0 + r;
      const bool rv   = gr < K;
      {
        const int gc   = i0 + cc;
        int bytes      = rv ? (n - gc) * 8 : 0;
        bytes          = bytes < 0 ? 0 : (bytes > 16 ? 16 : bytes);
        const double *src = (bytes > 0) ? (P + (size_t) gr * ldp + gc) : P;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(sA + r * PITCH + cc)), "l"(src), "r"(bytes) : "memory");
      }
      {
        const int gc   = j0 + cc;
        int bytes      = rv ? (n - gc) * 8 : 0;
        bytes          = bytes < 0 ? 0 : (bytes > 16 ? 16 : bytes);
        const double *src = (bytes > 0) ? (P + (size_t) gr * ldp + gc) : P;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(sB + r * PITCH + cc)), "l"(src), "r"(bytes) : "memory");
      }
    }
  };

  double acc[8][4][2];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

  const int nkb = (K + BK - 1) / BK;
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nkb) load_stage(s, s);
    cp_async_commit();
  }

  if (threadIdx.x == 0 && blockIdx.x == 0) g_aprobe[g_api++] = clock64();  // prologue_issued
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
      double af[8], bf[4];
      const double *pa = sA + (ks * 4 + lr) * PITCH + wm * 64 + lc;
      const double *pb = sB + (ks * 4 + lr) * PITCH + wn * 32 + lc;
#pragma unroll
      for (int a = 0; a < 8; ++a) af[a] = pa[a * 8];
#pragma unroll
      for (int b = 0; b < 4; ++b) bf[b] = pb[b * 8];
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) dmma884(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
    }
  }
  cp_async_wait<0>();
  if (threadIdx.x == 0 && blockIdx.x == 0) g_aprobe[g_api++] = clock64();  // mainloop_done

  // epilogue: lane holds C[row = lc][cols 2 lr, 2 lr + 1] of each 8 x 8 tile
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const int gi = i0 + wm * 64 + a * 8 + lc;
    if (gi >= n) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int gj = j0 + wn * 32 + b * 8 + 2 * lr;
      double *pc   = C + (size_t) gi * ldc + gj;
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
  if (threadIdx.x == 0 && blockIdx.x == 0) g_aprobe[g_api++] = clock64();  // end
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

int dsyrk_ata_general(ncm_sd_gpu_ctx *c, int K, int n, const double *dP, int ldp, double *dC, int ldc, double alpha, double beta) {
  if (n <= 0 || K <= 0) return NCM_SD_GPU_OK;
  if ((ldp & 1) || (ldc & 1) || (((uintptr_t) dP) & 15) || (((uintptr_t) dC) & 15))
    return c->fail(NCM_SD_GPU_EINVAL, "ata: operands must be 16-byte aligned with even leading dimensions");
  static bool attr_set = false;
  if (!attr_set) {
    NCM_CUDA_OK(c, cudaFuncSetAttribute(ata_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) ATA_SMEM));
    attr_set = true;
  }
  const int nt     = (n + BM - 1) / BM;
  const int ntiles = nt * (nt + 1) / 2;
  ata_kernel<<<ntiles, ATA_THREADS, ATA_SMEM, c->stream>>>(dP, ldp, K, n, dC, ldc, alpha, beta, nt);
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
