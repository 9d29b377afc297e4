// Blocked Cholesky solve  M x = b  for the NNLS passive-set systems.
//
// Replaces ncm_matrix_cholesky_solve (dposv 'U', ncm_matrix.c:1199-1210) as called from
// _ncm_nnls_solve_normal_cholesky (ncm_nnls.c:655-666): M = U^T U with U upper triangular,
// row-major, only the upper triangle of M is read.  Right-looking by block rows of NB = 64:
//   1. chol_diag_kernel   factor the 64 x 64 diagonal block in registers (one CTA of 16 x 16 threads,
//                         4 x 4 cyclic register tile each, one barrier per column), carrying the
//                         right-hand side along as a 65th column  (y_k = U_kk^-T b_k)
//   2. chol_panel_kernel  U[k, k+1:] = U_kk^-T M[k, k+1:], one thread per column, forward
//                         substitution in registers in 16-row sub-blocks (16 independent
//                         accumulators while sweeping the solved part), U_kk broadcast from
//                         shared memory; also b[j] -= sum_r U[r][j] y_r  (forward solve folded in)
//   3. ata_kernel         trailing update M[k+1:, k+1:] -= U[k, k+1:]^T U[k, k+1:]  on DMMA
// followed by a blocked back substitution U x = y (one launch per block row; the 64 x 64
// triangular solve is done by a single warp with shuffles, the rows above by the other CTAs).
#include "ctx.h"

int dsyrk_ata_general(ncm_sd_gpu_ctx *c, int K, int n, const double *dP, int ldp, double *dC, int ldc, double alpha, double beta);

namespace {

__device__ long long g_pprobe[64];
__device__ int g_ppi;
constexpr int NB = 64;

// M[k0:k0+nb, k0:k0+nb] -> U_kk in place; dinv[k0 + r] = 1 / U_rr; rhs[k0:k0+nb] -> y_k;
// info = first non-positive pivot (1-based), 0 otherwise.
//
// The 64 x 64 block (plus the right-hand side as column 64) lives in shared memory and is processed in
// four 16-column sub-blocks:  A) warp 0 factors the 16 x 16 diagonal sub-block in registers, lane c
// holding column c, pivots and scaled rows exchanged with shuffles (the serial chain is
// shuffle -> mul -> fma -> rsqrt per pivot);  B) every remaining column (and the rhs) is solved
// against it, one thread per column;  C) the trailing sub-block gets its rank-16 update.
#ifdef CHOL_PROBE
__device__ long long g_probe[64];
__device__ int g_pi;
#endif
constexpr int DS = NB + 2;   // row pitch of the shared block: columns 0..63 + rhs column 64 (+ pad)

__global__ void __launch_bounds__(256) chol_diag_kernel(double *__restrict__ M, int ldm, int n, int k0, double *__restrict__ rhs,
                                                        double *__restrict__ dinv, int *__restrict__ info) {
  __shared__ double S[NB][DS];
  __shared__ double sDinv[NB];
  __shared__ int sBad;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nb = min(NB, n - k0);
  if (tid == 0) sBad = 0;
  for (int e = tid; e < NB * (NB + 1); e += blockDim.x) {
    const int r = e / (NB + 1), cidx = e - r * (NB + 1);
    double v;
    if (cidx == NB)
      v = (r < nb && rhs != nullptr) ? rhs[k0 + r] : 0.0;
    else if (r < nb && cidx < nb)
      v = (cidx >= r) ? M[(size_t) (k0 + r) * ldm + k0 + cidx] : 0.0;
    else
      v = (r == cidx) ? 1.0 : 0.0;
    S[r][cidx] = v;
  }
  __syncthreads();
#ifdef CHOL_PROBE
    if (tid == 0) g_probe[g_pi++] = clock64();   // load
#endif

#pragma unroll 1
  for (int kb = 0; kb < NB / 16; ++kb) {
    const int b0 = 16 * kb;
    // ---- A: 16 x 16 diagonal sub-block, warp 0, lane c < 16 = column b0 + c ----
    if (warp == 0) {
      double col[16];
      const int c = lane;
#pragma unroll
      for (int r = 0; r < 16; ++r) col[r] = (c < 16 && r <= c) ? S[b0 + r][b0 + c] : 0.0;
      int bad = 0;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        double piv = __shfl_sync(0xffffffffu, col[j], j);
        if (!(piv > 0.0)) {
          if (bad == 0) bad = k0 + b0 + j + 1;
          piv = 1.0;
        }
        const double inv = rsqrt(piv);
        const double u   = col[j] * inv;           // lane c: U[j][c] (c > j); lane j: sqrt(piv)
        col[j]           = u;
        if (c == j) sDinv[b0 + j] = inv;
#pragma unroll
        for (int r = j + 1; r < 16; ++r) {
          const double ur = __shfl_sync(0xffffffffu, u, r);   // U[j][r]
          if (r <= c) col[r] = fma(-ur, u, col[r]);
        }
      }
      if (c < 16) {
#pragma unroll
        for (int r = 0; r < 16; ++r) S[b0 + r][b0 + c] = col[r];   // rows r > c hold zeros (unused lower part)
      }
      if (lane == 0 && bad != 0 && sBad == 0) sBad = bad;
    }
    __syncthreads();
#ifdef CHOL_PROBE
    if (tid == 0) g_probe[g_pi++] = clock64();   // A
#endif
    // ---- B: row panel  X = D^-T S[b0:b0+16, c]  for c = b0+16 .. 63 and the rhs column ----
    {
      const int ncols = NB - (b0 + 16) + 1;   // + rhs
      if (tid < ncols) {
        const int c = (tid == ncols - 1) ? NB : b0 + 16 + tid;
        double x[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          double t = S[b0 + r][c];
#pragma unroll
          for (int s = 0; s < r; ++s) t = fma(-S[b0 + s][b0 + r], x[s], t);
          x[r] = t * sDinv[b0 + r];
        }
#pragma unroll
        for (int r = 0; r < 16; ++r) S[b0 + r][c] = x[r];
      }
    }
    __syncthreads();
#ifdef CHOL_PROBE
    if (tid == 0) g_probe[g_pi++] = clock64();   // B
#endif
    // ---- C: trailing update  S[r][c] -= sum_s X[s][r] X[s][c],  b0+16 <= r <= c < 64, and the rhs column ----
    {
      const int m  = NB - (b0 + 16);   // remaining rows
      const int nc = m + 1;            // remaining columns + rhs
      for (int e = tid; e < m * nc; e += blockDim.x) {
        const int rr = e / nc, cc = e - rr * nc;
        if (cc < rr) continue;
        const int r = b0 + 16 + rr;
        const int c = (cc == m) ? NB : b0 + 16 + cc;
        double acc = S[r][c];
#pragma unroll
        for (int s = 0; s < 16; ++s) acc = fma(-S[b0 + s][r], S[b0 + s][c], acc);
        S[r][c] = acc;
      }
    }
    __syncthreads();
#ifdef CHOL_PROBE
    if (tid == 0) g_probe[g_pi++] = clock64();   // C
#endif
  }

  for (int e = tid; e < nb * nb; e += blockDim.x) {
    const int r = e / nb, cidx = e - r * nb;
    if (cidx >= r) M[(size_t) (k0 + r) * ldm + k0 + cidx] = S[r][cidx];
  }
  if (tid < nb) {
    dinv[k0 + tid] = sDinv[tid];
    if (rhs != nullptr) rhs[k0 + tid] = S[tid][NB];
  }
  if (tid == 0 && sBad != 0) atomicCAS(info, 0, sBad);
}

// one thread per trailing column j: x = U_kk^-T M[k0:k0+nb, j]; rhs[j] -= x . y_k
__global__ void __launch_bounds__(128) chol_panel_kernel(double *__restrict__ M, int ldm, int n, int k0, double *__restrict__ rhs,
                                                         const double *__restrict__ dinv) {
  __shared__ __align__(16) double sU[NB][NB];   // sU[s][r] = U_kk[s][r] (upper)
  __shared__ double sD[NB];
  __shared__ double sY[NB];
  const int tid = threadIdx.x;
  if (threadIdx.x == 0 && blockIdx.x == 0) g_pprobe[g_ppi++] = clock64();  // pstart
  for (int e = tid; e < NB * NB; e += blockDim.x) {
    const int s = e >> 6, r = e & 63;
    sU[s][r]    = (r >= s) ? M[(size_t) (k0 + s) * ldm + k0 + r] : 0.0;
  }
  if (tid < NB) {
    sD[tid] = dinv[k0 + tid];
    sY[tid] = rhs != nullptr ? rhs[k0 + tid] : 0.0;
  }
  __syncthreads();
  const int j = k0 + NB + blockIdx.x * blockDim.x + tid;
  if (j >= n) return;
  double x[NB];
  if (threadIdx.x == 0 && blockIdx.x == 0) g_pprobe[g_ppi++] = clock64();  // ploaded
  double dot = 0.0;
#pragma unroll
  for (int blk = 0; blk < NB / 16; ++blk) {
    double t[16];
#pragma unroll
    for (int rr = 0; rr < 16; ++rr) t[rr] = M[(size_t) (k0 + 16 * blk + rr) * ldm + j];
    // sweep the already solved unknowns: 16 independent accumulators
#pragma unroll
    for (int s = 0; s < 16 * blk; ++s) {
#pragma unroll
      for (int rr = 0; rr < 16; ++rr) t[rr] = fma(-sU[s][16 * blk + rr], x[s], t[rr]);
    }
    // triangular part of the sub-block
#pragma unroll
    for (int rr = 0; rr < 16; ++rr) {
      const int r = 16 * blk + rr;
#pragma unroll
      for (int ss = 0; ss < rr; ++ss) t[rr] = fma(-sU[16 * blk + ss][r], x[16 * blk + ss], t[rr]);
      x[r] = t[rr] * sD[r];
      M[(size_t) (k0 + r) * ldm + j] = x[r];
      dot = fma(x[r], sY[r], dot);
    }
  }
  if (rhs != nullptr) rhs[j] -= dot;
  if (threadIdx.x == 0 && blockIdx.x == 0) g_pprobe[g_ppi++] = clock64();  // pdone
}

// Back substitution step for block row kb (k0 = kb * NB), given x of block kb + 1 already in y:
//   CTA 0      : y[k0:k0+nb] -= U[k0:k0+nb, k1:k1+nb1] x[k1:k1+nb1]; then solve U_kk x_k = y_k in place
//   CTAs 1..   : rows above k0: y[r] -= U[r, k1:k1+nb1] x[k1:k1+nb1]
__global__ void __launch_bounds__(256) chol_backsolve_kernel(const double *__restrict__ M, int ldm, int n, int k0, double *__restrict__ y,
                                                             const double *__restrict__ dinv) {
  const int k1   = k0 + NB;
  const int nb1  = max(0, min(NB, n - k1));
  const int nb   = min(NB, n - k0);
  const int tid  = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ double sx[NB];
  __shared__ double sy[NB];
  __shared__ double sU[NB][NB + 1];
  if (tid < NB) sx[tid] = (tid < nb1) ? y[k1 + tid] : 0.0;
  if (blockIdx.x == 0) {
    for (int e = tid; e < NB * NB; e += blockDim.x) {
      const int r = e >> 6, cidx = e & 63;
      sU[r][cidx] = (r < nb && cidx < nb && cidx >= r) ? M[(size_t) (k0 + r) * ldm + k0 + cidx] : 0.0;
    }
  }
  __syncthreads();
  if (blockIdx.x > 0) {
    if (nb1 == 0) return;
    const int row = (blockIdx.x - 1) * 8 + warp;   // 8 warps, one row each
    if (row < k0) {
      const double *u = M + (size_t) row * ldm + k1;
      double sacc     = (lane < nb1 ? u[lane] : 0.0) * sx[lane];
      sacc            = fma(lane + 32 < nb1 ? u[lane + 32] : 0.0, sx[lane + 32], sacc);
      for (int off = 16; off > 0; off >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, off);
      if (lane == 0) y[row] -= sacc;
    }
    return;
  }
  // CTA 0: update own rows with x of the next block ...
  for (int r = warp; r < nb; r += 8) {
    double sacc = 0.0;
    if (nb1 > 0) {
      const double *u = M + (size_t) (k0 + r) * ldm + k1;
      sacc            = (lane < nb1 ? u[lane] : 0.0) * sx[lane];
      sacc            = fma(lane + 32 < nb1 ? u[lane + 32] : 0.0, sx[lane + 32], sacc);
      for (int off = 16; off > 0; off >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, off);
    }
    if (lane == 0) sy[r] = y[k0 + r] - sacc;
  }
  __syncthreads();
  // ... then one warp solves the 64 x 64 upper system: lane holds rows lane and lane + 32
  if (warp == 0) {
    double y0 = lane < nb ? sy[lane] : 0.0, y1 = lane + 32 < nb ? sy[lane + 32] : 0.0;
    const double d0 = lane < nb ? dinv[k0 + lane] : 1.0, d1 = lane + 32 < nb ? dinv[k0 + lane + 32] : 1.0;
    for (int r = NB - 1; r >= 32; --r) {
      const double xr = __shfl_sync(0xffffffffu, y1 * d1, r - 32);
      if (lane + 32 == r) y1 = xr;
      if (lane + 32 < r) y1 = fma(-sU[lane + 32][r], xr, y1);
      y0 = fma(-sU[lane][r], xr, y0);
    }
    for (int r = 31; r >= 0; --r) {
      const double xr = __shfl_sync(0xffffffffu, y0 * d0, r);
      if (lane == r) y0 = xr;
      if (lane < r) y0 = fma(-sU[lane][r], xr, y0);
    }
    if (lane < nb) y[k0 + lane] = y0;
    if (lane + 32 < nb) y[k0 + lane + 32] = y1;
  }
}

}   // namespace

// In-place blocked Cholesky of the upper triangle of dM (n x n, ld = ldm, even) and, when dRhs != nullptr,
// solution of M x = rhs in place.  dinv: scratch of n doubles.  info_host: 0, or 1-based index of the
// first non-positive pivot (result then undefined).
int dpotrf_upper_solve(ncm_sd_gpu_ctx *c, int n, double *dM, int ldm, double *dRhs, double *dDinv, int *dInfo, int *info_host) {
  NCM_CUDA_OK(c, cudaMemsetAsync(dInfo, 0, sizeof(int), c->stream));
  const int nblk = (n + NB - 1) / NB;
  for (int kb = 0; kb < nblk; ++kb) {
    const int k0 = kb * NB;
    chol_diag_kernel<<<1, 256, 0, c->stream>>>(dM, ldm, n, k0, dRhs, dDinv, dInfo);
    c->n_launches++;
    const int m = n - k0 - NB;
    if (m > 0) {
      chol_panel_kernel<<<(m + 127) / 128, 128, 0, c->stream>>>(dM, ldm, n, k0, dRhs, dDinv);
      c->n_launches++;
      int rc = dsyrk_ata_general(c, NB, m, dM + (size_t) k0 * ldm + k0 + NB, ldm, dM + (size_t) (k0 + NB) * ldm + k0 + NB, ldm, -1.0, 1.0);
      if (rc != NCM_SD_GPU_OK) return rc;
    }
  }
  if (dRhs != nullptr) {
    for (int kb = nblk - 1; kb >= 0; --kb) {
      const int k0    = kb * NB;
      const int nctas = 1 + (k0 + 7) / 8;
      chol_backsolve_kernel<<<nctas, 256, 0, c->stream>>>(dM, ldm, n, k0, dRhs, dDinv);
      c->n_launches++;
    }
  }
  NCM_CUDA_OK(c, cudaGetLastError());
  if (info_host != nullptr) {
    NCM_CUDA_OK(c, ncm_memcpy_async(c, info_host, dInfo, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    NCM_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  }
  return NCM_SD_GPU_OK;
}
