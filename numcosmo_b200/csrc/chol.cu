// Blocked Cholesky solve  M x = b  for the NNLS passive-set systems.
//
// Replaces ncm_matrix_cholesky_solve (dposv 'U', ncm_matrix.c:1199-1210) as called from
// _ncm_nnls_solve_normal_cholesky (ncm_nnls.c:655-666): M = U^T U with U upper triangular,
// row-major, only the upper triangle of M is read.  Right-looking by block rows of NB = 64:
//   1. chol_diag_kernel   factor the 64 x 64 diagonal block (one CTA), carrying the right-hand
//                         side along as a 65th column  (y_k = U_kk^-T b_k)
//   2. chol_panel_kernel  U[k, k+1:] = U_kk^-T M[k, k+1:], one thread per column, forward
//                         substitution in registers, U_kk broadcast from shared memory;
//                         also b[j] -= sum_r U[r][j] y_r  (forward solve folded in)
//   3. ata_kernel         trailing update M[k+1:, k+1:] -= U[k, k+1:]^T U[k, k+1:]  on DMMA
// followed by a blocked back substitution U x = y (one launch per block row).
#include "ctx.h"

int dsyrk_ata_general(ncm_sd_gpu_ctx *c, int K, int n, const double *dP, int ldp, double *dC, int ldc, double alpha, double beta);

namespace {

constexpr int NB = 64;
constexpr int DP1 = NB + 2;   // smem pitch of the diagonal block (+ rhs column), even

// M[k0:k0+nb, k0:k0+nb] -> U_kk in place; dinv[k0 + r] = 1 / U_rr; rhs[k0:k0+nb] -> y_k; info = first bad pivot (1-based)
__global__ void __launch_bounds__(256) chol_diag_kernel(double *__restrict__ M, int ldm, int n, int k0, double *__restrict__ rhs,
                                                        double *__restrict__ dinv, int *__restrict__ info) {
  __shared__ double s[NB][DP1];
  __shared__ int bad;
  const int nb  = min(NB, n - k0);
  const int tid = threadIdx.x;
  if (tid == 0) bad = 0;
  for (int e = tid; e < NB * (NB + 1); e += blockDim.x) {
    const int r = e / (NB + 1), cidx = e % (NB + 1);
    double v;
    if (cidx == NB)
      v = (r < nb && rhs != nullptr) ? rhs[k0 + r] : 0.0;
    else if (r < nb && cidx < nb)
      v = (cidx >= r) ? M[(size_t) (k0 + r) * ldm + k0 + cidx] : 0.0;
    else
      v = (r == cidx) ? 1.0 : 0.0;
    s[r][cidx] = v;
  }
  __syncthreads();
  for (int j = 0; j < nb; ++j) {
    if (tid == 0) {
      const double dd = s[j][j];
      if (!(dd > 0.0)) {
        if (bad == 0) bad = k0 + j + 1;
        s[j][j] = 1.0;   // keep going with a harmless pivot; the caller discards the result
      } else {
        s[j][j] = sqrt(dd);
      }
    }
    __syncthreads();
    const double piv = s[j][j];
    // scale row j (columns j+1 .. nb-1 and the rhs column NB)
    for (int cidx = j + 1 + tid; cidx <= NB; cidx += blockDim.x)
      if (cidx < nb || cidx == NB) s[j][cidx] = s[j][cidx] / piv;
    __syncthreads();
    // rank-1 update of the trailing upper triangle and of the rhs column
    const int rem = nb - j - 1;
    for (int e = tid; e < rem * (rem + 1); e += blockDim.x) {
      // e -> (r, c) with r in [0, rem), c in [r, rem]  where c == rem denotes the rhs column
      const int r = e / (rem + 1), cc = e % (rem + 1);
      if (cc < r) continue;
      const int gr = j + 1 + r;
      const int gc = (cc == rem) ? NB : j + 1 + cc;
      s[gr][gc]    = fma(-s[j][gr], s[j][gc], s[gr][gc]);
    }
    __syncthreads();
  }
  for (int e = tid; e < nb * nb; e += blockDim.x) {
    const int r = e / nb, cidx = e % nb;
    if (cidx >= r) M[(size_t) (k0 + r) * ldm + k0 + cidx] = s[r][cidx];
  }
  if (tid < nb) {
    dinv[k0 + tid] = 1.0 / s[tid][tid];
    if (rhs != nullptr) rhs[k0 + tid] = s[tid][NB];
  }
  if (tid == 0 && bad != 0 && atomicCAS(info, 0, bad) == 0) {}
}

// one thread per trailing column j: x = U_kk^-T M[k0:k0+nb, j]; rhs[j] -= x . y_k
__global__ void __launch_bounds__(128) chol_panel_kernel(double *__restrict__ M, int ldm, int n, int k0, double *__restrict__ rhs,
                                                         const double *__restrict__ dinv) {
  __shared__ double sU[NB * (NB + 1) / 2];   // packed: column r of U_kk (rows s < r) at r (r - 1) / 2 + s
  __shared__ double sD[NB];
  __shared__ double sY[NB];
  const int tid = threadIdx.x;
  for (int e = tid; e < NB * (NB - 1) / 2; e += blockDim.x) {
    int r = (int) ((1.0 + sqrt(1.0 + 8.0 * e)) * 0.5);
    while (r * (r - 1) / 2 > e) --r;
    while ((r + 1) * r / 2 <= e) ++r;
    const int sidx = e - r * (r - 1) / 2;
    sU[e]          = M[(size_t) (k0 + sidx) * ldm + k0 + r];
  }
  if (tid < NB) {
    sD[tid] = dinv[k0 + tid];
    sY[tid] = rhs != nullptr ? rhs[k0 + tid] : 0.0;
  }
  __syncthreads();
  const int j = k0 + NB + blockIdx.x * blockDim.x + tid;
  if (j >= n) return;
  double x[NB];
  double dot = 0.0;
#pragma unroll
  for (int r = 0; r < NB; ++r) {
    double t = M[(size_t) (k0 + r) * ldm + j];
#pragma unroll
    for (int sidx = 0; sidx < r; ++sidx) t = fma(-sU[r * (r - 1) / 2 + sidx], x[sidx], t);
    x[r] = t * sD[r];
    M[(size_t) (k0 + r) * ldm + j] = x[r];
    dot = fma(x[r], sY[r], dot);
  }
  if (rhs != nullptr) rhs[j] -= dot;
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
  if (tid < NB) sx[tid] = (tid < nb1) ? y[k1 + tid] : 0.0;
  __syncthreads();
  if (blockIdx.x > 0) {
    if (nb1 == 0) return;
    // 8 warps, one row each per pass
    const int row = (blockIdx.x - 1) * 8 + warp;
    if (row < k0) {
      const double *u = M + (size_t) row * ldm + k1;
      double sacc     = (lane < nb1 ? u[lane] : 0.0) * sx[lane];
      sacc            = fma(lane + 32 < nb1 ? u[lane + 32] : 0.0, sx[lane + 32], sacc);
      for (int off = 16; off > 0; off >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, off);
      if (lane == 0) y[row] -= sacc;
    }
    return;
  }
  // CTA 0: update own rows, then triangular solve
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
  // column-oriented back substitution on the nb x nb upper block
  for (int r = nb - 1; r >= 0; --r) {
    if (tid == 0) sy[r] = sy[r] * dinv[k0 + r];
    __syncthreads();
    if (tid < r) sy[tid] = fma(-M[(size_t) (k0 + tid) * ldm + k0 + r], sy[r], sy[tid]);
    __syncthreads();
  }
  if (tid < nb) y[k0 + tid] = sy[tid];
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
      const int k0   = kb * NB;
      const int nctas = 1 + (k0 + 7) / 8;
      chol_backsolve_kernel<<<nctas, 256, 0, c->stream>>>(dM, ldm, n, k0, dRhs, dDinv);
      c->n_launches++;
    }
  }
  NCM_CUDA_OK(c, cudaGetLastError());
  if (info_host != nullptr) {
    NCM_CUDA_OK(c, ncm_memcpy_async(c,info_host, dInfo, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    NCM_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  }
  return NCM_SD_GPU_OK;
}
