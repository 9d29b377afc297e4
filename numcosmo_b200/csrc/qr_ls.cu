// Least-squares solve min |A[:, P] x - f| by Householder QR: the reference's last resort when dsysv meets an exactly singular
// pivot block.
//
// Replaces ncm_lapack_dgels ('N', nrows, |P|, 1, ...) as called from _ncm_nnls_solve_normal_QR (ncm_nnls.c:608-638, the gathered
// columns come from _ncm_nnls_prepare_usys_QR, :520-542): dgeqrf + dormqr + dtrtrs.  The reflectors are LAPACK's dlarfg ones
// (beta = -sign (alpha) hypot (alpha, |x|), tau = (beta - alpha) / beta, v = x / (alpha - beta)), applied column by column as the
// unblocked dgeqr2 does, to the right-hand side as well; then R x = (Q^T f)[0 : n] by back substitution.  info > 0 (an exactly zero
// diagonal entry of R, dtrtrs' check): nothing is solved, as dgels.
//
// One cooperative grid (one CTA per SM at most), the gathered matrix column-major in global memory: the norm of the pivot column is
// taken redundantly in every CTA, the trailing columns (one warp each, the right-hand side is column n) are reflected across the
// grid, one grid barrier per column.
#include <algorithm>
#include "ctx.h"
#include "coop.cuh"

namespace {

using namespace ncm_coop;

// Q[j * m + i] = A[i * lda + idx[j]]  (column-major gather of the passive columns); Q[n * m + i] = f[i]
__global__ void qr_gather_kernel(const double *__restrict__ A, int lda, int m, const int *__restrict__ idx, int n, const double *__restrict__ f,
                                 double *__restrict__ Q) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i >= m) return;
  Q[(size_t) j * m + i] = j < n ? A[(size_t) i * lda + idx[j]] : f[i];
}

__global__ void __launch_bounds__(COOP_T, 1) qr_ls_kernel(double *Q, int m, int n, double *x, int *info, unsigned int *bar) {
  __shared__ double red[COOP_T / 32];
  GridSync sync{bar, 0u, gridDim.x};
  const int tid = threadIdx.x, lane = tid & 31;
  const int gwarp = (blockIdx.x * COOP_T + tid) >> 5, gwarps = (gridDim.x * COOP_T) >> 5;
  for (int k = 0; k < n; ++k) {
    double *ck = Q + (size_t) k * m;
    // dlarfg (m - k, alpha = ck[k], x = ck[k + 1 :]) -- every CTA alike; the owner writes the reflector after the others have used it
    double ss = 0.0;
    for (int i = k + 1 + tid; i < m; i += COOP_T) ss = fma(ck[i], ck[i], ss);
    ss = cta_sum(ss, red);
    const double alpha = ck[k], xnorm = sqrt(ss);
    double tau = 0.0, scal = 0.0, beta = alpha;
    if (xnorm != 0.0) {
      beta = -copysign(hypot(alpha, xnorm), alpha);
      tau  = (beta - alpha) / beta;
      scal = 1.0 / (alpha - beta);
    }
    // H = I - tau v v^T with v = [1; scal * x]: applied to the trailing columns and to the right-hand side (column n)
    if (tau != 0.0) {
      for (int j = k + 1 + gwarp; j <= n; j += gwarps) {
        double *cj = Q + (size_t) j * m;
        double w   = 0.0;
        for (int i = k + 1 + lane; i < m; i += 32) w = fma(ck[i] * scal, cj[i], w);
        for (int off = 16; off > 0; off >>= 1) w += __shfl_xor_sync(0xffffffffu, w, off);
        w += cj[k];
        const double tw = tau * w;
        for (int i = k + 1 + lane; i < m; i += 32) cj[i] = fma(-tw, ck[i] * scal, cj[i]);
        __syncwarp();
        if (lane == 0) cj[k] -= tw;
      }
    }
    sync();
    if (blockIdx.x == 0 && tid == 0) ck[k] = beta;   // R(k,k); the part below the diagonal is not needed again
  }
  sync();
  if (blockIdx.x != 0) return;
  // dtrtrs: exact zero on the diagonal of R -> info, no solve
  __shared__ int s_info;
  if (tid == 0) {
    int bad = 0;
    for (int k = 0; k < n && bad == 0; ++k)
      if (Q[(size_t) k * m + k] == 0.0) bad = k + 1;
    s_info = bad;
    *info  = bad;
  }
  __syncthreads();
  if (s_info != 0) return;
  // R x = c, c = first n entries of the reflected right-hand side (column n); column-oriented back substitution (dtrsv 'U' 'N')
  double *c = Q + (size_t) n * m;
  for (int k = n - 1; k >= 0; --k) {
    const double *ck = Q + (size_t) k * m;
    if (tid == 0) c[k] = c[k] / ck[k];
    __syncthreads();
    const double xk = c[k];
    for (int i = tid; i < k; i += COOP_T) c[i] = fma(-xk, ck[i], c[i]);
    __syncthreads();
  }
  for (int i = tid; i < n; i += COOP_T) x[i] = c[i];
}

}   // namespace

// x = argmin |A[:, idx] x - f| (A: m x lda row-major on the device, idx: n ascending column indices on the device); dX [n].
// info_host: 0, or the 1-based index of an exactly zero diagonal entry of R.
int dgels_cols_solve(ncm_sd_gpu_ctx *c, int m, int n, const double *dA, int lda, const int *dIdx, const double *dF, double *dX, int *info_host) {
  if (m < n) return c->fail(NCM_SD_GPU_EINVAL, "dgels: fewer rows than passive columns");
  if (!c->qrWork.reserve(((size_t) m * (n + 1) + 64) * sizeof(double))) return c->fail(NCM_SD_GPU_ENOMEM, "dgels: out of device memory");
  double *Q         = c->qrWork.as<double>() + 8;
  int *info         = c->qrWork.as<int>();
  unsigned int *bar = reinterpret_cast<unsigned int *>(c->qrWork.as<int>() + 4);
  cudaStream_t st   = c->stream;
  NCM_CUDA_OK(c, cudaMemsetAsync(c->qrWork.p, 0, 64, st));
  qr_gather_kernel<<<dim3((m + 255) / 256, n + 1), 256, 0, st>>>(dA, lda, m, dIdx, n, dF, Q);
  int nctas = std::min(c->n_sm, std::max(1, (n + COOP_T / 32) / (COOP_T / 32)));
  void *params[] = {(void *) &Q, (void *) &m, (void *) &n, (void *) &dX, (void *) &info, (void *) &bar};
  NCM_CUDA_OK(c, cudaLaunchCooperativeKernel((const void *) qr_ls_kernel, dim3(nctas), dim3(COOP_T), params, 0, st));
  c->n_launches += 2;
  NCM_CUDA_OK(c, cudaGetLastError());
  int h_info = 0;
  NCM_CUDA_OK(c, ncm_memcpy_async(c, &h_info, info, sizeof(int), cudaMemcpyDeviceToHost, st));
  NCM_CUDA_OK(c, cudaStreamSynchronize(st));
  if (info_host) *info_host = h_info;
  return NCM_SD_GPU_OK;
}
