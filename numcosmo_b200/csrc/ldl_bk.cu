// Symmetric-indefinite solve of a passive-set system: Bunch-Kaufman L D L^T with partial (diagonal) pivoting, then the
// triangular / block-diagonal solves -- the reference's fallback when dposv reports a non-positive pivot.
//
// Replaces ncm_lapack_dsysv ('U' row-major == 'L' column-major, ncm_lapack.c:58,798) as called from _ncm_nnls_solve_normal_LU
// (ncm_nnls.c:573-606): LAPACK dsytrf + dsytrs.  The pivoting rule, the 1 x 1 / 2 x 2 block choice (alpha = (1 + sqrt 17) / 8), the
// interchange pattern and the order of the solves are those of the unblocked dsytf2 / dsytrs 'L' variants, so that in exact arithmetic
// (and in floating point whenever no comparison is decided by rounding) the pivot sequence equals LAPACK's.  The kernel Gram matrices
// this path meets are singular to working precision (that is why dposv failed): their solutions are rounding-driven on the CPU as
// well, so agreement with a given LAPACK build is at the level of the residual, not of the solution (tests/test_gpu_ldl.py).
//
// Storage: the gathered system is row-major upper, S[c * ld + r] = A(r, c) for r >= c -- i.e. column c of the column-major lower
// triangle is contiguous.  Two kernels share the algorithm:
//   * n <= BK_SMEM_MAX_N: one CTA, the packed triangle and the right-hand side in shared memory;
//   * otherwise: a cooperative grid (one CTA per SM), matrix in global memory.  The pivot search runs redundantly in every CTA (the
//     column is at most n doubles), interchanges and the rank-1 / rank-2 trailing update are spread over the grid, two grid barriers
//     per pivot step; the (sequential) solves run in CTA 0.
#include <algorithm>
#include "ctx.h"
#include "coop.cuh"

namespace {

using namespace ncm_coop;
constexpr int BK_T = COOP_T;           // threads per CTA
constexpr int BK_SMEM_MAX_N = 232;     // packed triangle (n (n + 1) / 2 doubles) + vectors within 227 KB
constexpr double BK_ALPHA = 0.6403882032022076;   // (1 + sqrt (17)) / 8, as dsytf2 computes it

struct GlobalAcc {
  double *S;
  int ld;
  __device__ __forceinline__ double &operator()(int r, int c) const { return S[(size_t) c * ld + r]; }
};
struct PackedAcc {
  double *S;
  int n;
  __device__ __forceinline__ double &operator()(int r, int c) const { return S[c * n - (c * (c - 1)) / 2 + (r - c)]; }
};

// dsytf2 'L' + dsytrs 'L'.  ipiv: 0-based; 1 x 1 block at k: ipiv[k] = kp >= 0; 2 x 2 block at (k, k + 1): ipiv[k] = ipiv[k + 1] = -(kp + 1).
// info: 0, or 1 + the first index with an exactly zero (or NaN) pivot column (dsytf2's INFO; dsysv then does not solve).
template <class Acc>
__device__ void bk_factor_solve(Acc A, int n, double *b, int *ipiv, int *info_out, GridSync sync, char *red_raw) {
  AbsMax *red_m = reinterpret_cast<AbsMax *>(red_raw);
  double *red_d = reinterpret_cast<double *>(red_raw);
  const int tid = threadIdx.x;
  const int gtid = blockIdx.x * BK_T + tid, gthreads = sync.nctas * BK_T;
  const int gwarp = gtid >> 5, gwarps = gthreads >> 5, lane = tid & 31;
  int info = 0;
  int k = 0;
  while (k < n) {
    // ---- pivot choice (every CTA, redundantly and identically) ----
    const double akk = A(k, k), absakk = fabs(akk);
    AbsMax m = {0.0, -1};
    for (int i = k + 1 + tid; i < n; i += BK_T) {
      const double v = fabs(A(i, k));
      if (m.i < 0 || v > m.v) m = {v, i};
    }
    m = cta_absmax(m, red_m);
    const double colmax = m.i >= 0 ? m.v : 0.0;
    const int imax = m.i;
    int kp = k, kstep = 1;
    bool singular = false;
    if (!(fmax(absakk, colmax) > 0.0) || absakk != absakk) {   // zero column or NaN: dsytf2 sets INFO and moves on
      singular = true;
    } else if (!(absakk >= BK_ALPHA * colmax)) {
      // rowmax: largest off-diagonal of row / column imax within the trailing matrix
      AbsMax r = {0.0, -1};
      for (int j = k + tid; j < imax; j += BK_T) {
        const double v = fabs(A(imax, j));
        if (r.i < 0 || v > r.v) r = {v, j};
      }
      for (int i = imax + 1 + tid; i < n; i += BK_T) {
        const double v = fabs(A(i, imax));
        if (r.i < 0 || v > r.v) r = {v, i};
      }
      r = cta_absmax(r, red_m);
      const double rowmax = r.v;
      if (absakk >= BK_ALPHA * colmax * (colmax / rowmax)) {
        kp = k;
      } else if (fabs(A(imax, imax)) >= BK_ALPHA * rowmax) {
        kp = imax;
      } else {
        kp    = imax;
        kstep = 2;
      }
    }
    if (singular && info == 0) info = k + 1;
    const int kk = k + kstep - 1;
    // ---- interchange rows / columns kk <-> kp of the trailing matrix (spread over the grid) ----
    if (kp != kk) {
      sync();   // every CTA has taken its (identical) decision from the un-swapped matrix
      for (int i = kp + 1 + gtid; i < n; i += gthreads) {
        const double t = A(i, kk);
        A(i, kk)       = A(i, kp);
        A(i, kp)       = t;
      }
      for (int j = kk + 1 + gtid; j < kp; j += gthreads) {
        const double t = A(j, kk);
        A(j, kk)       = A(kp, j);
        A(kp, j)       = t;
      }
      if (gtid == 0) {
        const double t = A(kk, kk);
        A(kk, kk)      = A(kp, kp);
        A(kp, kp)      = t;
        if (kstep == 2) {
          const double u = A(k + 1, k);
          A(k + 1, k)    = A(kp, k);
          A(kp, k)       = u;
        }
      }
    }
    if (gtid == 0) {
      if (kstep == 1)
        ipiv[k] = kp;
      else
        ipiv[k] = ipiv[k + 1] = -(kp + 1);
    }
    sync();
    // ---- trailing update (one warp per column, lanes down the column) ----
    if (!singular) {
      if (kstep == 1) {
        if (k < n - 1) {
          const double r1 = 1.0 / A(k, k);
          for (int j = k + 1 + gwarp; j < n; j += gwarps) {
            const double temp = -r1 * A(j, k);   // dsyr: A(i,j) += x(i) * (alpha x(j))
            for (int i = j + lane; i < n; i += 32) A(i, j) = fma(A(i, k), temp, A(i, j));
          }
        }
      } else if (k < n - 2) {
        double d21       = A(k + 1, k);
        const double d11 = A(k + 1, k + 1) / d21, d22 = A(k, k) / d21;
        const double t   = 1.0 / (d11 * d22 - 1.0);
        d21              = t / d21;
        for (int j = k + 2 + gwarp; j < n; j += gwarps) {
          const double wk = d21 * (d11 * A(j, k) - A(j, k + 1)), wkp1 = d21 * (d22 * A(j, k + 1) - A(j, k));
          for (int i = j + lane; i < n; i += 32) A(i, j) = (A(i, j) - A(i, k) * wk) - A(i, k + 1) * wkp1;
        }
      }
    }
    sync();
    // ---- the multipliers replace the pivot column(s); nobody reads them again before the solve (CTA k mod nctas does it) ----
    if (!singular && blockIdx.x == (unsigned) (k % (int) sync.nctas)) {
      if (kstep == 1) {
        const double r1 = 1.0 / A(k, k);
        for (int i = k + 1 + tid; i < n; i += BK_T) A(i, k) *= r1;
      } else {
        double d21       = A(k + 1, k);
        const double d11 = A(k + 1, k + 1) / d21, d22 = A(k, k) / d21;
        const double t   = 1.0 / (d11 * d22 - 1.0);
        d21              = t / d21;
        for (int j = k + 2 + tid; j < n; j += BK_T) {
          const double ak = A(j, k), akp1 = A(j, k + 1);
          A(j, k)     = d21 * (d11 * ak - akp1);
          A(j, k + 1) = d21 * (d22 * akp1 - ak);
        }
      }
    }
    k += kstep;
  }
  sync();
  if (gtid == 0) *info_out = info;
  if (blockIdx.x != 0 || info != 0) return;
  // ---- dsytrs 'L': L D L^T x = b (CTA 0) ----
  k = 0;
  while (k < n) {
    const int p = ipiv[k];
    if (p >= 0) {
      if (tid == 0 && p != k) {
        const double t = b[k];
        b[k]           = b[p];
        b[p]           = t;
      }
      __syncthreads();
      const double bk = b[k];
      for (int i = k + 1 + tid; i < n; i += BK_T) b[i] = fma(-A(i, k), bk, b[i]);
      __syncthreads();   // every warp has read b[k] before it is overwritten (racecheck, r02y)
      if (tid == 0) b[k] = bk * (1.0 / A(k, k));   // dscal (1 / A(k,k))
      __syncthreads();
      k += 1;
    } else {
      const int kp = -p - 1;
      if (tid == 0 && kp != k + 1) {
        const double t = b[k + 1];
        b[k + 1]       = b[kp];
        b[kp]          = t;
      }
      __syncthreads();
      const double bk = b[k], bk1 = b[k + 1];
      for (int i = k + 2 + tid; i < n; i += BK_T) b[i] = fma(-A(i, k + 1), bk1, fma(-A(i, k), bk, b[i]));
      __syncthreads();   // b[k], b[k + 1] are read by every warp above
      if (tid == 0) {
        const double akm1k = A(k + 1, k), akm1 = A(k, k) / akm1k, ak = A(k + 1, k + 1) / akm1k, denom = akm1 * ak - 1.0;
        const double bkm1 = bk / akm1k, bkk = bk1 / akm1k;
        b[k]     = (ak * bkm1 - bkk) / denom;
        b[k + 1] = (akm1 * bkk - bkm1) / denom;
      }
      __syncthreads();
      k += 2;
    }
  }
  k = n - 1;
  while (k >= 0) {
    const int p = ipiv[k];
    if (p >= 0) {
      if (k < n - 1) {
        double s = 0.0;
        for (int i = k + 1 + tid; i < n; i += BK_T) s = fma(A(i, k), b[i], s);
        s = cta_sum(s, red_d);
        if (tid == 0) b[k] -= s;
      }
      __syncthreads();
      if (tid == 0 && p != k) {
        const double t = b[k];
        b[k]           = b[p];
        b[p]           = t;
      }
      __syncthreads();
      k -= 1;
    } else {
      const int kp = -p - 1;
      if (k < n - 1) {
        double s = 0.0, s1 = 0.0;
        for (int i = k + 1 + tid; i < n; i += BK_T) {
          const double bi = b[i];
          s  = fma(A(i, k), bi, s);
          s1 = fma(A(i, k - 1), bi, s1);
        }
        s  = cta_sum(s, red_d);
        s1 = cta_sum(s1, red_d);
        if (tid == 0) {
          b[k] -= s;
          b[k - 1] -= s1;
        }
      }
      __syncthreads();
      if (tid == 0 && kp != k) {
        const double t = b[k];
        b[k]           = b[kp];
        b[kp]          = t;
      }
      __syncthreads();
      k -= 2;
    }
  }
}

__global__ void __launch_bounds__(BK_T, 1) bk_global_kernel(double *S, int ld, int n, double *b, int *ipiv, int *info, unsigned int *bar) {
  __shared__ __align__(16) char red[BK_T / 32 * sizeof(AbsMax)];
  GridSync sync{bar, 0u, gridDim.x};
  bk_factor_solve(GlobalAcc{S, ld}, n, b, ipiv, info, sync, red);
}

__global__ void __launch_bounds__(BK_T, 1) bk_smem_kernel(const double *__restrict__ S, int ld, int n, double *b_g, int *ipiv_g, int *info) {
  extern __shared__ __align__(16) double sm[];
  __shared__ __align__(16) char red[BK_T / 32 * sizeof(AbsMax)];
  double *P = sm;                                   // packed lower triangle by columns
  double *b = sm + (size_t) n * (n + 1) / 2;
  int *ipiv = reinterpret_cast<int *>(b + n);
  PackedAcc A{P, n};
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = warp; c < n; c += BK_T / 32)
    for (int r = c + lane; r < n; r += 32) A(r, c) = S[(size_t) c * ld + r];
  for (int i = threadIdx.x; i < n; i += BK_T) b[i] = b_g[i];
  __syncthreads();
  GridSync sync{nullptr, 0u, 1u};
  bk_factor_solve(A, n, b, ipiv, info, sync, red);
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += BK_T) {
    b_g[i]    = b[i];
    ipiv_g[i] = ipiv[i];
  }
}

}   // namespace

// Solve S x = rhs for the symmetric (possibly indefinite) matrix held in the upper triangle of the row-major dS; dS is destroyed,
// dRhs is overwritten by x.  info_host: 0, or dsytf2's INFO (1-based index of an exactly singular pivot; nothing is solved then).
int dsysv_upper_solve(ncm_sd_gpu_ctx *c, int n, double *dS, int lds, double *dRhs, int *info_host) {
  if (n <= 0) {
    if (info_host) *info_host = 0;
    return NCM_SD_GPU_OK;
  }
  if (!c->bkWork.reserve((size_t) (n + 64) * sizeof(int))) return c->fail(NCM_SD_GPU_ENOMEM, "dsysv: out of device memory");
  int *ipiv          = c->bkWork.as<int>() + 16;
  int *info          = c->bkWork.as<int>();
  unsigned int *bar  = reinterpret_cast<unsigned int *>(c->bkWork.as<int>() + 8);
  cudaStream_t st    = c->stream;
  NCM_CUDA_OK(c, cudaMemsetAsync(c->bkWork.p, 0, 16 * sizeof(int), st));
  if (n <= BK_SMEM_MAX_N) {
    const size_t smem = ((size_t) n * (n + 1) / 2 + n) * sizeof(double) + (size_t) n * sizeof(int) + 16;
    static bool attr_set[NCM_MAX_DEVICES] = {};
    const int dev = c->device < 0 || c->device >= NCM_MAX_DEVICES ? 0 : c->device;
    if (!attr_set[dev]) {
      NCM_CUDA_OK(c, cudaFuncSetAttribute(bk_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 4096));
      attr_set[dev] = true;
    }
    bk_smem_kernel<<<1, BK_T, smem, st>>>(dS, lds, n, dRhs, ipiv, info);
  } else {
    // enough CTAs that a column of the trailing update has a warp of its own, at most one per SM (co-resident: cooperative launch)
    int nctas = std::min(c->n_sm, std::max(1, (n + BK_T / 32 - 1) / (BK_T / 32)));
    void *params[] = {(void *) &dS, (void *) &lds, (void *) &n, (void *) &dRhs, (void *) &ipiv, (void *) &info, (void *) &bar};
    NCM_CUDA_OK(c, cudaLaunchCooperativeKernel((const void *) bk_global_kernel, dim3(nctas), dim3(BK_T), params, 0, st));
  }
  c->n_launches++;
  NCM_CUDA_OK(c, cudaGetLastError());
  int h_info = 0;
  NCM_CUDA_OK(c, ncm_memcpy_async(c, &h_info, info, sizeof(int), cudaMemcpyDeviceToHost, st));
  NCM_CUDA_OK(c, cudaStreamSynchronize(st));
  if (info_host) *info_host = h_info;
  return NCM_SD_GPU_OK;
}
