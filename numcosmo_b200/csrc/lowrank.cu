// Passive-set solves of the NNLS by low-rank modification of a base factor.
//
// ncm_nnls_solve (ncm_nnls.c:728-751, 825-868) solves M[P,P] x = b[P] for a sequence of passive sets that differ from one
// another by a handful of indices (configs[1]: 2048 -> 1901 -> 1890 -> 1889 -> 1901 -> 1898 -> 1899); the reference calls dposv
// on every one of them (ncm_matrix.c:1199-1210).  Here one set B is factorised (M[B,B] = U^T U, chol_fused.cu / chol.cu), the
// triangular inverse W = U^-1 is formed once by recursive doubling (all DMMA GEMMs, no pivot chain), and every later set
// P = (B \ D) u A with |D| + |A| <= LR_KMAX is solved through the bordered / constrained system
//
//      [ M_BB   M_BA   E_D ] [x_B]   [b_B]          R = [M_BA  E_D],  z = [x_A; mu]
//      [ M_AB   M_AA    0  ] [x_A] = [b_A]          T = W^T [R  b_B]      (one triangular GEMM, n_B x (k + 1))
//      [ E_D^T   0      0  ] [mu ]   [ 0 ]          H = T_R^T T_R - diag(M_AA, 0),   H z = T_R^T t_b - [b_A; 0]
//                                                   x_B = W (t_b - T_R z)
//
// (x_D = 0 is enforced by the multipliers mu).  H is k x k, quasi-definite (-Schur complement of the border, +G_DD): it is
// factorised as L J L^T, J = diag(-I_A, +I_D), without pivoting in shared memory.  One step of iterative refinement on the
// residual of the TRUE system (b - M x, symmetric product with the stored upper triangle) follows; the size of that
// correction is returned so that the caller falls back to a fresh factorisation when the base is too ill-conditioned.
// Cost per solve: O(n_B^2 k) flops at GEMM rates instead of the n^3/3 latency chain of a factorisation.
#include <algorithm>
#include "ctx.h"

int dsyrk_ata_general(ncm_sd_gpu_ctx *c, int K, int n, const double *dP, int ldp, double *dC, int ldc, double alpha, double beta);
int gemv_t(ncm_sd_gpu_ctx *c, const double *dA, int lda, int nrows, int ncols, const double *dv, double *dOut, DevBuf &tmp);

namespace {

// ---- DMMA GEMM tile: C[i0:i0+64, j0:j0+64] = alpha * op(A) B over k in [k_lo, k_hi) -------------------------------------
// A: TRANSA ? (K x M row-major, op(A) = A^T) : (M x K row-major);  B: K x N row-major;  C: M x N row-major.
constexpr int GT = 64, GBK = 16, GTHREADS = 128, GSTAGES = 3;
constexpr int GPA = GBK + 4;                    // pitch of the [64][16] A tile (NN): lc * 20 + lr hits 16 distinct bank pairs per half-warp
constexpr int GPB = GT + 4;                     // pitch of the [16][64] tiles (B, and A when TRANSA)
constexpr int GSLAB_A = GT * GPA;               // 1280 doubles >= GBK * GPB = 1088
constexpr int GSLAB_B = GBK * GPB;
constexpr size_t GEMM_SMEM = (size_t) GSTAGES * (GSLAB_A + GSLAB_B) * sizeof(double);

template <bool TRANSA>
__device__ __forceinline__ void gemm_tile(const double *__restrict__ A, int lda, const double *__restrict__ B, int ldb, double *__restrict__ C, int ldc,
                                          int M, int N, int i0, int j0, int k_lo, int k_hi, double alpha, double *smem) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1;
  const int lr = lane & 3, lc = lane >> 2;
  const int rm = wm * 32, cn = wn * 32;
  k_lo &= ~(GBK - 1);
  const int nkb = k_hi > k_lo ? (k_hi - k_lo + GBK - 1) / GBK : 0;

  auto load_stage = [&](int kb, int st) {
    double *sA = smem + (size_t) st * (GSLAB_A + GSLAB_B);
    double *sB = sA + GSLAB_A;
    const int k0 = k_lo + kb * GBK;
    if (TRANSA) {
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int chunk = tid + it * GTHREADS;
        const int r = chunk >> 5, cc = (chunk & 31) * 2;
        const int gk = k0 + r, gc = i0 + cc;
        int bytes = gk < k_hi ? (M - gc) * 8 : 0;
        bytes     = bytes < 0 ? 0 : (bytes > 16 ? 16 : bytes);
        const double *src = bytes > 0 ? A + (size_t) gk * lda + gc : A;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(sA + r * GPB + cc)), "l"(src), "r"(bytes) : "memory");
      }
    } else {
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int chunk = tid + it * GTHREADS;
        const int r = chunk >> 3, cc = (chunk & 7) * 2;
        const int gr = i0 + r, gk = k0 + cc;
        int bytes = gr < M ? (k_hi - gk) * 8 : 0;
        bytes     = bytes < 0 ? 0 : (bytes > 16 ? 16 : bytes);
        const double *src = bytes > 0 ? A + (size_t) gr * lda + gk : A;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(sA + r * GPA + cc)), "l"(src), "r"(bytes) : "memory");
      }
    }
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int chunk = tid + it * GTHREADS;
      const int r = chunk >> 5, cc = (chunk & 31) * 2;
      const int gk = k0 + r, gc = j0 + cc;
      int bytes = gk < k_hi ? (N - gc) * 8 : 0;
      bytes     = bytes < 0 ? 0 : (bytes > 16 ? 16 : bytes);
      const double *src = bytes > 0 ? B + (size_t) gk * ldb + gc : B;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(sB + r * GPB + cc)), "l"(src), "r"(bytes) : "memory");
    }
  };

#pragma unroll
  for (int s = 0; s < GSTAGES - 1; ++s) {
    if (s < nkb) load_stage(s, s);
    cp_async_commit();
  }
  double acc[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

  for (int kb = 0; kb < nkb; ++kb) {
    cp_async_wait<GSTAGES - 2>();
    __syncthreads();
    {
      const int nk = kb + GSTAGES - 1;
      if (nk < nkb) load_stage(nk, nk % GSTAGES);
      cp_async_commit();
    }
    const double *sA = smem + (size_t) (kb % GSTAGES) * (GSLAB_A + GSLAB_B);
    const double *sB = sA + GSLAB_A;
#pragma unroll
    for (int ks = 0; ks < GBK / 4; ++ks) {
      double af[4], bf[4];
      if (TRANSA) {
        const double *pa = sA + (ks * 4 + lr) * GPB + rm + lc;
#pragma unroll
        for (int a = 0; a < 4; ++a) af[a] = pa[a * 8];
      } else {
        const double *pa = sA + (rm + lc) * GPA + ks * 4 + lr;
#pragma unroll
        for (int a = 0; a < 4; ++a) af[a] = pa[a * 8 * GPA];
      }
      const double *pb = sB + (ks * 4 + lr) * GPB + cn + lc;
#pragma unroll
      for (int b = 0; b < 4; ++b) bf[b] = pb[b * 8];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) dmma884(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
    }
  }
  cp_async_wait<0>();
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int gi = i0 + rm + a * 8 + lc;
    if (gi >= M) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int gj = j0 + cn + b * 8 + 2 * lr;
      double *pc   = C + (size_t) gi * ldc + gj;
      if (gj + 1 < N)
        *reinterpret_cast<double2 *>(pc) = make_double2(alpha * acc[a][b][0], alpha * acc[a][b][1]);
      else if (gj < N)
        pc[0] = alpha * acc[a][b][0];
    }
  }
}

// ---- triangular inverse W = U^-1 (upper, row-major) by recursive doubling -----------------------------------------------
// level 0: the 64 x 64 diagonal blocks, one CTA each (thread t back-substitutes column t)
__global__ void __launch_bounds__(64) trinv_diag_kernel(const double *__restrict__ U, double *__restrict__ W, int ld, int n) {
  // upper triangle: U; strictly lower triangle: X = U^-1 transposed (X[i][t] at sU[t][i]); diagonal of X in sDinv
  __shared__ double sU[64][65];
  __shared__ double sDinv[64];
  const int k0 = blockIdx.x * 64, nb = min(64, n - k0), t = threadIdx.x;
  for (int r = 0; r < 64; ++r) {
    double v = (r == t) ? 1.0 : 0.0;
    if (r < nb && t < nb && t >= r) v = U[(size_t) (k0 + r) * ld + k0 + t];
    sU[r][t] = v;
  }
  __syncthreads();
  sDinv[t] = 1.0 / sU[t][t];
  __syncthreads();
  const double xtt = sDinv[t];
  for (int i = 62; i >= 0; --i) {
    if (i < t) {
      double s0 = sU[i][t] * xtt, s1 = 0.0;
      int k = i + 1;
      for (; k + 1 < t; k += 2) {
        s0 = fma(sU[i][k], sU[t][k], s0);
        s1 = fma(sU[i][k + 1], sU[t][k + 1], s1);
      }
      if (k < t) s0 = fma(sU[i][k], sU[t][k], s0);
      sU[t][i] = -(s0 + s1) * sDinv[i];
    }
  }
  __syncthreads();
  for (int r = 0; r < nb; ++r)
    if (t < nb) W[(size_t) (k0 + r) * ld + k0 + t] = (t > r) ? sU[t][r] : (t == r ? xtt : 0.0);
}

// level s: for every pair of adjacent s-blocks (r0 = 2 p s, r1 = r0 + s)  S12 = U12 W22 ;  W12 = - W11 S12
__global__ void __launch_bounds__(GTHREADS) trinv_step1_kernel(const double *__restrict__ U, const double *__restrict__ W, double *__restrict__ S, int ld, int n, int s) {
  extern __shared__ __align__(16) double smem[];
  const int r0 = 2 * blockIdx.z * s, r1 = r0 + s;
  const int m2 = min(s, n - r1);
  const int i0 = blockIdx.y * GT, j0 = blockIdx.x * GT;
  if (m2 <= 0 || j0 >= m2) return;
  // B = W22 is upper triangular: k <= j
  gemm_tile<false>(U + (size_t) r0 * ld + r1, ld, W + (size_t) r1 * ld + r1, ld, S + (size_t) r0 * ld + r1, ld, s, m2, i0, j0, 0, min(m2, j0 + GT), 1.0, smem);
}
__global__ void __launch_bounds__(GTHREADS) trinv_step2_kernel(double *__restrict__ W, const double *__restrict__ S, int ld, int n, int s) {
  extern __shared__ __align__(16) double smem[];
  const int r0 = 2 * blockIdx.z * s, r1 = r0 + s;
  const int m2 = min(s, n - r1);
  const int i0 = blockIdx.y * GT, j0 = blockIdx.x * GT;
  if (m2 <= 0 || j0 >= m2) return;
  // A = W11 is upper triangular: k >= i
  gemm_tile<false>(W + (size_t) r0 * ld + r0, ld, S + (size_t) r0 * ld + r1, ld, W + (size_t) r0 * ld + r1, ld, s, m2, i0, j0, i0, s, -1.0, smem);
}

// C = A^T B, A: K x M upper triangular when `upper_a` (k <= i)
__global__ void __launch_bounds__(GTHREADS) gemm_tn_kernel(const double *__restrict__ A, int lda, const double *__restrict__ B, int ldb, double *__restrict__ C, int ldc,
                                                           int M, int N, int K, int upper_a) {
  extern __shared__ __align__(16) double smem[];
  const int i0 = blockIdx.y * GT, j0 = blockIdx.x * GT;
  gemm_tile<true>(A, lda, B, ldb, C, ldc, M, N, i0, j0, 0, upper_a ? min(K, i0 + GT) : K, 1.0, smem);
}

// ---- index plumbing -----------------------------------------------------------------------------------------------------
__device__ __forceinline__ double msym(const double *__restrict__ M, int ldm, int a, int b) {
  return a <= b ? M[(size_t) a * ldm + b] : M[(size_t) b * ldm + a];
}

// V = [ M[B, A] | E_D | b_B ]   (n_B x (na + nd + 1), row-major, ld = ldv)
__global__ void lr_gather_V_kernel(const double *__restrict__ M, int ldm, const double *__restrict__ b, const int *__restrict__ idxB, int nB,
                                   const int *__restrict__ idxA, int na, const int *__restrict__ posD, int nd, double *__restrict__ V, int ldv) {
  const int i = blockIdx.x * blockDim.y + threadIdx.y;
  if (i >= nB) return;
  const int gi = idxB[i], kc = na + nd + 1;
  for (int j = threadIdx.x; j < kc; j += blockDim.x) {
    double v;
    if (j < na)
      v = msym(M, ldm, gi, idxA[j]);
    else if (j < na + nd)
      v = (posD[j - na] == i) ? 1.0 : 0.0;
    else {
      v = b[gi];   // rows in D are constraint rows: their right-hand side is absorbed by the multipliers, so it is dropped (less cancellation)
      for (int q = 0; q < nd; ++q)
        if (posD[q] == i) v = 0.0;
    }
    V[(size_t) i * ldv + j] = v;
  }
}

// y[i] = base[i * bstride] - sum_j T[i][j] z[j]   (warp per row)
__global__ void lr_y_kernel(const double *__restrict__ T, int ldt, int nB, int k, const double *__restrict__ z, const double *__restrict__ base, int bstride,
                            double *__restrict__ y) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i    = blockIdx.x * (blockDim.x >> 5) + warp;
  if (i >= nB) return;
  double s = 0.0;
  for (int j = lane; j < k; j += 32) s = fma(T[(size_t) i * ldt + j], z[j], s);
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) y[i] = base[(size_t) i * bstride] - s;
}

// out[i] = sum_{k >= i} A[i][k] v[k]   (upper triangle, warp per row): x = W y, and the row part of the symmetric product
__global__ void trmv_upper_row_kernel(const double *__restrict__ A, int lda, int n, const double *__restrict__ v, double *__restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i    = blockIdx.x * (blockDim.x >> 5) + warp;
  if (i >= n) return;
  const double *a = A + (size_t) i * lda;
  double s0 = 0.0, s1 = 0.0;
  int k = (i & ~31) + lane;   // aligned start: coalesced 256-byte segments
  if (k >= i && k < n) s0 = a[k] * v[k];
  k += 32;
  for (; k + 32 < n; k += 64) {
    s0 = fma(a[k], v[k], s0);
    s1 = fma(a[k + 32], v[k + 32], s1);
  }
  for (; k < n; k += 32) s0 = fma(a[k], v[k], s0);
  double s = s0 + s1;
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) out[i] = s;
}

// part[blk][j] = sum_{r in block, r < j + incl} A[r][j] v[r]   (upper triangle, transposed product; thread per column)
__global__ void trmv_upper_col_partial_kernel(const double *__restrict__ A, int lda, int n, const double *__restrict__ v, double *__restrict__ part, int rows_per_block,
                                              int incl) {
  const int j  = blockIdx.x * blockDim.x + threadIdx.x;
  const int r0 = blockIdx.y * rows_per_block;
  if (j >= n) return;
  const int r1 = min(min(n, r0 + rows_per_block), j + incl);
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int r = r0;
  for (; r + 3 < r1; r += 4) {
    s0 = fma(A[(size_t) r * lda + j], v[r], s0);
    s1 = fma(A[(size_t) (r + 1) * lda + j], v[r + 1], s1);
    s2 = fma(A[(size_t) (r + 2) * lda + j], v[r + 2], s2);
    s3 = fma(A[(size_t) (r + 3) * lda + j], v[r + 3], s3);
  }
  for (; r < r1; ++r) s0 = fma(A[(size_t) r * lda + j], v[r], s0);
  part[(size_t) blockIdx.y * n + j] = (s0 + s1) + (s2 + s3);
}
// out[j] = (base ? base[j] : 0) + sign * (sum_p part[p][j] + (extra ? extra[j] : 0))
__global__ void lr_reduce_kernel(const double *__restrict__ part, int nparts, int n, const double *__restrict__ extra, const double *__restrict__ base, double sign,
                                 double *__restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  double s = 0.0;
  for (int p = 0; p < nparts; ++p) s += part[(size_t) p * n + j];
  if (extra) s += extra[j];
  out[j] = (base ? base[j] : 0.0) + sign * s;
}

// xfull[B[i]] = xB[i]; then (second launch) the positions in D are forced to 0 and xfull[A[j]] = z[j]
__global__ void lr_scatter_kernel(double *__restrict__ xfull, const int *__restrict__ idxB, int nB, const double *__restrict__ xB) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nB) xfull[idxB[i]] = xB[i];
}
__global__ void lr_scatter_fix_kernel(double *__restrict__ xfull, const int *__restrict__ idxB, const int *__restrict__ idxA, int na, const double *__restrict__ z,
                                      const int *__restrict__ posD, int nd) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nd) xfull[idxB[posD[i]]] = 0.0;
  if (i < na) xfull[idxA[i]] = z[i];
}
__global__ void lr_zero_at_kernel(double *__restrict__ v, const int *__restrict__ pos, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[pos[i]] = 0.0;
}
__global__ void lr_gather_kernel(const double *__restrict__ src, const int *__restrict__ idx, int n, double *__restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}
// out[j] = x0[P[j]] + dx[P[j]] for the new passive set; stats = {max |dx|, max |x|}   (single CTA)
__global__ void lr_finalize_kernel(const double *__restrict__ x0, const double *__restrict__ dx, const int *__restrict__ idxP, int np, double *__restrict__ out,
                                   double *__restrict__ stats) {
  __shared__ double s_dx[32], s_x[32];
  double mdx = 0.0, mx = 0.0;
  for (int j = threadIdx.x; j < np; j += blockDim.x) {
    const int g    = idxP[j];
    const double d = dx ? dx[g] : 0.0, v = x0[g] + d;
    out[j] = v;
    mdx    = fmax(mdx, fabs(d));
    mx     = fmax(mx, fabs(v));
  }
  for (int off = 16; off > 0; off >>= 1) {
    mdx = fmax(mdx, __shfl_xor_sync(0xffffffffu, mdx, off));
    mx  = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  }
  if ((threadIdx.x & 31) == 0) {
    s_dx[threadIdx.x >> 5] = mdx;
    s_x[threadIdx.x >> 5]  = mx;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int) (blockDim.x >> 5); ++w) {
      mdx = fmax(mdx, s_dx[w]);
      mx  = fmax(mx, s_x[w]);
    }
    stats[0] = mdx;
    stats[1] = mx;
  }
}

// ---- the k x k quasi-definite system: H = L J L^T in shared memory, packed lower -----------------------------------------
constexpr int LR_KMAX = 200;
constexpr int LR_ST = 512;
constexpr size_t LR_SMALL_SMEM = ((size_t) LR_KMAX * (LR_KMAX + 1) / 2 + 3 * LR_KMAX + 8) * sizeof(double);

__device__ __forceinline__ int pidx(int i, int j) { return i * (i + 1) / 2 + j; }   // i >= j

// mode 0: build H = H0[:k,:k] - diag(M_AA, 0) and rhs = H0[:k, k] - [b_A; 0], factor, solve, keep the factor in Lg / dinvg.
// mode 1: reload the factor, rhs = rhs_in - [rfull_A; 0], solve.                 info: 0 or the 1-based index of a pivot of the wrong sign
__global__ void __launch_bounds__(LR_ST) lr_small_kernel(int mode, int na, int nd, const double *__restrict__ H0, int ldh, const double *__restrict__ M, int ldm,
                                                         const int *__restrict__ idxA, const double *__restrict__ b, const double *__restrict__ rhs_in,
                                                         const double *__restrict__ rfull, double *__restrict__ Lg, double *__restrict__ dinvg, double *__restrict__ z,
                                                         int *__restrict__ info) {
  extern __shared__ __align__(16) double sm[];
  const int k = na + nd, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int npk = k * (k + 1) / 2;
  double *L = sm, *rhs = sm + (size_t) LR_KMAX * (LR_KMAX + 1) / 2, *dinv = rhs + LR_KMAX, *colj = dinv + LR_KMAX;
  if (k == 0) return;
  if (mode == 0) {
    for (int i = warp; i < k; i += LR_ST / 32)
      for (int j = lane; j <= i; j += 32) {
        double v = H0[(size_t) j * ldh + i];
        if (i < na) v -= M[(size_t) idxA[j] * ldm + idxA[i]];   // idxA ascending: (j, i) is in the stored upper triangle
        L[pidx(i, j)] = v;
      }
    for (int i = tid; i < k; i += LR_ST) rhs[i] = H0[(size_t) i * ldh + k] - (i < na ? b[idxA[i]] : 0.0);
    __syncthreads();
    for (int j = 0; j < k; ++j) {
      const double sgn = j < na ? -1.0 : 1.0;
      double p         = sgn * L[pidx(j, j)];
      if (!(p > 0.0)) {
        if (tid == 0 && *info == 0) *info = j + 1;
        p = 1.0;
      }
      const double ljj = sqrt(p), inv = 1.0 / ljj;
      for (int i = j + 1 + tid; i < k; i += LR_ST) {
        const double v = L[pidx(i, j)] * (sgn * inv);
        L[pidx(i, j)]  = v;
        colj[i]        = v;
      }
      __syncthreads();
      if (tid == 0) {
        L[pidx(j, j)] = ljj;
        dinv[j]       = inv;
      }
      // trailing update  H[i][l] -= sgn L[i][j] L[l][j],  j < l <= i
      for (int i = j + 1 + warp; i < k; i += LR_ST / 32) {
        const double cij = sgn * colj[i];
        double *row      = L + pidx(i, 0);
        for (int l = j + 1 + lane; l <= i; l += 32) row[l] = fma(-cij, colj[l], row[l]);
      }
      __syncthreads();
    }
    for (int e = tid; e < npk; e += LR_ST) Lg[e] = L[e];
    for (int i = tid; i < k; i += LR_ST) dinvg[i] = dinv[i];
  } else {
    for (int e = tid; e < npk; e += LR_ST) L[e] = Lg[e];
    for (int i = tid; i < k; i += LR_ST) {
      dinv[i] = dinvg[i];
      rhs[i]  = rhs_in[i] - (i < na ? rfull[idxA[i]] : 0.0);
    }
    __syncthreads();
  }
  if (warp == 0) {
    // L u = rhs ; v = J u ; L^T z = v
    for (int j = 0; j < k; ++j) {
      const double u = rhs[j] * dinv[j];
      __syncwarp();
      for (int i = j + 1 + lane; i < k; i += 32) rhs[i] = fma(-L[pidx(i, j)], u, rhs[i]);
      if (lane == 0) rhs[j] = (j < na) ? -u : u;
      __syncwarp();
    }
    for (int j = k - 1; j >= 0; --j) {
      const double zz = rhs[j] * dinv[j];
      __syncwarp();
      const double *row = L + pidx(j, 0);
      for (int i = lane; i < j; i += 32) rhs[i] = fma(-row[i], zz, rhs[i]);
      if (lane == 0) rhs[j] = zz;
      __syncwarp();
    }
    for (int i = lane; i < k; i += 32) z[i] = rhs[i];
  }
}

cudaError_t set_smem_attrs() {
  static bool done[NCM_MAX_DEVICES] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  dev = dev < 0 || dev >= NCM_MAX_DEVICES ? 0 : dev;
  if (done[dev]) return cudaSuccess;
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(trinv_step1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) GEMM_SMEM)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(trinv_step2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) GEMM_SMEM)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(gemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) GEMM_SMEM)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(lr_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) LR_SMALL_SMEM)) != cudaSuccess) return e;
  done[dev] = true;
  return cudaSuccess;
}

}   // namespace

int lowrank_kmax() { return LR_KMAX; }

// W = U^-1 for the upper-triangular factor U (n x n, row-major, ld); S is scratch of the same shape.
int trinv_upper(ncm_sd_gpu_ctx *c, int n, const double *dU, double *dW, double *dS, int ld) {
  NCM_CUDA_OK(c, set_smem_attrs());
  NCM_CUDA_OK(c, cudaMemsetAsync(dW, 0, (size_t) n * ld * sizeof(double), c->stream));
  trinv_diag_kernel<<<(n + 63) / 64, 64, 0, c->stream>>>(dU, dW, ld, n);
  c->n_launches++;
  for (int s = 64; s < n; s *= 2) {
    const int pairs = (n - s + 2 * s - 1) / (2 * s);
    dim3 grid((s + GT - 1) / GT, (s + GT - 1) / GT, pairs);
    trinv_step1_kernel<<<grid, GTHREADS, GEMM_SMEM, c->stream>>>(dU, dW, dS, ld, n, s);
    trinv_step2_kernel<<<grid, GTHREADS, GEMM_SMEM, c->stream>>>(dW, dS, ld, n, s);
    c->n_launches += 2;
  }
  NCM_CUDA_OK(c, cudaGetLastError());
  return NCM_SD_GPU_OK;
}

// out = A^T v over the upper triangle of A (rows r <= j, or r < j when !inclusive): two-pass, deterministic
static int trmv_upper_col(ncm_sd_gpu_ctx *c, const double *dA, int lda, int n, const double *dv, bool inclusive, const double *extra, const double *base, double sign,
                          double *dOut, DevBuf &tmp) {
  int nblk = (c->n_sm * 2 * 256 + n - 1) / n;
  if (nblk > (n + 63) / 64) nblk = (n + 63) / 64;
  if (nblk < 1) nblk = 1;
  const int rpb = (n + nblk - 1) / nblk;
  nblk          = (n + rpb - 1) / rpb;
  if (!tmp.reserve((size_t) nblk * n * sizeof(double))) return c->fail(NCM_SD_GPU_ENOMEM, "lowrank: out of device memory");
  dim3 grid((n + 255) / 256, nblk);
  trmv_upper_col_partial_kernel<<<grid, 256, 0, c->stream>>>(dA, lda, n, dv, tmp.as<double>(), rpb, inclusive ? 1 : 0);
  lr_reduce_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(tmp.as<double>(), nblk, n, extra, base, sign, dOut);
  c->n_launches += 2;
  return NCM_SD_GPU_OK;
}

// Solve M[P,P] x = b[P] for P = (B \ D) u A through the base inverse W (see the header).  idxB / idxA / posD / idxP are device
// arrays (ascending); the result (np doubles, in the order of P) and {max |dx|, max |x|} are left in bufs.out / bufs.stats.
int lowrank_solve(ncm_sd_gpu_ctx *c, const double *dM, int ldm, int n, const double *db, int nB, int na, int nd, int np, const LowrankBufs &w, int ldv,
                  DevBuf &tmp) {
  NCM_CUDA_OK(c, set_smem_attrs());
  const int k = na + nd, kc = k + 1;
  cudaStream_t st = c->stream;
  NCM_CUDA_OK(c, cudaMemsetAsync(w.info, 0, sizeof(int), st));
  {
    dim3 blk(32, 8);
    lr_gather_V_kernel<<<(nB + 7) / 8, blk, 0, st>>>(dM, ldm, db, w.idxB, nB, w.idxA, na, w.posD, nd, w.V, ldv);
  }
  {
    dim3 grid((kc + GT - 1) / GT, (nB + GT - 1) / GT);
    gemm_tn_kernel<<<grid, GTHREADS, GEMM_SMEM, st>>>(w.W, ldm, w.V, ldv, w.T, ldv, nB, kc, nB, 1);
  }
  c->n_launches += 2;
  int rc = dsyrk_ata_general(c, nB, kc, w.T, ldv, w.H, ldv, 1.0, 0.0);
  if (rc != NCM_SD_GPU_OK) return rc;
  if (k > 0) {
    lr_small_kernel<<<1, LR_ST, LR_SMALL_SMEM, st>>>(0, na, nd, w.H, ldv, dM, ldm, w.idxA, db, nullptr, nullptr, w.Lg, w.dinvg, w.z, w.info);
    c->n_launches++;
  }
  lr_y_kernel<<<(nB + 7) / 8, 256, 0, st>>>(w.T, ldv, nB, k, w.z, w.T + k, ldv, w.y);
  trmv_upper_row_kernel<<<(nB + 7) / 8, 256, 0, st>>>(w.W, ldm, nB, w.y, w.xB);
  NCM_CUDA_OK(c, cudaMemsetAsync(w.xfull, 0, (size_t) n * sizeof(double), st));
  lr_scatter_kernel<<<(nB + 255) / 256, 256, 0, st>>>(w.xfull, w.idxB, nB, w.xB);
  if (k > 0) lr_scatter_fix_kernel<<<(std::max(na, nd) + 255) / 256, 256, 0, st>>>(w.xfull, w.idxB, w.idxA, na, w.z, w.posD, nd);
  c->n_launches += 4;

  // one step of iterative refinement on the true system: r = b - Msym xfull
  trmv_upper_row_kernel<<<(n + 7) / 8, 256, 0, st>>>(dM, ldm, n, w.xfull, w.row);
  c->n_launches++;
  rc = trmv_upper_col(c, dM, ldm, n, w.xfull, false, w.row, db, -1.0, w.rfull, tmp);
  if (rc != NCM_SD_GPU_OK) return rc;
  lr_gather_kernel<<<(nB + 255) / 256, 256, 0, st>>>(w.rfull, w.idxB, nB, w.rB);
  if (nd > 0) lr_zero_at_kernel<<<(nd + 255) / 256, 256, 0, st>>>(w.rB, w.posD, nd);
  c->n_launches += 2;
  rc = trmv_upper_col(c, w.W, ldm, nB, w.rB, true, nullptr, nullptr, 1.0, w.tr, tmp);   // t_r = W^T r_B
  if (rc != NCM_SD_GPU_OK) return rc;
  if (k > 0) {
    rc = gemv_t(c, w.T, ldv, nB, k, w.tr, w.rhsz, tmp);                                   // T_R^T t_r
    if (rc != NCM_SD_GPU_OK) return rc;
    lr_small_kernel<<<1, LR_ST, LR_SMALL_SMEM, st>>>(1, na, nd, w.H, ldv, dM, ldm, w.idxA, db, w.rhsz, w.rfull, w.Lg, w.dinvg, w.z, w.info);
    c->n_launches++;
  }
  lr_y_kernel<<<(nB + 7) / 8, 256, 0, st>>>(w.T, ldv, nB, k, w.z, w.tr, 1, w.y);
  trmv_upper_row_kernel<<<(nB + 7) / 8, 256, 0, st>>>(w.W, ldm, nB, w.y, w.xB);
  NCM_CUDA_OK(c, cudaMemsetAsync(w.dxfull, 0, (size_t) n * sizeof(double), st));
  lr_scatter_kernel<<<(nB + 255) / 256, 256, 0, st>>>(w.dxfull, w.idxB, nB, w.xB);
  if (k > 0) lr_scatter_fix_kernel<<<(std::max(na, nd) + 255) / 256, 256, 0, st>>>(w.dxfull, w.idxB, w.idxA, na, w.z, w.posD, nd);
  lr_finalize_kernel<<<1, 1024, 0, st>>>(w.xfull, w.dxfull, w.idxP, np, w.out, w.out + np);
  c->n_launches += 5;
  NCM_CUDA_OK(c, cudaGetLastError());
  return NCM_SD_GPU_OK;
}
