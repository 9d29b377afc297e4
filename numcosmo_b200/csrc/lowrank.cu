// Passive-set solves of the NNLS by low-rank modification of a base factor.
//
// ncm_nnls_solve (ncm_nnls.c:728-751, 825-868) solves M[P,P] x = b[P] for a sequence of passive sets that differ from one
// another by a handful of indices (configs[1]: 2048 -> 1901 -> 1890 -> 1889 -> 1901 -> 1898 -> 1899); the reference calls dposv
// on every one of them (ncm_matrix.c:1199-1210).  Here one set B is factorised (M[B,B] = U^T U, chol_fused.cu / chol.cu), the
// triangular inverse W = U^-1 is formed once by recursive doubling (all DMMA GEMMs, no pivot chain), and every later set
// P = (B \ D) u A with |D| + |A| <= LR_KMAX is solved through the bordered / constrained system
//
//      [ M_BB   M_BA   E_D ] [x_B]   [b_B]          R = [M_BA  E_D],  z = [x_A; mu]
//      [ M_AB   M_AA    0  ] [x_A] = [b_A]          T = W^T [R  b_B]      (one triangular GEMM, n_B x (k + 1), split over K)
//      [ E_D^T   0      0  ] [mu ]   [ 0 ]          H = T_R^T T_R - diag(M_AA, 0),   H z = T_R^T t_b - [b_A; 0]
//                                                   x_B = W (t_b - T_R z)
//
// (x_D = 0 is enforced by the multipliers mu).  H is k x k, quasi-definite (-Schur complement of the border, +G_DD): it is
// factorised as L J L^T, J = diag(-I_A, +I_D), without pivoting in shared memory (8-column panels, the right-hand side rides
// along as one more row so the forward substitution is free).  One step of iterative refinement on the residual of the TRUE
// system (b - M x, M symmetrised once per NNLS call) follows; the size of that correction is returned so that the caller falls
// back to a fresh factorisation when the base is too ill-conditioned.
// Cost per solve: O(n_B^2 k) flops at GEMM rates instead of the n^3/3 latency chain of a factorisation.
#include <algorithm>
#include "ctx.h"
#include "gemm_tile.cuh"

namespace {

using namespace ncm_gemm;

// ---- triangular inverse W = U^-1 (upper, row-major) by recursive doubling -----------------------------------------------
// level 0: the 64 x 64 diagonal blocks, one CTA each (thread t back-substitutes column t)
__global__ void __launch_bounds__(256) trinv_diag_kernel(const double *__restrict__ U, double *__restrict__ W, int ld, int n) {
  // upper triangle: U; strictly lower triangle: X = U^-1 transposed (X[i][t] at sU[t][i]); diagonal of X in sDinv.
  // Column t of the inverse is back-substituted by FOUR adjacent lanes (each a quarter of every dot product, two shuffles to add
  // them up): the columns are independent, so a warp only synchronises with itself; 256 threads also move the tile.
  __shared__ double sU[64][65];
  __shared__ double sDinv[64];
  const int k0 = blockIdx.x * 64, nb = min(64, n - k0);
  {
    const int t = threadIdx.x & 63, ty = threadIdx.x >> 6;
#pragma unroll 4
    for (int r = ty; r < 64; r += 4) {
      double v = (r == t) ? 1.0 : 0.0;
      if (r < nb && t < nb && t >= r) v = U[(size_t) (k0 + r) * ld + k0 + t];
      sU[r][t] = v;
    }
  }
  __syncthreads();
  if (threadIdx.x < 64) sDinv[threadIdx.x] = 1.0 / sU[threadIdx.x][threadIdx.x];
  __syncthreads();
  const int t = threadIdx.x >> 2, part = threadIdx.x & 3;
  const double xtt = sDinv[t];
  for (int i = 62; i >= 0; --i) {
    // all four lanes of a column take the same branch; the columns of one warp (8 of them) may differ: shuffles under the full mask
    double s = 0.0;
    if (i < t) {
      for (int k = i + 1 + part; k < t; k += 4) s = fma(sU[i][k], sU[t][k], s);
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if (i < t && part == 0) sU[t][i] = -(fma(sU[i][t], xtt, s)) * sDinv[i];
    __syncwarp();
  }
  __syncthreads();
  {
    const int tc = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const double xd = sDinv[tc];
#pragma unroll 4
    for (int r = ty; r < nb; r += 4)
      if (tc < nb) W[(size_t) (k0 + r) * ld + k0 + tc] = (tc > r) ? sU[tc][r] : (tc == r ? xd : 0.0);
  }
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

// small levels (s <= 128): both products of a level in ONE launch.  A CTA takes one 64-column strip of a pair's off-diagonal block:
// it forms its strip of S12 = U12 W22 tile by tile, then W12 = -W11 S12 from that strip alone (no other CTA's tiles are needed), so
// the two half-empty launches of the level (16 CTAs of work each) become one.
__global__ void __launch_bounds__(GTHREADS) trinv_step12_kernel(const double *__restrict__ U, double *__restrict__ W, double *__restrict__ S, int ld, int n, int s) {
  extern __shared__ __align__(16) double smem[];
  const int r0 = 2 * blockIdx.z * s, r1 = r0 + s;
  const int m2 = min(s, n - r1);
  const int j0 = blockIdx.x * GT;
  if (m2 <= 0 || j0 >= m2) return;
  for (int i0 = 0; i0 < s; i0 += GT) {
    gemm_tile<false>(U + (size_t) r0 * ld + r1, ld, W + (size_t) r1 * ld + r1, ld, S + (size_t) r0 * ld + r1, ld, s, m2, i0, j0, 0, min(m2, j0 + GT), 1.0, smem);
    __syncthreads();   // the staging buffers are reused by the next tile
  }
  __threadfence_block();
  __syncthreads();     // the strip of S12 written above is read back (through L2) below
  for (int i0 = 0; i0 < s; i0 += GT) {
    gemm_tile<false>(W + (size_t) r0 * ld + r0, ld, S + (size_t) r0 * ld + r1, ld, W + (size_t) r0 * ld + r1, ld, s, m2, i0, j0, i0, s, -1.0, smem);
    __syncthreads();
  }
}

// dst[j][i] = src[i][j] for i <= j (32 x 32 tiles of the upper triangle); in place (dst == src) it symmetrises the matrix
__global__ void transpose_upper_kernel(const double *src, double *dst, int ld, int n, int nt) {
  __shared__ double tile[32][33];
  int t = blockIdx.x, ti = 0;   // linear index -> upper tile (ti <= tj)
  while (t >= nt - ti) {
    t -= nt - ti;
    ++ti;
  }
  const int tj = ti + t;
  const int i0 = ti * 32, j0 = tj * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int gi = i0 + r, gj = j0 + threadIdx.x;
    tile[r][threadIdx.x] = (gi < n && gj < n && gj >= gi) ? src[(size_t) gi * ld + gj] : 0.0;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int gj = j0 + r, gi = i0 + threadIdx.x;   // writes dst[gj][gi]
    if (gj < n && gi < n && (gj > gi || (dst != src && gj == gi))) dst[(size_t) gj * ld + gi] = tile[threadIdx.x][r];
  }
}

// T partial = W[k-chunk]^T V[k-chunk]: A = W is K x M upper triangular (k <= i), split over K in chunks of KC_T rows
constexpr int KC_T = 256, KC_H = 128;
__global__ void __launch_bounds__(GTHREADS) gemm_tn_splitk_kernel(const double *__restrict__ A, int lda, const double *__restrict__ B, int ldb, double *__restrict__ Cpart,
                                                                  int ldc, size_t cstride, int M, int N, int K) {
  extern __shared__ __align__(16) double smem[];
  const int i0 = blockIdx.y * GT, j0 = blockIdx.x * GT;
  const int k_hi = min(K, i0 + GT), k_lo = blockIdx.z * KC_T;
  if (k_lo >= k_hi) return;
  gemm_tile<true>(A, lda, B, ldb, Cpart + (size_t) blockIdx.z * cstride, ldc, M, N, i0, j0, k_lo, min(k_hi, k_lo + KC_T), 1.0, smem);
}
// T[i][j] = sum_c Tpart[c][i][j], c < number of K chunks that reach row tile i (fixed order: deterministic)
__global__ void splitk_reduce_kernel(const double *__restrict__ part, size_t cstride, int ld, int M, int N, int K, double *__restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= N || i >= M) return;
  const int k_hi = min(K, (i / GT) * GT + GT);
  const int nc   = (k_hi + KC_T - 1) / KC_T;
  double s = 0.0;
  for (int c = 0; c < nc; ++c) s += part[(size_t) c * cstride + (size_t) i * ld + j];
  out[(size_t) i * ld + j] = s;
}
// H partial = T[k-chunk]^T T[k-chunk], upper 64-tiles of the kc x kc result
__global__ void __launch_bounds__(GTHREADS) syrk_splitk_kernel(const double *__restrict__ T, int ldt, double *__restrict__ Hpart, int ldh, size_t hstride, int kc, int K) {
  extern __shared__ __align__(16) double smem[];
  int t = blockIdx.x, ti = 0;
  const int nt = (kc + GT - 1) / GT;
  while (t >= nt - ti) {
    t -= nt - ti;
    ++ti;
  }
  const int tj = ti + t;
  const int k_lo = blockIdx.y * KC_H;
  gemm_tile<true>(T, ldt, T, ldt, Hpart + (size_t) blockIdx.y * hstride, ldh, kc, kc, ti * GT, tj * GT, k_lo, min(K, k_lo + KC_H), 1.0, smem);
}

// H0[i][j] = sum_c Hpart[c][i][j] over the upper triangle (j >= i) of the kc x kc matrix, chunks in fixed order
__global__ void hreduce_kernel(const double *__restrict__ part, size_t hstride, int nhc, int ldh, int kc, double *__restrict__ H0) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= kc || j < i) return;
  double s0 = 0.0, s1 = 0.0;
  int c = 0;
  for (; c + 1 < nhc; c += 2) {
    s0 += part[(size_t) c * hstride + (size_t) i * ldh + j];
    s1 += part[(size_t) (c + 1) * hstride + (size_t) i * ldh + j];
  }
  if (c < nhc) s0 += part[(size_t) c * hstride + (size_t) i * ldh + j];
  H0[(size_t) i * ldh + j] = s0 + s1;
}

// ---- index plumbing and memory-bound products -----------------------------------------------------------------------------
// V = [ M[B, A] | E_D | b_B (0 on the rows of D) ]   (n_B x (na + nd + 1), row-major, ld = ldv); M is full symmetric
__global__ void lr_gather_V_kernel(const double *__restrict__ M, int ldm, const double *__restrict__ b, const int *__restrict__ bsel, const int *__restrict__ idxB, int nB,
                                   const int *__restrict__ idxA, int na, const int *__restrict__ posD, int nd, double *__restrict__ V, int ldv) {
  const int i = blockIdx.x * blockDim.y + threadIdx.y;
  if (i >= nB) return;
  const int gi = idxB[i], kc = na + nd + 1;
  const double *mrow = M + (size_t) gi * ldm;
  for (int j = threadIdx.x; j < kc; j += blockDim.x) {
    double v;
    if (j < na)
      v = mrow[idxA[j]];
    else if (j < na + nd)
      v = (posD[j - na] == i) ? 1.0 : 0.0;
    else
      v = bsel[i] >= 0 ? b[gi] : 0.0;   // rows of D are constraint rows: their right-hand side is absorbed by the multipliers
    V[(size_t) i * ldv + j] = v;
  }
}

// Pure removals (A empty): W^T E_D is a gather of rows of W, no product.  T[i][j] = W[posD[j]][i] for j < k (zero above the diagonal of
// W), T[i][k] = tb[i] = (W^T b_B)[i], formed once per base.  32 x 32 tiles through shared memory: reads along the rows of W, writes along
// the rows of T.
__global__ void lr_gather_T_kernel(const double *__restrict__ W, int ldw, int nB, const int *__restrict__ posD, int k, const double *__restrict__ tb,
                                   double *__restrict__ T, int ldt) {
  __shared__ double tile[32][33];
  const int i0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {   // r: column j of T = row posD[j] of W
    const int j = j0 + r, i = i0 + threadIdx.x;
    double v = 0.0;
    if (i < nB) {
      if (j < k) {
        const int pr = posD[j];
        v = i >= pr ? W[(size_t) pr * ldw + i] : 0.0;
      } else if (j == k) {
        v = tb[i];
      }
    }
    tile[r][threadIdx.x] = v;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int i = i0 + r, j = j0 + threadIdx.x;
    if (i < nB && j <= k) T[(size_t) i * ldt + j] = tile[threadIdx.x][r];
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

// out[i] = (base ? base[i] : 0) + sign * sum_{k in range(i)} A[i][k] v[k];  range: UPPER [i, n), LOWER [0, i], FULL [0, n).
// ROW_WPR warps share a row (64-element chunks dealt round-robin, four chunks in flight per warp), 8 / ROW_WPR rows per CTA:
// a single warp per row is latency-bound on its 15 KB (measured 11.5 us per product at n = 1900).
enum { ROW_UPPER = 0, ROW_LOWER = 1, ROW_FULL = 2 };
constexpr int ROW_WPR = 4;
template <int MODE>
__global__ void __launch_bounds__(256) row_dot_kernel(const double *__restrict__ A, int lda, int n, const double *__restrict__ v, const double *__restrict__ base, double sign,
                                                      double *__restrict__ out) {
  __shared__ double red[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i    = blockIdx.x * (8 / ROW_WPR) + warp / ROW_WPR, part = warp % ROW_WPR;
  double s = 0.0;
  if (i < n) {
    const double *a = A + (size_t) i * lda;
    const int lo = MODE == ROW_UPPER ? i : 0, hi = MODE == ROW_LOWER ? i + 1 : n;   // [lo, hi)
    constexpr int STEP = 64 * ROW_WPR;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0, s4 = 0.0, s5 = 0.0, s6 = 0.0, s7 = 0.0;
    int k = (lo & ~63) + 64 * part + 2 * lane;   // rows are 16-byte aligned (lda even), v too
    if (k < lo + 64) {                           // chunk that may straddle lo
      if (k >= lo && k < hi) s0 = a[k] * v[k];
      if (k + 1 >= lo && k + 1 < hi) s1 = a[k + 1] * v[k + 1];
      k += STEP;
    }
    for (; k + 3 * STEP + 1 < hi; k += 4 * STEP) {
      const double2 a0 = *reinterpret_cast<const double2 *>(a + k), a1 = *reinterpret_cast<const double2 *>(a + k + STEP);
      const double2 a2 = *reinterpret_cast<const double2 *>(a + k + 2 * STEP), a3 = *reinterpret_cast<const double2 *>(a + k + 3 * STEP);
      const double2 v0 = *reinterpret_cast<const double2 *>(v + k), v1 = *reinterpret_cast<const double2 *>(v + k + STEP);
      const double2 v2 = *reinterpret_cast<const double2 *>(v + k + 2 * STEP), v3 = *reinterpret_cast<const double2 *>(v + k + 3 * STEP);
      s0 = fma(a0.x, v0.x, s0);
      s1 = fma(a0.y, v0.y, s1);
      s2 = fma(a1.x, v1.x, s2);
      s3 = fma(a1.y, v1.y, s3);
      s4 = fma(a2.x, v2.x, s4);
      s5 = fma(a2.y, v2.y, s5);
      s6 = fma(a3.x, v3.x, s6);
      s7 = fma(a3.y, v3.y, s7);
    }
    for (; k < hi; k += STEP) {
      s0 = fma(a[k], v[k], s0);
      if (k + 1 < hi) s1 = fma(a[k + 1], v[k + 1], s1);
    }
    s = ((s0 + s1) + (s2 + s3)) + ((s4 + s5) + (s6 + s7));
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  }
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (i < n && part == 0 && lane == 0) {
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < ROW_WPR; ++q) t += red[warp + q];
    out[i] = (base ? base[i] : 0.0) + sign * t;
  }
}

// part[blk][j] = sum_{r in 16-row block} T[r][j] v[r]   (j < k): the right-hand side T_R^T t_r of the refinement system
constexpr int TT_ROWS = 16;
__global__ void __launch_bounds__(256) lr_tTt_partial_kernel(const double *__restrict__ T, int ldt, int nB, int k, const double *__restrict__ v, double *__restrict__ part,
                                                              int ldp) {
  const int r0 = blockIdx.x * TT_ROWS, r1 = min(nB, r0 + TT_ROWS);
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    double s0 = 0.0, s1 = 0.0;
    int r = r0;
    for (; r + 1 < r1; r += 2) {
      s0 = fma(T[(size_t) r * ldt + j], v[r], s0);
      s1 = fma(T[(size_t) (r + 1) * ldt + j], v[r + 1], s1);
    }
    if (r < r1) s0 = fma(T[(size_t) r * ldt + j], v[r], s0);
    part[(size_t) blockIdx.x * ldp + j] = s0 + s1;
  }
}

// xfull[bsel[i]] = xB[i] for the rows of B that stay, xfull[A[j]] = z[j]   (xfull zeroed before)
__global__ void lr_scatter_kernel(double *__restrict__ xfull, const int *__restrict__ bsel, int nB, const double *__restrict__ xB, const int *__restrict__ idxA, int na,
                                  const double *__restrict__ z) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nB && bsel[i] >= 0) xfull[bsel[i]] = xB[i];
  if (i < na) xfull[idxA[i]] = z[i];
}
// rB[i] = rfull[bsel[i]] (0 on the rows of D)
__global__ void lr_gather_r_kernel(const double *__restrict__ rfull, const int *__restrict__ bsel, int nB, double *__restrict__ rB) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nB) rB[i] = bsel[i] >= 0 ? rfull[bsel[i]] : 0.0;
}
// out[j] = x + dx of the new passive set in its own order: psrc[j] >= 0 -> position in B, else -(1 + position in A);
// out[np], out[np + 1] = max |dx|, max |x|   (single CTA)
__global__ void lr_finalize_kernel(const int *__restrict__ psrc, int np, const double *__restrict__ xB, const double *__restrict__ dxB, const double *__restrict__ z1,
                                   const double *__restrict__ z2, double *__restrict__ out) {
  __shared__ double s_dx[32], s_x[32];
  double mdx = 0.0, mx = 0.0;
  for (int j = threadIdx.x; j < np; j += blockDim.x) {
    const int p    = psrc[j];
    const double x = p >= 0 ? xB[p] : z1[-1 - p], d = dxB == nullptr ? 0.0 : (p >= 0 ? dxB[p] : z2[-1 - p]);
    out[j] = x + d;
    mdx    = fmax(mdx, fabs(d));
    mx     = fmax(mx, fabs(x + d));
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
    out[np]     = mdx;
    out[np + 1] = mx;
  }
}

// ---- the k x k quasi-definite system: H = L J L^T in shared memory, packed lower, 8-column panels -------------------------
constexpr int LR_KMAX = 200;
constexpr int LR_ST = 512;
constexpr int LR_PW = 8;
constexpr int LR_PKR = LR_KMAX + 8;   // row pitch of the column-major panel copies
// packed lower matrix of order kr = k + 1 (the right-hand side is row k), two panel copies [8][PKR], dinv[k], vec[k], red[8]
constexpr size_t LR_SMALL_SMEM = ((size_t) (LR_KMAX + 1) * (LR_KMAX + 2) / 2 + 2 * (size_t) LR_PKR * LR_PW + 2 * LR_KMAX + 8 + 80) * sizeof(double);

__device__ __forceinline__ int pidx(int i, int j) { return i * (i + 1) / 2 + j; }   // i >= j

// mode 0: H = H0[:k,:k] - diag(M_AA, 0), rhs = H0[:k, k] - [b_A; 0]; factor (rhs as row k: the forward substitution comes with the
//         factorisation), back-substitute, keep the factor in Lg.
//         kold > 0 (a multiple of 8, only with na = 0): the leading kold x kold block of H is the one factorised by the previous call
//         (the removed set only grew and the new members were appended): its factor is reloaded from Lg and only the rows below it
//         are eliminated -- the old panels cost a publish and a few row solves instead of a diagonal factorisation and an update
//         of everything behind them.
// mode 1: reload the factor, rhs = sum_p tpart[p] - [rfull_A; 0], blocked forward + back substitution.
// info: 0 or the 1-based index of a pivot of the wrong sign.
constexpr int LR_NPK_MAX = (LR_KMAX + 1) * (LR_KMAX + 2) / 2;   // Lg: packed factor [LR_NPK_MAX], then the inverse diagonal at a FIXED offset
__global__ void __launch_bounds__(LR_ST) lr_small_kernel(int mode, int kold, int na, int nd, const double *__restrict__ H0, int ldh,
                                                         const double *__restrict__ M, int ldm, const int *__restrict__ idxA, const double *__restrict__ b,
                                                         const double *__restrict__ tpart, int ldp, int ntp, const double *__restrict__ rfull, double *__restrict__ Lg,
                                                         double *__restrict__ z, int *__restrict__ info) {
  extern __shared__ __align__(16) double sm[];
  const int k = na + nd, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int kr = k + 1;
  const int npk = kr * (kr + 1) / 2;
  double *L    = sm;
  double *pan  = sm + (size_t) (LR_KMAX + 1) * (LR_KMAX + 2) / 2;   // [8][PKR]: L[p0 + r][p0 + c] at pan[c * PKR + r]
  double *pas  = pan + (size_t) LR_PKR * LR_PW;                     // sign-folded copy: J_c L[..][p0 + c]
  double *dinv = pas + (size_t) LR_PKR * LR_PW;
  double *vec  = dinv + LR_KMAX;
  double *red  = vec + LR_KMAX;
  double *dgs  = red + 8;   // published diagonal block [8][8] + its inverse diagonal [8]
  if (k == 0) return;
  if (mode == 0) {
    for (int e = tid; e < pidx(kold, 0); e += LR_ST) L[e] = Lg[e];   // rows < kold: the factor of the previous, nested system
    for (int i = tid; i < kold; i += LR_ST) dinv[i] = Lg[LR_NPK_MAX + i];
    for (int j = warp; j < k; j += LR_ST / 32)          // row j of the upper-stored H0, lanes along it: coalesced
      for (int i = max(j, kold) + lane; i < kr; i += 32) {
        double v = H0[(size_t) j * ldh + i];
        if (i < na) v -= M[(size_t) idxA[j] * ldm + idxA[i]];
        if (i == k && j < na) v -= b[idxA[j]];
        L[pidx(i, j)] = v;
      }
    if (tid == 0) L[pidx(k, k)] = 0.0;
    __syncthreads();
    for (int p0 = 0; p0 < k; p0 += LR_PW) {
      const int pw = min(LR_PW, k - p0);
      // warp 0 factors the pw x pw diagonal block in registers (all lanes redundantly: no communication on the pivot chain) and
      // publishes it; then every row i >= p0, owned by one thread, is solved against it
      const int i = p0 + tid;
      const bool old_panel = p0 + LR_PW <= kold;   // already factorised: publish it, solve the new rows against it
      if (warp == 0 && old_panel) {
        if (lane < LR_PW) {
#pragma unroll
          for (int q = 0; q < LR_PW; ++q) dgs[lane * LR_PW + q] = q <= lane ? L[pidx(p0 + lane, p0 + q)] : 0.0;
          dgs[LR_PW * LR_PW + lane] = dinv[p0 + lane];
        }
      } else if (warp == 0) {
        double dg[LR_PW][LR_PW], dv[LR_PW];
        int bad = 0;
#pragma unroll
        for (int r = 0; r < LR_PW; ++r)
#pragma unroll
          for (int q = 0; q <= r; ++q) dg[r][q] = (r < pw && q < pw) ? L[pidx(p0 + r, p0 + q)] : (r == q ? 1.0 : 0.0);
#pragma unroll
        for (int c = 0; c < LR_PW; ++c) {
          const double sgc = (p0 + c) < na ? -1.0 : 1.0;
          double p         = sgc * dg[c][c];
          if (!(p > 0.0)) {
            if (c < pw && bad == 0) bad = p0 + c + 1;
            p = 1.0;
          }
          const double inv = rsqrt(p);
          dg[c][c]         = p * inv;
          dv[c]            = inv;
#pragma unroll
          for (int r = c + 1; r < LR_PW; ++r) dg[r][c] *= sgc * inv;
#pragma unroll
          for (int r = c + 1; r < LR_PW; ++r)
#pragma unroll
            for (int q = c + 1; q <= r; ++q) dg[r][q] = fma(-sgc * dg[r][c], dg[q][c], dg[r][q]);
        }
        if (lane == 0) {
          if (bad != 0 && *info == 0) *info = bad;
#pragma unroll
          for (int r = 0; r < LR_PW; ++r) {
#pragma unroll
            for (int q = 0; q <= r; ++q) dgs[r * LR_PW + q] = dg[r][q];
            dgs[LR_PW * LR_PW + r] = dv[r];
          }
        }
      }
      __syncthreads();
      if (tid < kr - p0) {
        double row[LR_PW];
        const bool old_row = i < kold;   // its entries in this (old) panel are final
        if (old_row) {
#pragma unroll
          for (int c = 0; c < LR_PW; ++c) row[c] = (tid >= pw || c <= tid) ? L[pidx(i, p0 + c)] : 0.0;
        } else if (tid < pw) {
#pragma unroll
          for (int c = 0; c < LR_PW; ++c) row[c] = c <= tid ? dgs[tid * LR_PW + c] : 0.0;
          dinv[i] = dgs[LR_PW * LR_PW + tid];
        } else {
          double dg[LR_PW][LR_PW], dv[LR_PW];
#pragma unroll
          for (int r = 0; r < LR_PW; ++r) {
#pragma unroll
            for (int q = 0; q < r; ++q) dg[r][q] = dgs[r * LR_PW + q];
            dv[r] = dgs[LR_PW * LR_PW + r];
          }
#pragma unroll
          for (int c = 0; c < LR_PW; ++c) row[c] = c < pw ? L[pidx(i, p0 + c)] : 0.0;
          // row . (L_dd J)^-T : forward over the block columns
#pragma unroll
          for (int c = 0; c < LR_PW; ++c) {
            const double sgc = (p0 + c) < na ? -1.0 : 1.0;
            row[c] *= sgc * dv[c];
#pragma unroll
            for (int q = c + 1; q < LR_PW; ++q) row[q] = fma(-sgc * row[c], dg[q][c], row[q]);
          }
        }
#pragma unroll
        for (int c = 0; c < LR_PW; ++c) {
          const double sgc = (p0 + c) < na ? -1.0 : 1.0;
          if (!old_row && c < pw && (tid >= pw || c <= tid)) L[pidx(i, p0 + c)] = row[c];
          pan[c * LR_PKR + tid] = c < pw ? row[c] : 0.0;
          pas[c * LR_PKR + tid] = c < pw ? sgc * row[c] : 0.0;
        }
      }
      __syncthreads();
      // trailing update: H[i][l] -= sum_c J_c L[i][c] L[l][c],  p0 + pw <= l <= i < kr, l < k
      // (a warp takes two rows at a time and splits every dot product in two: the eight dependent FMAs of one element were the
      // longest chain of the kernel -- ncu r02f: a third of its stall samples sat on them)
      const int t0 = p0 + pw, u0 = max(t0, old_panel ? kold : 0);   // behind an old panel only the new rows change
      for (int ii = u0 + 2 * warp; ii < kr; ii += 2 * (LR_ST / 32)) {
        const bool two = ii + 1 < kr;
        double pi0[LR_PW], pi1[LR_PW];
#pragma unroll
        for (int c = 0; c < LR_PW; ++c) {
          pi0[c] = pas[c * LR_PKR + (ii - p0)];
          pi1[c] = two ? pas[c * LR_PKR + (ii + 1 - p0)] : 0.0;
        }
        double *row0 = L + pidx(ii, 0), *row1 = L + pidx(ii + 1, 0);
        const int lmax = min(two ? ii + 1 : ii, k - 1);
        for (int l = t0 + lane; l <= lmax; l += 32) {
          double a0 = 0.0, b0 = 0.0, a1 = 0.0, b1 = 0.0;
#pragma unroll
          for (int c = 0; c < LR_PW; c += 2) {
            const double pl0 = pan[c * LR_PKR + (l - p0)], pl1 = pan[(c + 1) * LR_PKR + (l - p0)];
            a0 = fma(pi0[c], pl0, a0);
            b0 = fma(pi0[c + 1], pl1, b0);
            a1 = fma(pi1[c], pl0, a1);
            b1 = fma(pi1[c + 1], pl1, b1);
          }
          if (l <= ii) row0[l] -= a0 + b0;
          if (two) row1[l] -= a1 + b1;
        }
      }
      __syncthreads();
    }
    for (int e = tid; e < npk; e += LR_ST) Lg[e] = L[e];
    for (int i = tid; i < k; i += LR_ST) {
      Lg[LR_NPK_MAX + i] = dinv[i];
      vec[i]      = L[pidx(k, i)];   // row k = J L^-1 rhs
    }
    __syncthreads();
  } else {
    for (int e = tid; e < npk; e += LR_ST) L[e] = Lg[e];
    for (int i = tid; i < k; i += LR_ST) {
      dinv[i] = Lg[LR_NPK_MAX + i];
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      int p = 0;
      for (; p + 3 < ntp; p += 4) {
        s0 += tpart[(size_t) p * ldp + i];
        s1 += tpart[(size_t) (p + 1) * ldp + i];
        s2 += tpart[(size_t) (p + 2) * ldp + i];
        s3 += tpart[(size_t) (p + 3) * ldp + i];
      }
      for (; p < ntp; ++p) s0 += tpart[(size_t) p * ldp + i];
      vec[i] = ((s0 + s1) + (s2 + s3)) - (i < na ? rfull[idxA[i]] : 0.0);
    }
    __syncthreads();
    // L u = rhs by panels; vec <- v = J u
    for (int p0 = 0; p0 < k; p0 += LR_PW) {
      const int pw = min(LR_PW, k - p0), t0 = p0 + pw;
      if (tid == 0) {
        double u[LR_PW];
#pragma unroll
        for (int c = 0; c < LR_PW; ++c) {
          double s = c < pw ? vec[p0 + c] : 0.0;
#pragma unroll
          for (int q = 0; q < c; ++q)
            if (c < pw) s = fma(-L[pidx(p0 + c, p0 + q)], u[q], s);
          u[c] = c < pw ? s * dinv[p0 + c] : 0.0;
        }
#pragma unroll
        for (int c = 0; c < LR_PW; ++c) {
          pan[c] = u[c];
          if (c < pw) vec[p0 + c] = (p0 + c) < na ? -u[c] : u[c];
        }
      }
      __syncthreads();
      for (int i = t0 + tid; i < k; i += LR_ST) {
        const double *row = L + pidx(i, p0);
        double s = vec[i];
#pragma unroll
        for (int c = 0; c < LR_PW; ++c)
          if (c < pw) s = fma(-row[c], pan[c], s);
        vec[i] = s;
      }
      __syncthreads();
    }
  }
  // L^T z = v by panels from the last: warp c of the first eight reduces the contributions of the rows below the panel to column p0 + c
  const int npan = (k + LR_PW - 1) / LR_PW;
  for (int pb = npan - 1; pb >= 0; --pb) {
    const int p0 = pb * LR_PW, pw = min(LR_PW, k - p0), t0 = p0 + pw;
    if (warp < LR_PW) {
      double s = 0.0;
      if (warp < pw)
        for (int i = t0 + lane; i < k; i += 32) s = fma(L[pidx(i, p0 + warp)], vec[i], s);
      for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
      if (lane == 0) red[warp] = s;
    }
    __syncthreads();
    if (tid == 0) {
      double zc[LR_PW];
#pragma unroll
      for (int c = LR_PW - 1; c >= 0; --c) {
        double s = c < pw ? vec[p0 + c] - red[c] : 0.0;
#pragma unroll
        for (int q = c + 1; q < LR_PW; ++q)
          if (q < pw) s = fma(-L[pidx(p0 + q, p0 + c)], zc[q], s);
        zc[c] = c < pw ? s * dinv[p0 + c] : 0.0;
      }
#pragma unroll
      for (int c = 0; c < LR_PW; ++c)
        if (c < pw) vec[p0 + c] = zc[c];
    }
    __syncthreads();
  }
  for (int i = tid; i < k; i += LR_ST) z[i] = vec[i];
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
  if ((e = cudaFuncSetAttribute(trinv_step12_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) GEMM_SMEM)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(gemm_tn_splitk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) GEMM_SMEM)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(syrk_splitk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) GEMM_SMEM)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(lr_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) LR_SMALL_SMEM)) != cudaSuccess) return e;
  done[dev] = true;
  return cudaSuccess;
}

}   // namespace

int lowrank_kmax() { return LR_KMAX; }

// out = A v over the upper triangle of the n x n row-major A (rows i, columns >= i)
int row_dot_upper(ncm_sd_gpu_ctx *c, const double *dA, int lda, int n, const double *dv, double *dOut) {
  row_dot_kernel<ROW_UPPER><<<(n * ROW_WPR + 7) / 8, 256, 0, c->stream>>>(dA, lda, n, dv, nullptr, 1.0, dOut);
  c->n_launches++;
  NCM_CUDA_OK(c, cudaGetLastError());
  return NCM_SD_GPU_OK;
}
// doubles of device scratch lowrank_solve needs for the split-K partial sums of an n_B x ldv product and of the k x k matrix
size_t lowrank_part_doubles(int n, int ldv) {
  const size_t tpart = (size_t) ((n + KC_T - 1) / KC_T) * n * ldv;
  const size_t hpart = (size_t) ((n + KC_H - 1) / KC_H + 1) * ldv * ldv;
  const size_t ttp   = (size_t) ((n + TT_ROWS - 1) / TT_ROWS) * ldv;
  return tpart + hpart + ttp;
}

// lower triangle := upper triangle transposed (n x n, row-major, ld)
int symmetrize_upper(ncm_sd_gpu_ctx *c, int n, double *dM, int ld) {
  const int nt = (n + 31) / 32;
  transpose_upper_kernel<<<nt * (nt + 1) / 2, dim3(32, 8), 0, c->stream>>>(dM, dM, ld, n, nt);
  c->n_launches++;
  NCM_CUDA_OK(c, cudaGetLastError());
  return NCM_SD_GPU_OK;
}

// W = U^-1 for the upper-triangular factor U (n x n, row-major, ld); S is scratch of the same shape.  Wt (optional) receives W^T
// (lower triangle and diagonal; its upper triangle is not written).
int trinv_upper_on(ncm_sd_gpu_ctx *c, cudaStream_t st, int n, const double *dU, double *dW, double *dS, int ld, double *dWt) {
  NCM_CUDA_OK(c, set_smem_attrs());
  NCM_CUDA_OK(c, cudaMemsetAsync(dW, 0, (size_t) n * ld * sizeof(double), st));
  trinv_diag_kernel<<<(n + 63) / 64, 256, 0, st>>>(dU, dW, ld, n);
  c->n_launches++;
  for (int s = 64; s < n; s *= 2) {
    const int pairs = (n - s + 2 * s - 1) / (2 * s);
    if (s <= 128) {
      trinv_step12_kernel<<<dim3((s + GT - 1) / GT, 1, pairs), GTHREADS, GEMM_SMEM, st>>>(dU, dW, dS, ld, n, s);
      c->n_launches++;
      continue;
    }
    dim3 grid((s + GT - 1) / GT, (s + GT - 1) / GT, pairs);
    trinv_step1_kernel<<<grid, GTHREADS, GEMM_SMEM, st>>>(dU, dW, dS, ld, n, s);
    trinv_step2_kernel<<<grid, GTHREADS, GEMM_SMEM, st>>>(dW, dS, ld, n, s);
    c->n_launches += 2;
  }
  if (dWt != nullptr) {
    const int nt = (n + 31) / 32;
    transpose_upper_kernel<<<nt * (nt + 1) / 2, dim3(32, 8), 0, st>>>(dW, dWt, ld, n, nt);
    c->n_launches++;
  }
  NCM_CUDA_OK(c, cudaGetLastError());
  return NCM_SD_GPU_OK;
}

int trinv_upper(ncm_sd_gpu_ctx *c, int n, const double *dU, double *dW, double *dS, int ld, double *dWt) {
  return trinv_upper_on(c, c->stream, n, dU, dW, dS, ld, dWt);
}

// Solve M[P,P] x = b[P] for P = (B \ D) u A through the base inverse W (see the header).  M is full symmetric.  Device index arrays
// (all ascending): idxB [nB], idxA [na], posD [nd]; bsel [nB] = idxB[i] or -1 on the rows of D; psrc [np] = position of P[j] in B or
// -(1 + position in A).  The np results in the order of P, then {max |dx|, max |x|}, are left in bufs.out.
// kold: the first kold entries of posD are, in this order, the whole removed set of the previous solve from this base (0: unrelated).
int lowrank_solve(ncm_sd_gpu_ctx *c, const double *dM, int ldm, int n, const double *db, int nB, int na, int nd, int np, const LowrankBufs &w, int ldv,
                  bool refine, int kold) {
  NCM_CUDA_OK(c, set_smem_attrs());
  const int k = na + nd, kc = k + 1;
  cudaStream_t st = c->stream;
  const size_t tstride = (size_t) nB * ldv, hstride = (size_t) ldv * ldv;
  const int ntc = (nB + KC_T - 1) / KC_T, nhc = (nB + KC_H - 1) / KC_H, ntp = (nB + TT_ROWS - 1) / TT_ROWS;
  double *Tpart = w.part, *Hpart = Tpart + (size_t) ntc * tstride, *H0 = Hpart + (size_t) nhc * hstride, *tTtpart = H0 + hstride;
  const int ctiles = (kc + GT - 1) / GT, rtiles = (nB + GT - 1) / GT;
  NCM_CUDA_OK(c, cudaMemsetAsync(w.info, 0, sizeof(int), st));
  if (refine) NCM_CUDA_OK(c, cudaMemsetAsync(w.xfull, 0, (size_t) n * sizeof(double), st));
  if (k == 0) {
    // the base set itself: x_B = W (W^T b_B) -- two triangular matrix-vector products instead of the forward / back substitution
    // (the fused factorisation then runs without its right-hand side column and its back-substitution phase)
    lr_gather_r_kernel<<<(nB + 255) / 256, 256, 0, st>>>(db, w.bsel, nB, w.rB);
    row_dot_kernel<ROW_LOWER><<<(nB * ROW_WPR + 7) / 8, 256, 0, st>>>(w.Wt, ldm, nB, w.rB, nullptr, 1.0, w.tb);   // t_b = W^T b_B: kept for the base's later sets
    row_dot_kernel<ROW_UPPER><<<(nB * ROW_WPR + 7) / 8, 256, 0, st>>>(w.W, ldm, nB, w.tb, nullptr, 1.0, w.xB);
    c->n_launches += 3;
    if (refine) {
      lr_scatter_kernel<<<(nB + 255) / 256, 256, 0, st>>>(w.xfull, w.bsel, nB, w.xB, w.idxA, 0, w.z);
      row_dot_kernel<ROW_FULL><<<(n * ROW_WPR + 7) / 8, 256, 0, st>>>(dM, ldm, n, w.xfull, db, -1.0, w.rfull);
      lr_gather_r_kernel<<<(nB + 255) / 256, 256, 0, st>>>(w.rfull, w.bsel, nB, w.rB);
      row_dot_kernel<ROW_LOWER><<<(nB * ROW_WPR + 7) / 8, 256, 0, st>>>(w.Wt, ldm, nB, w.rB, nullptr, 1.0, w.tr);
      row_dot_kernel<ROW_UPPER><<<(nB * ROW_WPR + 7) / 8, 256, 0, st>>>(w.W, ldm, nB, w.tr, nullptr, 1.0, w.dxB);
      c->n_launches += 5;
    }
    lr_finalize_kernel<<<1, 1024, 0, st>>>(w.psrc, np, w.xB, refine ? w.dxB : nullptr, w.z, w.z2, w.out);
    c->n_launches++;
    NCM_CUDA_OK(c, cudaGetLastError());
    return NCM_SD_GPU_OK;
  }
  if (na == 0 && w.tb_valid) {
    lr_gather_T_kernel<<<dim3((nB + 31) / 32, (kc + 31) / 32), dim3(32, 8), 0, st>>>(w.W, ldm, nB, w.posD, k, w.tb, w.T, ldv);
    c->n_launches++;
  } else {
    lr_gather_V_kernel<<<(nB + 7) / 8, dim3(32, 8), 0, st>>>(dM, ldm, db, w.bsel, w.idxB, nB, w.idxA, na, w.posD, nd, w.V, ldv);
    gemm_tn_splitk_kernel<<<dim3(ctiles, rtiles, ntc), GTHREADS, GEMM_SMEM, st>>>(w.W, ldm, w.V, ldv, Tpart, ldv, tstride, nB, kc, nB);
    splitk_reduce_kernel<<<dim3((kc + 63) / 64, nB), 64, 0, st>>>(Tpart, tstride, ldv, nB, kc, nB, w.T);
    c->n_launches += 3;
  }
  if (k > 0) {
    syrk_splitk_kernel<<<dim3(ctiles * (ctiles + 1) / 2, nhc), GTHREADS, GEMM_SMEM, st>>>(w.T, ldv, Hpart, ldv, hstride, kc, nB);
    hreduce_kernel<<<dim3((kc + 63) / 64, kc), 64, 0, st>>>(Hpart, hstride, nhc, ldv, kc, H0);
    lr_small_kernel<<<1, LR_ST, LR_SMALL_SMEM, st>>>(0, na == 0 ? (kold & ~7) : 0, na, nd, H0, ldv, dM, ldm, w.idxA, db, nullptr, 0, 0, nullptr, w.Lg, w.z, w.info);
    c->n_launches += 3;
  }
  lr_y_kernel<<<(nB + 7) / 8, 256, 0, st>>>(w.T, ldv, nB, k, w.z, w.T + k, ldv, w.y);
  row_dot_kernel<ROW_UPPER><<<(nB * ROW_WPR + 7) / 8, 256, 0, st>>>(w.W, ldm, nB, w.y, nullptr, 1.0, w.xB);
  c->n_launches += 2;
  if (!refine) {   // the base has already shown a negligible correction (see nnls.cu): x is final
    lr_finalize_kernel<<<1, 1024, 0, st>>>(w.psrc, np, w.xB, nullptr, w.z, nullptr, w.out);
    c->n_launches++;
    NCM_CUDA_OK(c, cudaGetLastError());
    return NCM_SD_GPU_OK;
  }
  lr_scatter_kernel<<<(std::max(nB, na) + 255) / 256, 256, 0, st>>>(w.xfull, w.bsel, nB, w.xB, w.idxA, na, w.z);
  // one step of iterative refinement on the true system: r = b - M xfull
  row_dot_kernel<ROW_FULL><<<(n * ROW_WPR + 7) / 8, 256, 0, st>>>(dM, ldm, n, w.xfull, db, -1.0, w.rfull);
  lr_gather_r_kernel<<<(nB + 255) / 256, 256, 0, st>>>(w.rfull, w.bsel, nB, w.rB);
  row_dot_kernel<ROW_LOWER><<<(nB * ROW_WPR + 7) / 8, 256, 0, st>>>(w.Wt, ldm, nB, w.rB, nullptr, 1.0, w.tr);   // t_r = W^T r_B
  c->n_launches += 4;
  if (k > 0) {
    lr_tTt_partial_kernel<<<ntp, 256, 0, st>>>(w.T, ldv, nB, k, w.tr, tTtpart, ldv);
    lr_small_kernel<<<1, LR_ST, LR_SMALL_SMEM, st>>>(1, 0, na, nd, H0, ldv, dM, ldm, w.idxA, db, tTtpart, ldv, ntp, w.rfull, w.Lg, w.z2, w.info);
    c->n_launches += 2;
  }
  lr_y_kernel<<<(nB + 7) / 8, 256, 0, st>>>(w.T, ldv, nB, k, w.z2, w.tr, 1, w.y);
  row_dot_kernel<ROW_UPPER><<<(nB * ROW_WPR + 7) / 8, 256, 0, st>>>(w.W, ldm, nB, w.y, nullptr, 1.0, w.dxB);
  lr_finalize_kernel<<<1, 1024, 0, st>>>(w.psrc, np, w.xB, w.dxB, w.z, w.z2, w.out);
  c->n_launches += 3;
  NCM_CUDA_OK(c, cudaGetLastError());
  return NCM_SD_GPU_OK;
}
