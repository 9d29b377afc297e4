// 64 x 64 DMMA.8x8x4 GEMM tile shared by the low-rank NNLS solves (lowrank.cu) and the distributed Cholesky (dist_chol.cu).
#pragma once
#include "common.cuh"

namespace ncm_gemm {

// ---- DMMA GEMM tile: C[i0:i0+64, j0:j0+64] = alpha * op(A) B over k in [k_lo, k_hi) -------------------------------------
// A: TRANSA ? (K x M row-major, op(A) = A^T) : (M x K row-major);  B: K x N row-major;  C: M x N row-major.
constexpr int GT = 64, GBK = 16, GTHREADS = 128, GSTAGES = 3;
constexpr int GPA = GBK + 4;                    // pitch of the [64][16] A tile (NN): lc * 20 + lr hits 16 distinct bank pairs per half-warp
constexpr int GPB = GT + 4;                     // pitch of the [16][64] tiles (B, and A when TRANSA)
constexpr int GSLAB_A = GT * GPA;               // 1280 doubles >= GBK * GPB = 1088
constexpr int GSLAB_B = GBK * GPB;
constexpr size_t GEMM_SMEM = (size_t) GSTAGES * (GSLAB_A + GSLAB_B) * sizeof(double);   // 56.8 KB: three CTAs per SM

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

}   // namespace ncm_gemm
