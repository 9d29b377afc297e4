// Blocked Cholesky solve  M x = b  for the NNLS passive-set systems.
//
// Replaces ncm_matrix_cholesky_solve (dposv 'U', ncm_matrix.c:1199-1210) as called from
// _ncm_nnls_solve_normal_cholesky (ncm_nnls.c:655-666): M = U^T U with U upper triangular,
// row-major, only the upper triangle of M is read.  Right-looking by block rows of NB = 64:
//   1. chol_diag_kernel   factor the 64 x 64 diagonal block (one CTA, block in shared memory) in 8-column
//                         sub-blocks; the right-hand side rides along as a 65th column (y_k = U_kk^-T b_k)
//   2. chol_panel_kernel  U[k, k+1:] = U_kk^-T M[k, k+1:], one thread per column, forward substitution in
//                         registers in 8-row sub-blocks (8 independent accumulators while sweeping the
//                         solved part), U_kk broadcast from shared memory; also
//                         b[j] -= sum_r U[r][j] y_r  (the forward solve of the right-hand side folded in)
//   3. ata_kernel         trailing update M[k+1:, k+1:] -= U[k, k+1:]^T U[k, k+1:]  on DMMA
// followed by a blocked back substitution U x = y (one launch per block row; the 64 x 64 triangular solve
// is done by a single warp with shuffles, the rows above by the other CTAs).
//
// These kernels are latency-bound (measured on the box, tools/microbench/latency.cu: DFMA 8.3 cycles,
// rsqrt 67, shuffle 27, LDS 29, __syncthreads 29): the serial pivot chain rsqrt -> mul -> fma costs about
// 83 cycles per column whatever the layout, so the design goal is to keep everything else (global loads,
// barriers, shared-memory round trips) off that chain.
#include <cstdlib>
#include <string>
#include "ctx.h"

int dsyrk_ata_general(ncm_sd_gpu_ctx *c, int K, int n, const double *dP, int ldp, double *dC, int ldc, double alpha, double beta);
int dsyrk_ata_first_rows64(ncm_sd_gpu_ctx *c, int K, int n, const double *dP, int ldp, double *dC, int ldc);
int dsyrk_ata_general_on(ncm_sd_gpu_ctx *c, cudaStream_t st, int K, int n, const double *dP, int ldp, double *dC, int ldc, double alpha, double beta);
int dsyrk_ata_first_rows64_on(ncm_sd_gpu_ctx *c, cudaStream_t st, int K, int n, const double *dP, int ldp, double *dC, int ldc);
int dsyrk_ata_head_rows_on(ncm_sd_gpu_ctx *c, cudaStream_t st, int K, int n, int rows, const double *dP, int ldp, double *dC, int ldc);

namespace {

constexpr int NB = 64;
#ifdef CHOL_PROBE   // phase timestamps for tools/microbench (compiled out of the library)
__device__ long long g_probe[128];
#define NCM_PROBE(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) g_probe[i] = clock64(); } while (0)
#else
#define NCM_PROBE(i) do { } while (0)
#endif
constexpr int DS = NB + 2;   // row pitch of the shared block: columns 0..63, rhs column 64, pad

// M[k0:k0+nb, k0:k0+nb] -> U_kk in place; dinv[k0 + r] = 1 / U_rr; rhs[k0:k0+nb] -> y_k;
// info = first non-positive pivot (1-based), 0 otherwise.
//
// Per 8-column sub-block:  every thread that owns one of the remaining columns (or the rhs) factors the
// 8 x 8 diagonal sub-block redundantly in its own registers (no communication on the pivot chain), solves
// its column against it and writes it back; then all 256 threads apply the rank-8 update to the rest.
__global__ void __launch_bounds__(256) chol_diag_kernel(double *__restrict__ M, int ldm, int n, int k0, double *__restrict__ rhs,
                                                        double *__restrict__ dinv, int *__restrict__ info) {
  __shared__ __align__(16) double S[NB][DS];
  __shared__ double sDinv[NB];
  __shared__ int sBad;
  const int tid = threadIdx.x;
  const int nb  = min(NB, n - k0);
  if (tid == 0) sBad = 0;
  NCM_PROBE(0);
  {
    // 256 threads x 16 independent loads: row = tid / 4, columns (tid % 4) * 16 .. + 15
    if (nb == NB) {
      // full block: 8 independent 16-byte loads per thread, a warp reads one 512-byte row per instruction
      // (the lower triangle is loaded and ignored)
      double2 v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int chunk = q * 256 + tid;   // 32 chunks of 16 B per row
        v[q] = *reinterpret_cast<const double2 *>(M + (size_t) (k0 + (chunk >> 5)) * ldm + k0 + (chunk & 31) * 2);
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int chunk = q * 256 + tid;
        *reinterpret_cast<double2 *>(&S[chunk >> 5][(chunk & 31) * 2]) = v[q];
      }
    } else {
      const int r = tid >> 2, cb = (tid & 3) * 16;
      double v[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const int cidx = cb + q;
        v[q] = (r < nb && cidx < nb && cidx >= r) ? M[(size_t) (k0 + r) * ldm + k0 + cidx] : ((r == cidx && r >= nb) ? 1.0 : 0.0);
      }
#pragma unroll
      for (int q = 0; q < 16; ++q) S[r][cb + q] = v[q];
    }
    if (tid < NB) S[tid][NB] = (tid < nb && rhs != nullptr) ? rhs[k0 + tid] : 0.0;
  }
  __syncthreads();
  NCM_PROBE(1);

#pragma unroll 1
  for (int kb = 0; kb < NB / 8; ++kb) {
    const int b0    = 8 * kb;
    const int ncols = NB - b0 + 1;                    // remaining columns + rhs
    const bool solver = tid >= 8 && tid < ncols;      // threads 8 .. ncols-1: one remaining column (or the rhs) each
    const bool writer = tid >= 224 && tid < 232;      // an otherwise idle warp writes the factored sub-block back
    // Divergent if/else paths cost ~100 cycles per reconvergence point here, so both roles run the same
    // straight-line code and only their stores are predicated.
    if (solver || writer) {
      const int ci = (tid == ncols - 1) ? NB : (solver ? b0 + tid : b0);
      // 8 x 8 diagonal sub-block (upper) in registers; every thread reads the same addresses (broadcast)
      double dgl[8][8];
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int q = r; q < 8; ++q) dgl[r][q] = S[b0 + r][b0 + q];
      double col[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) col[r] = S[b0 + r][ci];
      double inv[8];
      int bad = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        double piv = dgl[j][j];
        if (!(piv > 0.0)) {
          bad = (bad == 0) ? k0 + b0 + j + 1 : bad;
          piv = 1.0;
        }
        inv[j]    = rsqrt(piv);
        dgl[j][j] = piv * inv[j];
#pragma unroll
        for (int q = j + 1; q < 8; ++q) dgl[j][q] *= inv[j];
#pragma unroll
        for (int r = j + 1; r < 8; ++r)
#pragma unroll
          for (int q = r; q < 8; ++q) dgl[r][q] = fma(-dgl[j][r], dgl[j][q], dgl[r][q]);
      }
      // x = D^-T s_c  (the writer threads run it on a dummy column and discard it)
      double x[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        double t = col[r];
#pragma unroll
        for (int s = 0; s < r; ++s) t = fma(-dgl[s][r], x[s], t);
        x[r] = t * inv[r];
      }
      const int wq = tid - 224;   // writer: column wq of the factor
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        double v = x[r];
        if (writer) {
#pragma unroll
          for (int q = r; q < 8; ++q) v = (q == wq) ? dgl[r][q] : v;
        }
        const int cc = writer ? b0 + wq : ci;
        if (solver || (writer && wq >= r)) S[b0 + r][cc] = v;
      }
      if (writer && wq == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) sDinv[b0 + j] = inv[j];
        if (bad != 0 && sBad == 0) sBad = bad;
      }
    }
    __syncthreads();
    NCM_PROBE(2 + 2 * kb);
    // rank-8 update of the remaining upper triangle and of the rhs column: 16 x 16 threads, cyclic 4 x 4 tiles,
    // straight-line with predicated stores
    {
      const int base = b0 + 8;
      const int ty = tid >> 4, tx = tid & 15;
      if (base < NB) {
        double xr[4][8], xc[4][8], xh[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = min(base + ty + 16 * i, NB - 1);
          const int c = min(base + tx + 16 * i, NB - 1);
#pragma unroll
          for (int s = 0; s < 8; ++s) {
            xr[i][s] = S[b0 + s][r];
            xc[i][s] = S[b0 + s][c];
          }
        }
#pragma unroll
        for (int s = 0; s < 8; ++s) xh[s] = S[b0 + s][NB];
        // all 20 accumulators are loaded before any store, so the 8-deep FMA chains overlap
        double acc[4][5];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rc = min(base + ty + 16 * i, NB - 1);
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = S[rc][min(base + tx + 16 * j, NB - 1)];
          acc[i][4] = S[rc][NB];
        }
#pragma unroll
        for (int s = 0; s < 8; ++s)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fma(-xr[i][s], xc[j][s], acc[i][j]);
            acc[i][4] = fma(-xr[i][s], xh[s], acc[i][4]);
          }
        // (operand rows b0 .. b0+7 are not written in this phase and every thread owns its (r, c) entries)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r   = base + ty + 16 * i;
          const bool rv = r < NB;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c = base + tx + 16 * j;
            if (rv && c < NB && c >= r) S[r][c] = acc[i][j];
          }
          if (rv && tx == 0) S[r][NB] = acc[i][4];
        }
      }
    }
    __syncthreads();
    NCM_PROBE(3 + 2 * kb);
  }

  {
    const int r = tid >> 2, cb = (tid & 3) * 16;
    if (r < nb) {
      double *dst = M + (size_t) (k0 + r) * ldm + k0 + cb;
#pragma unroll
      for (int q = 0; q < 16; q += 2) {
        const int cidx = cb + q;
        if (cidx >= r && cidx + 1 < nb)
          *reinterpret_cast<double2 *>(dst + q) = make_double2(S[r][cidx], S[r][cidx + 1]);
        else {
          if (cidx < nb && cidx >= r) dst[q] = S[r][cidx];
          if (cidx + 1 < nb && cidx + 1 >= r) dst[q + 1] = S[r][cidx + 1];
        }
      }
    }
  }
  if (tid < nb) {
    dinv[k0 + tid] = sDinv[tid];
    if (rhs != nullptr) rhs[k0 + tid] = S[tid][NB];
  }
  if (tid == 0 && sBad != 0) atomicCAS(info, 0, sBad);
  NCM_PROBE(18);
}

// one thread per trailing column j: x = U_kk^-T M[k0:k0+nb, j]; rhs[j] -= x . y_k
__global__ void __launch_bounds__(128) chol_panel_kernel(double *__restrict__ M, int ldm, int n, int k0, double *__restrict__ rhs,
                                                         const double *__restrict__ dinv) {
  __shared__ __align__(16) double sU[NB][NB];   // sU[s][r] = U_kk[s][r] (upper)
  __shared__ double sD[NB];
  __shared__ double sY[NB];
  const int tid = threadIdx.x;
  NCM_PROBE(49);
  {
    // 128 threads x 16 independent 16-byte loads; a warp reads one 512-byte row per instruction (coalesced)
    double2 v[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int chunk = q * 128 + tid;               // 0 .. 2047, 32 chunks of 16 B per row
      v[q] = *reinterpret_cast<const double2 *>(M + (size_t) (k0 + (chunk >> 5)) * ldm + k0 + (chunk & 31) * 2);
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int chunk = q * 128 + tid;
      *reinterpret_cast<double2 *>(&sU[chunk >> 5][(chunk & 31) * 2]) = v[q];
    }
  }
  if (tid < NB) {
    sD[tid] = dinv[k0 + tid];
    sY[tid] = rhs != nullptr ? rhs[k0 + tid] : 0.0;
  }
  NCM_PROBE(50);
  const int j    = k0 + NB + blockIdx.x * blockDim.x + tid;
  const bool jv  = j < n;
  const int jj   = jv ? j : n - 1;
  double t[8];
#pragma unroll
  for (int rr = 0; rr < 8; ++rr) t[rr] = M[(size_t) (k0 + rr) * ldm + jj];
  __syncthreads();
  NCM_PROBE(51);
  double x[NB];
  double dot = 0.0;
#pragma unroll
  for (int blk = 0; blk < NB / 8; ++blk) {
    double tn[8];
    if (blk + 1 < NB / 8) {
#pragma unroll
      for (int rr = 0; rr < 8; ++rr) tn[rr] = M[(size_t) (k0 + 8 * (blk + 1) + rr) * ldm + jj];   // prefetch next sub-block
    }
    // sweep the already solved unknowns: 8 independent accumulators
#pragma unroll
    for (int s = 0; s < 8 * blk; ++s) {
#pragma unroll
      for (int rr = 0; rr < 8; ++rr) t[rr] = fma(-sU[s][8 * blk + rr], x[s], t[rr]);
    }
    // triangular part of the sub-block
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) {
      const int r = 8 * blk + rr;
#pragma unroll
      for (int ss = 0; ss < rr; ++ss) t[rr] = fma(-sU[8 * blk + ss][r], x[8 * blk + ss], t[rr]);
      x[r] = t[rr] * sD[r];
      dot  = fma(x[r], sY[r], dot);
    }
    if (jv) {
#pragma unroll
      for (int rr = 0; rr < 8; ++rr) M[(size_t) (k0 + 8 * blk + rr) * ldm + j] = x[8 * blk + rr];
    }
    if (blk + 1 < NB / 8) {
#pragma unroll
      for (int rr = 0; rr < 8; ++rr) t[rr] = tn[rr];
    }
    NCM_PROBE(52 + blk);
  }
  if (jv && rhs != nullptr) rhs[j] -= dot;
  NCM_PROBE(60);
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
    // 16 independent loads per thread; a warp reads 32 consecutive doubles of one row per instruction
    double v[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int e = q * 256 + tid, r = e >> 6, cidx = e & 63;
      v[q] = (r < nb && cidx < nb && cidx >= r) ? M[(size_t) (k0 + r) * ldm + k0 + cidx] : 0.0;
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int e = q * 256 + tid;
      sU[e >> 6][e & 63] = v[q];
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
  // CTA 0: update own rows with x of the next block (all loads of the warp's 8 rows issued up front) ...
  {
    double u0[8], u1[8], yy[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int r     = warp + 8 * q;
      const double *u = M + (size_t) (k0 + min(r, nb - 1)) * ldm + k1;
      u0[q]           = (r < nb && lane < nb1) ? u[lane] : 0.0;
      u1[q]           = (r < nb && lane + 32 < nb1) ? u[lane + 32] : 0.0;
      yy[q]           = (r < nb && lane == 0) ? y[k0 + r] : 0.0;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      double sacc = fma(u1[q], sx[lane + 32], u0[q] * sx[lane]);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, off);
      const int r = warp + 8 * q;
      if (lane == 0 && r < nb) sy[r] = yy[q] - sacc;
    }
  }
  __syncthreads();
  // ... then one warp solves the 64 x 64 upper system: lane holds rows lane and lane + 32
  if (warp == 0) {
    double y0 = lane < nb ? sy[lane] : 0.0, y1 = lane + 32 < nb ? sy[lane + 32] : 0.0;
    const double d0 = lane < nb ? dinv[k0 + lane] : 1.0, d1 = lane + 32 < nb ? dinv[k0 + lane + 32] : 1.0;
#pragma unroll 8
    for (int r = NB - 1; r >= 32; --r) {
      const double xr = __shfl_sync(0xffffffffu, y1 * d1, r - 32);
      if (lane + 32 == r) y1 = xr;
      if (lane + 32 < r) y1 = fma(-sU[lane + 32][r], xr, y1);
      y0 = fma(-sU[lane][r], xr, y0);
    }
#pragma unroll 8
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
  // Panels are factored GP at a time: before panel p of a group is factored, only its own 64 rows are brought up to date
  // with the p panels of the group already done (a thin update, K = 64 p); after the last one the rest of the matrix receives
  // the whole group in one update of depth K = 64 GP -- GP times fewer passes over the trailing matrix and a deeper DMMA
  // pipeline in ata_kernel (which matters at n >> 4096, where this path runs: the K = 64 updates ran at 17 TFLOP/s at n = 16384).
  static const int gp_env = getenv("NCM_SD_GPU_CHOL_GROUP") != nullptr ? atoi(getenv("NCM_SD_GPU_CHOL_GROUP")) : 0;
  const int GP = gp_env > 0 ? gp_env : (n >= 12288 ? 8 : (n >= 6144 ? 4 : 2));   // measured at n = 16384: 68.0 / 62.2 / 60.1 ms for 2 / 4 / 8
  // Look-ahead (W = 64 GP columns per group, a multiple of the 128-wide update tile): the update of group g is split into the
  // head strip -- the W block rows group g + 1 factors -- and the tail.  Panels and head strips run on a highest-priority stream,
  // the tails (the bulk of the flops) on the context stream, so that the latency-bound panel kernels of group g + 1 execute
  // while the tail of group g keeps the tensor pipe busy:
  //     hi :  panels(g)  -> [wait tail(g-1)] head(g) -> panels(g+1) -> ...
  //     lo :               [wait panels(g)]  tail(g)  -> tail(g+1)  -> ...
  // ($NCM_SD_GPU_CHOL_LOOKAHEAD=0 keeps everything on one stream.)
  static const bool la_env = getenv("NCM_SD_GPU_CHOL_LOOKAHEAD") == nullptr || atoi(getenv("NCM_SD_GPU_CHOL_LOOKAHEAD")) != 0;
  const int W              = GP * NB;
  const bool lookahead     = la_env && (W % 128 == 0) && n >= 4 * W;
  cudaStream_t hi = c->stream, lo = c->stream;
  if (lookahead) {
    if (c->stream_hi == nullptr) {
      int least = 0, greatest = 0;
      NCM_CUDA_OK(c, cudaDeviceGetStreamPriorityRange(&least, &greatest));
      NCM_CUDA_OK(c, cudaStreamCreateWithPriority(&c->stream_hi, cudaStreamNonBlocking, greatest));
      NCM_CUDA_OK(c, cudaEventCreateWithFlags(&c->ev_panel, cudaEventDisableTiming));
      NCM_CUDA_OK(c, cudaEventCreateWithFlags(&c->ev_tail, cudaEventDisableTiming));
    }
    hi = c->stream_hi;
    NCM_CUDA_OK(c, cudaEventRecord(c->ev_tail, lo));   // fork: everything queued on the context stream so far (gather, memset) comes first
    NCM_CUDA_OK(c, cudaStreamWaitEvent(hi, c->ev_tail, 0));
  }
  bool done = false;
  for (int kb = 0; kb < nblk && !done; kb += GP) {
    const int k0 = kb * NB;
    for (int p = 0; p < GP; ++p) {
      const int kp = k0 + p * NB;
      if (kp >= n) {
        done = true;
        break;
      }
      if (p > 0) {
        int rc = dsyrk_ata_first_rows64_on(c, hi, p * NB, n - kp, dM + (size_t) k0 * ldm + kp, ldm, dM + (size_t) kp * ldm + kp, ldm);
        if (rc != NCM_SD_GPU_OK) return rc;
      }
      chol_diag_kernel<<<1, 256, 0, hi>>>(dM, ldm, n, kp, dRhs, dDinv, dInfo);
      c->n_launches++;
      const int m = n - kp - NB;
      if (m <= 0) {
        done = true;
        break;
      }
      chol_panel_kernel<<<(m + 127) / 128, 128, 0, hi>>>(dM, ldm, n, kp, dRhs, dDinv);
      c->n_launches++;
    }
    if (done) break;
    const int kg = k0 + W;
    const int mg = n - kg;
    if (mg <= 0) break;
    const double *dP = dM + (size_t) k0 * ldm + kg;
    double *dC       = dM + (size_t) kg * ldm + kg;
    if (!lookahead) {
      int rc = dsyrk_ata_general(c, W, mg, dP, ldm, dC, ldm, -1.0, 1.0);
      if (rc != NCM_SD_GPU_OK) return rc;
      continue;
    }
    NCM_CUDA_OK(c, cudaEventRecord(c->ev_panel, hi));
    NCM_CUDA_OK(c, cudaStreamWaitEvent(lo, c->ev_panel, 0));
    NCM_CUDA_OK(c, cudaStreamWaitEvent(hi, c->ev_tail, 0));   // tail(g-1) wrote the rows head(g) updates (first group: the fork event)
    int rc = dsyrk_ata_head_rows_on(c, hi, W, mg, mg < W ? ((mg + 127) / 128) * 128 : W, dP, ldm, dC, ldm);
    if (rc != NCM_SD_GPU_OK) return rc;
    if (mg > W) {
      rc = dsyrk_ata_general_on(c, lo, W, mg - W, dP + W, ldm, dC + (size_t) W * ldm + W, ldm, -1.0, 1.0);
      if (rc != NCM_SD_GPU_OK) return rc;
    }
    NCM_CUDA_OK(c, cudaEventRecord(c->ev_tail, lo));
  }
  if (lookahead) {   // join: the back substitution (and whatever the caller queues next) follows both streams
    NCM_CUDA_OK(c, cudaEventRecord(c->ev_panel, hi));
    NCM_CUDA_OK(c, cudaStreamWaitEvent(lo, c->ev_panel, 0));
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

// Dispatch: the single-launch data-flow kernel (chol_fused.cu) up to its maximum order, the launch-per-step
// schedule above beyond it (there the DMMA trailing updates carry the time).  $NCM_SD_GPU_CHOL=legacy forces the latter.
int dpotrf_upper_solve_any(ncm_sd_gpu_ctx *c, int n, double *dM, int ldm, double *dRhs, double *dDinv, int *dInfo, int *info_host) {
  static const bool legacy = [] {
    const char *e = getenv("NCM_SD_GPU_CHOL");
    return e != nullptr && std::string(e) == "legacy";
  }();
  if (!legacy && n <= chol_fused_max_n()) return dpotrf_upper_solve_fused(c, n, dM, ldm, dRhs, info_host);
  return dpotrf_upper_solve(c, n, dM, ldm, dRhs, dDinv, dInfo, info_host);
}
