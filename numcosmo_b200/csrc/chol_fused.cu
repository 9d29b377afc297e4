// Single-launch Cholesky solve  M x = b  for the NNLS passive-set systems (n <= 4096): one persistent
// cooperative kernel, one CTA per SM, tile-level data flow instead of one launch per step.
//
// Replaces ncm_matrix_cholesky_solve (dposv 'U', ncm_matrix.c:1199-1210) as called from
// _ncm_nnls_solve_normal_cholesky (ncm_nnls.c:655-666), like chol.cu, whose three-launches-per-block-row
// schedule spends most of its time in launch gaps and under-filled grids at |P| ~ 2000 (profiles/r01b).
//
// The upper triangle (row-major, M = U^T U) is cut in 64 x 64 tiles (I, J), I <= J; the right-hand side is one
// more block column (J = nb, width 1), so the forward substitution y = U^-T b needs no code of its own.
//
//   spine (CTA 0)    runs the serial chain of the factorisation without leaving the SM:
//                    factor (k, k) -> solve the panel block (k, k+1) -> update (k+1, k+1) -> factor ...
//                    and afterwards the chain of the back substitution  x_J = U_JJ^-1 (y_J - ... - U[J, J+1] x_{J+1}).
//   workers (1..G-1) tile t = (I, J) in row-major order of the upper triangle belongs to worker t mod (G - 1) for the
//                    whole factorisation, so all updates of a tile are applied by one CTA in phase order:
//                      panel (k, J), J >= k+2   U[k, J] = U_kk^-T M[k, J]     (after flagD[k])
//                      update(k, I, J)          M[I, J] -= U[k, I]^T U[k, J]  (after flagP[k][I], flagP[k][J]); DMMA.8x8x4
//                    a worker hands tile (I, I) to the spine after phase I-2 (flagTd[I]) and tile (I, I+1) after
//                    phase I-1 (flagTs[I]); in the back substitution it contributes U[I, J] x_J, J >= I+2 (flagQ[I][J]).
// Every CTA walks its work in one global order (phase k: panels of row k, then updates with row k) and every wait is on a
// strictly earlier item; all CTAs are co-resident (cooperative launch), so the schedule cannot deadlock.  Flags carry the
// epoch of the call and never need clearing.  Panel solves and the diagonal factorisation work on 8-row strips: a DMMA
// sweep with the rows already done, then an 8 x 8 triangular step in registers (no inverses are formed anywhere).
//
// Measured (tools/chol_trace.py, n = 2048) with three CTA barriers per 8-row strip: 21 us per 64-column phase = factor 11.2
// + store/flag 1.5 + panel 7.1 + update 3.8 of which the 8 x 8 pivot chain (rsqrt -> mul -> fma, one warp) is about 0.4 us per
// strip; diag_factor and panel_solve below are therefore warp-level data flow without CTA barriers.
#include <algorithm>
#include "ctx.h"

namespace {

constexpr int FB = 64;        // tile size
constexpr int FPITCH = FB + 4;   // operand pitch: (lane%4) * 68 + lane/4 hits 16 distinct 8-byte bank pairs per half-warp
constexpr int FT = 256;       // threads per CTA
constexpr long long SPIN_LIMIT = 1LL << 27;

struct FusedArgs {
  double *M;
  int ldm, n;
  double *rhs;     // may be null: factorisation only
  double *part;    // [nb][nb][64] contributions U[I, J] x_J of the back substitution
  double *dinv;    // [n] 1 / U_rr
  int *flagD;      // [nb]
  int *flagP;      // [nb x (nb + 1)]
  int *flagX;      // [nb]
  int *flagQ;      // [nb x nb]
  int *flagTd;     // [nb]  tile (I, I) carries all updates of the phases < I - 1
  int *flagTs;     // [nb]  tile (I, I + 1) carries all updates of the phases < I
  int *info;       // first non-positive pivot (1-based), 0 otherwise
  int *abort_flag; // set when a wait ran into SPIN_LIMIT (a bug, never data): every CTA then drains
  int epoch, nb, nbc;
  long long *trace;   // optional: per-CTA event records {t_ns, code} (tools/chol_trace.py); null in production
  int trace_cap;
};

__device__ __forceinline__ int ld_acquire(const int *p) {
  int v;
  asm volatile("ld.acquire.gpu.global.b32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// relaxed poll: unlike ld.acquire it does not hold back the warp's later memory operations while it is in flight
__device__ __forceinline__ int ld_relaxed(const int *p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.b32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_release(int *p, int v) { asm volatile("st.release.gpu.global.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// all threads call; returns after the flag carries this call's epoch (or the launch is aborting)
__device__ __forceinline__ void wait_flag(const FusedArgs &a, const int *f) {
  if (threadIdx.x == 0) {
    long long spins = 0;
    while (ld_acquire(f) != a.epoch) {
      if (++spins > SPIN_LIMIT) {
        atomicExch(a.abort_flag, 1);
        break;
      }
      if ((spins & 1023) == 0 && ld_acquire(a.abort_flag) != 0) break;
    }
  }
  __syncthreads();
}
// all threads call after their global writes
__device__ __forceinline__ void post_flag(const FusedArgs &a, int *f) {
  __syncthreads();
  // st.release.gpu orders the writes of the whole CTA (cumulative over the barrier above) before the flag: no separate fence
  if (threadIdx.x == 0) st_release(f, a.epoch);
}

// event record: code = type << 24 | a << 12 | b ; type 1 diag, 2 panel, 3 update, 4 backsolve, 5 backprod; +8 = end, +16 = after the waits
__device__ __forceinline__ void trace_ev(const FusedArgs &a, int type, int x, int y) {
  __shared__ int s_trace_n;   // event count of this CTA (shared memory: reading a counter back from global memory would put an
                              // L2 round trip on thread 0 at every event and distort the very chain being measured)
  if (a.trace != nullptr && threadIdx.x == 0) {
    long long *base = a.trace + (size_t) blockIdx.x * a.trace_cap * 2;
    if (type == 0) {          // reset (first call of the kernel)
      s_trace_n = 0;
      return;
    }
    const int n = s_trace_n;
    if (n + 1 < a.trace_cap) {
      long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      base[2 * (n + 1)]     = t;
      base[2 * (n + 1) + 1] = ((long long) type << 24) | ((long long) x << 12) | y;
      base[0]               = n + 1;
      s_trace_n             = n + 1;
    }
  }
}

#ifdef NCM_FUSED_PROBE   // tools/chol_trace.py --probe: clock64 stamps of every warp inside diag_factor (block k = 0 of a SMALL system, whose
                        // workers 140 + w are idle and lend their trace regions); compiled out of the library
__device__ __forceinline__ void probe_ev(const FusedArgs &a, int k, int w, int code) {
  if (a.trace != nullptr && k == 0 && (threadIdx.x & 31) == 0) {
    long long *base = a.trace + (size_t) (140 + w) * a.trace_cap * 2;
    const long long n = base[0];
    base[2 * (n + 1)]     = clock64();
    base[2 * (n + 1) + 1] = code;
    base[0]               = n + 1;
  }
}
#define NCM_PROBE_EV(code) probe_ev(a, k, w, (code))
#else
#define NCM_PROBE_EV(code) do { } while (0)
#endif

__device__ __forceinline__ int row_start(int I, int nbc) { return I * nbc - (I * (I - 1)) / 2; }


// ---- 8-row strip sweep ------------------------------------------------------------------------------------
// Left-looking step shared by the diagonal factorisation and the panel solve: for the strip of rows
// b0 .. b0+7 of X (64 columns, pitch FPITCH)
//     X[b0 + r][c] -= sum_{s < b0} Asrc[s][b0 + r] * X[s][c]
// on DMMA.8x8x4: warp w owns columns 8w .. 8w+7 (tiles w_first <= w < w_end), two independent accumulator
// chains.  Asrc holds the upper factor (U_kk); for the diagonal block Asrc == X.
__device__ __forceinline__ void strip_sweep(const double *Asrc, double *X, int b0, int w_first, int w_end) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int lr = lane & 3, lc = lane >> 2;
  if (w >= w_first && w < w_end && b0 > 0) {
    double *px = X + (b0 + lc) * FPITCH + 8 * w + 2 * lr;
    double c0 = px[0], c1 = px[1], d0 = 0.0, d1 = 0.0, e0 = 0.0, e1 = 0.0, f0 = 0.0, f1 = 0.0;
    const double *pa = Asrc + lr * FPITCH + b0 + lc;
    const double *pb = X + lr * FPITCH + 8 * w + lc;
    int s0 = 0;
    for (; s0 + 16 <= b0; s0 += 16) {   // four independent accumulator chains
      const double a0 = -pa[s0 * FPITCH], b_0 = pb[s0 * FPITCH];
      const double a1 = -pa[(s0 + 4) * FPITCH], b_1 = pb[(s0 + 4) * FPITCH];
      const double a2 = -pa[(s0 + 8) * FPITCH], b_2 = pb[(s0 + 8) * FPITCH];
      const double a3 = -pa[(s0 + 12) * FPITCH], b_3 = pb[(s0 + 12) * FPITCH];
      dmma884(c0, c1, a0, b_0);
      dmma884(d0, d1, a1, b_1);
      dmma884(e0, e1, a2, b_2);
      dmma884(f0, f1, a3, b_3);
    }
    if (s0 < b0) {
      const double a0 = -pa[s0 * FPITCH], b_0 = pb[s0 * FPITCH];
      const double a1 = -pa[(s0 + 4) * FPITCH], b_1 = pb[(s0 + 4) * FPITCH];
      dmma884(c0, c1, a0, b_0);
      dmma884(d0, d1, a1, b_1);
    }
    c0 += e0; c1 += e1;
    d0 += f0; d1 += f1;
    px[0] = c0 + d0;
    px[1] = c1 + d1;
  }
}

// ---- tiles between global and shared memory -------------------------------------------------------------------
// rows k0 .. k0+nbk-1 of block column J (width columns; the rhs block is one column of a.rhs) -> sX, zero padded
__device__ __forceinline__ void tile_fetch(const FusedArgs &a, int k, int J, double *sX) {
  const int tid = threadIdx.x;
  const int k0 = k * FB, J0 = J * FB;
  const int nbk = min(FB, a.n - k0);
  const bool isR = (a.rhs != nullptr) && (J == a.nbc - 1);
  const int width = isR ? 1 : min(FB, a.n - J0);
  double2 tv[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int chunk = q * FT + tid;
    const int r = chunk >> 5, cc = (chunk & 31) * 2;
    double2 v = make_double2(0.0, 0.0);
    if (r < nbk) {
      if (isR) {
        if (cc == 0) v.x = __ldcg(a.rhs + k0 + r);
      } else if (cc + 1 < width) {
        v = __ldcg(reinterpret_cast<const double2 *>(a.M + (size_t) (k0 + r) * a.ldm + J0 + cc));
      } else if (cc < width) {
        v.x = __ldcg(a.M + (size_t) (k0 + r) * a.ldm + J0 + cc);
      }
    }
    tv[q] = v;
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int chunk = q * FT + tid;
    *reinterpret_cast<double2 *>(sX + (chunk >> 5) * FPITCH + (chunk & 31) * 2) = tv[q];
  }
}
// same through cp.async (not for the rhs block); the caller commits and waits
__device__ __forceinline__ void tile_fetch_async(const FusedArgs &a, int k, int J, double *sX) {
  const int tid = threadIdx.x;
  const int k0 = k * FB, J0 = J * FB;
  const int nbk = min(FB, a.n - k0);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int chunk = q * FT + tid;
    const int r = chunk >> 5, cc = (chunk & 31) * 2;
    const int gc = J0 + cc;
    int bytes    = r < nbk ? (a.n - gc) * 8 : 0;
    bytes        = bytes < 0 ? 0 : (bytes > 16 ? 16 : bytes);
    const double *src = bytes > 0 ? a.M + (size_t) (k0 + r) * a.ldm + gc : a.M;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(sX + r * FPITCH + cc)), "l"(src), "r"(bytes) : "memory");
  }
}
__device__ __forceinline__ void tile_store(const FusedArgs &a, int k, int J, const double *sX) {
  const int tid = threadIdx.x;
  const int k0 = k * FB, J0 = J * FB;
  const int nbk = min(FB, a.n - k0);
  const bool isR = (a.rhs != nullptr) && (J == a.nbc - 1);
  const int width = isR ? 1 : min(FB, a.n - J0);
  if (isR) {
    if (tid < nbk) a.rhs[k0 + tid] = sX[tid * FPITCH];
    return;
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int chunk = q * FT + tid;
    const int r = chunk >> 5, cc = (chunk & 31) * 2;
    if (r < nbk) {
      double *dst = a.M + (size_t) (k0 + r) * a.ldm + J0 + cc;
      if (cc + 1 < width)
        *reinterpret_cast<double2 *>(dst) = *reinterpret_cast<const double2 *>(sX + r * FPITCH + cc);
      else if (cc < width)
        dst[0] = sX[r * FPITCH + cc];
    }
  }
}
// write the factored diagonal block (upper part; the lower triangle of M is scratch for every reader) and 1 / U_rr
__device__ __forceinline__ void diag_store(const FusedArgs &a, int k, const double *S, const double *sDinv) {
  const int tid = threadIdx.x;
  const int k0  = k * FB;
  const int nbk = min(FB, a.n - k0);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int chunk = q * FT + tid;
    const int r = chunk >> 5, cc = (chunk & 31) * 2;
    if (r < nbk && cc + 1 >= r) {
      double *dst = a.M + (size_t) (k0 + r) * a.ldm + k0 + cc;
      if (cc + 1 < nbk)
        *reinterpret_cast<double2 *>(dst) = *reinterpret_cast<const double2 *>(S + r * FPITCH + cc);
      else if (cc < nbk)
        dst[0] = S[r * FPITCH + cc];
    }
  }
  if (tid < nbk) a.dinv[k0 + tid] = sDinv[tid];
}

// one warp's share of tile_fetch_async: the whole 64 x 64 tile through the 32 lanes of the calling warp
__device__ __forceinline__ void tile_fetch_async_warp(const FusedArgs &a, int k, int J, double *sX) {
  const int lane = threadIdx.x & 31;
  const int k0 = k * FB, J0 = J * FB;
  const int nbk = min(FB, a.n - k0);
#pragma unroll 8
  for (int q = 0; q < (FB * FB / 2) / 32; ++q) {
    const int chunk = q * 32 + lane;
    const int r = chunk >> 5, cc = (chunk & 31) * 2;
    const int gc = J0 + cc;
    int bytes    = r < nbk ? (a.n - gc) * 8 : 0;
    bytes        = bytes < 0 ? 0 : (bytes > 16 ? 16 : bytes);
    const double *src = bytes > 0 ? a.M + (size_t) (k0 + r) * a.ldm + gc : a.M;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(sX + r * FPITCH + cc)), "l"(src), "r"(bytes) : "memory");
  }
}

// 8 x 8 triangular step shared by the diagonal factorisation and the panel solve: the calling lane owns column ci of X and
// solves its 8 entries of the strip b0 .. b0+7 against the (already final) pivot block of U at (b0, b0)
__device__ __forceinline__ void strip_column_solve(const double *sU, double *X, const double *sD, int b0, int ci) {
  double u[8][8], t[8], inv[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
#pragma unroll
    for (int q = r + 1; q < 8; ++q) u[r][q] = sU[(b0 + r) * FPITCH + b0 + q];
    t[r]   = X[(b0 + r) * FPITCH + ci];
    inv[r] = sD[b0 + r];
  }
  // column-oriented: as soon as x_s is known every later entry takes its term, so the dependent chain is mul -> fma -> mul ... (15 operations)
  // instead of the 36 of the row-by-row order; each t[r] still receives its terms in ascending s: same rounding
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    const double xs = t[s] * inv[s];
    X[(b0 + s) * FPITCH + ci] = xs;
#pragma unroll
    for (int r = s + 1; r < 8; ++r) t[r] = fma(-u[s][r], xs, t[r]);
  }
}

// ---- diagonal block ---------------------------------------------------------------------------------
// named CTA barriers (bar.arrive / bar.sync with a thread count): producer-consumer hand-off between warps with the memory
// ordering of a barrier and without a fence on the producer's path
__device__ __forceinline__ void bar_sync_n(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void bar_arrive_n(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// S (pitch FPITCH) holds the block (upper part valid, identity padding beyond nbk); on return S holds U_kk and
// sDinv the reciprocal pivots.
//
// Data flow over the 8 warps, no CTA barrier inside.  Warp w owns the 8 columns of column block w and keeps its tiles (0, w) ..
// (w, w) in DMMA accumulator fragments; the block is factored right-looking by 8-row strips:
//   strip j < w :  fragment (j, w) -> shared memory; wait for pivot block j (barrier P_j); 8 x 8 triangular step, one lane per
//                  column; wait until the whole strip j is solved (barrier S_j); rank-8 update of the tiles (j+1 .. w, w)
//   strip j = w :  fragment (w, w) -> shared memory -> every lane loads the 36 entries and factors the 8 x 8 pivot block in
//                  registers (the same arithmetic in every lane: nothing but rsqrt -> mul -> fma on the pivot chain);
//                  lane 0 writes it back; arrive on P_w.
// The next pivot warp (w = j + 1) only ARRIVES on S_j: its own tile (j, j+1) is all its diagonal tile needs, so the chain per
// strip is  P_j -> triangular step -> 2 DMMA -> fragment round trip -> pivot block  with everything else beside it (the
// left-looking version with shared-memory flags and fences: 2640 cycles per strip, 1000 of them in fences and sweeps).
// Warp 0 is free after the first pivot block: it polls the two flags the spine will need next (want_*: still to be
// fetched) and fetches those tiles with cp.async, so a worker that is on time never puts an L2 round trip on the chain.
__device__ void diag_factor(const FusedArgs &a, double *S, double *sDinv, int *sBad, int k0, int k, bool &want_panel, bool &want_next, double *sX,
                            double *sC, int *sPoll) {
  __shared__ int sDone;
  constexpr int NS = FB / 8;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int lr = lane & 3, lc = lane >> 2;
  if (tid == 0) {
    sPoll[0] = 0;
    sPoll[1] = 0;
    sBad[3]  = 0x7fffffff;   // lowest failing pivot of this block (several warps may report)
    sDone    = 0;
  }
  __syncthreads();
  const int bw = 8 * w;
  double acc[NS][4][2];   // four accumulator pairs per tile in the summation order of the left-looking sweep (see panel_solve)
#pragma unroll
  for (int m = 0; m < NS; ++m)
    if (m <= w) {
      const double2 v = *reinterpret_cast<const double2 *>(S + (8 * m + lc) * FPITCH + bw + 2 * lr);
      acc[m][0][0]    = v.x;
      acc[m][0][1]    = v.y;
#pragma unroll
      for (int c = 1; c < 4; ++c) acc[m][c][0] = acc[m][c][1] = 0.0;
    }
  NCM_PROBE_EV(0);
#pragma unroll
  for (int j = 0; j < NS; ++j) {
    const int b0 = 8 * j;
    if (j < w) {
      if (j > 0) {
        const double v0 = (acc[j][0][0] + acc[j][2][0]) + (acc[j][1][0] + acc[j][3][0]);
        const double v1 = (acc[j][0][1] + acc[j][2][1]) + (acc[j][1][1] + acc[j][3][1]);
        *reinterpret_cast<double2 *>(S + (b0 + lc) * FPITCH + bw + 2 * lr) = make_double2(v0, v1);
        __syncwarp();
      }
      bar_sync_n(1 + j, 32 * (NS - j));   // P_j
      NCM_PROBE_EV(300 + j);
      if (lane < 8) strip_column_solve(S, S, sDinv, b0, bw + lane);
      __syncwarp();
      NCM_PROBE_EV(400 + j);
      if (j + 2 < NS) {                   // S_j has more than one participant
        if (w == j + 1)
          bar_arrive_n(1 + NS + j, 32 * (NS - 1 - j));
        else
          bar_sync_n(1 + NS + j, 32 * (NS - 1 - j));
      }
      const double bx0 = S[(b0 + lr) * FPITCH + bw + lc], bx1 = S[(b0 + 4 + lr) * FPITCH + bw + lc];
#pragma unroll
      for (int m = j + 1; m < NS; ++m)
        if (m <= w) {
          const double a0 = -S[(b0 + lr) * FPITCH + 8 * m + lc], a1 = -S[(b0 + 4 + lr) * FPITCH + 8 * m + lc];
          dmma884(acc[m][2 * (j & 1)][0], acc[m][2 * (j & 1)][1], a0, bx0);
          dmma884(acc[m][2 * (j & 1) + 1][0], acc[m][2 * (j & 1) + 1][1], a1, bx1);
        }
      NCM_PROBE_EV(500 + j);
    } else if (j == w) {
      if (j > 0) {
        const double v0 = (acc[j][0][0] + acc[j][2][0]) + (acc[j][1][0] + acc[j][3][0]);
        const double v1 = (acc[j][0][1] + acc[j][2][1]) + (acc[j][1][1] + acc[j][3][1]);
        *reinterpret_cast<double2 *>(S + (b0 + lc) * FPITCH + bw + 2 * lr) = make_double2(v0, v1);
        __syncwarp();
      }
      NCM_PROBE_EV(600);
      double dgl[8][8];
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int q = r; q < 8; ++q) dgl[r][q] = S[(bw + r) * FPITCH + bw + q];
      double inv[8];
      int bad = 0;
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        const double piv = dgl[p][p];
        if (!(piv > 0.0)) bad = (bad == 0) ? k0 + bw + p + 1 : bad;   // the factor is garbage (NaN) from here on; info reports it
        inv[p]    = rsqrt(piv);
        dgl[p][p] = piv * inv[p];
#pragma unroll
        for (int q = p + 1; q < 8; ++q) dgl[p][q] *= inv[p];
#pragma unroll
        for (int r = p + 1; r < 8; ++r)
#pragma unroll
          for (int q = r; q < 8; ++q) dgl[r][q] = fma(-dgl[p][r], dgl[p][q], dgl[r][q]);
      }
#ifdef NCM_FUSED_PROBE
      if (dgl[7][7] == -1.2345) sBad[3] = 1;   // keeps the stamp below after the arithmetic
      NCM_PROBE_EV(700);
#endif
      __syncwarp();   // every lane has read the tile before lane 0 overwrites it
      if (lane == 0) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          // row r: entries r .. 7; bw + r is even exactly when r is, so the pairs starting at an even column are 16-byte aligned
          if (r & 1) S[(bw + r) * FPITCH + bw + r] = dgl[r][r];
#pragma unroll
          for (int q = (r + 1) & ~1; q < 8; q += 2)
            *reinterpret_cast<double2 *>(S + (bw + r) * FPITCH + bw + q) = make_double2(dgl[r][q], dgl[r][q + 1]);
        }
#pragma unroll
        for (int r = 0; r < 8; r += 2) *reinterpret_cast<double2 *>(sDinv + bw + r) = make_double2(inv[r], inv[r + 1]);
        if (bad != 0) atomicMin(sBad + 3, bad);
        if (j == NS - 1) atomicExch(&sDone, 1);
      }
      __syncwarp();
      if (j + 1 < NS) bar_arrive_n(1 + j, 32 * (NS - j));   // P_j
      NCM_PROBE_EV(800);
    }
  }
  if (w == 0 && (want_panel || want_next)) {
    bool wp = want_panel, wn = want_next;
    for (;;) {
      int f0 = 0, f1 = 0, done = 0;
      if (lane == 0) {
        done = atomicAdd(&sDone, 0);   // the last pivot block is out: stop polling, the caller takes over
        f0   = wp ? (ld_relaxed(a.flagTs + k) == a.epoch) : 0;
        f1   = wn ? (ld_relaxed(a.flagTd + k + 1) == a.epoch) : 0;
      }
      done = __shfl_sync(0xffffffffu, done, 0);
      f0   = __shfl_sync(0xffffffffu, f0, 0);
      f1   = __shfl_sync(0xffffffffu, f1, 0);
      if (done != 0 && f0 == 0 && f1 == 0) break;
      if (f0 != 0 || f1 != 0) {
        __threadfence();   // relaxed poll + fence = acquire: the tile data is ordered after the flag that announced it
        if (f0 != 0) {
          tile_fetch_async_warp(a, k, k + 1, sX);
          wp = false;
          if (lane == 0) sPoll[0] = 1;
        }
        if (f1 != 0) {
          tile_fetch_async_warp(a, k + 1, k + 1, sC);
          wn = false;
          if (lane == 0) sPoll[1] = 1;
        }
        cp_async_commit();
      }
      if (!(wp || wn)) break;
    }
  }
  __syncthreads();
  if (sPoll[0] != 0) want_panel = false;
  if (sPoll[1] != 0) want_next = false;
  if (tid == 0 && sBad[3] != 0x7fffffff && *sBad == 0) *sBad = sBad[3];
}

// ---- panel solve in shared memory: sX <- sU^-T sX ----------------------------------------------------------
// sU: U_kk (pitch FPITCH, identity beyond the valid rows), sD: 1 / U_rr, sX: the tile, solved in place.
// U_kk is complete, so the 8 column blocks are independent chains: warp w keeps its 64 x 8 column block in DMMA accumulator
// fragments and works right-looking down the 8 strips -- strip j: fragment -> shared memory, 8 x 8 triangular step (one lane per
// column), then the rank-8 update of the strips below, the next one first (it is the only one on the chain).  Warp-level
// synchronisation only, one CTA barrier at the end (the left-looking version with two CTA barriers per strip: 5.3 - 7.1 us per tile).
__device__ void panel_solve(const FusedArgs &a, const double *sU, double *sX, const double *sD, int width) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int lr = lane & 3, lc = lane >> 2;
  const int w_end = (width + 7) >> 3;
  if (w < w_end) {
    const int ci = 8 * w + lane;
    // four accumulator pairs per tile, fed exactly as the left-looking DMMA sweep of the earlier versions fed them (rows 16 g + 0..3 ->
    // c, + 4..7 -> d, + 8..11 -> e, + 12..15 -> f; result (c + e) + (d + f)): the factor stays bit-identical to those versions, on which
    // the accepted-sequence parity tests were established
    double acc[FB / 8][4][2];
#pragma unroll
    for (int m = 0; m < FB / 8; ++m) {
      const double2 v = *reinterpret_cast<const double2 *>(sX + (8 * m + lc) * FPITCH + 8 * w + 2 * lr);
      acc[m][0][0] = v.x;
      acc[m][0][1] = v.y;
#pragma unroll
      for (int c = 1; c < 4; ++c) acc[m][c][0] = acc[m][c][1] = 0.0;
    }
#pragma unroll
    for (int j = 0; j < FB / 8; ++j) {
      const int b0 = 8 * j;
      if (j > 0) {
        const double v0 = (acc[j][0][0] + acc[j][2][0]) + (acc[j][1][0] + acc[j][3][0]);
        const double v1 = (acc[j][0][1] + acc[j][2][1]) + (acc[j][1][1] + acc[j][3][1]);
        *reinterpret_cast<double2 *>(sX + (b0 + lc) * FPITCH + 8 * w + 2 * lr) = make_double2(v0, v1);
        __syncwarp();
      }
      if (lane < 8 && ci < width) strip_column_solve(sU, sX, sD, b0, ci);
      __syncwarp();
      if (j + 1 < FB / 8) {
        const double bx0 = sX[(b0 + lr) * FPITCH + 8 * w + lc], bx1 = sX[(b0 + 4 + lr) * FPITCH + 8 * w + lc];
#pragma unroll
        for (int m = j + 1; m < FB / 8; ++m) {
          const double a0 = -sU[(b0 + lr) * FPITCH + 8 * m + lc], a1 = -sU[(b0 + 4 + lr) * FPITCH + 8 * m + lc];
          dmma884(acc[m][2 * (j & 1)][0], acc[m][2 * (j & 1)][1], a0, bx0);
          dmma884(acc[m][2 * (j & 1) + 1][0], acc[m][2 * (j & 1) + 1][1], a1, bx1);
        }
      }
    }
  }
  __syncthreads();
}

// worker: panel block (k, J), J >= k + 2
__device__ void do_panel(const FusedArgs &a, int k, int J, double *sU, double *sX, double *sD) {
  const int tid = threadIdx.x;
  const int k0  = k * FB;
  const int nbk = min(FB, a.n - k0);
  const bool isR = (a.rhs != nullptr) && (J == a.nbc - 1);
  const int width = isR ? 1 : min(FB, a.n - J * FB);
  trace_ev(a, 2, k, J);
  tile_fetch(a, k, J, sX);   // last written by this CTA: fetched while waiting for U_kk
  wait_flag(a, a.flagD + k);
  trace_ev(a, 2 + 16, k, J);
  {
    double2 tv[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int chunk = q * FT + tid;
      const int r = chunk >> 5, cc = (chunk & 31) * 2;
      double2 v = make_double2(0.0, 0.0);
      if (r < nbk) {
        if (cc + 1 < nbk)
          v = __ldcg(reinterpret_cast<const double2 *>(a.M + (size_t) (k0 + r) * a.ldm + k0 + cc));
        else if (cc < nbk)
          v.x = __ldcg(a.M + (size_t) (k0 + r) * a.ldm + k0 + cc);
      } else {
        if (cc == r) v.x = 1.0;
        if (cc + 1 == r) v.y = 1.0;
      }
      tv[q] = v;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int chunk = q * FT + tid;
      *reinterpret_cast<double2 *>(sU + (chunk >> 5) * FPITCH + (chunk & 31) * 2) = tv[q];
    }
    if (tid < FB) sD[tid] = tid < nbk ? __ldcg(a.dinv + k0 + tid) : 1.0;
  }
  __syncthreads();
  panel_solve(a, sU, sX, sD, width);
  tile_store(a, k, J, sX);
  post_flag(a, a.flagP + k * (a.nb + 1) + J);
  __syncthreads();
  trace_ev(a, 2 + 8, k, J);
}

// ---- trailing update (k, I, J): M[I, J] -= U[k, I]^T U[k, J] ------------------------------------------------
// 8 warps as 2 x 4, warp tile 32 x 16 (4 x 2 DMMA tiles).  to_S: the result is left in S (pitch FPITCH, identity
// padded) for do_diag instead of being written to global memory.
__device__ void do_update(const FusedArgs &a, int k, int I, int J, double *sA, double *sB, double *S, bool to_S) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;
  const int lr = lane & 3, lc = lane >> 2;
  const int rm = wm * 32, cn = wn * 16;
  const int k0 = k * FB, I0 = I * FB, J0 = J * FB;
  const bool isR = (a.rhs != nullptr) && (J == a.nbc - 1);
  const int hI = min(FB, a.n - I0);                    // valid rows of the tile
  const int wJ = isR ? 1 : min(FB, a.n - J0);          // valid columns

  trace_ev(a, 3, I, J);
  wait_flag(a, a.flagP + k * (a.nb + 1) + I);
  if (J != I) wait_flag(a, a.flagP + k * (a.nb + 1) + J);
  trace_ev(a, 3 + 16, I, J);

  // operands: rows k0 .. k0+63 (always a full block row), columns of block I (A) and block J (B)
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int chunk = q * FT + tid;
    const int r = chunk >> 5, cc = (chunk & 31) * 2;
    {
      const int gc = I0 + cc;
      int bytes    = (a.n - gc) * 8;
      bytes        = bytes < 0 ? 0 : (bytes > 16 ? 16 : bytes);
      const double *src = bytes > 0 ? a.M + (size_t) (k0 + r) * a.ldm + gc : a.M;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(sA + r * FPITCH + cc)), "l"(src), "r"(bytes) : "memory");
    }
    if (J != I && !isR) {
      const int gc = J0 + cc;
      int bytes    = (a.n - gc) * 8;
      bytes        = bytes < 0 ? 0 : (bytes > 16 ? 16 : bytes);
      const double *src = bytes > 0 ? a.M + (size_t) (k0 + r) * a.ldm + gc : a.M;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(sB + r * FPITCH + cc)), "l"(src), "r"(bytes) : "memory");
    }
  }
  cp_async_commit();
  if (isR) {
    // B = y_k as column 0 (the other columns feed accumulators that are never stored)
    for (int e = tid; e < FB * 8; e += FT) sB[(e >> 3) * FPITCH + (e & 7)] = 0.0;
    __syncthreads();
    if (tid < FB) sB[tid * FPITCH] = __ldcg(a.rhs + k0 + tid);
  }

  const bool fast = !isR && hI == FB && wJ == FB && (J0 + FB <= a.n);
  double acc[4][2][2];
  if (fast) {
#pragma unroll
    for (int am = 0; am < 4; ++am)
#pragma unroll
      for (int bn = 0; bn < 2; ++bn) {
        const double2 v = __ldcg(reinterpret_cast<const double2 *>(a.M + (size_t) (I0 + rm + am * 8 + lc) * a.ldm + J0 + cn + bn * 8 + 2 * lr));
        acc[am][bn][0]  = v.x;
        acc[am][bn][1]  = v.y;
      }
  } else {
#pragma unroll
    for (int am = 0; am < 4; ++am)
#pragma unroll
      for (int bn = 0; bn < 2; ++bn)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int ti = rm + am * 8 + lc, tj = cn + bn * 8 + 2 * lr + e;
          double v = 0.0;
          if (ti < hI && tj < wJ) v = isR ? __ldcg(a.rhs + I0 + ti) : __ldcg(a.M + (size_t) (I0 + ti) * a.ldm + J0 + tj);
          acc[am][bn][e] = v;
        }
  }
  cp_async_wait<0>();
  __syncthreads();
  const double *pB = (J == I) ? sA : sB;
#pragma unroll
  for (int ks = 0; ks < FB / 4; ++ks) {
    double af[4], bf[2];
    const double *pa = sA + (ks * 4 + lr) * FPITCH + rm + lc;
    const double *pb = pB + (ks * 4 + lr) * FPITCH + cn + lc;
#pragma unroll
    for (int am = 0; am < 4; ++am) af[am] = -pa[am * 8];
#pragma unroll
    for (int bn = 0; bn < 2; ++bn) bf[bn] = pb[bn * 8];
#pragma unroll
    for (int am = 0; am < 4; ++am)
#pragma unroll
      for (int bn = 0; bn < 2; ++bn) dmma884(acc[am][bn][0], acc[am][bn][1], af[am], bf[bn]);
  }
  if (to_S) {
#pragma unroll
    for (int am = 0; am < 4; ++am)
#pragma unroll
      for (int bn = 0; bn < 2; ++bn)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int ti = rm + am * 8 + lc, tj = cn + bn * 8 + 2 * lr + e;
          S[ti * FPITCH + tj] = (ti < hI && tj < wJ) ? acc[am][bn][e] : (ti == tj ? 1.0 : 0.0);
        }
  } else if (fast) {
#pragma unroll
    for (int am = 0; am < 4; ++am)
#pragma unroll
      for (int bn = 0; bn < 2; ++bn)
        *reinterpret_cast<double2 *>(a.M + (size_t) (I0 + rm + am * 8 + lc) * a.ldm + J0 + cn + bn * 8 + 2 * lr) =
            make_double2(acc[am][bn][0], acc[am][bn][1]);
  } else {
#pragma unroll
    for (int am = 0; am < 4; ++am)
#pragma unroll
      for (int bn = 0; bn < 2; ++bn)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int ti = rm + am * 8 + lc, tj = cn + bn * 8 + 2 * lr + e;
          if (ti < hI && tj < wJ) {
            if (isR)
              a.rhs[I0 + ti] = acc[am][bn][e];
            else
              a.M[(size_t) (I0 + ti) * a.ldm + J0 + tj] = acc[am][bn][e];
          }
        }
  }
  __syncthreads();
  trace_ev(a, 3 + 8, I, J);
}


// ---- back substitution --------------------------------------------------------------------------------
// warp 0 polls count <= 64 consecutive flags in parallel (one L2 round trip instead of one per flag)
__device__ __forceinline__ void wait_flags_many(const FusedArgs &a, const int *f0, int count) {
  if (threadIdx.x < 32 && count > 0) {
    const int lane = threadIdx.x;
    long long spins = 0;
    while (true) {
      bool ok = true;
      if (lane < count) ok = ld_acquire(f0 + lane) == a.epoch;
      if (lane + 32 < count) ok = ok && (ld_acquire(f0 + lane + 32) == a.epoch);
      if (__all_sync(0xffffffffu, ok)) break;
      if (++spins > SPIN_LIMIT / 64) {
        if (lane == 0) atomicExch(a.abort_flag, 1);
        break;
      }
      if ((spins & 255) == 0 && ld_acquire(a.abort_flag) != 0) break;
    }
  }
  __syncthreads();
}
// contribution of tile (I, J), J >= I + 2:  part[I][J][r] = sum_c U[I0 + r][J0 + c] x_J[c]
__device__ void do_backprod(const FusedArgs &a, int I, int J, double *sx) {
  const int tid = threadIdx.x;
  const int I0 = I * FB, J0 = J * FB;
  const int wJ = min(FB, a.n - J0);
  trace_ev(a, 5, I, J);
  const int r = tid >> 2, q = tid & 3;   // 4 threads per row, 16 columns each
  const double *u = a.M + (size_t) (I0 + r) * a.ldm + J0 + q * 16;
  const bool full = J0 + FB <= a.n;
  double2 v[8];
  if (full) {   // the tile was finalised by this CTA: fetch it before waiting for x_J
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = __ldcg(reinterpret_cast<const double2 *>(u) + e);
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int cidx = q * 16 + 2 * e;
      v[e].x = cidx < wJ ? __ldcg(u + 2 * e) : 0.0;
      v[e].y = cidx + 1 < wJ ? __ldcg(u + 2 * e + 1) : 0.0;
    }
  }
  wait_flag(a, a.flagX + J);
  trace_ev(a, 5 + 16, I, J);
  if (tid < FB) sx[tid] = tid < wJ ? __ldcg(a.rhs + J0 + tid) : 0.0;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int e = 0; e < 8; ++e) s = fma(v[e].y, sx[q * 16 + 2 * e + 1], fma(v[e].x, sx[q * 16 + 2 * e], s));
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  if (q == 0) a.part[((size_t) I * a.nb + J) * FB + r] = s;
  post_flag(a, a.flagQ + I * a.nb + J);
  __syncthreads();
  trace_ev(a, 5 + 8, I, J);
}

// ---- spine: the serial chain of the factorisation on one SM ----------------------------------------------------
// S: current diagonal block, sX: panel block (k, k+1), sC: prefetched tile (k+1, k+1)
__device__ void spine_factor(const FusedArgs &a, double *S, double *sX, double *sC, double *sDinv, int *sBad /* [3]: bad pivot, 2 poll results */) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  
  const int lr = lane & 3, lc = lane >> 2;
  const int nb = a.nb, nbc = a.nbc;
  if (tid == 0) *sBad = 0;
  {   // tile (0, 0), identity padded
    const int nbk = min(FB, a.n);
    const int r = tid >> 2, cb = (tid & 3) * 16;
    double v[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int cidx = cb + q;
      v[q] = (r < nbk && cidx < nbk && cidx >= r) ? __ldcg(a.M + (size_t) r * a.ldm + cidx) : ((r == cidx && r >= nbk) ? 1.0 : 0.0);
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) S[r * FPITCH + cb + q] = v[q];
  }
  __syncthreads();
  for (int k = 0; k < nb; ++k) {
    const bool has_panel = k + 1 < nbc;
    const bool has_next  = k + 1 < nb;
    const bool panel_isR = has_panel && (a.rhs != nullptr) && (k + 1 == nbc - 1);
    // tiles the workers owe us: (k, k+1) with all updates of the phases < k, (k+1, k+1) with those of the phases < k;
    // both are fetched from inside the factorisation as soon as their flags are seen
    bool want_panel = has_panel && !panel_isR, want_next = has_next;
    if (want_panel && k == 0) {
      tile_fetch_async(a, k, k + 1, sX);
      want_panel = false;
    }
    if (want_next && k + 1 < 2) {
      tile_fetch_async(a, k + 1, k + 1, sC);
      want_next = false;
    }
    cp_async_commit();
    trace_ev(a, 1, k, k);
    diag_factor(a, S, sDinv, sBad, k * FB, k, want_panel, want_next, sX, sC, sBad + 1);
    trace_ev(a, 6, k, 0);
    diag_store(a, k, S, sDinv);
    trace_ev(a, 6, k, 1);
    if (tid == 0 && *sBad != 0) atomicCAS(a.info, 0, *sBad);
    post_flag(a, a.flagD + k);
    trace_ev(a, 1 + 8, k, k);
    if (has_panel) {
      trace_ev(a, 2, k, k + 1);
      if (panel_isR) {
        if (k > 0) wait_flag(a, a.flagTs + k);
        tile_fetch(a, k, k + 1, sX);
      } else if (want_panel) {
        wait_flag(a, a.flagTs + k);
        tile_fetch_async(a, k, k + 1, sX);
        cp_async_commit();
        cp_async_wait<0>();
      } else {
        cp_async_wait<0>();   // in flight or landed (together with the next diagonal tile if that was requested too)
      }
      __syncthreads();
      trace_ev(a, 2 + 16, k, k + 1);
      const int width = panel_isR ? 1 : min(FB, a.n - (k + 1) * FB);
      panel_solve(a, S, sX, sDinv, width);
      trace_ev(a, 7, k, 0);
      tile_store(a, k, k + 1, sX);
      trace_ev(a, 7, k, 1);
      post_flag(a, a.flagP + k * (nb + 1) + k + 1);
      trace_ev(a, 2 + 8, k, k + 1);
    }
    if (has_next) {
      // S <- tile (k+1, k+1) - U[k, k+1]^T U[k, k+1], identity padded
      trace_ev(a, 3, k + 1, k + 1);
      if (want_next) {
        if (k + 1 >= 2) wait_flag(a, a.flagTd + k + 1);
        tile_fetch_async(a, k + 1, k + 1, sC);
        cp_async_commit();
      }
      cp_async_wait<0>();
      __syncthreads();
      const int I0 = (k + 1) * FB;
      const int hI = min(FB, a.n - I0);
      // only the upper triangle of the diagonal tile is ever read (diag_factor): its 36 8 x 8 sub-tiles are dealt round-robin to the
      // 8 warps (4 or 5 each, balanced over the four DMMA pipes) instead of 64 sub-tiles in a 4 x 2 block per warp
      {
        int tr[5], tc[5];
        const int nt = warp < 4 ? 5 : 4;   // sub-tile t = warp + 8 i of the row-major enumeration of the upper triangle, t < 36
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          int t = warp + 8 * i, r = 0;
#pragma unroll
          for (int rr = 0; rr < 7; ++rr)
            if (t >= 8 - rr && r == rr) {
              t -= 8 - rr;
              ++r;
            }
          tr[i] = r;
          tc[i] = r + t;
        }
        double acc[5][2];
#pragma unroll
        for (int i = 0; i < 5; ++i)
          if (i < nt) {
            const double2 v = *reinterpret_cast<const double2 *>(sC + (8 * tr[i] + lc) * FPITCH + 8 * tc[i] + 2 * lr);
            acc[i][0]       = v.x;
            acc[i][1]       = v.y;
          }
#pragma unroll 4
        for (int ks = 0; ks < FB / 4; ++ks) {
          const double *prow = sX + (ks * 4 + lr) * FPITCH + lc;
#pragma unroll
          for (int i = 0; i < 5; ++i)
            if (i < nt) dmma884(acc[i][0], acc[i][1], -prow[8 * tr[i]], prow[8 * tc[i]]);
        }
        // S <- identity beyond the valid rows, the updated upper triangle elsewhere (the strictly lower sub-tiles are never read)
#pragma unroll
        for (int i = 0; i < 5; ++i)
          if (i < nt) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int ti = 8 * tr[i] + lc, tj = 8 * tc[i] + 2 * lr + e;
              S[ti * FPITCH + tj] = (ti < hI && tj < hI) ? acc[i][e] : (ti == tj ? 1.0 : 0.0);
            }
          }
      }
      __syncthreads();
      trace_ev(a, 3 + 8, k + 1, k + 1);
    }
  }
}

// ---- spine: the serial chain of the back substitution ----------------------------------------------------------
// x_J = U_JJ^-1 (y_J - sum_{J' >= J+2} part[J][J'] - U[J, J+1] x_{J+1}); x_{J+1} never leaves shared memory and the two
// tiles of the next step are fetched while this one waits for the workers' contributions.
__device__ void spine_backsolve(const FusedArgs &a, double *bufU0, double *bufU1, double *bufT0, double *bufT1, double *sy /* [6][64] */) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nb = a.nb;
  double *sxp = sy + 5 * FB;   // x_{J+1}
  auto fetchU = [&](int J, double (&v)[16]) {
    const int J0 = J * FB, nbk = min(FB, a.n - J0);
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int e = q * FT + tid, r = e >> 6, cidx = e & 63;
      v[q] = (r < nbk && cidx < nbk && cidx >= r) ? __ldcg(a.M + (size_t) (J0 + r) * a.ldm + J0 + cidx) : 0.0;
    }
  };
  auto storeU = [&](double *sU, const double (&v)[16]) {
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int e = q * FT + tid;
      sU[(e >> 6) * (FB + 1) + (e & 63)] = v[q];
    }
  };
  {
    double v[16];
    fetchU(nb - 1, v);
    storeU(bufU0, v);
  }
  if (tid < FB) sxp[tid] = 0.0;
  __syncthreads();
  for (int J = nb - 1; J >= 0; --J) {
    const int cur = (nb - 1 - J) & 1;
    double *sU = cur ? bufU1 : bufU0, *sUn = cur ? bufU0 : bufU1;
    double *sT = cur ? bufT1 : bufT0, *sTn = cur ? bufT0 : bufT1;
    const int J0  = J * FB;
    const int nbk = min(FB, a.n - J0);
    const bool has_next = J + 1 < nb;     // block J+1 exists: x_{J+1} in sxp, tile (J, J+1) in sT
    trace_ev(a, 4, J, J);
    double vn[16];
    if (J > 0) {
      fetchU(J - 1, vn);
      tile_fetch_async(a, J - 1, J, sTn);
    }
    cp_async_commit();
    wait_flag(a, a.flagP + J * (nb + 1) + (a.nbc - 1));   // y_J is final
    wait_flags_many(a, a.flagQ + J * nb + J + 2, nb - J - 2);
    trace_ev(a, 4 + 16, J, J);
    {
      const int r = tid & 63, g = tid >> 6;
      double v[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const int Jp = J + 2 + g + 4 * q;
        v[q]         = Jp < nb ? __ldcg(a.part + ((size_t) J * nb + Jp) * FB + r) : 0.0;
      }
      double s = 0.0;
#pragma unroll
      for (int q = 0; q < 16; ++q) s += v[q];
      sy[(1 + g) * FB + r] = s;
    }
    cp_async_wait<1>();   // tile (J, J+1), issued one step ago
    __syncthreads();
    {
      const int r = tid >> 2, q = tid & 3;
      double s = 0.0;
      if (has_next) {
#pragma unroll
        for (int e = 0; e < 16; ++e) s = fma(sT[r * FPITCH + q * 16 + e], sxp[q * 16 + e], s);
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if (q == 0) {
        const double y = r < nbk ? __ldcg(a.rhs + J0 + r) : 0.0;
        sy[r] = (((y - sy[FB + r]) - sy[2 * FB + r]) - sy[3 * FB + r]) - sy[4 * FB + r] - s;
      }
    }
    __syncthreads();
    if (warp == 0) {
      double y0 = lane < nbk ? sy[lane] : 0.0, y1 = lane + 32 < nbk ? sy[lane + 32] : 0.0;
      const double d0 = lane < nbk ? __ldcg(a.dinv + J0 + lane) : 1.0, d1 = lane + 32 < nbk ? __ldcg(a.dinv + J0 + lane + 32) : 1.0;
#pragma unroll 8
      for (int r = FB - 1; r >= 32; --r) {
        const double xr = __shfl_sync(0xffffffffu, y1 * d1, r - 32);
        if (lane + 32 == r) y1 = xr;
        if (lane + 32 < r) y1 = fma(-sU[(lane + 32) * (FB + 1) + r], xr, y1);
        y0 = fma(-sU[lane * (FB + 1) + r], xr, y0);
      }
#pragma unroll 8
      for (int r = 31; r >= 0; --r) {
        const double xr = __shfl_sync(0xffffffffu, y0 * d0, r);
        if (lane == r) y0 = xr;
        if (lane < r) y0 = fma(-sU[lane * (FB + 1) + r], xr, y0);
      }
      if (lane < nbk) a.rhs[J0 + lane] = y0;
      if (lane + 32 < nbk) a.rhs[J0 + lane + 32] = y1;
      sxp[lane]      = lane < nbk ? y0 : 0.0;
      sxp[lane + 32] = lane + 32 < nbk ? y1 : 0.0;
    }
    if (J > 0) storeU(sUn, vn);
    post_flag(a, a.flagX + J);
    __syncthreads();
    trace_ev(a, 4 + 8, J, J);
  }
  cp_async_wait<0>();
}


__global__ void __launch_bounds__(FT, 1) chol_fused_kernel(const FusedArgs a) {
  extern __shared__ __align__(16) double fsm[];
  double *b0    = fsm;                      // four [64][68] tile buffers
  double *b1    = b0 + FB * FPITCH;
  double *b2    = b1 + FB * FPITCH;
  double *b3    = b2 + FB * FPITCH;
  double *sDinv = b3 + FB * FPITCH;         // [64]
  double *sVec  = sDinv + FB;               // [6][64]
  int *sBad     = reinterpret_cast<int *>(sVec + 6 * FB);
  const int bid = blockIdx.x;
  const int Gw  = gridDim.x - 1;            // workers
  trace_ev(a, 0, 0, 0);
  const int nb = a.nb, nbc = a.nbc;
  const int T  = row_start(nb, nbc);

  if (bid == 0) {
    spine_factor(a, b0, b1, b2, sDinv, sBad);
    if (a.rhs != nullptr) spine_backsolve(a, b0, b1, b2, b3, sVec);
    return;
  }
  const int me = bid - 1;
  for (int k = 0; k < nb; ++k) {
    {   // panels of block row k; (k, k+1) is the spine's
      const int t0 = row_start(k, nbc);
      for (int J = k + 2; J < nbc; ++J)
        if ((t0 + (J - k)) % Gw == me) do_panel(a, k, J, b0, b1, sDinv);
    }
    if (k + 1 < nb) {   // updates with block row k, in tile order; (k+1, k+1) is the spine's
      const int t1 = row_start(k + 1, nbc);
      int t        = t1 + ((me - t1) % Gw + Gw) % Gw;
      int I        = k + 1;
      for (; t < T; t += Gw) {
        while (t >= row_start(I + 1, nbc)) ++I;
        const int J = I + (t - row_start(I, nbc));
        if (I == k + 1 && J == k + 1) continue;
        do_update(a, k, I, J, b0, b1, b2, false);
        if (J == I && I == k + 2) post_flag(a, a.flagTd + I);        // the spine applies phase I-1 itself
        if (J == I + 1 && I == k + 1) post_flag(a, a.flagTs + I);    // ready for the spine's panel solve
      }
    }
  }
  if (a.rhs == nullptr) return;
  for (int J = nb - 1; J >= 2; --J)
    for (int I = J - 2; I >= 0; --I)
      if ((row_start(I, nbc) + (J - I)) % Gw == me) do_backprod(a, I, J, sVec);
}

}   // namespace

static constexpr size_t FUSED_SMEM = (size_t) (4 * FB * FPITCH + 7 * FB) * sizeof(double) + 16;

// Largest order served by the single-launch path (flags and the contribution buffer are sized for it).
int chol_fused_max_n() { return 4096; }

// In-place factorisation of the upper triangle of dM (n x n, ld = ldm even, 16-byte aligned rows) and, when
// dRhs != nullptr, solution of M x = rhs in place.  info_host: 0 or the 1-based index of the first non-positive pivot.
int dpotrf_upper_solve_fused(ncm_sd_gpu_ctx *c, int n, double *dM, int ldm, double *dRhs, int *info_host) {
  return dpotrf_upper_solve_fused_on(c, c->stream, c->n_sm, n, dM, ldm, dRhs, info_host);
}

// The same on stream `st` with at most `max_ctas` CTAs (>= 2: the spine and one worker): a small block factorised next to other work
// (dist_chol.cu: the next diagonal block while the trailing update of the previous step still owns most SMs) does not have to wait
// for the whole device to drain, as a cooperative launch of one CTA per SM would.  One fused factorisation per context at a time
// (the dependency flags are the context's).
int dpotrf_upper_solve_fused_on(ncm_sd_gpu_ctx *c, cudaStream_t st, int max_ctas, int n, double *dM, int ldm, double *dRhs, int *info_host) {
  if (n <= 0 || n > chol_fused_max_n()) return c->fail(NCM_SD_GPU_EINVAL, "chol_fused: order out of range");
  const int nctas = std::max(2, std::min(max_ctas, c->n_sm));
  const int nb  = (n + FB - 1) / FB;
  const int nbc = nb + (dRhs != nullptr ? 1 : 0);
  const int nbm = chol_fused_max_n() / FB;
  const size_t n_flags = (size_t) nbm + (size_t) nbm * (nbm + 1) + nbm + (size_t) nbm * nbm + 2 * nbm + 8;
  if (c->chol_flags.cap == 0) {
    if (!c->chol_flags.reserve(n_flags * sizeof(int))) return c->fail(NCM_SD_GPU_ENOMEM, "chol_fused: out of device memory");
    NCM_CUDA_OK(c, cudaMemsetAsync(c->chol_flags.p, 0, c->chol_flags.cap, st));
    NCM_CUDA_OK(c, cudaFuncSetAttribute(chol_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) FUSED_SMEM));
    int per_sm = 0;
    NCM_CUDA_OK(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, chol_fused_kernel, FT, FUSED_SMEM));
    if (per_sm < 1) return c->fail(NCM_SD_GPU_ECUDA, "chol_fused: kernel does not fit on an SM");
    c->chol_epoch = 0;
  }
  if (!c->chol_part.reserve(((size_t) nb * nb * FB + (size_t) n + 64) * sizeof(double)))
    return c->fail(NCM_SD_GPU_ENOMEM, "chol_fused: out of device memory");
  FusedArgs a;
  a.M = dM; a.ldm = ldm; a.n = n; a.rhs = dRhs;
  a.part  = c->chol_part.as<double>();
  a.dinv  = a.part + (size_t) nb * nb * FB;
  int *f  = c->chol_flags.as<int>();
  a.info = f; a.abort_flag = f + 1;
  a.flagD = f + 8;
  a.flagP = a.flagD + nbm;
  a.flagX = a.flagP + (size_t) nbm * (nbm + 1);
  a.flagQ = a.flagX + nbm;
  a.flagTd = a.flagQ + (size_t) nbm * nbm;
  a.flagTs = a.flagTd + nbm;
  a.epoch = ++c->chol_epoch;
  a.nb = nb; a.nbc = nbc;
  a.trace = c->chol_trace; a.trace_cap = c->chol_trace_cap;
  NCM_CUDA_OK(c, cudaMemsetAsync(f, 0, 2 * sizeof(int), st));
  void *params[] = {&a};
  if (nctas == c->n_sm) {
    NCM_CUDA_OK(c, cudaLaunchCooperativeKernel((const void *) chol_fused_kernel, dim3(nctas), dim3(FT), params, FUSED_SMEM, st));
  } else {
    // A partial grid next to other kernels: a cooperative launch would wait until all its CTAs fit at once, i.e. (with the update CTAs
    // of the other stream owning whole SMs) until that kernel has drained.  An ordinary launch places the CTAs as SMs retire; they do
    // become co-resident -- far fewer CTAs than SMs, and the CTAs they wait behind belong to kernels that terminate -- so the flag
    // waits inside the kernel cannot deadlock.
    chol_fused_kernel<<<nctas, FT, FUSED_SMEM, st>>>(a);
    NCM_CUDA_OK(c, cudaGetLastError());
  }
  c->n_launches++;
  if (info_host != nullptr) {
    int h[2] = {0, 0};
    NCM_CUDA_OK(c, ncm_memcpy_async(c, h, f, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    NCM_CUDA_OK(c, cudaStreamSynchronize(st));
    if (h[1] != 0) return c->fail(NCM_SD_GPU_ECUDA, "chol_fused: a tile dependency was never satisfied (internal error)");
    *info_host = h[0];
  }
  return NCM_SD_GPU_OK;
}
