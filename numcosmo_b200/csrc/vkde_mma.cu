// VKDE batched eval / interpolation matrix on the FP64 tensor cores (DMMA.8x8x4), for d >= 21.
//
// Same contract as vkde.cu (replaces _ncm_stats_dist_vkde_eval_weights{,_m2lnp}, ncm_stats_dist_vkde.c:631-723, and
// _ncm_stats_dist_vkde_compute_IM, ncm_stats_dist_vkde.c:517-606):   chi2_i(x) = | U_i^-T (x - theta_i) |^2 / h^2 .
// vkde.cu does the forward substitution per (query, centre) pair in registers; one shared-memory operand per FMA makes
// it LSU-bound at large d (37 % of the FP64 peak at d = 30, profiles/r01_sweep).  Here the per-centre solve is a small
// dense product with the explicit inverse  W_i = L_i^-1  (L_i = U_i^T, lower triangular, formed at pack time):
//     Y[8 queries x 8 components] += A[8 x 4] B[4 x 8],   A[q][k] = x_q[k] - theta_i[k],   B[k][n] = W_i[n][k]
// Only the k-steps at or below the diagonal are issued (20 DMMA per centre and 8 queries at d = 30 instead of 32).
// Records are "fragment-major": theta[DP], then for every (component tile t, k-step ks <= 2t+1) the 32 doubles of the B
// fragment in lane order, so every operand load is one conflict-free 256-byte shared-memory read and a chunk of centres
// is one bulk async copy.  chi2 is the sum of squares of a query row of Y: 8 FMAs per thread and two quad shuffles; then
// lane (query, centre mod 4) evaluates the kernel function and its online log-sum-exp for 4 centres at a time.
//
// Rounding: y = W (x - theta) instead of the substitution changes the error from ~eps to ~cond(L_i) eps.  The pack kernel
// returns max_i |L_i|_1 |W_i|_1; the host selects this path only below VKDE_MMA_MAX_COND (otherwise vkde.cu runs), which
// keeps chi2 within 1e5 * 2.2e-16 of the substitution result, far inside the 1e-10 parity bar.
#include <cstdlib>
#include <cstring>
#include "ctx.h"


namespace {

template <int DP>
struct MmaCfg {
  static constexpr int NT    = DP / 8;              // component tiles
  static constexpr int KS    = DP / 4;              // k-steps
  static constexpr int NFRAG = NT * (NT + 1);       // sum_t 2 (t + 1)
  static constexpr int REC   = DP + NFRAG * 32;     // doubles per record
  static constexpr int MQ    = 2;                   // query tiles (8 rows) per warp
  static constexpr int WARPS = 8;
  static constexpr int TQ    = WARPS * MQ * 8;      // 128 queries per CTA
  static constexpr int CH    = DP <= 16 ? 16 : 8;   // centres per stage (multiple of 4)
};

__device__ __forceinline__ int frag_index(int t, int ks) { return t * (t + 1) + ks; }

// one warp per centre: W = L^-1 by columns (lane j owns column j), fragment-major store, condition estimate
__global__ void vkde_mma_pack_kernel(const double *__restrict__ sample, const double *__restrict__ U_all, double *__restrict__ rec, int n, int d,
                                     int dp, int rec_len, unsigned long long *__restrict__ cond_max_bits) {
  extern __shared__ double psm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int i = blockIdx.x * wpb + warp;
  double *Lm = psm + (size_t) warp * 2 * 32 * 33;   // L[r][c] (lower), pitch 33
  double *Wm = Lm + 32 * 33;
  if (i >= n) return;
  const double *U = U_all + (size_t) i * d * d;
  for (int e = lane; e < 32 * 33; e += 32) { Lm[e] = 0.0; Wm[e] = 0.0; }
  __syncwarp();
  for (int e = lane; e < d * d; e += 32) {
    const int r = e / d, cidx = e % d;     // U[r][cidx], r <= cidx valid; L[cidx][r] = U[r][cidx]
    if (cidx >= r) Lm[cidx * 33 + r] = U[e];
  }
  __syncwarp();
  // column j of W: W[j][j] = 1 / L[j][j]; W[r][j] = -(sum_{k=j}^{r-1} L[r][k] W[k][j]) / L[r][r]
  if (lane < d) {
    const int j = lane;
    Wm[j * 33 + j] = 1.0 / Lm[j * 33 + j];
    for (int r = j + 1; r < d; ++r) {
      double s = 0.0;
      for (int k = j; k < r; ++k) s = fma(Lm[r * 33 + k], Wm[k * 33 + j], s);
      Wm[r * 33 + j] = -s / Lm[r * 33 + r];
    }
  }
  __syncwarp();
  // 1-norm condition estimate |L|_1 |W|_1 (max column sums)
  double cl = 0.0, cw = 0.0;
  if (lane < d) {
    for (int r = lane; r < d; ++r) {
      cl += fabs(Lm[r * 33 + lane]);
      cw += fabs(Wm[r * 33 + lane]);
    }
  }
  for (int off = 16; off > 0; off >>= 1) {
    cl = fmax(cl, __shfl_xor_sync(0xffffffffu, cl, off));
    cw = fmax(cw, __shfl_xor_sync(0xffffffffu, cw, off));
  }
  if (lane == 0) {
    double cnd = cl * cw;
    if (!(cnd == cnd)) cnd = INFINITY;   // NaN -> never select this path
    atomicMax(cond_max_bits, (unsigned long long) __double_as_longlong(cnd));   // positive doubles order like their bit patterns
  }
  double *r_out = rec + (size_t) i * rec_len;
  for (int k = lane; k < dp; k += 32) r_out[k] = k < d ? sample[(size_t) i * d + k] : 0.0;
  const int nt = dp / 8;
  for (int t = 0; t < nt; ++t)
    for (int ks = 0; ks <= 2 * t + 1; ++ks) {
      const int nn = 8 * t + (lane >> 2), kk = 4 * ks + (lane & 3);   // B[k][n] = W[n][k]
      r_out[dp + frag_index(t, ks) * 32 + lane] = (nn < d && kk < d && kk <= nn) ? Wm[nn * 33 + kk] : 0.0;
    }
}

struct MmaArgs {
  const double *X;
  int ldx, q, d;
  const double *rec;
  const double *cvec;
  int n, per_split;
  double inv_h2;
  KernParams kp;
  double *part_m, *part_s;
  double *IM;
  int ldim;
  const double *rowscale;
  const int *only_if;   // repair pass of the linear-domain evaluation: return at once unless *only_if != 0
};

template <int DP, int MODE>
__global__ void __launch_bounds__(MmaCfg<DP>::WARPS * 32, 2) vkde_mma_kernel(const MmaArgs a) {
  using Cfg = MmaCfg<DP>;
  if (a.only_if != nullptr && *a.only_if == 0) return;
  constexpr int REC = Cfg::REC, CH = Cfg::CH, NT = Cfg::NT, KS = Cfg::KS, MQ = Cfg::MQ;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *srec   = reinterpret_cast<double *>(smem_raw);   // [2][CH * REC]
  double *scv    = srec + 2 * CH * REC;                    // [2][CH]
  uint64_t *bars = reinterpret_cast<uint64_t *>(scv + 2 * CH);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lr = lane & 3, lc = lane >> 2;
  const int q0 = blockIdx.x * Cfg::TQ + warp * (MQ * 8);

  const int c_begin = blockIdx.y * a.per_split;
  const int c_end   = min(a.n, c_begin + a.per_split);
  const int nch     = (c_end - c_begin + CH - 1) / CH;

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  auto issue = [&](int ch) {
    const int st  = ch & 1;
    const int c0  = c_begin + ch * CH;
    const int cnt = min(CH, c_end - c0);
    const uint32_t b_rec = (uint32_t) (cnt * REC * sizeof(double));
    const uint32_t b_cv  = (uint32_t) (((cnt + 1) & ~1) * sizeof(double));
    mbar_arrive_expect_tx(&bars[st], b_rec + b_cv);
    bulk_g2s(srec + st * CH * REC, a.rec + (size_t) c0 * REC, b_rec, &bars[st]);
    bulk_g2s(scv + st * CH, a.cvec + c0, b_cv, &bars[st]);
  };
  if (tid == 0 && nch > 0) issue(0);

  // A-operand source: this lane's coordinates k = 4 ks + lr of the queries q0 + 8 mq + lc
  double xf[MQ][KS];
#pragma unroll
  for (int mq = 0; mq < MQ; ++mq) {
    const int qi = q0 + 8 * mq + lc;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const int k = 4 * ks + lr;
      xf[mq][ks]  = (qi < a.q && k < a.d) ? a.X[(size_t) qi * a.ldx + k] : 0.0;
    }
  }
  Lse accL[MQ];
  double rs[MQ];
#pragma unroll
  for (int mq = 0; mq < MQ; ++mq) {
    lse_init(accL[mq]);
    const int qi = q0 + 8 * mq + lc;
    rs[mq] = (MODE == 1 && qi < a.q && a.rowscale != nullptr) ? a.rowscale[qi] : 1.0;
  }

  for (int ch = 0; ch < nch; ++ch) {
    if (tid == 0 && ch + 1 < nch) issue(ch + 1);
    mbar_wait(&bars[ch & 1], (ch >> 1) & 1);
    const int c0       = c_begin + ch * CH;
    const int cnt      = min(CH, c_end - c0);
    const double *base = srec + (ch & 1) * CH * REC;
    const double *cv   = scv + (ch & 1) * CH;

    for (int g = 0; g < (cnt + 3) / 4; ++g) {
      double chi[MQ];
#pragma unroll
      for (int mq = 0; mq < MQ; ++mq) chi[mq] = 0.0;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const double *r = base + (4 * g + cc) * REC;
        double acc[MQ][NT][2];
#pragma unroll
        for (int mq = 0; mq < MQ; ++mq)
#pragma unroll
          for (int t = 0; t < NT; ++t) acc[mq][t][0] = acc[mq][t][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          const double th = r[4 * ks + lr];
          double af[MQ];
#pragma unroll
          for (int mq = 0; mq < MQ; ++mq) af[mq] = xf[mq][ks] - th;
#pragma unroll
          for (int t = ks / 2; t < NT; ++t) {
            const double b = r[DP + frag_index(t, ks) * 32 + lane];
#pragma unroll
            for (int mq = 0; mq < MQ; ++mq) dmma884(acc[mq][t][0], acc[mq][t][1], af[mq], b);
          }
        }
#pragma unroll
        for (int mq = 0; mq < MQ; ++mq) {
          double s = 0.0;
#pragma unroll
          for (int t = 0; t < NT; ++t) s = fma(acc[mq][t][1], acc[mq][t][1], fma(acc[mq][t][0], acc[mq][t][0], s));
          s += __shfl_xor_sync(0xffffffffu, s, 1);
          s += __shfl_xor_sync(0xffffffffu, s, 2);
          chi[mq] = (lr == cc) ? s : chi[mq];
        }
      }
      // lane (lc, lr): query 8 mq + lc, centre 4 g + lr
      const int cl      = 4 * g + lr;
      const bool cvalid = cl < cnt;
      const double cvv  = cvalid ? cv[cl] : 0.0;
#pragma unroll
      for (int mq = 0; mq < MQ; ++mq) {
        const double chi2 = chi[mq] * a.inv_h2;
        if (MODE == 0) {
          if (cvalid) {
            if (a.kp.lin)
              lin_push(accL[mq], a.kp, chi2, cvv);   // cvv = exp(ln w_i - lnu_i - cmax)
            else
              lse_push(accL[mq], kern_lnK(a.kp, chi2) + cvv);
          }
        } else {
          const int qi = q0 + 8 * mq + lc;
          if (cvalid && qi < a.q) a.IM[(size_t) qi * a.ldim + (c0 + cl)] = kern_K(a.kp, chi2) * cvv * rs[mq];
        }
      }
    }
    __syncthreads();   // everyone is done with stage (ch & 1) before it is refilled
  }

  if (MODE == 0) {
#pragma unroll
    for (int mq = 0; mq < MQ; ++mq) {
      lse_warp_reduce_xor(accL[mq], 4);   // the four centre classes of a query sit in one quad
      const int qi = q0 + 8 * mq + lc;
      if (lr == 0 && qi < a.q) {
        a.part_m[(size_t) blockIdx.y * a.q + qi] = a.kp.lin ? *a.kp.cmax : accL[mq].m;
        a.part_s[(size_t) blockIdx.y * a.q + qi] = accL[mq].s;
      }
    }
  }
}

template <int DP, int MODE>
int mma_launch_t(ncm_sd_gpu_ctx *c, const MmaArgs &a, int n_splits) {
  using Cfg = MmaCfg<DP>;
  const size_t smem = (size_t) (2 * Cfg::CH * Cfg::REC + 2 * Cfg::CH) * sizeof(double) + 2 * sizeof(uint64_t);
  static bool attr_set[NCM_MAX_DEVICES] = {};   // function attributes are per device
  if (!attr_set[c->device % NCM_MAX_DEVICES]) {
    NCM_CUDA_OK(c, cudaFuncSetAttribute(vkde_mma_kernel<DP, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    attr_set[c->device % NCM_MAX_DEVICES] = true;
  }
  dim3 grid((a.q + Cfg::TQ - 1) / Cfg::TQ, n_splits);
  vkde_mma_kernel<DP, MODE><<<grid, Cfg::WARPS * 32, smem, c->stream>>>(a);
  c->n_launches++;
  NCM_CUDA_OK(c, cudaGetLastError());
  return NCM_SD_GPU_OK;
}

#define MMA_DISPATCH(DPV, CALL)                        \
  switch (DPV) {                                       \
    case 16: { constexpr int DP = 16; CALL; } break;   \
    case 24: { constexpr int DP = 24; CALL; } break;   \
    case 32: { constexpr int DP = 32; CALL; } break;   \
    default: return c->fail(NCM_SD_GPU_EINVAL, "vkde_mma: unsupported padded dimension"); \
  }

int mma_rec_len(int dp) { return dp + (dp / 8) * (dp / 8 + 1) * 32; }
int mma_ch(int dp) { return dp <= 16 ? 16 : 8; }


void splits_for(const ncm_sd_gpu_ctx *c, int q, int n, int ch, int &splits, int &per_split) {
  const int q_tiles = (q + 127) / 128;
  splits            = ncm_pick_splits(c->n_sm * 2, q_tiles, (n + ch - 1) / ch);
  per_split = ((n + splits - 1) / splits + ch - 1) / ch * ch;
  splits    = (n + per_split - 1) / per_split;
}

}   // namespace

// padded dimension of the tensor-core path, or 0 when d is served by the substitution kernel
// (measured at 65536^2, profiles/r01_sweep: the DMMA path wins from d ~ 21 on; below, the padding to the 8-wide component
// tiles costs more than the substitution kernel's LSU limit; $NCM_SD_GPU_VKDE=mma lowers the threshold to 13 for experiments)
int vkde_mma_pad_dim(int d) {
  static const int d_min = [] {
    const char *e = getenv("NCM_SD_GPU_VKDE");
    return (e != nullptr && strcmp(e, "mma") == 0) ? 13 : 21;
  }();
  if (d < d_min || d > 32) return 0;
  return d <= 16 ? 16 : (d <= 24 ? 24 : 32);
}

// Builds the fragment-major records; *cond_max_host = max_i |L_i|_1 |W_i|_1.
int vkde_mma_pack(ncm_sd_gpu_ctx *c, const double *dU_all, double *cond_max_host) {
  const int dp = vkde_mma_pad_dim(c->d);
  if (dp == 0) return c->fail(NCM_SD_GPU_EINVAL, "vkde_mma_pack: dimension not served");
  c->vrec_mma_len = mma_rec_len(dp);
  if (!c->vrec_mma.reserve((size_t) c->n_kernels * c->vrec_mma_len * sizeof(double) + 64)) return c->fail(NCM_SD_GPU_ENOMEM, "vkde_mma_pack: out of device memory");
  if (!c->nn_f.reserve(64)) return c->fail(NCM_SD_GPU_ENOMEM, "vkde_mma_pack: out of device memory");
  unsigned long long *dcond = c->nn_f.as<unsigned long long>();
  NCM_CUDA_OK(c, cudaMemsetAsync(dcond, 0, sizeof(unsigned long long), c->stream));
  const int wpb     = 4;
  const size_t smem = (size_t) wpb * 2 * 32 * 33 * sizeof(double);
  static bool attr_set[NCM_MAX_DEVICES] = {};   // function attributes are per device
  if (!attr_set[c->device % NCM_MAX_DEVICES]) {
    NCM_CUDA_OK(c, cudaFuncSetAttribute(vkde_mma_pack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    attr_set[c->device % NCM_MAX_DEVICES] = true;
  }
  vkde_mma_pack_kernel<<<(c->n_kernels + wpb - 1) / wpb, wpb * 32, smem, c->stream>>>(c->sample.as<double>(), dU_all, c->vrec_mma.as<double>(), c->n_kernels,
                                                                                   c->d, dp, c->vrec_mma_len, dcond);
  c->n_launches++;
  NCM_CUDA_OK(c, cudaGetLastError());
  unsigned long long bits = 0;
  NCM_CUDA_OK(c, cudaMemcpyAsync(&bits, dcond, sizeof(bits), cudaMemcpyDeviceToHost, c->stream));
  NCM_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  double cnd;
  static_assert(sizeof(cnd) == sizeof(bits), "");
  memcpy(&cnd, &bits, sizeof(cnd));
  *cond_max_host = cnd;
  return NCM_SD_GPU_OK;
}

int vkde_mma_eval_launch(ncm_sd_gpu_ctx *c, int q, const double *dX, int ldx, double *dOut, bool as_density) {
  const int dp = vkde_mma_pad_dim(c->d);
  int splits, per_split;
  splits_for(c, q, c->n_kernels, mma_ch(dp), splits, per_split);
  if (!c->part.reserve((size_t) 2 * splits * q * sizeof(double))) return c->fail(NCM_SD_GPU_ENOMEM, "vkde_mma_eval: out of device memory");
  MmaArgs a;
  a.X = dX; a.ldx = ldx; a.q = q; a.d = c->d;
  a.rec = c->vrec_mma.as<double>();
  a.n = c->n_kernels; a.per_split = per_split;
  a.inv_h2 = 1.0 / (c->href * c->href);
  ncm_fill_kp(c, a.kp, true);
  a.cvec = a.kp.lin ? c->clin.as<double>() : c->cterm.as<double>();
  a.part_m = c->part.as<double>(); a.part_s = a.part_m + (size_t) splits * q;
  a.IM = nullptr; a.ldim = 0; a.rowscale = nullptr; a.only_if = nullptr;
  int *flag = a.kp.lin ? reinterpret_cast<int *>(c->clin.as<double>() + c->clin_n + 1) : nullptr;
  if (flag != nullptr) NCM_CUDA_OK(c, cudaMemsetAsync(flag, 0, sizeof(int), c->stream));
  int rc = NCM_SD_GPU_OK;
  MMA_DISPATCH(dp, rc = (mma_launch_t<DP, 0>(c, a, splits)));
  if (rc != NCM_SD_GPU_OK) return rc;
  rc = lse_finalize_launch(c, a.part_m, a.part_s, nullptr, q, splits, -c->d * log(c->href), as_density, dOut, flag, nullptr);
  if (rc != NCM_SD_GPU_OK || flag == nullptr) return rc;
  // repair pass in the log domain (see vkde.cu): returns at once unless a linear-domain sum came out next to the underflow threshold
  a.kp.lin = 0;
  a.cvec = c->cterm.as<double>(); a.only_if = flag;
  MMA_DISPATCH(dp, rc = (mma_launch_t<DP, 0>(c, a, splits)));
  if (rc != NCM_SD_GPU_OK) return rc;
  return lse_finalize_launch(c, a.part_m, a.part_s, nullptr, q, splits, -c->d * log(c->href), as_density, dOut, nullptr, flag);
}

// cvec = 1 / exp(lnu_i + d ln h) prepared by the caller (vkde_im_launch)
int vkde_mma_im_launch(ncm_sd_gpu_ctx *c, const double *dInvNorm, const double *dRowScale) {
  const int dp = vkde_mma_pad_dim(c->d);
  const int q  = c->nrows;
  int splits, per_split;
  splits_for(c, q, c->n_kernels, mma_ch(dp), splits, per_split);
  MmaArgs a;
  a.X = c->sample.as<double>() + (size_t) c->row0 * c->d; a.ldx = c->d; a.q = q; a.d = c->d;
  a.rec = c->vrec_mma.as<double>(); a.cvec = dInvNorm;
  a.n = c->n_kernels; a.per_split = per_split;
  a.inv_h2 = 1.0 / (c->href * c->href);
  ncm_fill_kp(c, a.kp, false);
  a.only_if = nullptr;
  a.part_m = a.part_s = nullptr;
  a.IM = c->IM.as<double>(); a.ldim = (c->n_kernels + 7) & ~7;
  a.rowscale = dRowScale != nullptr ? dRowScale + c->row0 : nullptr;
  int rc = NCM_SD_GPU_OK;
  MMA_DISPATCH(dp, rc = (mma_launch_t<DP, 1>(c, a, splits)));
  return rc;
}
