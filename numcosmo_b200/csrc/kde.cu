// KDE (one shared covariance) kernels: batched eval_m2lnp / eval and the interpolation matrix on the
// FP64 tensor cores (DMMA.8x8x4).
//
// Replaces _ncm_stats_dist_kde_eval_weights{,_m2lnp} (ncm_stats_dist_kde.c:596-681) and
// _ncm_stats_dist_kde_compute_IM (ncm_stats_dist_kde.c:492-557).  With z = U^-T x (whitened, and
// centred on the whitened sample mean to keep |z| small -- chi2 is translation invariant)
//     chi2_ij = |a_i - b_j|^2 / h^2 = (|a_i|^2 + |b_j|^2 - 2 a_i . b_j) / h^2
// is one dense contraction.  Both norms ride along as two extra K columns, so a single GEMM yields
// the quantity the epilogue needs with no further adds:
//     A row i  = [ s_a * a_i , p_i , 1 ]      B column j = [ b_j , 1 , q_j ]      (K = d + 2, padded to 4)
//     Gauss eval :  s_a = 1/h^2,       p_i = 0 (added per row at the end), q_j = -|b_j|^2/(2h^2) + ln w_j  -> ln t_ij - alpha_i
//     Gauss IM   :  s_a = 1/h^2,       p_i = -|a_i|^2/(2h^2),              q_j = -|b_j|^2/(2h^2)           -> -chi2_ij / 2
//     Student-t  :  s_a = -2/(h^2 nu), p_i = |a_i|^2/(h^2 nu),             q_j = |b_j|^2/(h^2 nu)          -> chi2_ij / nu
// Centres are stored "fragment-major" (tile of 8 centres x k-step of 4 = 32 consecutive doubles in lane
// order), so a 64-centre chunk is one contiguous bulk async copy and every B-fragment load is a
// conflict-free 256-byte shared-memory read.  Each warp keeps the A fragments of its 16 query rows in
// registers for the whole kernel and folds every 16 x 64 accumulator tile into a running
// (max, sum) pair per row (online log-sum-exp), so nothing but Q doubles per centre-split is written.
#include "ctx.h"


namespace {

constexpr int KDE_THREADS = 256;          // 8 warps
constexpr int MI = 2;                     // m-tiles (8 rows) per warp  -> 16 queries per warp, 128 per CTA
constexpr int NI = 8;                     // n-tiles per chunk          -> 64 centres per chunk
constexpr int QT = (KDE_THREADS / 32) * MI * 8;
constexpr int CHK = NI * 8;

// ---- preparation -------------------------------------------------------------------------------------
__global__ void col_mean_kernel(const double *__restrict__ Z, int n, int d, double *__restrict__ mean) {
  // one block per column, fixed-order tree reduction (deterministic)
  __shared__ double sh[256];
  const int k = blockIdx.x;
  double s    = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += Z[(size_t) i * d + k];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) mean[k] = sh[0] / n;
}

__global__ void center_kernel(const double *__restrict__ Z, int n, int d, const double *__restrict__ mean, double *__restrict__ zc,
                              double *__restrict__ nrm2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int k = 0; k < d; ++k) {
    const double v        = Z[(size_t) i * d + k] - mean[k];
    zc[(size_t) i * d + k] = v;
    s                     = fma(v, v, s);
  }
  nrm2[i] = s;
}

// B operand, fragment-major: out[(tile * KS + ks) * 32 + lane] = Bext[8 tile + lane / 4][4 ks + lane % 4]
// mode 0 Gauss eval, 1 Gauss IM, 2 Student-t; cterm (Student-t eval) = ln w_j or NEG_BIG for padding
__global__ void build_b_kernel(const double *__restrict__ zc, const double *__restrict__ nrm2, const double *__restrict__ weights, int n,
                               int n_pad, int d, int KS, int mode, double inv_h2, double inv_nu, double *__restrict__ out,
                               double *__restrict__ cterm) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = n_pad / 8 * KS * 32;
  if (idx < total) {
    const int lane = idx & 31;
    const int ks   = (idx >> 5) % KS;
    const int tile = (idx >> 5) / KS;
    const int j    = tile * 8 + (lane >> 2);
    const int k    = ks * 4 + (lane & 3);
    double v       = 0.0;
    if (j < n) {
      if (k < d)
        v = zc[(size_t) j * d + k];
      else if (k == d)
        v = 1.0;
      else if (k == d + 1) {
        if (mode == 0)
          v = -0.5 * nrm2[j] * inv_h2 + log(weights[j]);
        else if (mode == 1)
          v = -0.5 * nrm2[j] * inv_h2;
        else
          v = nrm2[j] * inv_h2 * inv_nu;
      }
    } else if (k == d + 1 && mode == 0) {
      v = NCM_NEG_BIG;
    }
    out[idx] = v;
  }
  if (cterm != nullptr && idx < n_pad) cterm[idx] = (idx < n && weights != nullptr) ? log(weights[idx]) : NCM_NEG_BIG;
}

// A operand, row-major [q_pad x KP]; whiten = 1: rows are raw points x (v = U^-T x - zmean), else rows of zc
__global__ void build_a_kernel(const double *__restrict__ X, int ldx, int q, int q_pad, int d, int KP, int whiten,
                               const double *__restrict__ U, const double *__restrict__ zmean, int mode, double inv_h2, double inv_nu,
                               double *__restrict__ A, double *__restrict__ alpha) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= q_pad) return;
  double *a = A + (size_t) i * KP;
  if (i >= q) {
    for (int k = 0; k < KP; ++k) a[k] = 0.0;
    return;
  }
  double v[NCM_SD_GPU_MAX_DIM + 4];
  double na = 0.0;
  if (whiten) {
    // gsl_blas_dtrsv (CblasUpper, CblasTrans, CblasNonUnit, U, x): forward substitution with U^T
    for (int k = 0; k < d; ++k) {
      double t = X[(size_t) i * ldx + k];
      for (int j = 0; j < k; ++j) t = fma(-U[j * d + k], v[j], t);
      v[k] = t / U[k * d + k];
    }
    for (int k = 0; k < d; ++k) v[k] -= zmean[k];
  } else {
    for (int k = 0; k < d; ++k) v[k] = X[(size_t) i * ldx + k];
  }
  for (int k = 0; k < d; ++k) na = fma(v[k], v[k], na);
  const double sa = (mode == 2) ? -2.0 * inv_h2 * inv_nu : inv_h2;
  for (int k = 0; k < d; ++k) a[k] = sa * v[k];
  double p = 0.0;
  if (mode == 1) p = -0.5 * na * inv_h2;
  if (mode == 2) p = na * inv_h2 * inv_nu;
  a[d]     = p;
  a[d + 1] = 1.0;
  for (int k = d + 2; k < KP; ++k) a[k] = 0.0;
  if (alpha != nullptr) alpha[i] = (mode == 0) ? -0.5 * na * inv_h2 : 0.0;
}

struct KdeArgs {
  const double *A;       // [q_pad x KP]
  int q, q_pad;
  const double *bfrag;   // fragment-major centres
  const double *cterm;   // [n_pad] (Student-t eval)
  int n, n_pad;
  int per_split;         // centres per gridDim.y slice, multiple of 64
  KernParams kp;
  int mode;              // 0 Gauss eval, 1 Gauss IM, 2 ST eval, 3 ST IM
  double *part_m, *part_s;
  double *IM;
  int ldim;
  double im_scale;       // exp(-(lnnorm + d ln h))
  const double *rowscale;
  const int *only_if;    // repair pass of the linear-domain Student-t evaluation: return at once unless *only_if != 0
};

template <int KS, int MODE>   // MODE: 0 Gauss eval, 1 Gauss IM, 2 ST eval, 3 ST IM
__global__ void __launch_bounds__(KDE_THREADS) kde_kernel(const KdeArgs a) {
  if (a.only_if != nullptr && *a.only_if == 0) return;
  constexpr int KP    = KS * 4;
  constexpr int STAGE = CHK * KP;          // doubles per stage of B
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *sB     = reinterpret_cast<double *>(smem_raw);            // [2][STAGE]
  double *sC     = sB + 2 * STAGE;                                  // [2][CHK]  cterm
  uint64_t *bars = reinterpret_cast<uint64_t *>(sC + 2 * CHK);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lr = lane & 3, lc = lane >> 2;
  const int m0 = blockIdx.x * QT + warp * (MI * 8);

  const int c_begin = blockIdx.y * a.per_split;
  const int c_end   = min(a.n_pad, c_begin + a.per_split);
  const int nch     = (c_end - c_begin) / CHK;

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  __syncthreads();

  auto issue = [&](int ch) {
    const int st = ch & 1;
    const int c0 = c_begin + ch * CHK;
    uint32_t bytes = STAGE * sizeof(double);
    if (MODE == 2) bytes += CHK * sizeof(double);
    mbar_arrive_expect_tx(&bars[st], bytes);
    bulk_g2s(sB + st * STAGE, a.bfrag + (size_t) c0 * KP, STAGE * sizeof(double), &bars[st]);
    if (MODE == 2) bulk_g2s(sC + st * CHK, a.cterm + c0, CHK * sizeof(double), &bars[st]);
  };
  if (tid == 0 && nch > 0) issue(0);

  // A fragments of this warp's 16 rows: lane holds A[m0 + 8 mi + lc][4 ks + lr]
  double af[MI][KS];
#pragma unroll
  for (int mi = 0; mi < MI; ++mi) {
    const int row = m0 + mi * 8 + lc;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) af[mi][ks] = (row < a.q_pad) ? a.A[(size_t) row * KP + ks * 4 + lr] : 0.0;
  }

  Lse st[MI];
#pragma unroll
  for (int mi = 0; mi < MI; ++mi) lse_init(st[mi]);

  double rs[MI];
#pragma unroll
  for (int mi = 0; mi < MI; ++mi) {
    const int row = m0 + mi * 8 + lc;
    rs[mi] = ((MODE & 1) && a.rowscale != nullptr && row < a.q) ? a.rowscale[row] * a.im_scale : a.im_scale;
  }

  for (int ch = 0; ch < nch; ++ch) {
    if (tid == 0 && ch + 1 < nch) issue(ch + 1);
    mbar_wait(&bars[ch & 1], (ch >> 1) & 1);
    const double *b  = sB + (ch & 1) * STAGE;
    const double *ct = sC + (ch & 1) * CHK;
    const int c0     = c_begin + ch * CHK;

    double acc[MI][NI][2];
#pragma unroll
    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
      for (int ni = 0; ni < NI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int ni = 0; ni < NI; ++ni) {
        const double bf = b[(ni * KS + ks) * 32 + lane];
#pragma unroll
        for (int mi = 0; mi < MI; ++mi) dmma884(acc[mi][ni][0], acc[mi][ni][1], af[mi][ks], bf);
      }
    }

    if (MODE == 2 && a.kp.lin) {
      // linear domain: s += exp(ln w_j - cmax) (1 + chi2/nu)^(-(nu + d)/2), the power by rsqrt and multiplications (common.cuh)
#pragma unroll
      for (int ni = 0; ni < NI; ++ni) {
        const double2 cw = *reinterpret_cast<const double2 *>(ct + ni * 8 + 2 * lr);
#pragma unroll
        for (int mi = 0; mi < MI; ++mi) {
          st[mi].s = fma(cw.x, st_pow_u(a.kp, fmax(acc[mi][ni][0], 0.0)), st[mi].s);
          st[mi].s = fma(cw.y, st_pow_u(a.kp, fmax(acc[mi][ni][1], 0.0)), st[mi].s);
        }
      }
    } else if (MODE == 0 || MODE == 2) {
      if (MODE == 2) {
        // ln t = kappa log1p(chi2 / nu) + ln w_j
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) {
          const double2 cw = *reinterpret_cast<const double2 *>(ct + ni * 8 + 2 * lr);
#pragma unroll
          for (int mi = 0; mi < MI; ++mi) {
            acc[mi][ni][0] = fma(a.kp.kappa, log1p_nonneg_fast(fmax(acc[mi][ni][0], 0.0)), cw.x);
            acc[mi][ni][1] = fma(a.kp.kappa, log1p_nonneg_fast(fmax(acc[mi][ni][1], 0.0)), cw.y);
          }
        }
      }
#pragma unroll
      for (int mi = 0; mi < MI; ++mi) {
        double bm = acc[mi][0][0];
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) bm = fmax(bm, fmax(acc[mi][ni][0], acc[mi][ni][1]));
        if (bm > st[mi].m) {
          st[mi].s *= exp_nonpos_fast(st[mi].m - bm);
          st[mi].m = bm;
        }
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) {
          s0 += exp_nonpos_fast(acc[mi][ni][0] - st[mi].m);
          s1 += exp_nonpos_fast(acc[mi][ni][1] - st[mi].m);
        }
        st[mi].s += s0 + s1;
      }
    } else {
#pragma unroll
      for (int mi = 0; mi < MI; ++mi) {
        const int row = m0 + mi * 8 + lc;
        if (row < a.q) {
#pragma unroll
          for (int ni = 0; ni < NI; ++ni) {
            const int col = c0 + ni * 8 + 2 * lr;
            double k0, k1;
            if (MODE == 1) {
              k0 = exp_nonpos_fast(acc[mi][ni][0]);
              k1 = exp_nonpos_fast(acc[mi][ni][1]);
            } else if (a.kp.m2 > 0) {
              k0 = st_pow_u(a.kp, fmax(acc[mi][ni][0], 0.0));
              k1 = st_pow_u(a.kp, fmax(acc[mi][ni][1], 0.0));
            } else {
              k0 = exp_nonpos_fast(a.kp.kappa * log1p_nonneg_fast(fmax(acc[mi][ni][0], 0.0)));
              k1 = exp_nonpos_fast(a.kp.kappa * log1p_nonneg_fast(fmax(acc[mi][ni][1], 0.0)));
            }
            double *dst = a.IM + (size_t) row * a.ldim + col;
            if (col + 1 < a.n)
              *reinterpret_cast<double2 *>(dst) = make_double2(k0 * rs[mi], k1 * rs[mi]);
            else if (col < a.n)
              dst[0] = k0 * rs[mi];
          }
        }
      }
    }
    __syncthreads();
  }

  if (MODE == 0 || MODE == 2) {
#pragma unroll
    for (int mi = 0; mi < MI; ++mi) {
      // merge the 4 lanes that share a row
      lse_warp_reduce_xor(st[mi], 4);
      const int row = m0 + mi * 8 + lc;
      if (lr == 0 && row < a.q) {
        a.part_m[(size_t) blockIdx.y * a.q + row] = (MODE == 2 && a.kp.lin) ? *a.kp.cmax : st[mi].m;
        a.part_s[(size_t) blockIdx.y * a.q + row] = st[mi].s;
      }
    }
  }
}

template <int KS, int MODE>
int kde_launch_t(ncm_sd_gpu_ctx *c, const KdeArgs &a, int splits) {
  const size_t smem = (size_t) (2 * CHK * KS * 4 + 2 * CHK) * sizeof(double) + 2 * sizeof(uint64_t);
  static bool attr_set[NCM_MAX_DEVICES] = {};   // function attributes are per device
  if (!attr_set[c->device % NCM_MAX_DEVICES]) {
    NCM_CUDA_OK(c, cudaFuncSetAttribute(kde_kernel<KS, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    attr_set[c->device % NCM_MAX_DEVICES] = true;
  }
  dim3 grid((a.q + QT - 1) / QT, splits);
  kde_kernel<KS, MODE><<<grid, KDE_THREADS, smem, c->stream>>>(a);
  c->n_launches++;
  NCM_CUDA_OK(c, cudaGetLastError());
  return NCM_SD_GPU_OK;
}

template <int MODE>
int kde_launch_ks(ncm_sd_gpu_ctx *c, int KS, const KdeArgs &a, int splits) {
  switch (KS) {
    case 1: return kde_launch_t<1, MODE>(c, a, splits);
    case 2: return kde_launch_t<2, MODE>(c, a, splits);
    case 3: return kde_launch_t<3, MODE>(c, a, splits);
    case 4: return kde_launch_t<4, MODE>(c, a, splits);
    case 5: return kde_launch_t<5, MODE>(c, a, splits);
    case 6: return kde_launch_t<6, MODE>(c, a, splits);
    case 7: return kde_launch_t<7, MODE>(c, a, splits);
    case 8: return kde_launch_t<8, MODE>(c, a, splits);
    case 9: return kde_launch_t<9, MODE>(c, a, splits);
    default: return c->fail(NCM_SD_GPU_EINVAL, "kde: unsupported dimension");
  }
}

void fill_kp(const ncm_sd_gpu_ctx *c, KernParams &kp, bool eval = false) { ncm_fill_kp(c, kp, eval); }

int pick_splits(const ncm_sd_gpu_ctx *c, int q_tiles, int n_pad) { return ncm_pick_splits(c->n_sm * 2, q_tiles, n_pad / CHK); }

}   // namespace

static inline int kde_ks(int d) { return (d + 2 + 3) / 4; }
static inline int kde_npad(int n) { return (n + CHK - 1) / CHK * CHK; }

// zc, zmean, |zc|^2 from the uploaded whitened points (rows 0..n_kernels-1 are the centres)
int kde_prepare(ncm_sd_gpu_ctx *c, const double *dInvU) {
  const int d = c->d;
  c->kp       = kde_ks(d) * 4;
  const int n_pad = kde_npad(c->n_kernels);
  if (!c->zc.reserve((size_t) c->n_obs * d * sizeof(double)) || !c->zmean.reserve((size_t) (d + c->n_obs + 8) * sizeof(double)) ||
      !c->bfrag.reserve((size_t) 2 * n_pad * c->kp * sizeof(double)) || !c->cterm.reserve((size_t) (n_pad + 8) * sizeof(double)))
    return c->fail(NCM_SD_GPU_ENOMEM, "kde_prepare: out of device memory");
  col_mean_kernel<<<d, 256, 0, c->stream>>>(dInvU, c->n_kernels, d, c->zmean.as<double>());
  center_kernel<<<(c->n_obs + 255) / 256, 256, 0, c->stream>>>(dInvU, c->n_obs, d, c->zmean.as<double>(), c->zc.as<double>(),
                                                               c->zmean.as<double>() + d + 4);
  c->n_launches += 2;
  NCM_CUDA_OK(c, cudaGetLastError());
  return NCM_SD_GPU_OK;
}

// (re)build the eval B operand after a change of weights / bandwidth
int kde_set_weights(ncm_sd_gpu_ctx *c) {
  const int n_pad = kde_npad(c->n_kernels);
  const int KS    = c->kp / 4;
  const int total = n_pad / 8 * KS * 32;
  const int mode  = c->kind == NCM_SD_GPU_KERNEL_GAUSS ? 0 : 2;
  const int nthr  = total > n_pad ? total : n_pad;
  build_b_kernel<<<(nthr + 255) / 256, 256, 0, c->stream>>>(c->zc.as<double>(), c->zmean.as<double>() + c->d + 4, c->weights.as<double>(),
                                                          c->n_kernels, n_pad, c->d, KS, mode, 1.0 / (c->href * c->href), 1.0 / c->nu,
                                                          c->bfrag.as<double>(), c->cterm.as<double>());
  c->n_launches++;
  NCM_CUDA_OK(c, cudaGetLastError());
  return mode == 2 ? build_clin(c, c->cterm.as<double>(), n_pad) : NCM_SD_GPU_OK;
}

int kde_eval_launch(ncm_sd_gpu_ctx *c, int q, const double *dX, int ldx, double *dOut, bool as_density) {
  const int KS = c->kp / 4, KP = c->kp;
  const int q_pad = (q + QT - 1) / QT * QT;
  const int n_pad = kde_npad(c->n_kernels);
  const int q_tiles = q_pad / QT;
  int splits    = pick_splits(c, q_tiles, n_pad);
  int per_split = ((n_pad / CHK + splits - 1) / splits) * CHK;
  splits        = (n_pad + per_split - 1) / per_split;
  if (!c->qA.reserve((size_t) q_pad * (KP + 1) * sizeof(double)) || !c->part.reserve((size_t) 2 * splits * q * sizeof(double)))
    return c->fail(NCM_SD_GPU_ENOMEM, "kde_eval: out of device memory");
  const int mode  = c->kind == NCM_SD_GPU_KERNEL_GAUSS ? 0 : 2;
  double *dA      = c->qA.as<double>();
  double *dAlpha  = dA + (size_t) q_pad * KP;
  build_a_kernel<<<(q_pad + 127) / 128, 128, 0, c->stream>>>(dX, ldx, q, q_pad, c->d, KP, 1, c->kde_U.as<double>(), c->zmean.as<double>(), mode,
                                                           1.0 / (c->href * c->href), 1.0 / c->nu, dA, dAlpha);
  c->n_launches++;
  KdeArgs a;
  a.A = dA; a.q = q; a.q_pad = q_pad;
  a.bfrag = c->bfrag.as<double>();
  a.n = c->n_kernels; a.n_pad = n_pad; a.per_split = per_split;
  fill_kp(c, a.kp, mode == 2);
  a.cterm = a.kp.lin ? c->clin.as<double>() : c->cterm.as<double>();
  a.mode = mode;
  a.part_m = c->part.as<double>(); a.part_s = a.part_m + (size_t) splits * q;
  a.IM = nullptr; a.ldim = 0; a.im_scale = 1.0; a.rowscale = nullptr; a.only_if = nullptr;
  int *flag = a.kp.lin ? reinterpret_cast<int *>(c->clin.as<double>() + c->clin_n + 1) : nullptr;
  if (flag != nullptr) NCM_CUDA_OK(c, cudaMemsetAsync(flag, 0, sizeof(int), c->stream));
  int rc = (mode == 0) ? kde_launch_ks<0>(c, KS, a, splits) : kde_launch_ks<2>(c, KS, a, splits);
  if (rc != NCM_SD_GPU_OK) return rc;
  // m2lnp = -2 (gamma + log1p(lambda) - d ln h), gamma already includes -lnnorm (kde.c:679, _kernel_gauss.c:332)
  const double shift = -c->lnnorm - c->d * log(c->href);
  rc = lse_finalize_launch(c, a.part_m, a.part_s, mode == 0 ? dAlpha : nullptr, q, splits, shift, as_density, dOut, flag, nullptr);
  if (rc != NCM_SD_GPU_OK || flag == nullptr) return rc;
  // repair pass in the log domain (see vkde.cu)
  a.kp.lin = 0;
  a.cterm = c->cterm.as<double>(); a.only_if = flag;
  rc = kde_launch_ks<2>(c, KS, a, splits);
  if (rc != NCM_SD_GPU_OK) return rc;
  return lse_finalize_launch(c, a.part_m, a.part_s, nullptr, q, splits, shift, as_density, dOut, nullptr, flag);
}

int kde_im_launch(ncm_sd_gpu_ctx *c, const double *dRowScale) {
  const int KS = c->kp / 4, KP = c->kp;
  const int q     = c->nrows;
  const int q_pad = (q + QT - 1) / QT * QT;
  const int n_pad = kde_npad(c->n_kernels);
  const int mode  = c->kind == NCM_SD_GPU_KERNEL_GAUSS ? 1 : 2;
  if (!c->qA.reserve((size_t) q_pad * (KP + 1) * sizeof(double))) return c->fail(NCM_SD_GPU_ENOMEM, "kde_im: out of device memory");
  double *dA  = c->qA.as<double>();
  double *dB  = c->bfrag.as<double>() + (size_t) n_pad * KP;   // second half of bfrag: the IM operand
  const double inv_h2 = 1.0 / (c->href * c->href);
  {
    const int total = n_pad / 8 * KS * 32;
    build_b_kernel<<<(total + 255) / 256, 256, 0, c->stream>>>(c->zc.as<double>(), c->zmean.as<double>() + c->d + 4, nullptr, c->n_kernels, n_pad,
                                                             c->d, KS, mode, inv_h2, 1.0 / c->nu, dB, nullptr);
    build_a_kernel<<<(q_pad + 127) / 128, 128, 0, c->stream>>>(c->zc.as<double>() + (size_t) c->row0 * c->d, c->d, q, q_pad, c->d, KP, 0, nullptr,
                                                             nullptr, mode, inv_h2, 1.0 / c->nu, dA, nullptr);
    c->n_launches += 2;
  }
  const int q_tiles = q_pad / QT;
  int splits    = pick_splits(c, q_tiles, n_pad);
  int per_split = ((n_pad / CHK + splits - 1) / splits) * CHK;
  splits        = (n_pad + per_split - 1) / per_split;
  KdeArgs a;
  a.A = dA; a.q = q; a.q_pad = q_pad;
  a.bfrag = dB; a.cterm = nullptr;
  a.n = c->n_kernels; a.n_pad = n_pad; a.per_split = per_split;
  fill_kp(c, a.kp);
  a.mode = mode == 1 ? 1 : 3;
  a.part_m = a.part_s = nullptr;
  a.IM = c->IM.as<double>(); a.ldim = (c->n_kernels + 7) & ~7;
  a.im_scale = exp(-(c->lnnorm + c->d * log(c->href)));   // ncm_stats_dist_kde.c:556
  a.rowscale = dRowScale != nullptr ? dRowScale + c->row0 : nullptr;
  a.only_if = nullptr;
  return (mode == 1) ? kde_launch_ks<1>(c, KS, a, splits) : kde_launch_ks<3>(c, KS, a, splits);
}
