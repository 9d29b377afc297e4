// VKDE (per-centre covariance) kernels: batched eval_m2lnp / eval and the interpolation matrix.
//
// Replaces _ncm_stats_dist_vkde_eval_weights{,_m2lnp} (ncm_stats_dist_vkde.c:631-723) and
// _ncm_stats_dist_vkde_compute_IM (ncm_stats_dist_vkde.c:517-606):
//     chi2_i(x) = | U_i^-T (x - theta_i) |^2 / h^2
// One thread owns one query point (coordinates in registers); the CTA streams the centres'
// packed records  { theta_i[DP], L_i = U_i^T packed lower, diagonal stored as 1/U_kk }
// through shared memory with bulk async copies (cp.async.bulk + mbarrier, double buffered);
// every lane of a warp reads the same record element (shared-memory broadcast) while doing
// the forward substitution in registers, then the kernel function and an online
// log-sum-exp.  The centre range is split across gridDim.y so that small query batches
// still fill 148 SMs; partial (max, sum) pairs are merged by lse_finalize_kernel.
#include <cstdlib>
#include <cstring>
#include "ctx.h"

static constexpr double VKDE_MMA_MAX_COND = 1.0e5;   // see vkde_mma.cu: chi2 then stays within ~1e5 eps of the substitution

namespace {

template <int DP>
struct VkdeCfg {
  static constexpr int REC = (DP + DP * (DP + 1) / 2 + 1) & ~1;   // doubles per record (even => 16 B multiple)
  static constexpr int CH  = DP <= 8 ? 64 : DP <= 12 ? 32 : DP <= 16 ? 16 : DP <= 24 ? 8 : 4;   // centres per stage
  static constexpr int TQ  = 128;                                  // threads per CTA
  static constexpr int QPT = DP <= 20 ? 2 : 1;                     // queries per thread: each factor element read from shared memory feeds QPT
                                                                   // FMAs (the kernel is LSU-bound otherwise: 84 % LSU wavefronts at d = 10, ncu r01c)
};

// ---- record packing -------------------------------------------------------------------------------
__global__ void vkde_pack_kernel(const double *__restrict__ sample, const double *__restrict__ U_all, double *__restrict__ rec, int n,
                                 int d, int dp, int rec_len) {
  const int i = blockIdx.x;
  if (i >= n) return;
  const double *U = U_all + (size_t) i * d * d;
  double *r       = rec + (size_t) i * rec_len;
  for (int k = threadIdx.x; k < dp; k += blockDim.x) r[k] = k < d ? sample[(size_t) i * d + k] : 0.0;
  const int tri = dp * (dp + 1) / 2;
  for (int t = threadIdx.x; t < tri; t += blockDim.x) {
    // t = k (k + 1) / 2 + j, j <= k
    int k = (int) ((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
    while (k * (k + 1) / 2 > t) --k;
    while ((k + 1) * (k + 2) / 2 <= t) ++k;
    const int j = t - k * (k + 1) / 2;
    double v;
    if (k < d && j < d)
      v = (j == k) ? 1.0 / U[k * d + k] : U[j * d + k];   // L[k][j] = U[j][k]
    else
      v = (j == k) ? 1.0 : 0.0;
    r[dp + t] = v;
  }
  if (threadIdx.x == 0 && dp + tri < rec_len) r[dp + tri] = 0.0;
}

struct VkdeArgs {
  const double *X;        // queries [q x ldx]
  int ldx, q, d;
  const double *rec;      // packed records
  const double *cvec;     // eval: ln w_i - lnu_i ; IM: 1 / exp(lnu_i + d ln h)
  int n;                  // centres
  int per_split;          // centres per gridDim.y slice (multiple of CH)
  double inv_h2;
  KernParams kp;
  // eval outputs
  double *part_m, *part_s;   // [gridDim.y x q]
  // IM outputs
  double *IM;             // [q x ldim]
  int ldim;
  const double *rowscale; // [q] or null
  const int *only_if;     // when set: the launch is a repair pass and returns at once unless *only_if != 0
};

template <int DP, int MODE>   // MODE 0: log-sum-exp partials ; MODE 1: IM entries
__global__ void __launch_bounds__(VkdeCfg<DP>::TQ) vkde_kernel(const VkdeArgs a) {
  using Cfg = VkdeCfg<DP>;
  if (a.only_if != nullptr && *a.only_if == 0) return;
  constexpr int REC = Cfg::REC, CH = Cfg::CH;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *srec      = reinterpret_cast<double *>(smem_raw);             // [2][CH * REC]
  double *scv       = srec + 2 * CH * REC;                              // [2][CH]
  uint64_t *bars    = reinterpret_cast<uint64_t *>(scv + 2 * CH);       // [2]

  const int tid = threadIdx.x;
  constexpr int QPT = Cfg::QPT;
  int qi[QPT];
  bool qv[QPT];
#pragma unroll
  for (int u = 0; u < QPT; ++u) {
    qi[u] = (blockIdx.x * QPT + u) * Cfg::TQ + tid;
    qv[u] = qi[u] < a.q;
  }

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

  double x[QPT][DP];
#pragma unroll
  for (int u = 0; u < QPT; ++u)
#pragma unroll
    for (int k = 0; k < DP; ++k) x[u][k] = (qv[u] && k < a.d) ? a.X[(size_t) qi[u] * a.ldx + k] : 0.0;

  Lse acc[QPT];
  double rs[QPT];
#pragma unroll
  for (int u = 0; u < QPT; ++u) {
    lse_init(acc[u]);
    rs[u] = (MODE == 1 && qv[u] && a.rowscale != nullptr) ? a.rowscale[qi[u]] : 1.0;
  }

  for (int ch = 0; ch < nch; ++ch) {
    if (tid == 0 && ch + 1 < nch) issue(ch + 1);
    mbar_wait(&bars[ch & 1], (ch >> 1) & 1);
    const int c0       = c_begin + ch * CH;
    const int cnt      = min(CH, c_end - c0);
    const double *base = srec + (ch & 1) * CH * REC;
    const double *cv   = scv + (ch & 1) * CH;

    for (int c = 0; c < cnt; ++c) {
      const double *r = base + c * REC;
      const double *L = r + DP;
      double y[QPT][DP];
      double chi2[QPT];
#pragma unroll
      for (int u = 0; u < QPT; ++u) chi2[u] = 0.0;
#pragma unroll
      for (int k = 0; k < DP; ++k) {
        const double th = r[k];
        double t[QPT];
#pragma unroll
        for (int u = 0; u < QPT; ++u) t[u] = x[u][k] - th;
#pragma unroll
        for (int j = 0; j < k; ++j) {
          const double l = L[k * (k + 1) / 2 + j];
#pragma unroll
          for (int u = 0; u < QPT; ++u) t[u] = fma(-l, y[u][j], t[u]);
        }
        const double dinv = L[k * (k + 1) / 2 + k];
#pragma unroll
        for (int u = 0; u < QPT; ++u) {
          y[u][k] = t[u] * dinv;
          chi2[u] = fma(y[u][k], y[u][k], chi2[u]);
        }
      }
      const double cvc = cv[c];
#pragma unroll
      for (int u = 0; u < QPT; ++u) {
        const double c2 = chi2[u] * a.inv_h2;
        if (MODE == 0) {
          if (a.kp.lin)
            lin_push(acc[u], a.kp, c2, cvc);   // cvc = exp(ln w_i - lnu_i - cmax)
          else
            lse_push(acc[u], kern_lnK(a.kp, c2) + cvc);
        } else {
          if (qv[u]) a.IM[(size_t) qi[u] * a.ldim + (c0 + c)] = kern_K(a.kp, c2) * cvc * rs[u];
        }
      }
    }
    __syncthreads();   // everyone is done with stage (ch & 1) before it is refilled
  }

  if (MODE == 0) {
#pragma unroll
    for (int u = 0; u < QPT; ++u)
      if (qv[u]) {
        a.part_m[(size_t) blockIdx.y * a.q + qi[u]] = a.kp.lin ? *a.kp.cmax : acc[u].m;
        a.part_s[(size_t) blockIdx.y * a.q + qi[u]] = acc[u].s;
      }
  }
}

template <int DP, int MODE>
int launch_t(ncm_sd_gpu_ctx *c, const VkdeArgs &a, int n_splits) {
  using Cfg = VkdeCfg<DP>;
  const size_t smem = (size_t) (2 * Cfg::CH * Cfg::REC + 2 * Cfg::CH) * sizeof(double) + 2 * sizeof(uint64_t);
  static bool attr_set[NCM_MAX_DEVICES] = {};   // function attributes are per device
  if (!attr_set[c->device % NCM_MAX_DEVICES]) {
    NCM_CUDA_OK(c, cudaFuncSetAttribute(vkde_kernel<DP, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    attr_set[c->device % NCM_MAX_DEVICES] = true;
  }
  dim3 grid((a.q + Cfg::TQ * Cfg::QPT - 1) / (Cfg::TQ * Cfg::QPT), n_splits);
  vkde_kernel<DP, MODE><<<grid, Cfg::TQ, smem, c->stream>>>(a);
  c->n_launches++;
  NCM_CUDA_OK(c, cudaGetLastError());
  return NCM_SD_GPU_OK;
}

template <int DP>
int ch_of() { return VkdeCfg<DP>::CH; }
template <int DP>
int qtile_of() { return VkdeCfg<DP>::TQ * VkdeCfg<DP>::QPT; }

#define VKDE_DISPATCH(DPV, CALL)                 \
  switch (DPV) {                                 \
    case 2: { constexpr int DP = 2; CALL; } break;     \
    case 3: { constexpr int DP = 3; CALL; } break;     \
    case 4: { constexpr int DP = 4; CALL; } break;     \
    case 5: { constexpr int DP = 5; CALL; } break;     \
    case 6: { constexpr int DP = 6; CALL; } break;     \
    case 7: { constexpr int DP = 7; CALL; } break;     \
    case 8: { constexpr int DP = 8; CALL; } break;     \
    case 10: { constexpr int DP = 10; CALL; } break;   \
    case 12: { constexpr int DP = 12; CALL; } break;   \
    case 14: { constexpr int DP = 14; CALL; } break;   \
    case 16: { constexpr int DP = 16; CALL; } break;   \
    case 20: { constexpr int DP = 20; CALL; } break;   \
    case 24: { constexpr int DP = 24; CALL; } break;   \
    case 28: { constexpr int DP = 28; CALL; } break;   \
    case 32: { constexpr int DP = 32; CALL; } break;   \
    default: return c->fail(NCM_SD_GPU_EINVAL, "vkde: unsupported padded dimension"); \
  }

}   // namespace

int vkde_pad_dim(int d) {
  static const int sizes[] = {2, 3, 4, 5, 6, 7, 8, 10, 12, 14, 16, 20, 24, 28, 32};
  for (int s : sizes)
    if (d <= s) return s;
  return -1;
}

static int vkde_rec_len(int dp) { return (dp + dp * (dp + 1) / 2 + 1) & ~1; }

static int vkde_ch(ncm_sd_gpu_ctx *c, int dp) {
  int ch = 0;
  VKDE_DISPATCH(dp, ch = ch_of<DP>());
  return ch;
}
static int vkde_qtile(ncm_sd_gpu_ctx *c, int dp) {
  int qt = 0;
  VKDE_DISPATCH(dp, qt = qtile_of<DP>());
  return qt;
}

int vkde_pack(ncm_sd_gpu_ctx *c, const double *dU_all) {
  c->dp       = vkde_pad_dim(c->d);
  c->vrec_len = vkde_rec_len(c->dp);
  if (!c->vrec.reserve((size_t) c->n_kernels * c->vrec_len * sizeof(double))) return c->fail(NCM_SD_GPU_ENOMEM, "vkde_pack: out of device memory");
  vkde_pack_kernel<<<c->n_kernels, 128, 0, c->stream>>>(c->sample.as<double>(), dU_all, c->vrec.as<double>(), c->n_kernels, c->d, c->dp,
                                                      c->vrec_len);
  c->n_launches++;
  NCM_CUDA_OK(c, cudaGetLastError());
  // tensor-core path (vkde_mma.cu) for d >= 13 when every factor is conditioned well enough for its explicit inverse;
  // $NCM_SD_GPU_VKDE = subst | mma overrides the choice (parity experiments)
  c->vkde_mma = false;
  static const char *force = getenv("NCM_SD_GPU_VKDE");
  const bool no_mma = force != nullptr && strcmp(force, "subst") == 0;
  if (vkde_mma_pad_dim(c->d) != 0 && !no_mma) {
    int rc = vkde_mma_pack(c, dU_all, &c->vkde_cond);
    if (rc != NCM_SD_GPU_OK) return rc;
    c->vkde_mma = (c->vkde_cond <= VKDE_MMA_MAX_COND) || (force != nullptr && strcmp(force, "mma") == 0);
  }
  return NCM_SD_GPU_OK;
}

// merge the per-split partials:  out = -2 (m + log s + shift)  or  exp(m + log s + shift)
// flag_out (linear-domain partials): raised when a sum is not safely above the underflow threshold -- the caller's repair pass then
// recomputes everything in the log domain.  only_if: this launch IS the repair pass.
__global__ void lse_finalize_kernel(const double *__restrict__ pm, const double *__restrict__ ps, const double *__restrict__ row_add, int q,
                                    int n_splits, double shift, int as_density, double *__restrict__ out, int *__restrict__ flag_out,
                                    const int *__restrict__ only_if) {
  if (only_if != nullptr && *only_if == 0) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= q) return;
  Lse a;
  a.m = pm[i];
  a.s = ps[i];
  for (int s = 1; s < n_splits; ++s) lse_merge(a, pm[(size_t) s * q + i], ps[(size_t) s * q + i]);
  if (flag_out != nullptr && !(a.s >= 1.0e-200)) atomicExch(flag_out, 1);   // also NaN
  const double ln = a.m + log(a.s) + shift + (row_add != nullptr ? row_add[i] : 0.0);
  out[i]          = as_density ? exp(ln) : -2.0 * ln;
}

int lse_finalize_launch(ncm_sd_gpu_ctx *c, const double *pm, const double *ps, const double *row_add, int q, int n_splits, double shift,
                        bool as_density, double *dOut, int *dFlagOut, const int *dOnlyIf) {
  lse_finalize_kernel<<<(q + 255) / 256, 256, 0, c->stream>>>(pm, ps, row_add, q, n_splits, shift, as_density ? 1 : 0, dOut, dFlagOut, dOnlyIf);
  c->n_launches++;
  NCM_CUDA_OK(c, cudaGetLastError());
  return NCM_SD_GPU_OK;
}

static void fill_kp(const ncm_sd_gpu_ctx *c, KernParams &kp, bool eval = false) { ncm_fill_kp(c, kp, eval); }

// choose the number of centre splits so that the grid is about two waves of resident CTAs
static int pick_splits(const ncm_sd_gpu_ctx *c, int q_tiles, int n, int ch, int ctas_per_sm) {
  return ncm_pick_splits(c->n_sm * ctas_per_sm, q_tiles, (n + ch - 1) / ch);
}

int vkde_eval_launch(ncm_sd_gpu_ctx *c, int q, const double *dX, int ldx, double *dOut, bool as_density) {
  if (c->vkde_mma) return vkde_mma_eval_launch(c, q, dX, ldx, dOut, as_density);
  const int dp      = c->dp;
  const int ch      = vkde_ch(c, dp);
  const int qtile   = vkde_qtile(c, dp);
  const int q_tiles = (q + qtile - 1) / qtile;
  int splits        = pick_splits(c, q_tiles, c->n_kernels, ch, 4);
  int per_split     = ((c->n_kernels + splits - 1) / splits + ch - 1) / ch * ch;
  splits            = (c->n_kernels + per_split - 1) / per_split;
  if (!c->part.reserve((size_t) 2 * splits * q * sizeof(double))) return c->fail(NCM_SD_GPU_ENOMEM, "vkde_eval: out of device memory");
  VkdeArgs a;
  a.X = dX; a.ldx = ldx; a.q = q; a.d = c->d;
  a.rec = c->vrec.as<double>();
  a.n = c->n_kernels; a.per_split = per_split;
  a.inv_h2 = 1.0 / (c->href * c->href);
  fill_kp(c, a.kp, true);
  a.cvec = a.kp.lin ? c->clin.as<double>() : c->cterm.as<double>();
  a.part_m = c->part.as<double>(); a.part_s = a.part_m + (size_t) splits * q;
  a.IM = nullptr; a.ldim = 0; a.rowscale = nullptr; a.only_if = nullptr;
  int *flag = a.kp.lin ? reinterpret_cast<int *>(c->clin.as<double>() + c->clin_n + 1) : nullptr;
  if (flag != nullptr) NCM_CUDA_OK(c, cudaMemsetAsync(flag, 0, sizeof(int), c->stream));
  int rc = NCM_SD_GPU_OK;
  VKDE_DISPATCH(dp, rc = (launch_t<DP, 0>(c, a, splits)));
  if (rc != NCM_SD_GPU_OK) return rc;
  // m2lnp = -2 (gamma + log1p(lambda) - d ln h), ncm_stats_dist_vkde.c:721
  rc = lse_finalize_launch(c, a.part_m, a.part_s, nullptr, q, splits, -c->d * log(c->href), as_density, dOut, flag, nullptr);
  if (rc != NCM_SD_GPU_OK || flag == nullptr) return rc;
  // repair pass of the linear-domain evaluation: the same launches in the log domain, which return at once unless a sum came out
  // too close to the underflow threshold (a query absurdly far from every centre)
  a.kp.lin = 0;
  a.cvec = c->cterm.as<double>(); a.only_if = flag;
  VKDE_DISPATCH(dp, rc = (launch_t<DP, 0>(c, a, splits)));
  if (rc != NCM_SD_GPU_OK) return rc;
  return lse_finalize_launch(c, a.part_m, a.part_s, nullptr, q, splits, -c->d * log(c->href), as_density, dOut, nullptr, flag);
}

__global__ void vkde_invnorm_kernel(const double *__restrict__ lnu, int n, double dlnh, double *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = 1.0 / exp(lnu[i] + dlnh);   // 1 / norm_i, ncm_stats_dist_vkde.c:601-603
}

int vkde_im_launch(ncm_sd_gpu_ctx *c, const double *dRowScale) {
  const int dp = c->dp;
  const int ch = vkde_ch(c, dp);
  const int q  = c->nrows;
  if (!c->nn_tmp.reserve((size_t) (c->n_kernels + 2) * sizeof(double))) return c->fail(NCM_SD_GPU_ENOMEM, "vkde_im: out of device memory");
  vkde_invnorm_kernel<<<(c->n_kernels + 255) / 256, 256, 0, c->stream>>>(c->lnu.as<double>(), c->n_kernels, c->d * log(c->href),
                                                                        c->nn_tmp.as<double>());
  c->n_launches++;
  if (c->vkde_mma) return vkde_mma_im_launch(c, c->nn_tmp.as<double>(), dRowScale);
  const int qtile   = vkde_qtile(c, dp);
  const int q_tiles = (q + qtile - 1) / qtile;
  int splits        = pick_splits(c, q_tiles, c->n_kernels, ch, 4);
  int per_split     = ((c->n_kernels + splits - 1) / splits + ch - 1) / ch * ch;
  splits            = (c->n_kernels + per_split - 1) / per_split;
  VkdeArgs a;
  a.X = c->sample.as<double>() + (size_t) c->row0 * c->d; a.ldx = c->d; a.q = q; a.d = c->d;
  a.rec = c->vrec.as<double>(); a.cvec = c->nn_tmp.as<double>();
  a.n = c->n_kernels; a.per_split = per_split;
  a.inv_h2 = 1.0 / (c->href * c->href);
  fill_kp(c, a.kp);
  a.part_m = a.part_s = nullptr;
  a.IM = c->IM.as<double>(); a.ldim = (c->n_kernels + 7) & ~7;
  a.rowscale = dRowScale != nullptr ? dRowScale + c->row0 : nullptr;
  a.only_if = nullptr;
  int rc = NCM_SD_GPU_OK;
  VKDE_DISPATCH(dp, rc = (launch_t<DP, 1>(c, a, splits)));
  return rc;
}
