// Counter-based proposal sampling (throughput mode): one Philox4x32-10 stream per proposal row.
//
// Same distribution as ncm_stats_dist_sample (ncm_stats_dist.c:1565-1627) + kernel->sample
// (ncm_stats_dist_kernel_gauss.c:335-355, ncm_stats_dist_kernel_st.c:388-414):
//   i ~ Categorical(w) by bisection on the normalised cumulative weights, z ~ N(0, I_d),
//   x = theta_i + s U_i^T (h z),  s = 1 (Gauss) or sqrt(nu / chi2_nu) (Student-t).
// NOT stream-compatible with the reference's serial MT19937 (hard part a):
// the parity path keeps the draws on the host and only uses ncm_sd_gpu_sample_apply.
#include "ctx.h"

namespace {

struct Philox {
  uint32_t key[2];
  uint32_t ctr[4];
  uint32_t out[4];
  int have;
  __device__ void init(unsigned long long seed, unsigned long long row, unsigned long long offset) {
    key[0] = (uint32_t) seed;
    key[1] = (uint32_t) (seed >> 32);
    ctr[0] = (uint32_t) offset;
    ctr[1] = (uint32_t) (offset >> 32);
    ctr[2] = (uint32_t) row;
    ctr[3] = (uint32_t) (row >> 32);
    have   = 0;
  }
  __device__ void round10() {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
    if (++ctr[0] == 0) ++ctr[1];
    have = 4;
  }
  __device__ uint32_t next() {
    if (have == 0) round10();
    return out[--have];
  }
  // uniform in (0, 1) with 53 random bits
  __device__ double uniform_pos() {
    const uint32_t a = next(), b = next();
    const unsigned long long v = (((unsigned long long) a << 32) | b) >> 11;
    return ((double) v + 0.5) * (1.0 / 9007199254740992.0);
  }
  __device__ double normal() {
    const double u1 = uniform_pos(), u2 = uniform_pos();
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
  }
  // Marsaglia-Tsang gamma(a, 1), a > 0
  __device__ double gamma(double a) {
    double boost = 1.0;
    if (a < 1.0) {
      boost = pow(uniform_pos(), 1.0 / a);
      a += 1.0;
    }
    const double dd = a - 1.0 / 3.0, cc = (1.0 / 3.0) / sqrt(dd);
    for (int it = 0; it < 64; ++it) {
      double x, v;
      do {
        x = normal();
        v = 1.0 + cc * x;
      } while (v <= 0.0);
      v = v * v * v;
      const double u = uniform_pos();
      if (u < 1.0 - 0.0331 * x * x * x * x || log(u) < 0.5 * x * x + dd * (1.0 - v + log(v))) return boost * dd * v;
    }
    return boost * dd;
  }
};

__global__ void wcum_kernel(const double *__restrict__ w, int n, double *__restrict__ wcum) {
  // wcum[0] = 0, wcum[i + 1] = sum_{j <= i} w_j, then scaled by 1 / total (ncm_stats_dist.c:1571-1585);
  // single block: per-thread chunk sums, serial scan of the 1024 chunk totals, then chunk-local prefix sums
  __shared__ double tot[1024];
  const int t = threadIdx.x, nt = blockDim.x;
  const int chunk = (n + nt - 1) / nt;
  const int b = t * chunk, e = min(n, b + chunk);
  double s = 0.0;
  for (int i = b; i < e; ++i) s += w[i];
  tot[t] = s;
  __syncthreads();
  if (t == 0) {
    double run = 0.0;
    for (int i = 0; i < nt; ++i) {
      const double v = tot[i];
      tot[i]         = run;
      run += v;
    }
    wcum[n + 1] = run;   // total, scratch slot
  }
  __syncthreads();
  const double inv = 1.0 / wcum[n + 1];
  double run       = tot[t];
  if (t == 0) wcum[0] = 0.0;
  for (int i = b; i < e; ++i) {
    run += w[i];
    wcum[i + 1] = run * inv;
  }
}

__global__ void philox_sample_kernel(const double *__restrict__ centres, const double *__restrict__ U_all, const double *__restrict__ wcum, int n,
                                     int d, int kind, double nu, double href, unsigned long long seed, unsigned long long offset, int q,
                                     double *__restrict__ X, int ldx, int *__restrict__ kidx) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= q) return;
  Philox g;
  g.init(seed, (unsigned long long) r, offset);
  const double p = g.uniform_pos();
  int ilo = 0, ihi = n;
  while (ihi > ilo + 1) {
    const int mi = (ihi + ilo) / 2;
    if (wcum[mi] > p)
      ihi = mi;
    else
      ilo = mi;
  }
  const int i = ilo;
  double z[NCM_SD_GPU_MAX_DIM];
  for (int k = 0; k < d; ++k) z[k] = g.normal() * href;
  double s = 1.0;
  if (kind == NCM_SD_GPU_KERNEL_ST) s = sqrt(nu / (2.0 * g.gamma(0.5 * nu)));
  const double *U = U_all + (size_t) i * d * d;
  for (int k = 0; k < d; ++k) {
    double t = 0.0;
    for (int j = 0; j <= k; ++j) t = fma(U[j * d + k], z[j], t);
    X[(size_t) r * ldx + k] = fma(s, t, centres[(size_t) i * d + k]);
  }
  if (kidx != nullptr) kidx[r] = i;
}

}   // namespace

int sample_philox_launch(ncm_sd_gpu_ctx *c, int q, unsigned long long seed, unsigned long long offset, double *dX, int ldx, int *dIdx) {
  if (!c->nn_x.reserve((size_t) (c->n_kernels + 8) * sizeof(double))) return c->fail(NCM_SD_GPU_ENOMEM, "sample: out of device memory");
  wcum_kernel<<<1, 1024, 0, c->stream>>>(c->weights.as<double>(), c->n_kernels, c->nn_x.as<double>());
  philox_sample_kernel<<<(q + 127) / 128, 128, 0, c->stream>>>(c->sample.as<double>(), c->Ufull.as<double>(), c->nn_x.as<double>(), c->n_kernels,
                                                             c->d, c->kind, c->nu, c->href, seed, offset, q, dX, ldx, dIdx);
  c->n_launches += 2;
  NCM_CUDA_OK(c, cudaGetLastError());
  return NCM_SD_GPU_OK;
}
