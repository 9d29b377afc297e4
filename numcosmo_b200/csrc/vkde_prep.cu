// VKDE prepare_kernel on the device: per-centre k nearest neighbours, local covariance, Cholesky factor.
//
// Replaces the OpenMP loop of _ncm_stats_dist_vkde_build_cov_array_kdtree (ncm_stats_dist_vkde.c:362-496):
//   * exact kNN of every centre among the n_obs whitened points, ordered by (distance, index) as the kd-tree's
//     red-black list orders them (kdtree.c:192-321, rb_knn_list.c:31-40; the centre itself is neighbour 0),
//   * NcmStatsVec online mean / covariance of the RAW neighbours appended in that order
//     (ncm_stats_vec.c:510-551, read-out :2375-2399),
//   * upper Cholesky factor of the covariance (ncm_matrix_cholesky_decomp 'U').
// Every floating-point operation is issued with the explicit round-to-nearest intrinsics (__dmul_rn,
// __dadd_rn, ...) in the order of the host mirror (numcosmo_b200/host/stats_dist.cc), so no FMA contraction can
// change a distance ordering or a covariance bit: the factors are bit-identical to the host path, which the
// tests compare with the oracle.  A non-positive pivot only raises a flag; the host then applies the
// reference's nearPD / diagonal fallback (kde.c:344-367) to that one matrix.
//
// knn_kernel   one CTA per centre: distances into shared memory, bitonic sort of (distance, index), first k indices out.
// cov_kernel   one warp per centre: lanes own the (i, j) pairs of the covariance, neighbours streamed in order.
#include "ctx.h"

namespace {

__device__ __forceinline__ bool key_less(double d1, int i1, double d2, int i2) { return (d1 < d2) || (d1 == d2 && i1 < i2); }

__global__ void __launch_bounds__(512) knn_kernel(const double *__restrict__ Z, int n_obs, int d, int npow2, int k, int *__restrict__ nbr) {
  extern __shared__ __align__(16) unsigned char knn_smem[];
  double *sd = reinterpret_cast<double *>(knn_smem);   // [npow2]
  int *si    = reinterpret_cast<int *>(sd + npow2);     // [npow2]
  double *st = reinterpret_cast<double *>(si + npow2);  // [d] target
  const int c = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  for (int r = tid; r < d; r += nt) st[r] = Z[(size_t) c * d + r];
  __syncthreads();
  for (int m = tid; m < npow2; m += nt) {
    double dist = INFINITY;
    if (m < n_obs) {
      // kdtree.c:27-38 distance(): sum of squared differences in index order (point - target), no contraction
      const double *p = Z + (size_t) m * d;
      dist = 0.0;
      for (int r = 0; r < d; ++r) {
        const double df = __dsub_rn(p[r], st[r]);
        dist = __dadd_rn(dist, __dmul_rn(df, df));
      }
    }
    sd[m] = dist;
    si[m] = m;
  }
  __syncthreads();
  // bitonic sort, ascending by (distance, index)
  for (int size = 2; size <= npow2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < (npow2 >> 1); t += nt) {
        const int lo = 2 * t - (t & (stride - 1));   // index with bit `stride` cleared
        const int hi = lo + stride;
        const bool up = ((lo & size) == 0);
        const double d1 = sd[lo], d2 = sd[hi];
        const int i1 = si[lo], i2 = si[hi];
        const bool sw = up ? key_less(d2, i2, d1, i1) : key_less(d1, i1, d2, i2);
        if (sw) {
          sd[lo] = d2; sd[hi] = d1;
          si[lo] = i2; si[hi] = i1;
        }
      }
      __syncthreads();
    }
  }
  for (int j = tid; j < k; j += nt) nbr[(size_t) c * k + j] = si[j];
}

// one warp per centre
__global__ void __launch_bounds__(256) cov_kernel(const double *__restrict__ X /* raw sample [n_obs x d] */, int d, int k, int n_kernels,
                                                  const int *__restrict__ nbr, double *__restrict__ U_all, int *__restrict__ fail) {
  extern __shared__ __align__(16) unsigned char cov_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int c = blockIdx.x * wpb + warp;
  const int npair = d * (d - 1) / 2;
  // per-warp scratch: x[d], mean_old[d], mean_new[d], var[d], C[d*d]
  double *base = reinterpret_cast<double *>(cov_smem) + (size_t) warp * (4 * d + d * d);
  double *sx = base, *mo = sx + d, *mn = mo + d, *var = mn + d, *Cm = var + d;
  if (c >= n_kernels) return;
  for (int i = lane; i < d; i += 32) { mo[i] = 0.0; mn[i] = 0.0; var[i] = 0.0; }
  for (int e = lane; e < d * d; e += 32) Cm[e] = 0.0;
  __syncwarp();
  double weight = 0.0;
  for (int q = 0; q < k; ++q) {
    const double *x = X + (size_t) nbr[(size_t) c * k + q] * d;
    const double curweight = __dadd_rn(weight, 1.0);
    for (int i = lane; i < d; i += 32) {
      const double x_i     = x[i];
      const double mean_i  = mo[i];
      const double delta_i = __dsub_rn(x_i, mean_i);
      const double R_i     = __ddiv_rn(__dmul_rn(delta_i, 1.0), curweight);
      const double dvar    = __dmul_rn(__dmul_rn(weight, delta_i), R_i);
      sx[i]  = x_i;
      mn[i]  = __dadd_rn(mean_i, R_i);
      var[i] = __dadd_rn(var[i], dvar);
    }
    __syncwarp();
    // pairs (i < j): dC = (x_i - mean_i_new) * (x_j - mean_j_old)
    for (int p = lane; p < npair; p += 32) {
      // p -> (i, j), i < j, row-major enumeration of the strict upper triangle
      int i = 0, rem = p;
      while (rem >= d - 1 - i) { rem -= d - 1 - i; ++i; }
      const int j = i + 1 + rem;
      const double dC = __dmul_rn(__dmul_rn(1.0, __dsub_rn(sx[i], mn[i])), __dsub_rn(sx[j], mo[j]));
      Cm[i * d + j] = __dadd_rn(Cm[i * d + j], dC);
    }
    __syncwarp();
    for (int i = lane; i < d; i += 32) mo[i] = mn[i];
    __syncwarp();
    weight = curweight;
  }
  // read-out: cov = C (diag = var) * bias_wt, bias_wt = 1 / (weight - weight2 / weight), weight2 = weight (unit weights)
  const double bias = __ddiv_rn(1.0, __dsub_rn(weight, __ddiv_rn(weight, weight)));
  for (int e = lane; e < d * d; e += 32) {
    const int i = e / d, j = e % d;
    const double v = (i == j) ? var[i] : (i < j ? Cm[i * d + j] : Cm[j * d + i]);
    Cm[e] = __dmul_rn(v, bias);
  }
  __syncwarp();
  // Cholesky by rows (host mirror ncm_b200_cholesky_upper): U_ii = sqrt(a_ii - sum_k U_ki^2), U_ij = (a_ij - sum_k U_ki U_kj) / U_ii
  int bad = 0;
  for (int i = 0; i < d; ++i) {
    double s = Cm[i * d + i];
    for (int kk = 0; kk < i; ++kk) s = __dsub_rn(s, __dmul_rn(Cm[kk * d + i], Cm[kk * d + i]));
    if (!(s > 0.0) || !isfinite(s)) { bad = 1; break; }
    const double uii = __dsqrt_rn(s);
    __syncwarp();
    if (lane == 0) Cm[i * d + i] = uii;
    for (int j = i + 1 + lane; j < d; j += 32) {
      double t = Cm[i * d + j];
      for (int kk = 0; kk < i; ++kk) t = __dsub_rn(t, __dmul_rn(Cm[kk * d + i], Cm[kk * d + j]));
      Cm[i * d + j] = __ddiv_rn(t, uii);
    }
    __syncwarp();
  }
  // the lower triangle keeps the covariance (as LAPACK leaves it); the host mirror only reads the upper part
  double *out = U_all + (size_t) c * d * d;
  for (int e = lane; e < d * d; e += 32) out[e] = Cm[e];
  if (lane == 0) fail[c] = bad;
}

}   // namespace

// dZ: whitened points [n_obs x d] (device), dX: raw points [n_obs x d] (device).  Outputs on the device:
// dU_all [n_kernels x d x d], dFail [n_kernels].  Returns NCM_SD_GPU_EINVAL when the shared-memory sort does not fit.
int vkde_prepare_dev(ncm_sd_gpu_ctx *c, int n_obs, int n_kernels, int k, const double *dZ, const double *dX, int *dNbr, double *dU_all, int *dFail) {
  const int d = c->d;
  int npow2 = 1;
  while (npow2 < n_obs) npow2 <<= 1;
  const size_t smem_knn = (size_t) npow2 * (sizeof(double) + sizeof(int)) + (size_t) d * sizeof(double) + 16;
  if (smem_knn > 200 * 1024) return c->fail(NCM_SD_GPU_EINVAL, "vkde_prepare: n_obs too large for the shared-memory neighbour sort");
  static size_t knn_attr = 0, cov_attr = 0;
  if (smem_knn > knn_attr) {
    NCM_CUDA_OK(c, cudaFuncSetAttribute(knn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_knn));
    knn_attr = smem_knn;
  }
  const int threads = npow2 >= 1024 ? 512 : (npow2 >= 512 ? 256 : 128);
  knn_kernel<<<n_kernels, threads, smem_knn, c->stream>>>(dZ, n_obs, d, npow2, k, dNbr);
  const int wpb = 8;
  const size_t smem_cov = (size_t) wpb * (4 * d + d * d) * sizeof(double);
  if (smem_cov > cov_attr) {
    NCM_CUDA_OK(c, cudaFuncSetAttribute(cov_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_cov));
    cov_attr = smem_cov;
  }
  cov_kernel<<<(n_kernels + wpb - 1) / wpb, wpb * 32, smem_cov, c->stream>>>(dX, d, k, n_kernels, dNbr, dU_all, dFail);
  c->n_launches += 2;
  NCM_CUDA_OK(c, cudaGetLastError());
  return NCM_SD_GPU_OK;
}
