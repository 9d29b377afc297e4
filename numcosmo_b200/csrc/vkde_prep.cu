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
// knn_kernel   one CTA per centre: distances into shared memory, radix select of the k-th smallest, sort of the k winners.
// cov_kernel   one warp per centre: lanes own the (i, j) pairs of the covariance (sums in registers), neighbours streamed in order.
#include "ctx.h"

namespace {

__device__ __forceinline__ bool key_less(double d1, int i1, double d2, int i2) { return (d1 < d2) || (d1 == d2 && i1 < i2); }

// All squared distances centre x point, D[c][m] = sum_r (z_m[r] - z_c[r])^2 in index order with explicit round-to-nearest
// operations (kdtree.c:27-38 distance()): 64 x 64 tile per CTA, both point sets staged in shared memory, 4 x 4 pairs per
// thread.  One pass over the points per 64 centres instead of one per centre (64 GB -> 1 GB of L2 traffic at N = 16384).
// centres cbeg .. cbeg + n_kernels - 1 of Z (a rank's share of the centres in multi-rank mode); row c of D belongs to centre cbeg + c
__global__ void __launch_bounds__(256) dist_kernel(const double *__restrict__ Z, int n_obs, int cbeg, int n_kernels, int d, double *__restrict__ D, size_t ldd) {
  extern __shared__ __align__(16) unsigned char dist_smem[];
  double *sc = reinterpret_cast<double *>(dist_smem);   // [64][d + 1] centres
  double *sp = sc + 64 * (d + 1);                        // [64][d + 1] points
  const int c0 = blockIdx.y * 64, m0 = blockIdx.x * 64;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  for (int e = tid; e < 64 * d; e += 256) {
    const int r = e / d, q = e % d;
    sc[r * (d + 1) + q] = (c0 + r < n_kernels) ? Z[(size_t) (cbeg + c0 + r) * d + q] : 0.0;
    sp[r * (d + 1) + q] = (m0 + r < n_obs) ? Z[(size_t) (m0 + r) * d + q] : 0.0;
  }
  __syncthreads();
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  for (int r = 0; r < d; ++r) {
    double cv[4], pv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) cv[i] = sc[(ty + 16 * i) * (d + 1) + r];
#pragma unroll
    for (int j = 0; j < 4; ++j) pv[j] = sp[(tx + 16 * j) * (d + 1) + r];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double df = __dsub_rn(pv[j], cv[i]);
        acc[i][j]       = __dadd_rn(acc[i][j], __dmul_rn(df, df));
      }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int cidx = c0 + ty + 16 * i;
    if (cidx >= n_kernels) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + tx + 16 * j;
      if (m < n_obs) D[(size_t) cidx * ldd + m] = acc[i][j];
    }
  }
}

// Exact k nearest neighbours of one centre (CTA per centre): distances to shared memory, radix SELECT of the k-th smallest
// distance on the order-preserving bit patterns (6 digit passes over n_obs keys instead of a full bitonic sort of n_obs:
// 46 ms -> a few ms at n_obs = 16384), ordered compaction of the k winners (ties at the threshold are taken in index
// order, as the (distance, index) ordering of the reference's neighbour list demands), bitonic sort of those k only.
__global__ void __launch_bounds__(512) knn_kernel(const double *__restrict__ Z, const double *__restrict__ D, size_t ldd, int n_obs, int n_smem, int d,
                                                  int kpow2, int k, int *__restrict__ nbr, int cbeg) {
  extern __shared__ __align__(16) unsigned char knn_smem[];
  // distances: a shared-memory copy when it fits (n_smem = n_obs), else the row of the precomputed matrix in global memory
  // (n_smem = 0; the six selection passes and the compaction then stream it from L2)
  double *ssd  = reinterpret_cast<double *>(knn_smem);           // [n_smem]
  double *kd   = ssd + n_smem;                                   // [kpow2] selected distances
  int *ki      = reinterpret_cast<int *>(kd + kpow2);            // [kpow2] selected indices
  int *hist    = ki + kpow2;                                     // [2048]
  int *sscan   = hist + 2048;                                    // [16] warp totals
  double *st   = reinterpret_cast<double *>(sscan + 16);         // [d] target  (8-byte aligned: all counts above are even)
  __shared__ unsigned long long s_prefix;
  __shared__ int s_rank;
  const int c = blockIdx.x, tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const double *sd;
  if (n_smem == 0) {
    sd = D + (size_t) c * ldd;   // host guarantees D != nullptr here
  } else if (D != nullptr) {     // distances precomputed by dist_kernel
    const double *row = D + (size_t) c * ldd;
    for (int m = tid; m < n_obs; m += nt) ssd[m] = row[m];
    sd = ssd;
  } else {
    for (int r = tid; r < d; r += nt) st[r] = Z[(size_t) (cbeg + c) * d + r];
    __syncthreads();
    for (int m = tid; m < n_obs; m += nt) {
      // kdtree.c:27-38 distance(): sum of squared differences in index order (point - target), no contraction
      const double *p = Z + (size_t) m * d;
      double dist = 0.0;
      for (int r = 0; r < d; ++r) {
        const double df = __dsub_rn(p[r], st[r]);
        dist = __dadd_rn(dist, __dmul_rn(df, df));
      }
      ssd[m] = dist;
    }
    sd = ssd;
  }
  if (tid == 0) {
    s_prefix = 0ull;
    s_rank   = k;      // 1-based rank of the wanted key among the keys that match the prefix
  }
  __syncthreads();
  // radix select: digits of 11, 11, 11, 11, 11, 9 bits from the top (non-negative doubles order like their bit patterns)
  int shift = 64;
  unsigned long long mask = 0ull;
#pragma unroll 1
  for (int pass = 0; pass < 6; ++pass) {
    const int bits = pass < 5 ? 11 : 9;
    shift -= bits;
    const int nbins = 1 << bits;
    for (int b = tid; b < nbins; b += nt) hist[b] = 0;
    __syncthreads();
    const unsigned long long prefix = s_prefix;
    for (int m = tid; m < n_obs; m += nt) {
      const unsigned long long key = (unsigned long long) __double_as_longlong(sd[m]);
      if ((key & mask) == prefix) atomicAdd(&hist[(int) ((key >> shift) & (unsigned long long) (nbins - 1))], 1);
    }
    __syncthreads();
    // find the bin where the cumulative count reaches the rank: thread t scans bins [t * per, (t + 1) * per)
    const int per = (nbins + nt - 1) / nt;
    int local = 0;
    for (int b = tid * per; b < min(nbins, (tid + 1) * per); ++b) local += hist[b];
    int incl = local;
    for (int off = 1; off < 32; off <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += v;
    }
    if (lane == 31) sscan[warp] = incl;
    __syncthreads();
    int base = 0;
    for (int w = 0; w < warp; ++w) base += sscan[w];
    const int excl = base + incl - local;     // keys in the bins before this thread's range
    const int rank = s_rank;
    __syncthreads();
    if (rank > excl && rank <= excl + local) {   // exactly one thread
      int cum = excl;
      for (int b = tid * per; b < min(nbins, (tid + 1) * per); ++b) {
        if (rank <= cum + hist[b]) {
          s_prefix = prefix | ((unsigned long long) b << shift);
          s_rank   = rank - cum;
          break;
        }
        cum += hist[b];
      }
    }
    mask |= ((unsigned long long) (nbins - 1)) << shift;
    __syncthreads();
  }
  // threshold key T = s_prefix: take every key < T and, in index order, the first (rank) keys == T
  const unsigned long long T = s_prefix;
  const int need_eq          = s_rank;
  // ordered compaction: chunks of nt consecutive indices, block-wide exclusive scan of the two predicates
  __shared__ int s_eq_taken, s_out;
  if (tid == 0) { s_eq_taken = 0; s_out = 0; }
  __syncthreads();
  for (int m0 = 0; m0 < n_obs; m0 += nt) {
    const int m = m0 + tid;
    unsigned long long key = ~0ull;
    if (m < n_obs) key = (unsigned long long) __double_as_longlong(sd[m]);
    const int is_lt = (m < n_obs && key < T) ? 1 : 0;
    const int is_eq = (m < n_obs && key == T) ? 1 : 0;
    // inclusive scans inside the warp
    int a_lt = is_lt, a_eq = is_eq;
    for (int off = 1; off < 32; off <<= 1) {
      const int v1 = __shfl_up_sync(0xffffffffu, a_lt, off), v2 = __shfl_up_sync(0xffffffffu, a_eq, off);
      if (lane >= off) { a_lt += v1; a_eq += v2; }
    }
    if (lane == 31) { sscan[warp] = a_lt; hist[warp] = a_eq; }
    __syncthreads();
    int b_lt = 0, b_eq = 0, t_lt = 0, t_eq = 0;
    for (int w = 0; w < (nt >> 5); ++w) {
      if (w < warp) { b_lt += sscan[w]; b_eq += hist[w]; }
      t_lt += sscan[w];
      t_eq += hist[w];
    }
    const int eq_before = s_eq_taken + b_eq + a_eq - is_eq;   // equal keys with a smaller index
    const int take_eq   = is_eq && eq_before < need_eq;
    // position: all takers in index order; count takers before me in this chunk
    // takers = lt + (eq with eq_before < need_eq): eq takers before me in chunk = min(b_eq + a_eq - is_eq, max(0, need_eq - s_eq_taken))
    const int eq_room   = max(0, need_eq - s_eq_taken);
    const int eq_tk_bef = min(b_eq + a_eq - is_eq, eq_room);
    const int pos       = s_out + (b_lt + a_lt - is_lt) + eq_tk_bef;
    if (is_lt || take_eq) {
      kd[pos] = sd[m];
      ki[pos] = m;
    }
    __syncthreads();
    if (tid == 0) {
      s_out += t_lt + min(t_eq, eq_room);
      s_eq_taken += t_eq;
    }
    __syncthreads();
  }
  for (int j = k + tid; j < kpow2; j += nt) {
    kd[j] = INFINITY;
    ki[j] = 0x7fffffff;
  }
  __syncthreads();
  // bitonic sort of the k winners, ascending by (distance, index)
  for (int size = 2; size <= kpow2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < (kpow2 >> 1); t += nt) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool up = ((lo & size) == 0);
        const double d1 = kd[lo], d2 = kd[hi];
        const int i1 = ki[lo], i2 = ki[hi];
        const bool sw = up ? key_less(d2, i2, d1, i1) : key_less(d1, i1, d2, i2);
        if (sw) {
          kd[lo] = d2; kd[hi] = d1;
          ki[lo] = i2; ki[hi] = i1;
        }
      }
      __syncthreads();
    }
  }
  for (int j = tid; j < k; j += nt) nbr[(size_t) c * k + j] = ki[j];
}

// one warp per centre.  The covariance pairs (i < j) are dealt to the lanes once (NP per lane, decoded up front) and their
// running sums live in registers; per neighbour the lanes < d update mean / variance and publish  a_i = x_i - mean_i(new),
// b_j = x_j - mean_j(old)  in shared memory, then every lane adds a_i * b_j to its pairs.  The next neighbour's row is
// fetched while the current one is processed.  Arithmetic and order are those of NcmStatsVec (ncm_stats_vec.c:510-551).
template <int NP>
__global__ void __launch_bounds__(256) cov_kernel(const double *__restrict__ X /* raw sample [n_obs x d] */, int d, int k, int n_kernels,
                                                  const int *__restrict__ nbr, double *__restrict__ U_all, int *__restrict__ fail) {
  extern __shared__ __align__(16) unsigned char cov_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int c = blockIdx.x * wpb + warp;
  const int npair = d * (d - 1) / 2;
  // per-warp scratch: a[d], b[d], C[d*d]
  double *base = reinterpret_cast<double *>(cov_smem) + (size_t) warp * (2 * d + d * d);
  double *sa = base, *sb = sa + d, *Cm = sb + d;
  if (c >= n_kernels) return;
  int pi[NP], pj[NP];
  double Cr[NP];
#pragma unroll
  for (int q = 0; q < NP; ++q) {
    const int p = lane + 32 * q;
    int i = 0, rem = p < npair ? p : 0;
    while (rem >= d - 1 - i) { rem -= d - 1 - i; ++i; }
    pi[q] = i;
    pj[q] = i + 1 + rem;
    Cr[q] = 0.0;
  }
  double mean = 0.0, var = 0.0, weight = 0.0;   // lane i < d owns coordinate i
  const int *nb = nbr + (size_t) c * k;
  double x_next = (lane < d) ? X[(size_t) nb[0] * d + lane] : 0.0;
  for (int q = 0; q < k; ++q) {
    const double x_i = x_next;
    if (q + 1 < k && lane < d) x_next = X[(size_t) nb[q + 1] * d + lane];
    const double curweight = __dadd_rn(weight, 1.0);
    if (lane < d) {
      const double delta_i = __dsub_rn(x_i, mean);
      const double R_i     = __ddiv_rn(__dmul_rn(delta_i, 1.0), curweight);
      const double dvar    = __dmul_rn(__dmul_rn(weight, delta_i), R_i);
      sb[lane] = delta_i;                       // x_i - mean_i (old)
      mean     = __dadd_rn(mean, R_i);
      var      = __dadd_rn(var, dvar);
      sa[lane] = __dsub_rn(x_i, mean);          // x_i - mean_i (new)
    }
    __syncwarp();
#pragma unroll
    for (int qq = 0; qq < NP; ++qq) {
      if (lane + 32 * qq < npair) Cr[qq] = __dadd_rn(Cr[qq], __dmul_rn(__dmul_rn(1.0, sa[pi[qq]]), sb[pj[qq]]));
    }
    __syncwarp();
    weight = curweight;
  }
  // read-out: cov = C (diag = var) * bias_wt, bias_wt = 1 / (weight - weight2 / weight), weight2 = weight (unit weights)
  const double bias = __ddiv_rn(1.0, __dsub_rn(weight, __ddiv_rn(weight, weight)));
#pragma unroll
  for (int qq = 0; qq < NP; ++qq) {
    if (lane + 32 * qq < npair) {
      const double v = __dmul_rn(Cr[qq], bias);
      Cm[pi[qq] * d + pj[qq]] = v;
      Cm[pj[qq] * d + pi[qq]] = v;
    }
  }
  if (lane < d) Cm[lane * d + lane] = __dmul_rn(var, bias);
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
// centres [cbeg, cbeg + n_kernels) of the sample (dU_all / dFail already point at centre cbeg; dNbr is scratch for n_kernels lists)
int vkde_prepare_dev(ncm_sd_gpu_ctx *c, int n_obs, int n_kernels, int k, const double *dZ, const double *dX, int *dNbr, double *dU_all, int *dFail, int cbeg) {
  const int d = c->d;
  int kpow2 = 1;
  while (kpow2 < k) kpow2 <<= 1;
  const size_t smem_fixed = (size_t) kpow2 * (sizeof(double) + sizeof(int)) + (2048 + 16) * sizeof(int) + (size_t) d * sizeof(double) + 16;
  if (smem_fixed > 200 * 1024) return c->fail(NCM_SD_GPU_EINVAL, "vkde_prepare: too many neighbours for the shared-memory sort");
  int n_smem = n_obs;   // distances of one centre in shared memory when they fit
  if (smem_fixed + (size_t) n_obs * sizeof(double) > 220 * 1024) n_smem = 0;
  const size_t smem_knn = smem_fixed + (size_t) n_smem * sizeof(double);
  static size_t knn_attr_tab[NCM_MAX_DEVICES] = {};   // function attributes are per device
  size_t &knn_attr = knn_attr_tab[c->device % NCM_MAX_DEVICES];
  const size_t cov_attr = 0;   // the covariance kernel's attribute is set at every launch (four instantiations)
  if (smem_knn > knn_attr) {
    NCM_CUDA_OK(c, cudaFuncSetAttribute(knn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_knn));
    knn_attr = smem_knn;
  }
  const int threads = n_obs >= 1024 ? 512 : (n_obs >= 512 ? 256 : 128);
  // large problems: all distances first (one pass over the points per 64 centres), then the selection reads its row
  const double *dD = nullptr;
  size_t ldd       = 0;
  if ((size_t) n_kernels * n_obs >= ((size_t) 1 << 20)) {
    ldd = ((size_t) n_obs + 7) & ~(size_t) 7;
    if (c->dist.reserve((size_t) n_kernels * ldd * sizeof(double))) {
      const size_t smem_dist = (size_t) 2 * 64 * (d + 1) * sizeof(double);
      dim3 dgrid((n_obs + 63) / 64, (n_kernels + 63) / 64);
      dist_kernel<<<dgrid, 256, smem_dist, c->stream>>>(dZ, n_obs, cbeg, n_kernels, d, c->dist.as<double>(), ldd);
      c->n_launches++;
      dD = c->dist.as<double>();
    }   // else: not enough memory for the matrix, every CTA computes its own distances
  }
  if (n_smem == 0 && dD == nullptr)
    return c->fail(NCM_SD_GPU_ENOMEM, "vkde_prepare: the distance matrix does not fit in device memory and one row does not fit in shared memory");
  knn_kernel<<<n_kernels, threads, smem_knn, c->stream>>>(dZ, dD, ldd, n_obs, n_smem, d, kpow2, k, dNbr, cbeg);
  const int wpb = 8;
  const size_t smem_cov = (size_t) wpb * (2 * d + d * d) * sizeof(double);
  const int npair = d * (d - 1) / 2;
  const int np    = (npair + 31) / 32;
  const dim3 cgrid((n_kernels + wpb - 1) / wpb);
#define NCM_COV_LAUNCH(NPV)                                                                                                   \
  do {                                                                                                                        \
    if (smem_cov > cov_attr) {                                                                                                \
      NCM_CUDA_OK(c, cudaFuncSetAttribute(cov_kernel<NPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_cov));   \
    }                                                                                                                         \
    cov_kernel<NPV><<<cgrid, wpb * 32, smem_cov, c->stream>>>(dX, d, k, n_kernels, dNbr, dU_all, dFail);                     \
  } while (0)
  if (np <= 2) NCM_COV_LAUNCH(2);
  else if (np <= 4) NCM_COV_LAUNCH(4);
  else if (np <= 8) NCM_COV_LAUNCH(8);
  else NCM_COV_LAUNCH(16);
#undef NCM_COV_LAUNCH
  c->n_launches += 2;
  NCM_CUDA_OK(c, cudaGetLastError());
  return NCM_SD_GPU_OK;
}
