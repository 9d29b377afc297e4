// C ABI of libncm_sd_gpu (include/ncm_sd_gpu.h): context, uploads, batched evaluation, IM, NNLS,
// sampling.  Everything below runs on the context's CUDA stream; there is no CPU path.
#include <cstring>
#include <new>
#include <algorithm>
#include "ctx.h"
#include "nccl_shim.h"

int dsyrk_ata_general(ncm_sd_gpu_ctx *c, int K, int n, const double *dP, int ldp, double *dC, int ldc, double alpha, double beta);
int dpotrf_upper_solve(ncm_sd_gpu_ctx *c, int n, double *dM, int ldm, double *dRhs, double *dDinv, int *dInfo, int *info_host);
int sample_philox_launch(ncm_sd_gpu_ctx *c, int q, unsigned long long seed, unsigned long long offset, double *dX, int ldx, int *dIdx);



bool DevBuf::reserve(size_t bytes) {
  if (bytes <= cap) return true;
  if (p != nullptr) cudaFree(p);
  p   = nullptr;
  cap = 0;
  size_t want = bytes + bytes / 8 + 256;
  if (cudaMalloc(&p, want) != cudaSuccess) {
    cudaGetLastError();
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
      cudaGetLastError();
      p = nullptr;
      return false;
    }
    want = bytes;
  }
  cap = want;
  return true;
}
void DevBuf::release() {
  if (p != nullptr) cudaFree(p);
  p   = nullptr;
  cap = 0;
}
bool PinBuf::reserve(size_t bytes) {
  if (bytes <= cap) return true;
  if (p != nullptr) cudaFreeHost(p);
  p   = nullptr;
  cap = 0;
  const size_t want = bytes + bytes / 8 + 256;
  if (cudaMallocHost(&p, want) != cudaSuccess) {
    cudaGetLastError();
    p = nullptr;
    return false;
  }
  cap = want;
  return true;
}
void PinBuf::release() {
  if (p != nullptr) cudaFreeHost(p);
  p   = nullptr;
  cap = 0;
}

namespace {

__global__ void cterm_kernel(const double *__restrict__ w, const double *__restrict__ lnu, int n, int n_alloc, double *__restrict__ out) {
  // ln w_i - lnu_i ; ln 0 = -inf is mapped to a large negative finite value (term vanishes in exp, _kernel_gauss.c:270)
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_alloc) return;
  if (i < n) {
    const double wi = w[i];
    out[i]          = (wi > 0.0) ? log(wi) - lnu[i] : NCM_NEG_BIG;
  } else {
    out[i] = NCM_NEG_BIG;
  }
}

__global__ void sample_apply_kernel(const double *__restrict__ centres, int ldc, const double *__restrict__ U_all, int per_kernel_U, int d,
                                    const int *__restrict__ kidx, const double *__restrict__ Z, int ldz, const double *__restrict__ scale,
                                    double href, int q, double *__restrict__ X, int ldx) {
  // x = theta_i + s * U_i^T (h z):  (U^T y)_k = sum_{j <= k} U[j][k] y_j     (gsl_blas_dtrmv Upper/Trans)
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= q) return;
  const int i     = kidx[r];
  const double *U = U_all + (per_kernel_U ? (size_t) i * d * d : 0);
  const double s  = scale != nullptr ? scale[r] : 1.0;
  for (int k = 0; k < d; ++k) {
    double t = 0.0;
    for (int j = 0; j <= k; ++j) t = fma(U[j * d + k], Z[(size_t) r * ldz + j] * href, t);
    X[(size_t) r * ldx + k] = fma(s, t, centres[(size_t) i * ldc + k]);
  }
}

// clin[i] = exp(cterm[i] - cmax), cmax = max_i cterm[i] (single CTA: two passes over at most a few 10^5 entries); clin[n] = cmax
__global__ void clin_kernel(const double *__restrict__ cterm, int n, double *__restrict__ clin) {
  __shared__ double red[32];
  __shared__ double s_max;
  double m = NCM_NEG_BIG;
  for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmax(m, cterm[i]);
  for (int off = 16; off > 0; off >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, off));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int) (blockDim.x >> 5); ++w) m = fmax(m, red[w]);
    s_max   = m;
    clin[n] = m;
  }
  __syncthreads();
  const double cm = s_max;
  for (int i = threadIdx.x; i < n; i += blockDim.x) clin[i] = exp(cterm[i] - cm);   // zero weights (NEG_BIG) -> exactly 0
}

int check_ready(ncm_sd_gpu_ctx *c, bool need_weights) {
  if (c == nullptr) return NCM_SD_GPU_EINVAL;
  if (c->type < 0 || c->n_kernels <= 0) return c->fail(NCM_SD_GPU_EINVAL, "no centres uploaded");
  if (need_weights && !c->have_weights) return c->fail(NCM_SD_GPU_EINVAL, "weights / bandwidth not set");
  return NCM_SD_GPU_OK;
}

}   // namespace

bool st_linear_enabled(const ncm_sd_gpu_ctx *c) {
  static const bool env_on = [] {
    const char *e = getenv("NCM_SD_GPU_ST_LINEAR");
    return !(e != nullptr && e[0] == '0');
  }();
  const double m = c->nu + c->d;
  return env_on && c->kind == NCM_SD_GPU_KERNEL_ST && c->nu >= 1.0 && c->nu == floor(c->nu) && m < 128.0;
}

void ncm_fill_kp(const ncm_sd_gpu_ctx *c, KernParams &kp, bool eval) {
  kp.kind   = c->kind;
  kp.nu     = c->nu;
  kp.kappa  = -0.5 * (c->nu + c->d);
  kp.inv_nu = 1.0 / c->nu;
  kp.m2     = st_linear_enabled(c) ? (int) (c->nu + c->d) : 0;
  kp.lin    = (eval && kp.m2 > 0 && c->clin_n > 0) ? 1 : 0;
  kp.cmax   = kp.lin ? c->clin.as<double>() + c->clin_n : nullptr;
}

// linear-domain weights of the Student-t evaluation from the log-domain ones (dCterm [n_alloc], padding = NEG_BIG)
int build_clin(ncm_sd_gpu_ctx *c, const double *dCterm, int n_alloc) {
  c->clin_n = 0;
  if (!st_linear_enabled(c)) return NCM_SD_GPU_OK;
  if (!c->clin.reserve((size_t) (n_alloc + 4) * sizeof(double))) return c->fail(NCM_SD_GPU_ENOMEM, "clin: out of device memory");
  clin_kernel<<<1, 1024, 0, c->stream>>>(dCterm, n_alloc, c->clin.as<double>());
  c->n_launches++;
  NCM_CUDA_OK(c, cudaGetLastError());
  c->clin_n = n_alloc;
  return NCM_SD_GPU_OK;
}

int update_cterm(ncm_sd_gpu_ctx *c) {
  if (c->type == NCM_SD_GPU_VKDE) {
    const int n_alloc = c->n_kernels + 2;
    if (!c->cterm.reserve((size_t) n_alloc * sizeof(double))) return c->fail(NCM_SD_GPU_ENOMEM, "cterm: out of device memory");
    cterm_kernel<<<(n_alloc + 255) / 256, 256, 0, c->stream>>>(c->weights.as<double>(), c->lnu.as<double>(), c->n_kernels, n_alloc,
                                                             c->cterm.as<double>());
    c->n_launches++;
    NCM_CUDA_OK(c, cudaGetLastError());
    return build_clin(c, c->cterm.as<double>(), n_alloc);
  }
  return kde_set_weights(c);
}

extern "C" {

int ncm_sd_gpu_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int ncm_sd_gpu_ctx_new(ncm_sd_gpu_ctx **out, int device) {
  if (out == nullptr) return NCM_SD_GPU_EINVAL;
  *out  = nullptr;
  int n = ncm_sd_gpu_device_count();
  if (n <= 0 || device < 0 || device >= n) return NCM_SD_GPU_ENODEV;
  if (cudaSetDevice(device) != cudaSuccess) return NCM_SD_GPU_ENODEV;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return NCM_SD_GPU_ENODEV;
  if (prop.major < 10) return NCM_SD_GPU_ENODEV;   // sm_100a code only
  ncm_sd_gpu_ctx *c = new (std::nothrow) ncm_sd_gpu_ctx();
  if (c == nullptr) return NCM_SD_GPU_ENOMEM;
  c->device = device;
  c->n_sm   = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&c->ev0) != cudaSuccess ||
      cudaEventCreate(&c->ev1) != cudaSuccess || cudaEventCreate(&c->ev2) != cudaSuccess || cudaEventCreate(&c->ev3) != cudaSuccess) {
    delete c;
    return NCM_SD_GPU_ECUDA;
  }
  *out = c;
  return NCM_SD_GPU_OK;
}

int ncm_sd_gpu_ctx_free(ncm_sd_gpu_ctx *c) {
  if (c == nullptr) return NCM_SD_GPU_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (c->nccl_comm != nullptr && nccl_api().ok) nccl_api().CommDestroy((ncclComm_t) c->nccl_comm);
  DevBuf *bufs[] = {&c->sample, &c->vrec, &c->lnu, &c->cterm, &c->weights, &c->Ufull, &c->zc, &c->zmean, &c->bfrag, &c->kde_U, &c->qX,
                    &c->qOut, &c->qA, &c->part, &c->IM, &c->rowscale, &c->M, &c->MU, &c->nn_b, &c->nn_x, &c->nn_r, &c->nn_g, &c->nn_tmp,
                    &c->nn_idx, &c->nn_f, &c->chol_flags, &c->chol_part, &c->vrec_mma, &c->dist, &c->lrW, &c->lrWt, &c->lrS, &c->lrV, &c->lrT, &c->lrPart, &c->lrSmall, &c->lrVec, &c->lrIdx, &c->gath, &c->dcPack, &c->dcW, &c->dcStage, &c->dcVec, &c->dcTiles, &c->bkWork, &c->qrWork, &c->clin};
  for (DevBuf *b : bufs) b->release();
  c->pinX.release();
  c->pinOut.release();
  c->pin_nn.release();
  cudaEventDestroy(c->ev0);
  cudaEventDestroy(c->ev1);
  cudaEventDestroy(c->ev2);
  cudaEventDestroy(c->ev3);
  if (c->ev_panel != nullptr) cudaEventDestroy(c->ev_panel);
  if (c->ev_tail != nullptr) cudaEventDestroy(c->ev_tail);
  for (cudaEvent_t e : {c->dc_evP, c->dc_evA, c->dc_evD, c->dc_evW, c->dc_evU})
    if (e != nullptr) cudaEventDestroy(e);
  if (c->dc_sW != nullptr) cudaStreamDestroy(c->dc_sW);
  if (c->dc_sP != nullptr) cudaStreamDestroy(c->dc_sP);
  if (c->stream_hi != nullptr) cudaStreamDestroy(c->stream_hi);
  cudaStreamDestroy(c->stream);
  delete c;
  return NCM_SD_GPU_OK;
}

const char *ncm_sd_gpu_last_error(const ncm_sd_gpu_ctx *c) { return c != nullptr ? c->err.c_str() : "null context"; }
void *ncm_sd_gpu_stream(ncm_sd_gpu_ctx *c) { return c != nullptr ? (void *) c->stream : nullptr; }

int ncm_sd_gpu_synchronize(ncm_sd_gpu_ctx *c) {
  if (c == nullptr) return NCM_SD_GPU_EINVAL;
  NCM_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return NCM_SD_GPU_OK;
}

int ncm_sd_gpu_set_kernel(ncm_sd_gpu_ctx *c, int kind, double nu, int d) {
  if (c == nullptr) return NCM_SD_GPU_EINVAL;
  if (kind != NCM_SD_GPU_KERNEL_GAUSS && kind != NCM_SD_GPU_KERNEL_ST) return c->fail(NCM_SD_GPU_EINVAL, "unknown kernel kind");
  if (d < 1 || d > NCM_SD_GPU_MAX_DIM) return c->fail(NCM_SD_GPU_EINVAL, "dimension out of range [1, 32]");
  if (kind == NCM_SD_GPU_KERNEL_ST && !(nu > 0.0)) return c->fail(NCM_SD_GPU_EINVAL, "nu must be positive");
  c->kind = kind;
  c->nu   = kind == NCM_SD_GPU_KERNEL_ST ? nu : 1.0;
  c->d    = d;
  c->type = -1;
  c->have_weights = false;
  c->clin_n = 0;
  return NCM_SD_GPU_OK;
}

int ncm_sd_gpu_upload_kde(ncm_sd_gpu_ctx *c, int n_obs, int n_kernels, const double *invU, int ld, const double *U, int ldu, double lnnorm) {
  if (c == nullptr) return NCM_SD_GPU_EINVAL;
  if (c->d <= 0) return c->fail(NCM_SD_GPU_EINVAL, "set_kernel first");
  if (n_kernels <= 0 || n_obs < n_kernels || invU == nullptr || U == nullptr || ld < c->d || ldu < c->d) return c->fail(NCM_SD_GPU_EINVAL, "upload_kde: bad arguments");
  cudaSetDevice(c->device);
  const int d = c->d;
  StageTimer t(c, NCM_SD_GPU_T_H2D);
  if (!c->qX.reserve((size_t) n_obs * d * sizeof(double)) || !c->kde_U.reserve((size_t) d * d * sizeof(double)) ||
      !c->weights.reserve((size_t) (n_kernels + 8) * sizeof(double)))
    return c->fail(NCM_SD_GPU_ENOMEM, "upload_kde: out of device memory");
  NCM_CUDA_OK(c, ncm_memcpy2d_async(c,c->qX.p, d * sizeof(double), invU, ld * sizeof(double), d * sizeof(double), n_obs, cudaMemcpyHostToDevice, c->stream));
  NCM_CUDA_OK(c, ncm_memcpy2d_async(c,c->kde_U.p, d * sizeof(double), U, ldu * sizeof(double), d * sizeof(double), d, cudaMemcpyHostToDevice, c->stream));
  c->type      = NCM_SD_GPU_KDE;
  c->n_obs     = n_obs;
  c->n_kernels = n_kernels;
  c->lnnorm    = lnnorm;
  c->row0      = 0;
  c->nrows     = n_obs;
  c->have_weights = false;
  int rc = kde_prepare(c, c->qX.as<double>());
  if (rc != NCM_SD_GPU_OK) return rc;
  NCM_CUDA_OK(c, cudaStreamSynchronize(c->stream));   // host buffers may be reused by the caller
  return NCM_SD_GPU_OK;
}

int ncm_sd_gpu_upload_vkde(ncm_sd_gpu_ctx *c, int n_obs, int n_kernels, const double *sample, int ld, const double *U_all, const double *lnnorms) {
  if (c == nullptr) return NCM_SD_GPU_EINVAL;
  if (c->d <= 0) return c->fail(NCM_SD_GPU_EINVAL, "set_kernel first");
  if (n_kernels <= 0 || n_obs < n_kernels || sample == nullptr || U_all == nullptr || lnnorms == nullptr || ld < c->d)
    return c->fail(NCM_SD_GPU_EINVAL, "upload_vkde: bad arguments");
  cudaSetDevice(c->device);
  const int d = c->d;
  StageTimer t(c, NCM_SD_GPU_T_H2D);
  if (!c->sample.reserve((size_t) n_obs * d * sizeof(double)) || !c->Ufull.reserve((size_t) n_kernels * d * d * sizeof(double)) ||
      !c->lnu.reserve((size_t) (n_kernels + 8) * sizeof(double)) || !c->weights.reserve((size_t) (n_kernels + 8) * sizeof(double)))
    return c->fail(NCM_SD_GPU_ENOMEM, "upload_vkde: out of device memory");
  NCM_CUDA_OK(c, ncm_memcpy2d_async(c,c->sample.p, d * sizeof(double), sample, ld * sizeof(double), d * sizeof(double), n_obs, cudaMemcpyHostToDevice, c->stream));
  NCM_CUDA_OK(c, ncm_memcpy_async(c,c->Ufull.p, U_all, (size_t) n_kernels * d * d * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  NCM_CUDA_OK(c, ncm_memcpy_async(c,c->lnu.p, lnnorms, (size_t) n_kernels * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  c->type      = NCM_SD_GPU_VKDE;
  c->n_obs     = n_obs;
  c->n_kernels = n_kernels;
  c->row0      = 0;
  c->nrows     = n_obs;
  c->have_weights = false;
  int rc = vkde_pack(c, c->Ufull.as<double>());
  if (rc != NCM_SD_GPU_OK) return rc;
  NCM_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return NCM_SD_GPU_OK;
}

int ncm_sd_gpu_vkde_prepare(ncm_sd_gpu_ctx *c, int n_obs, int n_kernels, const double *sample, int ld, const double *invUsample, int ldz, int k,
                            double *U_all_out, int *fail_out) {
  if (c == nullptr) return NCM_SD_GPU_EINVAL;
  if (c->d <= 0) return c->fail(NCM_SD_GPU_EINVAL, "set_kernel first");
  if (n_kernels <= 0 || n_obs < n_kernels || sample == nullptr || invUsample == nullptr || U_all_out == nullptr || fail_out == nullptr ||
      ld < c->d || ldz < c->d || k < 2 || k > n_obs)
    return c->fail(NCM_SD_GPU_EINVAL, "vkde_prepare: bad arguments");
  cudaSetDevice(c->device);
  const int d = c->d;
  // multi-rank (auto-shard) mode: the kNN search, covariance and factor of a centre do not depend on the other centres, so each rank
  // prepares a contiguous block of `cap` centres and the factors / failure flags are all-gathered (in place, blocks padded to cap)
  const bool shard = c->auto_shard && c->nranks > 1 && c->nccl_comm != nullptr && n_kernels >= 8 * c->nranks;
  const int G = shard ? c->nranks : 1, cap = (n_kernels + G - 1) / G;
  const int cbeg = shard ? std::min(n_kernels, c->rank * cap) : 0, ncl = shard ? std::max(0, std::min(n_kernels, cbeg + cap) - cbeg) : n_kernels;
  const size_t n_pad = (size_t) G * cap;
  if (!c->sample.reserve((size_t) n_obs * d * sizeof(double)) || !c->zc.reserve((size_t) n_obs * d * sizeof(double)) ||
      !c->Ufull.reserve(n_pad * d * d * sizeof(double)) || !c->lnu.reserve((size_t) (n_kernels + 8) * sizeof(double)) ||
      !c->weights.reserve((size_t) (n_kernels + 8) * sizeof(double)) || !c->nn_idx.reserve(((size_t) (ncl > 0 ? ncl : 1) * k + n_pad + 16) * sizeof(int)))
    return c->fail(NCM_SD_GPU_ENOMEM, "vkde_prepare: out of device memory");
  {
    StageTimer t(c, NCM_SD_GPU_T_H2D);
    NCM_CUDA_OK(c, ncm_memcpy2d_async(c, c->sample.p, d * sizeof(double), sample, ld * sizeof(double), d * sizeof(double), n_obs, cudaMemcpyHostToDevice, c->stream));
    NCM_CUDA_OK(c, ncm_memcpy2d_async(c, c->zc.p, d * sizeof(double), invUsample, ldz * sizeof(double), d * sizeof(double), n_obs, cudaMemcpyHostToDevice, c->stream));
  }
  int *dNbr  = c->nn_idx.as<int>();
  int *dFail = dNbr + (size_t) (ncl > 0 ? ncl : 1) * k;   // [n_pad]
  int rc;
  if (ncl > 0) {
    StageTimer t(c, NCM_SD_GPU_T_PREP);
    rc = vkde_prepare_dev(c, n_obs, ncl, k, c->zc.as<double>(), c->sample.as<double>(), dNbr, c->Ufull.as<double>() + (size_t) cbeg * d * d, dFail + cbeg, cbeg);
    if (rc != NCM_SD_GPU_OK) return rc;
  }
  if (shard) {
    StageTimer t(c, NCM_SD_GPU_T_COMM);
    NcclApi &api = nccl_api();
    ncclResult_t r = api.AllGather(c->Ufull.as<double>() + (size_t) c->rank * cap * d * d, c->Ufull.p, (size_t) cap * d * d, ncclDouble, (ncclComm_t) c->nccl_comm, c->stream);
    if (r == ncclSuccess) r = api.AllGather(dFail + (size_t) c->rank * cap, dFail, (size_t) cap, ncclInt32, (ncclComm_t) c->nccl_comm, c->stream);
    if (r != ncclSuccess) return c->fail(NCM_SD_GPU_ENCCL, std::string("ncclAllGather: ") + api.GetErrorString(r));
  }
  {
    StageTimer t(c, NCM_SD_GPU_T_D2H);
    NCM_CUDA_OK(c, ncm_memcpy_async(c, U_all_out, c->Ufull.p, (size_t) n_kernels * d * d * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    NCM_CUDA_OK(c, ncm_memcpy_async(c, fail_out, dFail, (size_t) n_kernels * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    NCM_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  }
  c->type      = NCM_SD_GPU_VKDE;
  c->n_obs     = n_obs;
  c->n_kernels = n_kernels;
  c->row0      = 0;
  c->nrows     = n_obs;
  c->have_weights = false;
  c->prep_pending = true;
  return NCM_SD_GPU_OK;
}

int ncm_sd_gpu_vkde_finish(ncm_sd_gpu_ctx *c, const double *lnnorms, int n_fixed, const int *fixed_idx, const double *fixed_U) {
  if (c == nullptr) return NCM_SD_GPU_EINVAL;
  if (!c->prep_pending || c->type != NCM_SD_GPU_VKDE) return c->fail(NCM_SD_GPU_EINVAL, "vkde_finish: call vkde_prepare first");
  if (lnnorms == nullptr || n_fixed < 0 || (n_fixed > 0 && (fixed_idx == nullptr || fixed_U == nullptr))) return c->fail(NCM_SD_GPU_EINVAL, "vkde_finish: bad arguments");
  cudaSetDevice(c->device);
  const int d = c->d;
  {
    StageTimer t(c, NCM_SD_GPU_T_H2D);
    NCM_CUDA_OK(c, ncm_memcpy_async(c, c->lnu.p, lnnorms, (size_t) c->n_kernels * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    for (int f = 0; f < n_fixed; ++f) {
      if (fixed_idx[f] < 0 || fixed_idx[f] >= c->n_kernels) return c->fail(NCM_SD_GPU_EINVAL, "vkde_finish: fixed index out of range");
      NCM_CUDA_OK(c, ncm_memcpy_async(c, c->Ufull.as<double>() + (size_t) fixed_idx[f] * d * d, fixed_U + (size_t) f * d * d, (size_t) d * d * sizeof(double),
                                      cudaMemcpyHostToDevice, c->stream));
    }
  }
  int rc = vkde_pack(c, c->Ufull.as<double>());
  if (rc != NCM_SD_GPU_OK) return rc;
  NCM_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  c->prep_pending = false;
  return NCM_SD_GPU_OK;
}

int ncm_sd_gpu_set_weights(ncm_sd_gpu_ctx *c, int n_kernels, const double *weights, double href) {
  int rc = check_ready(c, false);
  if (rc != NCM_SD_GPU_OK) return rc;
  if (n_kernels != c->n_kernels || weights == nullptr || !(href > 0.0)) return c->fail(NCM_SD_GPU_EINVAL, "set_weights: bad arguments");
  cudaSetDevice(c->device);
  {
    StageTimer t(c, NCM_SD_GPU_T_H2D);
    NCM_CUDA_OK(c, ncm_memcpy_async(c,c->weights.p, weights, (size_t) n_kernels * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  }
  c->href         = href;
  c->have_weights = true;
  rc = update_cterm(c);
  if (rc != NCM_SD_GPU_OK) return rc;
  NCM_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return NCM_SD_GPU_OK;
}

int ncm_sd_gpu_set_href(ncm_sd_gpu_ctx *c, double href) {
  int rc = check_ready(c, false);
  if (rc != NCM_SD_GPU_OK) return rc;
  if (!(href > 0.0)) return c->fail(NCM_SD_GPU_EINVAL, "href must be positive");
  c->href = href;
  if (c->have_weights) return update_cterm(c);
  return NCM_SD_GPU_OK;
}

int ncm_sd_gpu_get_weights(ncm_sd_gpu_ctx *c, int n_kernels, double *weights) {
  int rc = check_ready(c, true);
  if (rc != NCM_SD_GPU_OK) return rc;
  if (n_kernels != c->n_kernels || weights == nullptr) return c->fail(NCM_SD_GPU_EINVAL, "get_weights: bad arguments");
  NCM_CUDA_OK(c, ncm_memcpy_async(c,weights, c->weights.p, (size_t) n_kernels * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  NCM_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return NCM_SD_GPU_OK;
}

int ncm_sd_gpu_eval_m2lnp_dev(ncm_sd_gpu_ctx *c, int q, const double *dX, int ldx, double *dOut) {
  int rc = check_ready(c, true);
  if (rc != NCM_SD_GPU_OK) return rc;
  if (q <= 0) return NCM_SD_GPU_OK;
  cudaSetDevice(c->device);
  StageTimer t(c, NCM_SD_GPU_T_EVAL);
  return c->type == NCM_SD_GPU_VKDE ? vkde_eval_launch(c, q, dX, ldx, dOut, false) : kde_eval_launch(c, q, dX, ldx, dOut, false);
}

static int eval_host(ncm_sd_gpu_ctx *c, int q, const double *X, int ldx, double *out, bool as_density) {
  int rc = check_ready(c, true);
  if (rc != NCM_SD_GPU_OK) return rc;
  if (q < 0 || (q > 0 && (X == nullptr || out == nullptr)) || ldx < c->d) return c->fail(NCM_SD_GPU_EINVAL, "eval: bad arguments");
  if (q == 0) return NCM_SD_GPU_OK;
  cudaSetDevice(c->device);
  const int d = c->d;
  // auto-shard: this rank evaluates the query rows [q0, q1) and the blocks are all-gathered (padded to the largest block)
  const bool shard = c->auto_shard && c->nranks > 1 && q >= c->nranks;
  const int G = shard ? c->nranks : 1, g = shard ? c->rank : 0;
  const int q0 = (int) (((long long) q * g) / G), q1 = (int) (((long long) q * (g + 1)) / G), ql = q1 - q0;
  const int cap = (q + G - 1) / G;
  if (!c->qX.reserve((size_t) (ql > 0 ? ql : 1) * d * sizeof(double)) || !c->qOut.reserve((size_t) (cap + 8) * sizeof(double)) ||
      (shard && !c->gath.reserve((size_t) G * cap * sizeof(double))))
    return c->fail(NCM_SD_GPU_ENOMEM, "eval: out of device memory");
  if (ql > 0) {
    StageTimer t(c, NCM_SD_GPU_T_H2D);
    NCM_CUDA_OK(c, ncm_memcpy2d_async(c,c->qX.p, d * sizeof(double), X + (size_t) q0 * ldx, ldx * sizeof(double), d * sizeof(double), ql, cudaMemcpyHostToDevice, c->stream));
  }
  if (ql > 0) {
    StageTimer t(c, NCM_SD_GPU_T_EVAL);
    rc = c->type == NCM_SD_GPU_VKDE ? vkde_eval_launch(c, ql, c->qX.as<double>(), d, c->qOut.as<double>(), as_density)
                                    : kde_eval_launch(c, ql, c->qX.as<double>(), d, c->qOut.as<double>(), as_density);
    if (rc != NCM_SD_GPU_OK) return rc;
  }
  if (shard) {
    StageTimer t(c, NCM_SD_GPU_T_COMM);
    NcclApi &api   = nccl_api();
    ncclResult_t r = api.AllGather(c->qOut.p, c->gath.p, (size_t) cap, ncclDouble, (ncclComm_t) c->nccl_comm, c->stream);
    if (r != ncclSuccess) return c->fail(NCM_SD_GPU_ENCCL, std::string("ncclAllGather: ") + api.GetErrorString(r));
  }
  {
    StageTimer t(c, NCM_SD_GPU_T_D2H);
    if (shard) {
      for (int r = 0; r < G; ++r) {
        const int a = (int) (((long long) q * r) / G), b = (int) (((long long) q * (r + 1)) / G);
        if (b > a)
          NCM_CUDA_OK(c, ncm_memcpy_async(c, out + a, c->gath.as<double>() + (size_t) r * cap, (size_t) (b - a) * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
      }
    } else {
      NCM_CUDA_OK(c, ncm_memcpy_async(c,out, c->qOut.p, (size_t) q * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    }
    NCM_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  }
  return NCM_SD_GPU_OK;
}

int ncm_sd_gpu_eval_m2lnp(ncm_sd_gpu_ctx *c, int q, const double *X, int ldx, double *out) { return eval_host(c, q, X, ldx, out, false); }
int ncm_sd_gpu_eval(ncm_sd_gpu_ctx *c, int q, const double *X, int ldx, double *out) { return eval_host(c, q, X, ldx, out, true); }

int ncm_sd_gpu_set_row_shard(ncm_sd_gpu_ctx *c, int row0, int nrows) {
  int rc = check_ready(c, false);
  if (rc != NCM_SD_GPU_OK) return rc;
  if (row0 < 0 || nrows < 0 || row0 + nrows > c->n_obs) return c->fail(NCM_SD_GPU_EINVAL, "set_row_shard: rows out of range");
  c->row0  = row0;
  c->nrows = nrows;
  return NCM_SD_GPU_OK;
}

int ncm_sd_gpu_set_auto_shard(ncm_sd_gpu_ctx *c, int on) {
  if (c == nullptr) return NCM_SD_GPU_EINVAL;
  if (on && c->nccl_comm == nullptr) return c->fail(NCM_SD_GPU_EINVAL, "set_auto_shard: ncm_sd_gpu_comm_init first");
  c->auto_shard = on != 0;
  c->im_sharded = false;
  return NCM_SD_GPU_OK;
}

int ncm_sd_gpu_allgather_dev(ncm_sd_gpu_ctx *c, const double *dsend, double *drecv, int count) {
  if (c == nullptr) return NCM_SD_GPU_EINVAL;
  if (c->nccl_comm == nullptr || dsend == nullptr || drecv == nullptr || count < 0) return c->fail(NCM_SD_GPU_EINVAL, "allgather: communicator and buffers required");
  cudaSetDevice(c->device);
  StageTimer t(c, NCM_SD_GPU_T_COMM);
  NcclApi &api   = nccl_api();
  ncclResult_t r = api.AllGather(dsend, drecv, (size_t) count, ncclDouble, (ncclComm_t) c->nccl_comm, c->stream);
  if (r != ncclSuccess) return c->fail(NCM_SD_GPU_ENCCL, std::string("ncclAllGather: ") + api.GetErrorString(r));
  return NCM_SD_GPU_OK;
}

int ncm_sd_gpu_compute_IM(ncm_sd_gpu_ctx *c, const double *row_scale, double *IM_host) {
  int rc = check_ready(c, false);
  if (rc != NCM_SD_GPU_OK) return rc;
  cudaSetDevice(c->device);
  if (c->auto_shard && c->nranks > 1) {
    // whoever wants the matrix on the host (cross-validation modes) gets all rows on every rank; otherwise this rank's row block
    c->im_sharded = IM_host == nullptr;
    c->row0       = c->im_sharded ? (int) (((long long) c->n_obs * c->rank) / c->nranks) : 0;
    c->nrows      = c->im_sharded ? (int) (((long long) c->n_obs * (c->rank + 1)) / c->nranks) - c->row0 : c->n_obs;
  }
  const int ldim = (c->n_kernels + 7) & ~7;
  if (!c->IM.reserve((size_t) (c->nrows > 0 ? c->nrows : 1) * ldim * sizeof(double)) || !c->rowscale.reserve((size_t) (c->n_obs + 8) * sizeof(double)))
    return c->fail(NCM_SD_GPU_ENOMEM, "compute_IM: out of device memory");
  if (row_scale != nullptr) {
    StageTimer t(c, NCM_SD_GPU_T_H2D);
    NCM_CUDA_OK(c, ncm_memcpy_async(c,c->rowscale.p, row_scale, (size_t) c->n_obs * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  }
  if (c->nrows > 0) {
    StageTimer t(c, NCM_SD_GPU_T_IM);
    const double *drs = row_scale != nullptr ? c->rowscale.as<double>() : nullptr;
    rc = c->type == NCM_SD_GPU_VKDE ? vkde_im_launch(c, drs) : kde_im_launch(c, drs);
    if (rc != NCM_SD_GPU_OK) return rc;
  }
  if (IM_host != nullptr && c->nrows > 0) {
    StageTimer t(c, NCM_SD_GPU_T_D2H);
    NCM_CUDA_OK(c, ncm_memcpy2d_async(c,IM_host, (size_t) c->n_kernels * sizeof(double), c->IM.p, (size_t) ldim * sizeof(double),
                                     (size_t) c->n_kernels * sizeof(double), c->nrows, cudaMemcpyDeviceToHost, c->stream));
  }
  NCM_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return NCM_SD_GPU_OK;
}

int ncm_sd_gpu_nnls_solve(ncm_sd_gpu_ctx *c, double reltol, double *x_out, double *rnorm_out, ncm_sd_gpu_nnls_stats *stats) {
  int rc = check_ready(c, false);
  if (rc != NCM_SD_GPU_OK) return rc;
  if (c->IM.p == nullptr || x_out == nullptr || rnorm_out == nullptr) return c->fail(NCM_SD_GPU_EINVAL, "nnls_solve: compute_IM first");
  cudaSetDevice(c->device);
  const int ldim = (c->n_kernels + 7) & ~7;
  rc = nnls_solve_dev(c, c->nrows, c->n_kernels, c->IM.as<double>(), ldim, nullptr, reltol, x_out, rnorm_out, stats);
  if (rc != NCM_SD_GPU_OK) return rc;
  NCM_CUDA_OK(c, ncm_memcpy_async(c,c->weights.p, x_out, (size_t) c->n_kernels * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  NCM_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return NCM_SD_GPU_OK;
}

int ncm_sd_gpu_nnls_solve_host(ncm_sd_gpu_ctx *c, int nrows, int ncols, const double *A, int lda, const double *f, double reltol,
                               double *x_out, double *rnorm_out, ncm_sd_gpu_nnls_stats *stats) {
  if (c == nullptr) return NCM_SD_GPU_EINVAL;
  if (nrows <= 0 || ncols <= 0 || A == nullptr || f == nullptr || x_out == nullptr || rnorm_out == nullptr || lda < ncols)
    return c->fail(NCM_SD_GPU_EINVAL, "nnls_solve_host: bad arguments");
  cudaSetDevice(c->device);
  const int ldim = (ncols + 7) & ~7;
  if (!c->IM.reserve((size_t) nrows * ldim * sizeof(double)) || !c->nn_f.reserve((size_t) (nrows + 8) * sizeof(double)))
    return c->fail(NCM_SD_GPU_ENOMEM, "nnls_solve_host: out of device memory");
  NCM_CUDA_OK(c, cudaMemsetAsync(c->IM.p, 0, (size_t) nrows * ldim * sizeof(double), c->stream));
  NCM_CUDA_OK(c, ncm_memcpy2d_async(c,c->IM.p, (size_t) ldim * sizeof(double), A, (size_t) lda * sizeof(double), (size_t) ncols * sizeof(double), nrows,
                                   cudaMemcpyHostToDevice, c->stream));
  NCM_CUDA_OK(c, ncm_memcpy_async(c,c->nn_f.p, f, (size_t) nrows * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  return nnls_solve_dev(c, nrows, ncols, c->IM.as<double>(), ldim, c->nn_f.as<double>(), reltol, x_out, rnorm_out, stats);
}

int ncm_sd_gpu_sample_apply(ncm_sd_gpu_ctx *c, int q, const int *kidx, const double *Z, int ldz, const double *scale, double *X_out, int ldx) {
  int rc = check_ready(c, true);
  if (rc != NCM_SD_GPU_OK) return rc;
  if (q < 0 || (q > 0 && (kidx == nullptr || Z == nullptr || X_out == nullptr)) || ldz < c->d || ldx < c->d)
    return c->fail(NCM_SD_GPU_EINVAL, "sample_apply: bad arguments");
  if (q == 0) return NCM_SD_GPU_OK;
  if (c->type == NCM_SD_GPU_KDE) return c->fail(NCM_SD_GPU_EINVAL, "sample_apply: KDE centres are uploaded whitened; use the VKDE upload");
  cudaSetDevice(c->device);
  const int d = c->d;
  const size_t bz = (size_t) q * d * sizeof(double);
  if (!c->qX.reserve(2 * bz + (size_t) q * sizeof(double)) || !c->nn_idx.reserve((size_t) (q + 8) * sizeof(int)))
    return c->fail(NCM_SD_GPU_ENOMEM, "sample_apply: out of device memory");
  double *dZ = c->qX.as<double>(), *dXo = dZ + (size_t) q * d, *dS = dXo + (size_t) q * d;
  NCM_CUDA_OK(c, ncm_memcpy2d_async(c,dZ, d * sizeof(double), Z, ldz * sizeof(double), d * sizeof(double), q, cudaMemcpyHostToDevice, c->stream));
  NCM_CUDA_OK(c, ncm_memcpy_async(c,c->nn_idx.p, kidx, (size_t) q * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  if (scale != nullptr) NCM_CUDA_OK(c, ncm_memcpy_async(c,dS, scale, (size_t) q * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  sample_apply_kernel<<<(q + 127) / 128, 128, 0, c->stream>>>(c->sample.as<double>(), d, c->Ufull.as<double>(), 1, d, c->nn_idx.as<int>(), dZ, d,
                                                            scale != nullptr ? dS : nullptr, c->href, q, dXo, d);
  c->n_launches++;
  NCM_CUDA_OK(c, cudaGetLastError());
  NCM_CUDA_OK(c, ncm_memcpy2d_async(c,X_out, ldx * sizeof(double), dXo, d * sizeof(double), d * sizeof(double), q, cudaMemcpyDeviceToHost, c->stream));
  NCM_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return NCM_SD_GPU_OK;
}

int ncm_sd_gpu_sample_philox(ncm_sd_gpu_ctx *c, int q, unsigned long long seed, unsigned long long offset, double *X_out, int ldx, int *kidx_out) {
  int rc = check_ready(c, true);
  if (rc != NCM_SD_GPU_OK) return rc;
  if (c->type != NCM_SD_GPU_VKDE) return c->fail(NCM_SD_GPU_EINVAL, "sample_philox: VKDE upload required");
  if (q <= 0 || X_out == nullptr || ldx < c->d) return c->fail(NCM_SD_GPU_EINVAL, "sample_philox: bad arguments");
  cudaSetDevice(c->device);
  const int d = c->d;
  if (!c->qX.reserve((size_t) q * d * sizeof(double)) || !c->nn_idx.reserve((size_t) (q + 8) * sizeof(int)))
    return c->fail(NCM_SD_GPU_ENOMEM, "sample_philox: out of device memory");
  rc = sample_philox_launch(c, q, seed, offset, c->qX.as<double>(), d, c->nn_idx.as<int>());
  if (rc != NCM_SD_GPU_OK) return rc;
  NCM_CUDA_OK(c, ncm_memcpy2d_async(c,X_out, ldx * sizeof(double), c->qX.p, d * sizeof(double), d * sizeof(double), q, cudaMemcpyDeviceToHost, c->stream));
  if (kidx_out != nullptr) NCM_CUDA_OK(c, ncm_memcpy_async(c,kidx_out, c->nn_idx.p, (size_t) q * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  NCM_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return NCM_SD_GPU_OK;
}

int ncm_sd_gpu_comm_unique_id(char id_out[128]) {
  NcclApi &api = nccl_api();
  if (!api.ok) return NCM_SD_GPU_ENCCL;
  ncclUniqueId id;
  if (api.GetUniqueId(&id) != ncclSuccess) return NCM_SD_GPU_ENCCL;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  std::memcpy(id_out, &id, 128);
  return NCM_SD_GPU_OK;
}

int ncm_sd_gpu_comm_init(ncm_sd_gpu_ctx *c, int nranks, int rank, const char id_in[128]) {
  if (c == nullptr) return NCM_SD_GPU_EINVAL;
  NcclApi &api = nccl_api();
  if (!api.ok) return c->fail(NCM_SD_GPU_ENCCL, "libnccl.so.2 not loadable");
  if (nranks < 1 || rank < 0 || rank >= nranks) return c->fail(NCM_SD_GPU_EINVAL, "comm_init: bad rank");
  cudaSetDevice(c->device);
  ncclUniqueId id;
  std::memcpy(&id, id_in, 128);
  ncclComm_t comm;
  ncclResult_t r = api.CommInitRank(&comm, nranks, id, rank);
  if (r != ncclSuccess) return c->fail(NCM_SD_GPU_ENCCL, std::string("ncclCommInitRank: ") + api.GetErrorString(r));
  c->nccl_comm = (void *) comm;
  c->nranks    = nranks;
  c->rank      = rank;
  return NCM_SD_GPU_OK;
}

int ncm_sd_gpu_get_timers(ncm_sd_gpu_ctx *c, double ms[NCM_SD_GPU_T_LEN], long long *n_launches) {
  if (c == nullptr) return NCM_SD_GPU_EINVAL;
  if (ms != nullptr) std::memcpy(ms, c->t_ms, sizeof(c->t_ms));
  if (n_launches != nullptr) *n_launches = c->n_launches;
  return NCM_SD_GPU_OK;
}
int ncm_sd_gpu_reset_timers(ncm_sd_gpu_ctx *c) {
  if (c == nullptr) return NCM_SD_GPU_EINVAL;
  std::memset(c->t_ms, 0, sizeof(c->t_ms));
  c->n_launches = 0;
  c->h2d_bytes = c->d2h_bytes = 0;
  return NCM_SD_GPU_OK;
}
int ncm_sd_gpu_get_traffic(ncm_sd_gpu_ctx *c, long long *h2d_bytes, long long *d2h_bytes) {
  if (c == nullptr) return NCM_SD_GPU_EINVAL;
  if (h2d_bytes != nullptr) *h2d_bytes = c->h2d_bytes;
  if (d2h_bytes != nullptr) *d2h_bytes = c->d2h_bytes;
  return NCM_SD_GPU_OK;
}
int ncm_sd_gpu_enable_timers(ncm_sd_gpu_ctx *c, int enable) {
  if (c == nullptr) return NCM_SD_GPU_EINVAL;
  c->timers_on = enable != 0;
  return NCM_SD_GPU_OK;
}

int ncm_sd_gpu_dsyrk_ata_dev(ncm_sd_gpu_ctx *c, int nrows, int ncols, const double *dA, int lda, double *dM, int ldm) {
  if (c == nullptr) return NCM_SD_GPU_EINVAL;
  cudaSetDevice(c->device);
  return dsyrk_ata_general(c, nrows, ncols, dA, lda, dM, ldm, 1.0, 0.0);
}

int ncm_sd_gpu_dpotrf_upper_dev(ncm_sd_gpu_ctx *c, int n, double *dM, int ldm, int *info_host) {
  if (c == nullptr) return NCM_SD_GPU_EINVAL;
  cudaSetDevice(c->device);
  if (!c->nn_b.reserve((size_t) (n + 64) * sizeof(double)) || !c->nn_idx.reserve(64)) return c->fail(NCM_SD_GPU_ENOMEM, "dpotrf: out of device memory");
  return dpotrf_upper_solve_any(c, n, dM, ldm, nullptr, c->nn_b.as<double>(), c->nn_idx.as<int>(), info_host);
}

int ncm_sd_gpu_host_alloc(void **ptr, size_t bytes) {
  if (ptr == nullptr) return NCM_SD_GPU_EINVAL;
  *ptr = nullptr;
  if (cudaHostAlloc(ptr, bytes, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    *ptr = nullptr;
    return NCM_SD_GPU_ENOMEM;
  }
  return NCM_SD_GPU_OK;
}
int ncm_sd_gpu_host_free(void *ptr) {
  if (ptr != nullptr && cudaFreeHost(ptr) != cudaSuccess) {
    cudaGetLastError();
    return NCM_SD_GPU_ECUDA;
  }
  return NCM_SD_GPU_OK;
}

int ncm_sd_gpu_vkde_path(ncm_sd_gpu_ctx *c, int *uses_mma, double *cond_max) {
  if (c == nullptr) return NCM_SD_GPU_EINVAL;
  if (uses_mma != nullptr) *uses_mma = (c->type == NCM_SD_GPU_VKDE && c->vkde_mma) ? 1 : 0;
  if (cond_max != nullptr) *cond_max = c->vkde_cond;
  return NCM_SD_GPU_OK;
}

// debugging aid (tools/chol_trace.py): per-CTA event trace of the next single-launch Cholesky calls; dTrace = device buffer of
// n_sm * cap * 2 int64 (zero it before each call), or null to switch tracing off.  Not part of the public header.
int ncm_sd_gpu_chol_trace(ncm_sd_gpu_ctx *c, long long *dTrace, int cap) {
  if (c == nullptr) return NCM_SD_GPU_EINVAL;
  c->chol_trace     = dTrace;
  c->chol_trace_cap = cap;
  return NCM_SD_GPU_OK;
}

int ncm_sd_gpu_dposv_upper_dev(ncm_sd_gpu_ctx *c, int n, double *dM, int ldm, double *dRhs, int *info_host) {
  if (c == nullptr) return NCM_SD_GPU_EINVAL;
  if (dRhs == nullptr) return c->fail(NCM_SD_GPU_EINVAL, "dposv: rhs required");
  cudaSetDevice(c->device);
  if (!c->nn_b.reserve((size_t) (n + 64) * sizeof(double)) || !c->nn_idx.reserve(64)) return c->fail(NCM_SD_GPU_ENOMEM, "dposv: out of device memory");
  // with a communicator and n >= NCM_SD_GPU_DIST_CHOL_MIN_N (8192) every rank must make this call on identical data: the trailing
  // updates are then distributed over the ranks (dist_chol.cu)
  if (c->nccl_comm != nullptr && c->nranks > 1 && n >= dist_chol_min_n()) return dpotrf_upper_solve_dist(c, n, dM, ldm, dRhs, info_host);
  return dpotrf_upper_solve_any(c, n, dM, ldm, dRhs, c->nn_b.as<double>(), c->nn_idx.as<int>(), info_host);
}

int ncm_sd_gpu_dsysv_upper_dev(ncm_sd_gpu_ctx *c, int n, double *dM, int ldm, double *dRhs, int *info_host) {
  if (c == nullptr) return NCM_SD_GPU_EINVAL;
  if (n <= 0 || dM == nullptr || dRhs == nullptr || ldm < n) return c->fail(NCM_SD_GPU_EINVAL, "dsysv: bad arguments");
  cudaSetDevice(c->device);
  return dsysv_upper_solve(c, n, dM, ldm, dRhs, info_host);
}

int ncm_sd_gpu_dgels_cols_dev(ncm_sd_gpu_ctx *c, int m, int n, const double *dA, int lda, const int *dIdx, const double *dF, double *dX, int *info_host) {
  if (c == nullptr) return NCM_SD_GPU_EINVAL;
  if (m <= 0 || n <= 0 || dA == nullptr || dIdx == nullptr || dF == nullptr || dX == nullptr) return c->fail(NCM_SD_GPU_EINVAL, "dgels: bad arguments");
  cudaSetDevice(c->device);
  return dgels_cols_solve(c, m, n, dA, lda, dIdx, dF, dX, info_host);
}

int ncm_sd_gpu_dtrtri_upper_dev(ncm_sd_gpu_ctx *c, int n, const double *dU, int ld, double *dW, double *dScratch) {
  if (c == nullptr) return NCM_SD_GPU_EINVAL;
  if (n <= 0 || (ld & 7) || ld < n) return c->fail(NCM_SD_GPU_EINVAL, "dtrtri: ld must be a multiple of 8 and >= n");
  cudaSetDevice(c->device);
  return trinv_upper(c, n, dU, dW, dScratch, ld, nullptr);
}

}   // extern "C"
