// NNLS driver: block-pivoting active-set method on the normal equations, device linear algebra,
// host index-set decisions.
//
// Replaces ncm_nnls_solve with NCM_NNLS_UMETHOD_NORMAL (ncm_nnls.c:767-871):
//   M = A^T A (upper), b = A^T f                         ncm_nnls.c:791-792
//   _ncm_nnls_solve_feasible                             ncm_nnls.c:728-751
//   residuals / mgrad                                    ncm_nnls.c:710-726
//   outer loop with add_frac halving                     ncm_nnls.c:825-868
// and the NcmISet operations it relies on (ncm_iset.c:853-1077): the passive set is always
// handled in ascending index order, "invalid" means x_i < 1e-300, the most negative
// max_remove entries are evicted, and the largest add_frac share of the positive gradient
// entries is admitted.  Which systems get solved decides the final passive set, so this
// logic mirrors the reference decision by decision; only the arithmetic runs on the device.
//
// Which systems get solved is the reference's; HOW a system is solved is not always a fresh dposv: once a passive set B has been
// factorised, the following sets that differ from it by at most min(lowrank_kmax(), |B| / 8) indices are solved by low-rank
// modification of that factor (lowrank.cu: triangular inverse by recursive doubling, bordered / constrained k x k system, one
// refinement step on the true residual).  A refinement correction above 1e-7 of the solution falls back to a fresh
// factorisation.  NCM_SD_GPU_NNLS_REUSE=0 restores one dposv per system.
//
// When dposv reports a non-positive pivot the reference's fallback chain is followed (ncm_nnls.c:573-638, 655-666): the system is
// solved by the symmetric-indefinite L D L^T (ldl_bk.cu, dsysv) and, if that meets an exactly singular pivot, by Householder least
// squares on the passive columns (qr_ls.cu, dgels); counted in stats->n_lu / n_qr.
#include <algorithm>
#include <cstring>
#include <dlfcn.h>
#include <vector>
#include "ctx.h"
#include "nccl_shim.h"

int dsyrk_ata_general(ncm_sd_gpu_ctx *c, int K, int n, const double *dP, int ldp, double *dC, int ldc, double alpha, double beta);
int dpotrf_upper_solve(ncm_sd_gpu_ctx *c, int n, double *dM, int ldm, double *dRhs, double *dDinv, int *dInfo, int *info_host);
int gemv_t(ncm_sd_gpu_ctx *c, const double *dA, int lda, int nrows, int ncols, const double *dv, double *dOut, DevBuf &tmp);
int residual(ncm_sd_gpu_ctx *c, const double *dA, int lda, int nrows, int ncols, const double *dx, const double *df, double *dr,
             double *d_ss_part, int *nblocks_out);

// ---- NCCL through dlopen (nccl_shim.h) ----------------------------------------------------------------
NcclApi &nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried   = true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (h != nullptr) {
      api.GetUniqueId    = (decltype(api.GetUniqueId)) dlsym(h, "ncclGetUniqueId");
      api.CommInitRank   = (decltype(api.CommInitRank)) dlsym(h, "ncclCommInitRank");
      api.AllReduce      = (decltype(api.AllReduce)) dlsym(h, "ncclAllReduce");
      api.AllGather      = (decltype(api.AllGather)) dlsym(h, "ncclAllGather");
      api.Broadcast      = (decltype(api.Broadcast)) dlsym(h, "ncclBroadcast");
      api.CommDestroy    = (decltype(api.CommDestroy)) dlsym(h, "ncclCommDestroy");
      api.GetErrorString = (decltype(api.GetErrorString)) dlsym(h, "ncclGetErrorString");
      api.ok = api.GetUniqueId && api.CommInitRank && api.AllReduce && api.AllGather && api.Broadcast && api.CommDestroy && api.GetErrorString;
    }
  }
  return api;
}

static int allreduce_sum(ncm_sd_gpu_ctx *c, double *dbuf, size_t count) {
  // auto-shard mode with an unsharded IM (someone asked for the whole matrix on the host): every rank holds all rows, nothing to sum
  if (c->nranks <= 1 || (c->auto_shard && !c->im_sharded)) return NCM_SD_GPU_OK;
  StageTimer t(c, NCM_SD_GPU_T_COMM);
  NcclApi &api   = nccl_api();
  ncclResult_t r = api.AllReduce(dbuf, dbuf, count, ncclDouble, ncclSum, (ncclComm_t) c->nccl_comm, c->stream);
  if (r != ncclSuccess) return c->fail(NCM_SD_GPU_ENCCL, std::string("ncclAllReduce: ") + api.GetErrorString(r));
  return NCM_SD_GPU_OK;
}

namespace {

__global__ void gather_sym_kernel(const double *__restrict__ M, int ldm, const int *__restrict__ idx, int np, double *__restrict__ S, int lds,
                                  const double *__restrict__ b, double *__restrict__ rhs, double shift) {
  // S[i][j] = M[idx[i]][idx[j]] for j >= i (idx ascending => source is in the upper triangle)
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (i >= np || j >= np) return;
  if (j >= i) {
    double v = M[(size_t) idx[i] * ldm + idx[j]];
    if (i == j) v += shift;
    S[(size_t) i * lds + j] = v;
  }
  if (i == 0) rhs[j] = b[idx[j]];
}

__global__ void copy_upper_kernel(const double *__restrict__ M, int ldm, int n, double *__restrict__ S, int lds, const double *__restrict__ b,
                                  double *__restrict__ rhs, double shift) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (i >= n || j >= n) return;
  if (j >= i) {
    double v = M[(size_t) i * ldm + j];
    if (i == j) v += shift;
    S[(size_t) i * lds + j] = v;
  }
  if (i == 0) rhs[j] = b[j];
}

__global__ void sum_parts_kernel(const double *__restrict__ part, int n, double *__restrict__ out) {
  __shared__ double sh[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += part[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sh[0];
}

// GSL sort/subsetind_source.c semantics (first-come wins on ties), as NcmISet uses them
void sort_smallest_index(std::vector<int> &p, int k, const std::vector<double> &src, int n) {
  p.assign(k, 0);
  if (k == 0 || n == 0) return;
  int j         = 1;
  double xbound = src[0];
  p[0]          = 0;
  for (int i = 1; i < n; i++) {
    const double xi = src[i];
    if (j < k)
      j++;
    else if (xi >= xbound)
      continue;
    int i1;
    for (i1 = j - 1; i1 > 0; i1--) {
      if (xi > src[p[i1 - 1]]) break;
      p[i1] = p[i1 - 1];
    }
    p[i1]  = i;
    xbound = src[p[j - 1]];
  }
}

void sort_largest_index(std::vector<int> &p, int k, const std::vector<double> &src, int n) {
  p.assign(k, 0);
  if (k == 0 || n == 0) return;
  int j         = 1;
  double xbound = src[0];
  p[0]          = 0;
  for (int i = 1; i < n; i++) {
    const double xi = src[i];
    if (j < k)
      j++;
    else if (xi <= xbound)
      continue;
    int i1;
    for (i1 = j - 1; i1 > 0; i1--) {
      if (xi < src[p[i1 - 1]]) break;
      p[i1] = p[i1 - 1];
    }
    p[i1]  = i;
    xbound = src[p[j - 1]];
  }
}

struct NnlsWork {
  ncm_sd_gpu_ctx *c;
  int nrows, n, lda, ldm;
  const double *dA, *dF;
  double *dM, *dMU, *db, *drhs, *dx, *dr, *dr_try, *dg, *ddinv, *dss, *dscal;
  int *didx, *dinfo;
  double *h_buf;   // pinned: n doubles + 8
  double *h_g;     // pinned: n doubles (gradient computed along with a residual)
  int *h_idx;      // pinned: n ints
  ncm_sd_gpu_nnls_stats *st;
  // low-rank reuse of the last factorisation (lowrank.cu)
  bool lr_on = false, base_valid = false, w_valid = false, base_trusted = false;
  std::vector<int> baseP;
  std::vector<int> prevD, ordD;   // removed set (positions in the base, device order) whose k x k factor is resident; scratch
  std::vector<char> markB;
  LowrankBufs lb;
  int ldv = 0, nv = 0;
};

bool nnls_trace() {   // debugging aid: one stderr line per passive-set system (the oracle prints the same under ORC_NNLS_TRACE)
  static const bool on = getenv("NCM_SD_GPU_NNLS_TRACE") != nullptr;
  return on;
}
bool lowrank_enabled() {
  static const bool on = [] {
    const char *e = getenv("NCM_SD_GPU_NNLS_REUSE");
    return !(e != nullptr && e[0] == '0');
  }();
  return on;
}
bool base_solve_enabled() {   // NCM_SD_GPU_NNLS_BASE_SOLVE=0: the base system keeps its forward / back substitution
  static const bool on = [] {
    const char *e = getenv("NCM_SD_GPU_NNLS_BASE_SOLVE");
    return !(e != nullptr && e[0] == '0');
  }();
  return on;
}
constexpr int LR_MIN_N = 512;       // below this a fused factorisation (a few 64-column phases) is as cheap as the ~25 launches of an update
constexpr double LR_MAX_CORR = 1e-7;
// a base whose first low-rank solve needed a refinement correction below this is trusted: the later solves from it skip the refinement
// (its cost is about a third of a solve; the correction measures the conditioning of the base, which does not change between solves)
constexpr double LR_TRUST_CORR = 1e-12;

// P = (B \ D) u A: solve through the base inverse; returns 1 when the caller has to factorise afresh
int solve_lowrank(NnlsWork &w, const std::vector<int> &P, bool *done) {
  ncm_sd_gpu_ctx *c = w.c;
  *done             = false;
  const std::vector<int> &B = w.baseP;
  const int nB = (int) B.size(), np = (int) P.size();
  int kcap = std::min(lowrank_kmax(), nB / 8);
  // D = B \ P (positions in B), A = P \ B (global indices): one merge pass over the two ascending lists; bsel[i] = B[i] or -1 on the rows of D,
  // psrc[j] = position of P[j] in B or -(1 + position in A)
  const int kpad = lowrank_kmax() + 8;
  int *hA = w.h_idx, *hD = hA + kpad, *hSel = hD + kpad, *hSrc = hSel + w.nv;
  int na = 0, nd = 0;
  {
    int i = 0, j = 0;
    while (i < nB || j < np) {
      if (j >= np || (i < nB && B[i] < P[j])) {
        if (nd + na >= kcap) return NCM_SD_GPU_OK;
        hSel[i] = -1;
        hD[nd++] = i++;
      } else if (i >= nB || P[j] < B[i]) {
        if (nd + na >= kcap) return NCM_SD_GPU_OK;
        hSrc[j]  = -(1 + na);
        hA[na++] = P[j++];
      } else {
        hSel[i] = B[i];
        hSrc[j] = i;
        ++i;
        ++j;
      }
    }
  }
  // nested growth of the removed set (the rule inside ncm_nnls.c:728-751: every pass only removes): keep the previous members first, in
  // their previous order, and append the new ones -- the k x k factor of the previous solve is then the leading block of this one's
  int kold = 0;
  if (na == 0 && !w.prevD.empty() && (int) w.prevD.size() <= nd) {
    std::vector<char> &mark = w.markB;
    mark.assign(nB, 0);
    for (int q = 0; q < nd; ++q) mark[hD[q]] = 1;
    bool nested = true;
    for (int p : w.prevD)
      if (!mark[p]) {
        nested = false;
        break;
      }
    if (nested) {
      for (int p : w.prevD) mark[p] = 2;
      std::vector<int> &ord = w.ordD;
      ord.assign(w.prevD.begin(), w.prevD.end());
      for (int q = 0; q < nd; ++q)
        if (mark[hD[q]] == 1) ord.push_back(hD[q]);
      std::memcpy(hD, ord.data(), sizeof(int) * nd);
      kold = (int) w.prevD.size();
    }
  }
  StageTimer t(c, NCM_SD_GPU_T_LOWRANK);
  if (!w.w_valid) {
    int rc = trinv_upper(c, nB, w.dMU, w.lb.W, w.lb.S, w.ldm, w.lb.Wt);
    if (rc != NCM_SD_GPU_OK) return rc;
    w.w_valid = true;
    if (w.st) {
      w.st->n_trinv++;
      w.st->lowrank_flops += (double) nB * nB * nB / 3.0;
    }
  }
  NCM_CUDA_OK(c, ncm_memcpy_async(c, w.lb.idxA, hA, sizeof(int) * (size_t) (2 * kpad + 2 * w.nv), cudaMemcpyHostToDevice, c->stream));
  const bool refine = !w.base_trusted;
  int rc = lowrank_solve(c, w.dM, w.ldm, w.n, w.db, nB, na, nd, np, w.lb, w.ldv, refine, kold);
  if (rc != NCM_SD_GPU_OK) return rc;
  w.prevD.clear();   // valid again only once this solve's factor is known to be good (below)
  if (na + nd == 0) w.lb.tb_valid = true;   // the base's own solve left t_b = W^T b_B behind: pure removals now need no product with W
  NCM_CUDA_OK(c, ncm_memcpy_async(c, w.h_buf, w.lb.out, sizeof(double) * (size_t) (np + 2), cudaMemcpyDeviceToHost, c->stream));
  NCM_CUDA_OK(c, ncm_memcpy_async(c, w.h_buf + np + 2, w.lb.info, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  NCM_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  const double mdx = w.h_buf[np], mx = w.h_buf[np + 1];
  int info;
  std::memcpy(&info, w.h_buf + np + 2, sizeof(int));
  if (w.st) {
    if (na + nd > 0) w.st->n_lowrank++;   // the base set's own solve (k = 0) belongs to its factorisation, counted in n_chol
    w.st->lowrank_flops += 2.0 * nB * (double) nB * (na + nd + 1);
    w.st->max_lowrank_k = std::max(w.st->max_lowrank_k, na + nd);
  }
  if (nnls_trace()) fprintf(stderr, "gpu_nnls: lowrank |P| = %d |B| = %d k = %d + %d info = %d corr = %.3e%s\n", np, nB, na, nd, info, mdx / mx, refine ? "" : " (no refinement)");
  if (info != 0 || !(mdx <= LR_MAX_CORR * mx)) {   // also catches NaN
    if (w.st && na + nd > 0) w.st->n_lowrank_fallback++;
    return NCM_SD_GPU_OK;
  }
  if (refine && mdx <= LR_TRUST_CORR * mx) w.base_trusted = true;
  if (na == 0 && nd > 0) w.prevD.assign(hD, hD + nd);   // the factor left on the device belongs to this removed set, in this order
  if (w.st && kold > 0) w.st->n_lowrank_nested++;
  *done = true;
  return NCM_SD_GPU_OK;
}

// solve M[P,P] x_P = b[P]; result in h_buf[0..np)
int solve_unconstrained(NnlsWork &w, const std::vector<int> &P) {
  ncm_sd_gpu_ctx *c = w.c;
  const int np      = (int) P.size();
  if (np == 0) return NCM_SD_GPU_OK;
  if (w.lr_on && w.base_valid && np >= LR_MIN_N) {
    bool done = false;
    int rc    = solve_lowrank(w, P, &done);
    if (rc != NCM_SD_GPU_OK) return rc;
    if (done) return NCM_SD_GPU_OK;
  }
  w.base_valid = false;
  auto gather = [&]() -> int {
    StageTimer t(c, NCM_SD_GPU_T_NNLS_MISC);
    if (np == w.n) {
      dim3 grid((np + 255) / 256, np);
      copy_upper_kernel<<<grid, 256, 0, c->stream>>>(w.dM, w.ldm, np, w.dMU, w.ldm, w.db, w.drhs, 0.0);
    } else {
      std::memcpy(w.h_idx, P.data(), sizeof(int) * np);
      NCM_CUDA_OK(c, ncm_memcpy_async(c, w.didx, w.h_idx, sizeof(int) * np, cudaMemcpyHostToDevice, c->stream));
      dim3 grid((np + 255) / 256, np);
      gather_sym_kernel<<<grid, 256, 0, c->stream>>>(w.dM, w.ldm, w.didx, np, w.dMU, w.ldm, w.db, w.drhs, 0.0);
    }
    c->n_launches++;
    return NCM_SD_GPU_OK;
  };
  int rc = gather();
  if (rc != NCM_SD_GPU_OK) return rc;
  int info = 0;
  // a set that becomes the base of low-rank solves is only factorised here: its own solution comes from the triangular inverse the
  // following solves need anyway (x = W W^T b + one refinement step, lowrank.cu), not from a forward / back substitution
  const bool dist = c->nccl_comm != nullptr && c->nranks > 1 && np >= dist_chol_min_n();
  const bool as_base = w.lr_on && np >= LR_MIN_N && !dist;
  for (int pass = 0; pass < 2; ++pass) {
    const bool factor_only = as_base && pass == 0 && base_solve_enabled();
    {
      StageTimer t(c, NCM_SD_GPU_T_CHOL);
      // all ranks hold the same all-reduced matrix and take the same decisions: large systems are factorised together (dist_chol.cu)
      rc = dist ? dpotrf_upper_solve_dist(c, np, w.dMU, w.ldm, w.drhs, &info)
                : dpotrf_upper_solve_any(c, np, w.dMU, w.ldm, factor_only ? nullptr : w.drhs, w.ddinv, w.dinfo, &info);
      if (rc != NCM_SD_GPU_OK) return rc;
      if (dist && w.st) w.st->n_dist_chol++;
    }
    if (w.st) {
      w.st->n_chol++;
      w.st->chol_flops += (double) np * np * np / 3.0;
    }
    if (nnls_trace()) fprintf(stderr, "gpu_nnls: chol |P| = %d info = %d%s\n", np, info, factor_only ? " (factor only)" : "");
    if (info != 0) break;
    if (as_base) {   // the factor left in dMU becomes the base of the following low-rank solves
      std::memcpy(w.h_idx, P.data(), sizeof(int) * np);
      NCM_CUDA_OK(c, ncm_memcpy_async(c, w.lb.idxB, w.h_idx, sizeof(int) * np, cudaMemcpyHostToDevice, c->stream));
      w.baseP        = P;
      w.base_valid   = true;
      w.w_valid      = false;
      w.base_trusted = false;
      w.lb.tb_valid  = false;
      w.prevD.clear();
    }
    if (!factor_only) {
      NCM_CUDA_OK(c, ncm_memcpy_async(c, w.h_buf, w.drhs, sizeof(double) * np, cudaMemcpyDeviceToHost, c->stream));
      NCM_CUDA_OK(c, cudaStreamSynchronize(c->stream));
      return NCM_SD_GPU_OK;
    }
    bool done = false;
    rc        = solve_lowrank(w, P, &done);
    if (rc != NCM_SD_GPU_OK) return rc;
    if (done) return NCM_SD_GPU_OK;
    // the inverse of this base is not good enough for its own system: factorise again, this time with the substitutions
    w.base_valid = false;
    rc           = gather();
    if (rc != NCM_SD_GPU_OK) return rc;
  }
  if (info == 0) return c->fail(NCM_SD_GPU_ECUDA, "nnls: internal error (base solve)");
  // _ncm_nnls_solve_normal_LU, ncm_nnls.c:573-606: the system is gathered again (dposv destroyed it) and solved by dsysv
  rc = gather();
  if (rc != NCM_SD_GPU_OK) return rc;
  int info_lu = 0;
  {
    StageTimer t(c, NCM_SD_GPU_T_CHOL);
    rc = dsysv_upper_solve(c, np, w.dMU, w.ldm, w.drhs, &info_lu);
    if (rc != NCM_SD_GPU_OK) return rc;
  }
  if (w.st) w.st->n_lu++;
  if (nnls_trace()) fprintf(stderr, "gpu_nnls: lu |P| = %d info = %d\n", np, info_lu);
  if (info_lu > 0) {
    // _ncm_nnls_solve_normal_QR, ncm_nnls.c:608-638: least squares on the passive columns of A itself
    if (c->nranks > 1 && c->im_sharded)
      return c->fail(NCM_SD_GPU_ENOTPD, "nnls: the passive-set system is exactly singular and the dgels fallback needs the whole matrix on one rank");
    if (np == w.n) {
      for (int i = 0; i < np; ++i) w.h_idx[i] = i;
      NCM_CUDA_OK(c, ncm_memcpy_async(c, w.didx, w.h_idx, sizeof(int) * np, cudaMemcpyHostToDevice, c->stream));
    }
    int info_qr = 0;
    {
      StageTimer t(c, NCM_SD_GPU_T_CHOL);
      rc = dgels_cols_solve(c, w.nrows, np, w.dA, w.lda, w.didx, w.dF, w.drhs, &info_qr);
      if (rc != NCM_SD_GPU_OK) return rc;
    }
    if (w.st) w.st->n_qr++;
    if (nnls_trace()) fprintf(stderr, "gpu_nnls: qr |P| = %d info = %d\n", np, info_qr);
    if (info_qr != 0) return c->fail(NCM_SD_GPU_ENOTPD, "nnls: dgels met an exactly rank-deficient passive set (the reference asserts here, ncm_nnls.c:637)");
  }
  NCM_CUDA_OK(c, ncm_memcpy_async(c, w.h_buf, w.drhs, sizeof(double) * np, cudaMemcpyDeviceToHost, c->stream));
  NCM_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return NCM_SD_GPU_OK;
}

// ncm_nnls.c:728-751
int solve_feasible(NnlsWork &w, std::vector<int> &P, std::vector<double> &x, int max_remove) {
  std::vector<int> invalid, ptmp;
  std::vector<double> vals;
  while (true) {
    int rc = solve_unconstrained(w, P);
    if (rc != NCM_SD_GPU_OK) return rc;
    std::fill(x.begin(), x.end(), 0.0);
    for (size_t j = 0; j < P.size(); ++j) x[P[j]] = w.h_buf[j];
    invalid.clear();
    for (int i : P)
      if (x[i] < 1.0e-300) invalid.push_back(i);
    if (invalid.empty()) return NCM_SD_GPU_OK;
    // ncm_iset_remove_smallest_subset, ncm_iset.c:920-985
    const int rsize = (int) invalid.size();
    std::vector<char> drop(w.n, 0);
    if (max_remove >= rsize) {
      for (int i : invalid) drop[i] = 1;
    } else {
      vals.resize(rsize);
      for (int j = 0; j < rsize; ++j) vals[j] = x[invalid[j]];
      sort_smallest_index(ptmp, max_remove, vals, rsize);
      for (int j = 0; j < max_remove; ++j) drop[invalid[ptmp[j]]] = 1;
    }
    std::vector<int> Pn;
    Pn.reserve(P.size());
    for (int i : P)
      if (!drop[i]) Pn.push_back(i);
    P.swap(Pn);
  }
}

// rnorm of x (host) -> residual vector left in dr_out.  g_spec != nullptr: the gradient A^T r of that residual is computed in the same
// submission (ncm_nnls.c:722-726 computes it right after the residuals whenever the step is accepted, which is the rule; when the
// step is rejected the 20 us of device work are dropped) -- one host round trip instead of two, the very same kernels and values.
int compute_residuals(NnlsWork &w, const std::vector<double> &x, double *dr_out, double *rnorm, std::vector<double> *g_spec = nullptr) {
  ncm_sd_gpu_ctx *c = w.c;
  StageTimer t(c, NCM_SD_GPU_T_NNLS_MISC);
  std::memcpy(w.h_buf, x.data(), sizeof(double) * w.n);
  NCM_CUDA_OK(c, ncm_memcpy_async(c,w.dx, w.h_buf, sizeof(double) * w.n, cudaMemcpyHostToDevice, c->stream));
  int nb = 0;
  int rc = residual(c, w.dA, w.lda, w.nrows, w.n, w.dx, w.dF, dr_out, w.dss, &nb);
  if (rc != NCM_SD_GPU_OK) return rc;
  sum_parts_kernel<<<1, 256, 0, c->stream>>>(w.dss, nb, w.dscal);
  c->n_launches++;
  rc = allreduce_sum(c, w.dscal, 1);
  if (rc != NCM_SD_GPU_OK) return rc;
  NCM_CUDA_OK(c, ncm_memcpy_async(c,w.h_buf + w.n, w.dscal, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (g_spec != nullptr) {
    rc = gemv_t(c, w.dA, w.lda, w.nrows, w.n, dr_out, w.dg, c->nn_tmp);
    if (rc != NCM_SD_GPU_OK) return rc;
    rc = allreduce_sum(c, w.dg, w.n);
    if (rc != NCM_SD_GPU_OK) return rc;
    NCM_CUDA_OK(c, ncm_memcpy_async(c, w.h_g, w.dg, sizeof(double) * w.n, cudaMemcpyDeviceToHost, c->stream));
  }
  NCM_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  *rnorm = sqrt(w.h_buf[w.n]);
  if (g_spec != nullptr) g_spec->assign(w.h_g, w.h_g + w.n);
  return NCM_SD_GPU_OK;
}

int compute_mgrad(NnlsWork &w, const double *dr_in, std::vector<double> &g) {
  ncm_sd_gpu_ctx *c = w.c;
  StageTimer t(c, NCM_SD_GPU_T_NNLS_MISC);
  int rc = gemv_t(c, w.dA, w.lda, w.nrows, w.n, dr_in, w.dg, c->nn_tmp);
  if (rc != NCM_SD_GPU_OK) return rc;
  rc = allreduce_sum(c, w.dg, w.n);
  if (rc != NCM_SD_GPU_OK) return rc;
  NCM_CUDA_OK(c, ncm_memcpy_async(c,w.h_buf, w.dg, sizeof(double) * w.n, cudaMemcpyDeviceToHost, c->stream));
  NCM_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  g.assign(w.h_buf, w.h_buf + w.n);
  return NCM_SD_GPU_OK;
}

// ncm_iset_add_largest_subset, ncm_iset.c:994-1077
int add_largest_subset(std::vector<int> &P, int n, const std::vector<double> &v, double min, double add_frac) {
  const int csize = n - (int) P.size();
  if (csize == 0) return 0;
  std::vector<double> vc;
  std::vector<int> ic;
  std::vector<char> in(n, 0);
  for (int i : P) in[i] = 1;
  for (int j = 0; j < n; ++j) {
    if (!in[j] && v[j] > min) {
      vc.push_back(v[j]);
      ic.push_back(j);
    }
  }
  const int k = (int) vc.size();
  double a    = k * add_frac;
  if (a < 1.0) a = 1.0;
  if ((double) k < a) a = (double) k;
  const int adds = (int) a;
  if (adds > 0) {
    std::vector<int> p;
    sort_largest_index(p, adds, vc, k);
    for (int j = 0; j < adds; ++j) P.push_back(ic[p[j]]);
    std::sort(P.begin(), P.end());
  }
  return adds;
}

}   // namespace

int nnls_solve_dev(ncm_sd_gpu_ctx *c, int nrows, int ncols, const double *dA, int lda, const double *dF, double reltol, double *x_host,
                   double *rnorm_host, ncm_sd_gpu_nnls_stats *stats) {
  const int n   = ncols;
  const int ldm = (n + 7) & ~7;
  if (stats) std::memset(stats, 0, sizeof(*stats));
  const size_t mm = (size_t) n * ldm * sizeof(double);
  if (!c->M.reserve(mm) || !c->MU.reserve(mm) || !c->nn_b.reserve((size_t) (4 * n + 64) * sizeof(double)) ||
      !c->nn_x.reserve((size_t) (2 * n + 16) * sizeof(double)) || !c->nn_r.reserve((size_t) (2 * nrows + 16) * sizeof(double)) ||
      !c->nn_g.reserve((size_t) (nrows / 8 + n + 64) * sizeof(double)) || !c->nn_idx.reserve((size_t) (n + 16) * sizeof(int)) ||
      !c->pin_nn.reserve((size_t) (2 * n + 32) * sizeof(double) + (size_t) (3 * (n + 32) + 2 * (lowrank_kmax() + 8)) * sizeof(int)))
    return c->fail(NCM_SD_GPU_ENOMEM, "nnls: out of memory");

  NnlsWork w;
  w.c = c; w.nrows = nrows; w.n = n; w.lda = lda; w.ldm = ldm; w.dA = dA; w.dF = dF; w.st = stats;
  w.dM     = c->M.as<double>();
  w.dMU    = c->MU.as<double>();
  w.db     = c->nn_b.as<double>();
  w.drhs   = w.db + (n + 16);
  w.ddinv  = w.drhs + (n + 16);
  w.dscal  = w.ddinv + (n + 16);
  w.dx     = c->nn_x.as<double>();
  w.dg     = w.dx + (n + 8);
  w.dr     = c->nn_r.as<double>();
  w.dr_try = w.dr + (nrows + 8);
  w.dss    = c->nn_g.as<double>();
  w.didx   = c->nn_idx.as<int>();
  w.dinfo  = w.didx + (n + 8);
  w.h_buf  = c->pin_nn.as<double>();
  w.h_g    = w.h_buf + (n + 16);
  w.h_idx  = reinterpret_cast<int *>(w.h_g + (n + 16));

  w.lr_on = lowrank_enabled() && n >= LR_MIN_N && n <= chol_fused_max_n();
  if (w.lr_on) {
    const int kmax = lowrank_kmax(), kpad = kmax + 8;
    w.ldv          = kpad;
    w.nv           = (n + 16 + 7) & ~7;
    const size_t nv = (size_t) w.nv;
    if (!c->lrW.reserve(mm) || !c->lrWt.reserve(mm) || !c->lrS.reserve(mm) || !c->lrV.reserve((size_t) n * w.ldv * sizeof(double)) ||
        !c->lrT.reserve((size_t) n * w.ldv * sizeof(double)) || !c->lrPart.reserve(lowrank_part_doubles(n, w.ldv) * sizeof(double)) ||
        !c->lrSmall.reserve(((size_t) (kmax + 1) * (kmax + 2) / 2 + 3 * (size_t) kpad + 64) * sizeof(double)) ||
        !c->lrVec.reserve(9 * nv * sizeof(double)) || !c->lrIdx.reserve((3 * nv + 2 * (size_t) kpad + 16) * sizeof(int)))
      return c->fail(NCM_SD_GPU_ENOMEM, "nnls: out of memory");
    LowrankBufs &b = w.lb;
    b.W = c->lrW.as<double>(); b.Wt = c->lrWt.as<double>(); b.S = c->lrS.as<double>(); b.V = c->lrV.as<double>(); b.T = c->lrT.as<double>();
    b.part = c->lrPart.as<double>();
    b.Lg   = c->lrSmall.as<double>();
    b.z    = b.Lg + ((size_t) (kmax + 1) * (kmax + 2) / 2 + kpad);
    b.z2   = b.z + kpad;
    double *v = c->lrVec.as<double>();
    b.y = v; b.xB = v + nv; b.dxB = v + 2 * nv; b.xfull = v + 3 * nv; b.rfull = v + 4 * nv; b.rB = v + 5 * nv; b.tr = v + 6 * nv; b.out = v + 7 * nv; b.tb = v + 8 * nv;
    int *ix = c->lrIdx.as<int>();   // idxA | posD | bsel | psrc are uploaded together; idxB when a base is established
    b.idxA = ix; b.posD = ix + kpad; b.bsel = b.posD + kpad; b.psrc = b.bsel + nv; b.idxB = b.psrc + nv; b.info = b.idxB + nv;
  }
  int rc;
  {
    StageTimer t(c, NCM_SD_GPU_T_SYRK);
    rc = dsyrk_ata_general(c, nrows, n, dA, lda, w.dM, ldm, 1.0, 0.0);
    if (rc != NCM_SD_GPU_OK) return rc;
    if (stats) stats->syrk_flops = (double) nrows * n * n;
  }
  rc = allreduce_sum(c, w.dM, (size_t) n * ldm);   // timed as NCM_SD_GPU_T_COMM
  if (rc != NCM_SD_GPU_OK) return rc;
  if (w.lr_on) {   // the low-rank solves read rows of M[:, A] and take symmetric products: fill the lower triangle once
    StageTimer t(c, NCM_SD_GPU_T_NNLS_MISC);
    rc = symmetrize_upper(c, n, w.dM, ldm);
    if (rc != NCM_SD_GPU_OK) return rc;
  }
  {
    StageTimer t(c, NCM_SD_GPU_T_NNLS_MISC);
    rc = gemv_t(c, dA, lda, nrows, n, dF, w.db, c->nn_tmp);
    if (rc != NCM_SD_GPU_OK) return rc;
    rc = allreduce_sum(c, w.db, n);
    if (rc != NCM_SD_GPU_OK) return rc;
  }

  std::vector<int> P(n), P_try;
  for (int i = 0; i < n; ++i) P[i] = i;
  std::vector<double> x(n, 0.0), x_try(n, 0.0), mgrad;
  double rnorm = 0.0;

  std::vector<double> g_try;
  rc = solve_feasible(w, P, x, n);
  if (rc != NCM_SD_GPU_OK) return rc;
  rc = compute_residuals(w, x, w.dr, &rnorm, &mgrad);
  if (rc != NCM_SD_GPU_OK) return rc;

  while (true) {
    double add_frac = 1.0;
    bool finish     = false, have_g = false;
    double lrnorm   = 0.0;
    while (true) {
      P_try           = P;
      const int added = add_largest_subset(P_try, n, mgrad, rnorm * reltol, add_frac);
      add_frac *= 0.5;
      if (added == 0) {
        finish = true;
        break;
      }
      rc = solve_feasible(w, P_try, x_try, added);
      if (rc != NCM_SD_GPU_OK) return rc;
      rc = compute_residuals(w, x_try, w.dr_try, &lrnorm, &g_try);
      if (rc != NCM_SD_GPU_OK) return rc;
      if (rnorm - lrnorm > rnorm * reltol) {
        P = P_try;
        x = x_try;
        std::swap(w.dr, w.dr_try);
        rnorm = lrnorm;
        have_g = true;   // g_try is A^T of the residual just accepted
        if (stats) stats->n_outer++;
        break;
      }
      if (added == 1) {
        finish = true;
        break;
      }
    }
    if (finish) break;
    if (have_g) {
      mgrad.swap(g_try);
    } else {
      rc = compute_mgrad(w, w.dr, mgrad);
      if (rc != NCM_SD_GPU_OK) return rc;
    }
  }
  if (stats) stats->n_passive = (int) P.size();
  std::memcpy(x_host, x.data(), sizeof(double) * n);
  *rnorm_host = rnorm;
  return NCM_SD_GPU_OK;
}
