// Shared device/host helpers of libncm_sd_gpu (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cmath>

#define NCM_WARP 32
#define NCM_MAX_DEVICES 64
#define NCM_NEG_BIG (-1.0e300)   // running-max seed: finite, so (t - m) never evaluates inf - inf

// ---- shared-memory / mbarrier / bulk-copy (TMA 1-D) primitives ------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}

// global -> shared bulk async copy (UBLKCP), completion counted in bytes on an mbarrier.
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- cp.async (LDGSTS) 16-byte -----------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src_gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async16_zfill(void *dst_smem, const void *src_gmem, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---- FP64 tensor core: DMMA.8x8x4 -----------------------------------------------------------------
// A (8x4, row): lane l holds A[l/4][l%4];  B (4x8, col): lane l holds B[l%4][l/4];
// C (8x8): lane l holds C[l/4][2*(l%4) + {0,1}].
__device__ __forceinline__ void dmma884(double &c0, double &c1, const double a, const double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// ---- exp for non-positive arguments ------------------------------------------------------------------------
// The log-sum-exp terms exp(-|t - m|) and the Gaussian kernel exp(-chi2/2) are the most expensive FP64 operation of
// the path (libdevice exp: ~21 DFMA issue slots measured, tools/microbench/fp64_peaks.cu).  Here: k = rint(x log2 e) by
// the magic-number add, r = (x - k ln 2) / 4 in two FMAs (Cody-Waite), degree-9 Taylor polynomial on |r| <= 0.087
// (truncation 7e-18), two squarings, exponent patched in with integer adds: 15 FP64 operations, relative error <= 9e-16
// (tools/check_fast_exp.py).  Results below 2^-1020 are flushed to zero (they can never reach a sum that contains the
// arg-max term 1, and the kernel-matrix entries are compared relative to their maximum).
__device__ __forceinline__ double exp_nonpos_fast(const double x) {
  const double MAGIC = 6755399441055744.0;   // 1.5 * 2^52
  const double t  = fma(x, 1.4426950408889634074, MAGIC);
  const int k     = __double2loint(t);
  const double kf = t - MAGIC;
  double r = fma(kf, -6.93147180369123816490e-01, x);
  r        = fma(kf, -1.90821492927058770002e-10, r);
  r *= 0.25;
  double p = 2.75573192239858906526e-06;           // 1/9!
  p = fma(p, r, 2.48015873015873015873e-05);        // 1/8!
  p = fma(p, r, 1.98412698412698412698e-04);        // 1/7!
  p = fma(p, r, 1.38888888888888888889e-03);        // 1/6!
  p = fma(p, r, 8.33333333333333333333e-03);        // 1/5!
  p = fma(p, r, 4.16666666666666666667e-02);        // 1/4!
  p = fma(p, r, 1.66666666666666666667e-01);        // 1/3!
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  p *= p;
  p *= p;
  const double v = __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
  return (x > -707.0) ? v : 0.0;   // also covers x = -inf (ln of a zero weight); k >= -1020 whenever x > -707
}

// ---- log1p for non-negative arguments ------------------------------------------------------------------------
// Student-t kernel: ln Kbar = kappa log1p(chi2 / nu).  libdevice log1p costs ~42 DFMA issue slots (measured).  Only the
// ABSOLUTE error of ln Kbar matters (it is added to ln w - ln norm and exponentiated), so u = 1 + x may be rounded:
// u = 2^e m, m in [1, 2); j = top 5 mantissa bits, c_j = 1 + (j + 1/2)/32; r = m / c_j - 1 (one FMA with the tabulated
// 1/c_j, |r| <= 2^-6); ln u = e ln2 + ln c_j + (r - r^2/2 + ... + r^7/7).  14 FP64 operations and one 16-byte table read
// (512-byte table, L1-resident); absolute error <= 4e-15 for x up to 1e18 (tools/check_fast_exp.py).
static __device__ const double NCM_LOG_TAB[64] = {
  9.84615384615384670e-01, 1.55041865359652545e-02,
  9.55223880597014907e-01, 4.58095360312942013e-02,
  9.27536231884057982e-01, 7.52234212375875316e-02,
  9.01408450704225372e-01, 1.03796793681643559e-01,
  8.76712328767123239e-01, 1.31576357788719261e-01,
  8.53333333333333388e-01, 1.58605030176638573e-01,
  8.31168831168831224e-01, 1.84922338494011990e-01,
  8.10126582278481000e-01, 2.10564769107349642e-01,
  7.90123456790123413e-01, 2.35566071312766911e-01,
  7.71084337349397630e-01, 2.59957524436926046e-01,
  7.52941176470588225e-01, 2.83768173130644619e-01,
  7.35632183908045967e-01, 3.07025035294911874e-01,
  7.19101123595505598e-01, 3.29753286372467980e-01,
  7.03296703296703352e-01, 3.51976423157178198e-01,
  6.88172043010752743e-01, 3.73716409793584059e-01,
  6.73684210526315774e-01, 3.94993808240868993e-01,
  6.59793814432989678e-01, 4.15827895143710990e-01,
  6.46464646464646520e-01, 4.36236766774918072e-01,
  6.33663366336633671e-01, 4.56237433481587573e-01,
  6.21359223300970820e-01, 4.75845904869963920e-01,
  6.09523809523809579e-01, 4.95077266797851523e-01,
  5.98130841121495282e-01, 5.13945751102234283e-01,
  5.87155963302752326e-01, 5.32464798869471845e-01,
  5.76576576576576572e-01, 5.50647117952662302e-01,
  5.66371681415929196e-01, 5.68504735352668766e-01,
  5.56521739130434789e-01, 5.86049045003578239e-01,
  5.47008547008547064e-01, 6.03290851438084252e-01,
  5.37815126050420145e-01, 6.20240409751857569e-01,
  5.28925619834710758e-01, 6.36907462237069177e-01,
  5.20325203252032575e-01, 6.53301272012745682e-01,
  5.12000000000000011e-01, 6.69430653942629239e-01,
  5.03937007874015741e-01, 6.85304003098919368e-01,
};

__device__ __forceinline__ double log1p_nonneg_fast(const double x) {
  const double u = 1.0 + x;
  const int hi   = __double2hiint(u);
  const int e    = (hi >> 20) - 1023;                 // u >= 1: sign bit clear
  const int j    = (hi >> 15) & 31;
  const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(u));
  const double2 t = __ldg(reinterpret_cast<const double2 *>(NCM_LOG_TAB) + j);
  const double r  = fma(m, t.x, -1.0);
  double p = 1.0 / 7.0;
  p = fma(p, r, -1.0 / 6.0);
  p = fma(p, r, 0.2);
  p = fma(p, r, -0.25);
  p = fma(p, r, 1.0 / 3.0);
  p = fma(p, r, -0.5);
  p = fma(p, r, 1.0);
  p *= r;
  const double ef = (double) e;
  return fma(ef, 6.93147180369123816490e-01, t.y) + fma(ef, 1.90821492927058770002e-10, p);
}

// ---- log-sum-exp state -----------------------------------------------------------------------------
// (m, s): log-sum = m + log(s), s counts the arg-max term as 1 (the reference returns
// gamma = m and lambda = s - 1, ncm_stats_dist_kernel_gauss.c:246-289).
struct Lse {
  double m, s;
};

__device__ __forceinline__ void lse_init(Lse &a) {
  a.m = NCM_NEG_BIG;
  a.s = 0.0;
}

__device__ __forceinline__ void lse_push(Lse &a, const double t) {
  const double d = t - a.m;
  const double e = exp_nonpos_fast(-fabs(d));
  if (d > 0.0) {
    a.s = fma(a.s, e, 1.0);
    a.m = t;
  } else {
    a.s += e;
  }
}

__device__ __forceinline__ void lse_merge(Lse &a, const double m2, const double s2) {
  const double d = m2 - a.m;
  const double e = exp_nonpos_fast(-fabs(d));
  if (d > 0.0) {
    a.s = fma(a.s, e, s2);
    a.m = m2;
  } else {
    a.s = fma(s2, e, a.s);
  }
}

__device__ __forceinline__ void lse_warp_reduce_xor(Lse &a, const int width) {
  for (int off = width >> 1; off > 0; off >>= 1) {
    const double m2 = __shfl_xor_sync(0xffffffffu, a.m, off);
    const double s2 = __shfl_xor_sync(0xffffffffu, a.s, off);
    lse_merge(a, m2, s2);
  }
}

// kernel function in the log domain: ln Kbar(chi2)   (Appendix C of SURVEY.md)
//   Gauss: -chi2/2 ; ST: kappa * log1p(chi2 / nu), kappa = -(nu + d)/2
struct KernParams {
  int kind;        // 0 Gauss, 1 ST
  double nu;
  double kappa;    // -(nu + d) / 2
  double inv_nu;
  int m2;          // ST with integer nu: nu + d, so that Kbar = rsqrt(1 + chi2/nu)^m2 (0: use exp(kappa log1p))
  int lin;         // ST eval in the linear domain: sum_i exp(c_i - cmax) Kbar_i, no exp / log per pair (needs m2 > 0)
  const double *cmax;   // device scalar: max_i c_i of the current weights (lin)
};

// Student-t kernel with an integer number of degrees of freedom: (1 + chi2/nu)^(-(nu + d)/2) = r^(nu + d), r = rsqrt(1 + chi2/nu).
// One rsqrt and at most 2 log2(m) multiplications (m is uniform over the grid) instead of log1p + exp: relative error about 2 m x 2^-52
// with a 1-ulp rsqrt (1.6e-14 at m = 35; tests/test_st_integer_power_rule.py), the same order as the table-based log1p above.  Reference: pow (1 + chi2/nu, kappa), ncm_stats_dist_kernel_st.c:239-243.
__device__ __forceinline__ double st_pow_u(const KernParams &kp, const double u) {   // u = chi2 / nu >= 0
  // straight-line binary powering (m2 < 128): six squarings, the factors picked by the bits of m2 -- selects on a grid-uniform value,
  // no loop, so that the calls of an unrolled epilogue interleave (a loop per call serialised them: slower than log1p + exp)
  const double b1 = rsqrt(1.0 + u), b2 = b1 * b1, b4 = b2 * b2, b8 = b4 * b4, b16 = b8 * b8, b32 = b16 * b16, b64 = b32 * b32;
  const int m = kp.m2;
  const double lo = ((m & 1) ? b1 : 1.0) * ((m & 2) ? b2 : 1.0), mid = ((m & 4) ? b4 : 1.0) * ((m & 8) ? b8 : 1.0);
  const double hi = ((m & 16) ? b16 : 1.0) * ((m & 32) ? b32 : 1.0) * ((m & 64) ? b64 : 1.0);
  return (lo * mid) * hi;
}
__device__ __forceinline__ double st_pow_int(const KernParams &kp, const double chi2) { return st_pow_u(kp, fmax(chi2, 0.0) * kp.inv_nu); }

__device__ __forceinline__ double kern_lnK(const KernParams &kp, const double chi2) {
  return kp.kind == 0 ? -0.5 * chi2 : kp.kappa * log1p_nonneg_fast(chi2 * kp.inv_nu);
}
// Kbar(chi2) as the reference evaluates it (eval_unnorm): exp(-chi2/2) or pow(1 + chi2/nu, kappa)
__device__ __forceinline__ double kern_K(const KernParams &kp, const double chi2) {
  if (kp.kind == 0) return exp_nonpos_fast(-0.5 * chi2);
  if (kp.m2 > 0) return st_pow_int(kp, chi2);
  return exp_nonpos_fast(kp.kappa * log1p_nonneg_fast(chi2 * kp.inv_nu));   // kappa < 0
}

// Linear-domain partial of one (query, centre range): the caller stores (m, s) = (cmax, sum) so that the log-sum-exp merge of the
// splits and lse_finalize apply unchanged (equal m: the sums add).
__device__ __forceinline__ void lin_push(Lse &a, const KernParams &kp, const double chi2, const double cw) { a.s = fma(cw, st_pow_int(kp, chi2), a.s); }
