// Pieces shared by the cooperative (one CTA per SM, all co-resident) fallback solvers ldl_bk.cu and qr_ls.cu.
#pragma once
#include <cuda_runtime.h>

namespace ncm_coop {

constexpr int COOP_T = 512;   // threads per CTA

// grid-wide barrier on a monotonic counter (all CTAs co-resident: cooperative launch); ONE CTA: a plain CTA barrier
struct GridSync {
  unsigned int *count;
  unsigned int target;
  unsigned int nctas;
  __device__ __forceinline__ void operator()() {
    if (nctas == 1) {
      __syncthreads();
      return;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      target += nctas;
      __threadfence();
      atomicAdd(count, 1u);
      unsigned int v;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(count) : "memory");
      } while ((int) (v - target) < 0);
    }
    __syncthreads();
  }
};

// (max |v|, first index attaining it) over the CTA; idamax semantics (first occurrence wins); NaN poisons the maximum
struct AbsMax {
  double v;
  int i;
};
__device__ __forceinline__ AbsMax absmax_merge(AbsMax a, AbsMax b) {
  if (b.i < 0) return a;
  if (a.i < 0) return b;
  if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
  return a;
}
__device__ inline AbsMax cta_absmax(AbsMax m, AbsMax *red) {
  for (int off = 16; off > 0; off >>= 1) {
    AbsMax o;
    o.v = __shfl_xor_sync(0xffffffffu, m.v, off);
    o.i = __shfl_xor_sync(0xffffffffu, m.i, off);
    m   = absmax_merge(m, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();   // red may still be read from the previous call
  if (lane == 0) red[warp] = m;
  __syncthreads();
  AbsMax r = red[0];
  for (int w = 1; w < COOP_T / 32; ++w) r = absmax_merge(r, red[w]);
  return r;
}
__device__ inline double cta_sum(double s, double *red) {
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = s;
  __syncthreads();
  double r = 0.0;
  for (int w = 0; w < COOP_T / 32; ++w) r += red[w];
  return r;
}


}   // namespace ncm_coop
