// Microbenchmarks that fix the FP64 roofline denominators on the box:
//   DFMA issue rate, DMMA.8x8x4 issue rate, both concurrently, FP64 exp() rate,
//   cuBLAS DGEMM 8192^3 (P64), plus device properties.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o fp64_peaks fp64_peaks.cu 
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int ILP>
__global__ void __launch_bounds__(256) k_dfma(double *out, int iters, double a, double b) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += acc[i];
  if (s == 12345.678) out[0] = s;
}

template <int ILP>
__global__ void __launch_bounds__(256) k_dmma(double *out, int iters, double a, double b) {
  double c0[ILP], c1[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) { c0[i] = threadIdx.x + i; c1[i] = i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += c0[i] + c1[i];
  if (s == 12345.678) out[0] = s;
}

// DMMA and DFMA interleaved: does the tensor FP64 path share the DFMA datapath?
template <int ILP>
__global__ void __launch_bounds__(256) k_mix(double *out, int iters, double a, double b) {
  double c0[ILP], c1[ILP], f[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) { c0[i] = threadIdx.x + i; c1[i] = i; f[i] = i + 0.5; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
#pragma unroll
      for (int j = 0; j < 8; j++) f[i] = fma(f[i], a, b);   // 8 DFMA per DMMA = same FMA count per lane
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += c0[i] + c1[i] + f[i];
  if (s == 12345.678) out[0] = s;
}

__global__ void __launch_bounds__(256) k_exp(double *out, int iters, double x0) {
  double x = x0 - 1e-3 * threadIdx.x, s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  for (int it = 0; it < iters; it++) {
    s0 += exp(x); s1 += exp(x - 0.25); s2 += exp(x - 0.5); s3 += exp(x - 0.75);
    x -= 1e-6;
  }
  if (s0 + s1 + s2 + s3 == 12345.678) out[0] = s0;
}

__global__ void __launch_bounds__(256) k_log1p(double *out, int iters, double x0) {
  double x = x0 + 1e-3 * threadIdx.x, s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  for (int it = 0; it < iters; it++) {
    s0 += log1p(x); s1 += log1p(x + 0.25); s2 += log1p(x + 0.5); s3 += log1p(x + 0.75);
    x += 1e-6;
  }
  if (s0 + s1 + s2 + s3 == 12345.678) out[0] = s0;
}

template <typename F>
float time_ms(F f, int reps = 5) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("{\"device\": \"%s\", \"sms\": %d, \"l2_bytes\": %d, \"smem_per_sm\": %zu, \"smem_optin\": %zu, \"clock_khz\": %d, \"cc\": \"%d.%d\", \"global_mem\": %zu}\n",
         p.name, p.multiProcessorCount, p.l2CacheSize, p.sharedMemPerMultiprocessor, p.sharedMemPerBlockOptin, clk, p.major, p.minor, p.totalGlobalMem);
  fflush(stdout);
  double *out; CK(cudaMalloc(&out, 1024));
  const int sms = p.multiProcessorCount;
  const int iters = 4096;
  for (int bps = 1; bps <= 8; bps *= 2) {
    dim3 grid(sms * bps), block(256);
    float t1 = time_ms([&] { k_dfma<8><<<grid, block>>>(out, iters, 1.0000001, 1e-9); });
    double fl1 = 2.0 * 8 * iters * 256.0 * grid.x / (t1 * 1e-3) / 1e12;
    float t2 = time_ms([&] { k_dmma<8><<<grid, block>>>(out, iters, 1.0000001, 1e-9); });
    double fl2 = 2.0 * 256 * 8 * iters * 8.0 * grid.x / (t2 * 1e-3) / 1e12;   // 256 FMA per warp-DMMA, 8 warps
    float t3 = time_ms([&] { k_mix<4><<<grid, block>>>(out, iters, 1.0000001, 1e-9); });
    double fl3 = 2.0 * (256.0 + 8 * 32) * 4 * iters * 8.0 * grid.x / (t3 * 1e-3) / 1e12;
    float t4 = time_ms([&] { k_exp<<<grid, block>>>(out, iters, -0.1); });
    double ex = 4.0 * iters * 256.0 * grid.x / (t4 * 1e-3) / 1e12;
    float t5 = time_ms([&] { k_log1p<<<grid, block>>>(out, iters, 0.1); });
    double lg = 4.0 * iters * 256.0 * grid.x / (t5 * 1e-3) / 1e12;
    printf("{\"blocks_per_sm\": %d, \"dfma_tflops\": %.2f, \"dmma_tflops\": %.2f, \"mix_tflops\": %.2f, \"exp_tera_per_s\": %.4f, \"log1p_tera_per_s\": %.4f}\n",
           bps, fl1, fl2, fl3, ex, lg);
    fflush(stdout);
  }
  return 0;
}
