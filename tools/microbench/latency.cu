// Dependent-chain latencies on the box (cycles per op, single warp): DFMA, DMUL, rsqrt(double), sqrt, 1/x,
// shuffle, shared-memory load, __syncthreads with 8 warps.  Calibrates the latency-bound Cholesky kernels.
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
__global__ void lat(double *out, long long *cyc, double a, double b) {
  __shared__ double sm[256];
  __shared__ int idx[64];
  const int t = threadIdx.x;
  sm[t] = a + t;
  if (t < 64) idx[t] = (t * 17 + 1) & 63;
  __syncthreads();
  double x = a + t * 1e-9;
  long long t0, t1;
  int k = 0;
  // DFMA
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) x = fma(x, a, b);
  t1 = clock64(); if (t == 0) cyc[k] = t1 - t0; k++;
  // DMUL
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) x = x * a;
  t1 = clock64(); if (t == 0) cyc[k] = t1 - t0; k++;
  x = fabs(x) + 1.5;
  // rsqrt
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; i++) x = rsqrt(x) + 1.5;
  t1 = clock64(); if (t == 0) cyc[k] = t1 - t0; k++;
  // sqrt
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; i++) x = sqrt(x) + 1.5;
  t1 = clock64(); if (t == 0) cyc[k] = t1 - t0; k++;
  // reciprocal
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; i++) x = 1.0 / x + 1.5;
  t1 = clock64(); if (t == 0) cyc[k] = t1 - t0; k++;
  // shuffle
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) x = __shfl_sync(0xffffffffu, x, (t + 1) & 31);
  t1 = clock64(); if (t == 0) cyc[k] = t1 - t0; k++;
  // shared load chain (pointer chase)
  int p = t & 63;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) p = idx[p];
  t1 = clock64(); if (t == 0) cyc[k] = t1 - t0; k++;
  // exp
  x = -fabs(x) * 1e-3;
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; i++) x = exp(x) - 1.0;
  t1 = clock64(); if (t == 0) cyc[k] = t1 - t0; k++;
  // syncthreads
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) __syncthreads();
  t1 = clock64(); if (t == 0) cyc[k] = t1 - t0; k++;
  out[t] = x + p;
}
int main() {
  double *out; long long *cyc, h[16];
  cudaMalloc(&out, 4096); cudaMalloc(&cyc, 128);
  const char *names[] = {"dfma", "dmul", "rsqrt+add", "sqrt+add", "rcp+add", "shfl", "lds_chase", "exp+add", "syncthreads"};
  for (int threads : {32, 256}) {
    lat<<<1, threads>>>(out, cyc, 1.0000001, 1e-9); cudaDeviceSynchronize();
    lat<<<1, threads>>>(out, cyc, 1.0000001, 1e-9); cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, 9 * 8, cudaMemcpyDeviceToHost);
    printf("{\"threads\": %d", threads);
    for (int i = 0; i < 9; i++) printf(", \"%s\": %.1f", names[i], (double) h[i] / N);
    printf("}\n");
  }
  // wall-clock of an empty kernel chain: launch+dependency gap
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  for (int i = 0; i < 1000; i++) lat<<<1, 32>>>(out, cyc, 1.0, 0.0);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("{\"kernel_us_each_of_1000_back_to_back\": %.2f}\n", ms);
  return 0;
}
