// How fast can one warp run the 8 x 8 Cholesky pivot chain?  (cycles per 8 x 8 factorisation, clock64)
#include <cstdio>
#include <cuda_runtime.h>
template <int VAR>
__global__ void k(const double *in, double *out, long long *cyc, int iters) {
  double a[8][8];
  for (int r = 0; r < 8; ++r) for (int q = 0; q < 8; ++q) a[r][q] = in[r * 8 + q] + threadIdx.x * 1e-9;
  double acc = 0.0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    double d[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int q = r; q < 8; ++q) d[r][q] = a[r][q] + acc * 1e-30;
    double inv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      double piv = d[j][j];
      if (VAR == 0) {
        if (!(piv > 0.0)) piv = 1.0;
        inv[j] = rsqrt(piv);
      } else if (VAR == 1) {
        inv[j] = rsqrt(piv);
      } else {
        // float seed + 2 Newton steps in double, no special cases
        const float s = rsqrtf((float) piv);
        double y = (double) s;
        y = y * (1.5 - 0.5 * piv * y * y);
        y = y * (1.5 - 0.5 * piv * y * y);
        inv[j] = y;
      }
      d[j][j] = piv * inv[j];
#pragma unroll
      for (int q = j + 1; q < 8; ++q) d[j][q] *= inv[j];
#pragma unroll
      for (int r = j + 1; r < 8; ++r)
#pragma unroll
        for (int q = r; q < 8; ++q) d[r][q] = fma(-d[j][r], d[j][q], d[r][q]);
    }
    acc += d[7][7] + inv[3];
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[VAR] = (t1 - t0) / iters;
  out[threadIdx.x] = acc;
}
int main() {
  double h[64]; for (int r = 0; r < 8; r++) for (int q = 0; q < 8; q++) h[r * 8 + q] = (r == q) ? 10.0 + r : 0.5 / (1 + abs(r - q));
  double *in, *out; long long *cyc, hc[4];
  cudaMalloc(&in, 512); cudaMalloc(&out, 4096); cudaMalloc(&cyc, 64);
  cudaMemcpy(in, h, 512, cudaMemcpyHostToDevice);
  for (int threads : {32, 96}) {
    k<0><<<1, threads>>>(in, out, cyc, 200); k<1><<<1, threads>>>(in, out, cyc, 200); k<2><<<1, threads>>>(in, out, cyc, 200);
    cudaDeviceSynchronize();
    cudaMemcpy(hc, cyc, 24, cudaMemcpyDeviceToHost);
    printf("{\"threads\": %d, \"chol8_checked\": %lld, \"chol8_rsqrt\": %lld, \"chol8_floatseed\": %lld}\n", threads, hc[0], hc[1], hc[2]);
  }
  return 0;
}
