#include "dgemm_probe.cuh"
#include "chol_probe.cuh"
#include <vector>
bool DevBuf::reserve(size_t) { return false; }
void DevBuf::release() {}
int main() {
  const int n = 2048, ld = 2048;
  std::vector<double> h((size_t) n * ld, 0.0);
  for (int i = 0; i < n; i++) for (int j = i; j < n; j++) h[(size_t) i * ld + j] = (i == j) ? n + 1.0 : 0.5 / (1.0 + j - i);
  double *M, *rhs, *dinv; int *info;
  cudaMalloc(&M, sizeof(double) * n * ld); cudaMalloc(&rhs, sizeof(double) * n); cudaMalloc(&dinv, sizeof(double) * n); cudaMalloc(&info, 4);
  cudaMemcpy(M, h.data(), sizeof(double) * n * ld, cudaMemcpyHostToDevice);
  cudaMemset(rhs, 0, sizeof(double) * n); cudaMemset(info, 0, 4);
  cudaFuncSetAttribute(ata_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) ATA_SMEM);
  int zero = 0;
  for (int rep = 0; rep < 3; rep++) {
    cudaMemcpyToSymbol(g_api, &zero, 4); cudaMemcpyToSymbol(g_ppi, &zero, 4);
    chol_diag_kernel<<<1, 256>>>(M, ld, n, 0, rhs, dinv, info);
    chol_panel_kernel<<<(n - 64 + 127) / 128, 128>>>(M, ld, n, 0, rhs, dinv);
    const int m = n - 64, nt = (m + 127) / 128;
    ata_kernel<<<nt * (nt + 1) / 2, 256, ATA_SMEM>>>(M + 64, ld, 64, m, M + (size_t) 64 * ld + 64, ld, -1.0, 1.0, nt);
    cudaDeviceSynchronize();
    long long a[64], p[64]; cudaMemcpyFromSymbol(a, g_aprobe, sizeof(a)); cudaMemcpyFromSymbol(p, g_pprobe, sizeof(p));
    printf("rep %d: ata: prologue_issue %lld mainloop %lld epilogue %lld | panel: load %lld solve %lld\n", rep, a[1] - a[0], a[2] - a[1], a[3] - a[2], p[1] - p[0], p[2] - p[1]);
  }
  return 0;
}
