#define CHOL_PROBE
#include "../../numcosmo_b200/csrc/dgemm.cu"
#include "../../numcosmo_b200/csrc/chol.cu"
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
  for (int rep = 0; rep < 2; rep++) {
    chol_diag_kernel<<<1, 256>>>(M, ld, n, 0, rhs, dinv, info);
    chol_panel_kernel<<<(n - 64 + 127) / 128, 128>>>(M, ld, n, 0, rhs, dinv); cudaDeviceSynchronize();
    long long p[128]; cudaMemcpyFromSymbol(p, g_probe, sizeof(p));
    printf("rep %d: panel: loadU_issue %lld to_sync %lld |", rep, p[50] - p[49], p[51] - p[50]);
    for (int b = 0; b < 8; b++) printf(" %lld", p[52 + b] - p[51 + b]);
    printf(" | end %lld total %lld\n", p[60] - p[59], p[60] - p[49]);
  }
  return 0;
}
