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
  for (int rep = 0; rep < 3; rep++) {
    chol_diag_kernel<<<1, 256>>>(M, ld, n, 0, rhs, dinv, info); cudaDeviceSynchronize();
    long long p[128]; cudaMemcpyFromSymbol(p, g_probe, sizeof(p));
    printf("rep %d: load %lld |", rep, p[1] - p[0]);
    for (int kb = 0; kb < 8; kb++) printf(" [%lld %lld]", p[2 + 2 * kb] - p[1 + 2 * kb], p[3 + 2 * kb] - p[2 + 2 * kb]);
    printf(" | store %lld total %lld\n", p[18] - p[17], p[18] - p[0]);
  }
  return 0;
}
