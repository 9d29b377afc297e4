#define CHOL_PROBE
#include "../../numcosmo_b200/csrc/dgemm.cu"
#include "chol_fine.cuh"
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
    chol_diag_kernel<<<1, 256>>>(M, ld, n, 0, rhs, dinv, info); cudaDeviceSynchronize();
    long long p[128]; cudaMemcpyFromSymbol(p, g_probe, sizeof(p));
    printf("rep %d: enter->loaded %lld factor %lld subst %lld store %lld to_barrier %lld barrier %lld | phase(t0) %lld\n", rep, p[40] - p[39], p[41] - p[40], p[42] - p[41], p[43] - p[42], p[44] - p[43], p[45] - p[44], p[2] - p[1]);
  }
  return 0;
}
