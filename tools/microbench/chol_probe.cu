// Times the individual kernels of csrc/chol.cu and csrc/dgemm.cu (included for access to the static kernels).
#include "../../numcosmo_b200/csrc/dgemm.cu"
#include "../../numcosmo_b200/csrc/chol.cu"
#include <vector>
bool DevBuf::reserve(size_t) { return false; }
void DevBuf::release() {}
template <typename F> float tm(F f, int reps) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); for (int i = 0; i < reps; i++) f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms * 1000.f / reps;
}
int main() {
  const int n = 2048, ld = 2048;
  std::vector<double> h((size_t) n * ld, 0.0);
  for (int i = 0; i < n; i++) for (int j = i; j < n; j++) h[(size_t) i * ld + j] = (i == j) ? n + 1.0 : 0.5 / (1.0 + j - i);
  double *M, *rhs, *dinv; int *info;
  cudaMalloc(&M, sizeof(double) * n * ld); cudaMalloc(&rhs, sizeof(double) * n); cudaMalloc(&dinv, sizeof(double) * n); cudaMalloc(&info, 4);
  cudaMemcpy(M, h.data(), sizeof(double) * n * ld, cudaMemcpyHostToDevice);
  cudaMemset(rhs, 0, sizeof(double) * n); cudaMemset(info, 0, 4);
  ncm_sd_gpu_ctx c; c.stream = 0;
  
  printf("{\"diag_us\": %.2f", tm([&] { chol_diag_kernel<<<1, 256>>>(M, ld, n, 0, rhs, dinv, info); }, 200));
  printf(", \"diag_norhs_us\": %.2f", tm([&] { chol_diag_kernel<<<1, 256>>>(M, ld, n, 0, nullptr, dinv, info); }, 200));
  printf(", \"panel_us\": %.2f", tm([&] { chol_panel_kernel<<<(n - 64 + 127) / 128, 128>>>(M, ld, n, 0, rhs, dinv); }, 200));
  for (int m : {1984, 1024, 256}) {
        printf(", \"ata_k64_m%d_us\": %.2f", m, tm([&] { dsyrk_ata_general(&c, 64, m, M + 64, ld, M + (size_t) 64 * ld + 64, ld, -1.0, 1.0); }, 200));
  }
  printf(", \"backsolve_us\": %.2f", tm([&] { chol_backsolve_kernel<<<1 + (1024 + 7) / 8, 256>>>(M, ld, n, 1024, rhs, dinv); }, 200));
  printf(", \"empty_launch_us\": %.2f", tm([&] { chol_backsolve_kernel<<<1, 32>>>(M, ld, 0, 0, rhs, dinv); }, 1000));
  printf("}\n");
  return 0;
}
