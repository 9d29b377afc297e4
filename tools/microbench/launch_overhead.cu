// Launch overhead on the box: CPU enqueue cost vs GPU-side dependent-launch gap, stream launches vs CUDA graph.
#include <chrono>
#include <cstdio>
#include <cuda_runtime.h>
__global__ void tiny(int *p) { if (p != nullptr && threadIdx.x == 1000) *p = 1; }
__global__ void spin(long long cycles) { long long t0 = clock64(); while (clock64() - t0 < cycles) {} }
static double now_us() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
  cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int N = 2000;
  for (int rep = 0; rep < 2; rep++) {
    cudaStreamSynchronize(s);
    double c0 = now_us();
    cudaEventRecord(e0, s);
    for (int i = 0; i < N; i++) tiny<<<1, 32, 0, s>>>(nullptr);
    cudaEventRecord(e1, s);
    double c1 = now_us();
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep) printf("{\"stream_launch_cpu_us\": %.2f, \"stream_launch_gpu_us\": %.2f", (c1 - c0) / N, ms * 1000 / N);
  }
  // GPU-side gap when the CPU is far ahead: enqueue behind a long spin kernel
  spin<<<1, 1, 0, s>>>(40000000LL);
  cudaEventRecord(e0, s);
  for (int i = 0; i < N; i++) tiny<<<1, 32, 0, s>>>(nullptr);
  cudaEventRecord(e1, s); cudaEventSynchronize(e1);
  { float ms; cudaEventElapsedTime(&ms, e0, e1); printf(", \"queued_launch_gpu_us\": %.2f", ms * 1000 / N); }
  // CUDA graph of N kernel nodes
  cudaGraph_t g; cudaGraphExec_t ge;
  cudaStreamBeginCapture(s, cudaStreamCaptureModeGlobal);
  for (int i = 0; i < N; i++) tiny<<<1, 32, 0, s>>>(nullptr);
  cudaStreamEndCapture(s, &g);
  double i0 = now_us();
  cudaGraphInstantiate(&ge, g, 0);
  double i1 = now_us();
  cudaGraphLaunch(ge, s); cudaStreamSynchronize(s);
  cudaEventRecord(e0, s); cudaGraphLaunch(ge, s); cudaEventRecord(e1, s); cudaEventSynchronize(e1);
  { float ms; cudaEventElapsedTime(&ms, e0, e1); printf(", \"graph_node_gpu_us\": %.2f, \"graph_instantiate_us_per_node\": %.2f}\n", ms * 1000 / N, (i1 - i0) / N); }
  return 0;
}
