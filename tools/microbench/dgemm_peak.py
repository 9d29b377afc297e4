# Measures the FP64 GEMM peak (cuBLAS through torch.matmul) used as the P64 roofline denominator.
import json, time, torch
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda")
b = torch.randn(n, n, dtype=torch.float64, device="cuda")
for _ in range(2):
    torch.matmul(a, b)
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); c = torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(json.dumps({"fp64_dgemm_8192_tflops": 2.0 * n**3 / (best * 1e-3) / 1e12, "ms": best}), flush=True)
# sustained
t0 = time.time(); k = 0
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
while time.time() - t0 < 3.0:
    c = torch.matmul(a, b); k += 1
    if k % 4 == 0: torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(json.dumps({"fp64_dgemm_8192_tflops_sustained": 2.0 * n**3 * k / (ms * 1e-3) / 1e12, "iters": k}), flush=True)
