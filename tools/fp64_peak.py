#!/usr/bin/env python
"""FP64 GEMM peak of this GPU through cuBLAS (torch.matmul, 8192^3, best of 10 and a 3 s sustained loop) -- the library
denominator reported beside tools/dmma_peak.cu's register-loop DMMA ceiling for gemm_tall_kernel."""
import json
import time

import torch

n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda")
b = torch.randn(n, n, dtype=torch.float64, device="cuda")
c = torch.empty_like(a)
for _ in range(3):
    torch.matmul(a, b, out=c)
torch.cuda.synchronize()
best = 1e30
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
t0 = time.perf_counter(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); reps = 0
while time.perf_counter() - t0 < 3.0:
    torch.matmul(a, b, out=c); reps += 1
    if reps % 8 == 0:
        torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
print(json.dumps({"cublas_dgemm_8192_burst_tflops": 2.0 * n ** 3 / best / 1e9,
                  "cublas_dgemm_8192_sustained_tflops": 2.0 * n ** 3 * reps / e0.elapsed_time(e1) / 1e9, "reps": reps}))
